"""Copy the artefacts of one tools/gpu_round.sh pass (gpurun_out/<tag>/) into profiles/ and derive the summaries the
bench line and DESIGN.md cite: per-kernel launch shares and the ncu metrics of the dominant kernel.
usage: python tools/summarize_profiles.py gpurun_out/<tag>"""
import collections, csv, json, shutil, subprocess, sys

R = sys.argv[1]
for src, dst in [("bench.json", "r1_bench_n1.json"), ("bench_reference.json", "r1_bench_reference_n1.json"),
                 ("op_profile.txt", "r1_op_profile.txt"), ("gpu.txt", "r1_gpu.txt"), ("launches.csv", "r1_launch_list.csv")]:
    shutil.copy(f"{R}/{src}", f"profiles/{dst}")

# ---- launch list: share of the step per kernel (ncu times are cold-cache and serialised: compare shares) ----
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in csv.reader(open(f"{R}/launches.csv")):
    if len(r) < 6:
        continue
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val, unit = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
    us = val / 1000.0 if unit.startswith("n") else val * 1000.0 if unit.startswith("m") else val
    name = d["Kernel Name"].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
json.dump({
    "source": "ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 python bench.py --steps 1 --warmup 1 "
              "--batch 64 --diffusion-steps 3 --no-cpu-baseline",
    "note": "per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
    "total_us": round(tot, 1),
    "kernels": {k: {"launches": v[0], "total_us": round(v[1], 1), "share": round(v[1] / tot, 4)}
                for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
}, open("profiles/r1_launch_list_summary.json", "w"), indent=1)

# ---- full capture of the dominant kernel ----
raw = subprocess.run(["ncu", "-i", f"{R}/conv_umma.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warp_latency_per_inst_issued.ratio"] + [
    f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio"
    for k in ("no_instruction", "long_scoreboard", "wait", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "barrier", "not_selected")]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
caps, traffic = [], []
i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
for r in rows[2:]:
    caps.append({h: f"{v} {units[i]}" for i, (h, v) in enumerate(zip(hdr, r)) if h in WANT or h == "Kernel Name"})
    traffic.append(float(r[i_r].replace(",", "")) * SCALE[units[i_r]] + float(r[i_w].replace(",", "")) * SCALE[units[i_w]])
json.dump({
    "source": "ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 10 -c 3 python bench.py --steps 1 "
              "--warmup 1 --batch 64 --diffusion-steps 1 --no-cpu-baseline",
    "captures": caps, "conv_umma_dram_bytes_per_launch_avg": sum(traffic) / len(traffic),
}, open("profiles/ncu_summary.json", "w"), indent=1)
open("profiles/r1_conv_umma_ncu_details.csv", "w").write(
    subprocess.run(["ncu", "-i", f"{R}/conv_umma.ncu-rep", "--page", "details", "--csv"], capture_output=True, text=True).stdout)
print("profiles/ refreshed from", R, "| conv_umma dram bytes per launch:", round(sum(traffic) / len(traffic)))
