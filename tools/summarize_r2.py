"""Copy the artefacts of one tools/gpu_round2.sh pass (gpurun_out/<tag>/) into profiles/ and derive the round-2 summaries:
per-class table, launch list with DRAM traffic (the `traffic` field of bench.py), per-class ncu metrics, stall tables of four
captured launches, SASS opcode histogram.   usage: python tools/summarize_r2.py gpurun_out/<tag>"""
import collections, csv, gzip, json, os, re, shutil, subprocess, sys

R = sys.argv[1]
P = "profiles"
for src, dst in [("bench_uncond.json", "r2_bench_n1.json"), ("bench_reference.json", "r2_bench_reference_n1.json"),
                 ("bench_vqvae.json", "r2_bench_vqvae_n1.json"), ("bench_guided.json", "r2_bench_guided_n1.json"),
                 ("bench_uncond_fast128.json", "r2_bench_uncond_fast128_n1.json"), ("bench_uncond_fast.json", "r2_bench_uncond_fast_n1.json"),
                 ("bench_uncond_b1.json", "r2_bench_uncond_batch1.json"), ("bench_uncond_b4.json", "r2_bench_uncond_batch4.json"),
                 ("op_profile.txt", "r2_op_profile.txt"), ("op_profile_unet32_b32.txt", "r2_op_profile_unet32_b32.txt"),
                 ("gpu.txt", "r2_gpu.txt"), ("pytest_gpu.txt", "r2_pytest_gpu.txt"), ("smoke.txt", "r2_smoke.txt"),
                 ("classes.log", "r2_ncu_classes.log")]:
    if os.path.exists(f"{R}/{src}"):
        shutil.copy(f"{R}/{src}", f"{P}/{dst}")

# ---- per-class table from the per-op profile (CUDA events between launches, bench.profile_kernels) ----
rows = []
for l in open(f"{R}/op_profile.txt"):
    m = re.match(r'\s*(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d)\s+(\d+)\s+(\d)\s+(\d+)\s+\|\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)%', l)
    if m:
        rows.append([float(x) for x in m.groups()])
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6464.3
tot = sum(r[8] * r[9] for r in rows)
cls = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for cin, cout, tin, tout, k, d, skip, cs, n, ms, gbs, tf, share in rows:
    key = f"C_out={int(cout)}" if tout > 250 else "T<=250 (all C=512)"
    e = cls[key]
    e[0] += n * ms; e[1] += gbs * ms * n; e[2] += tf * ms * n; e[3] += int(n)
table = {k: {"launches": v[3], "ms_per_unet_step": round(v[0], 3), "share_of_conv_time": round(v[0] / tot, 4),
             "algorithmic_GBps": round(v[1] / v[0], 1), "frac_of_hbm_peak": round(v[1] / v[0] / peak, 3),
             "fp32_equivalent_TFLOPs": round(v[2] / v[0], 1)} for k, v in sorted(cls.items(), key=lambda kv: -kv[1][0])}
json.dump({"source": "tools/op_profile.py (unet64, batch 64, T = 64000): CUDA events between the launches of one UNet step",
           "conv_ms_per_unet_step": round(tot, 3), "hbm_peak_GBps": peak,
           "note": "tensor-core work = 3 x fp32-equivalent FLOPs for the bf16x3 layers (C_out = 64, and C_out = 128 at T >= 8000), 1 x for the fp16 ones (C_out >= 256, C_out = 128 at T = 4000)",
           "classes": table}, open(f"{P}/r2_class_table.json", "w"), indent=1)

# ---- launch list: duration + DRAM bytes of every launch of one sampler call with one diffusion step ----
hdr, per = None, collections.defaultdict(dict)
for r in csv.reader(open(f"{R}/launches.csv", errors="replace")):
    if len(r) < 6:
        continue
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        val = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = d["Metric Unit"]
    scale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per[(int(d["ID"]), d["Kernel Name"].split("(")[0])][d["Metric Name"]] = val * scale
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (_, name), m in per.items():
    a = agg[name]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tt = sum(v[1] for v in agg.values())
umma = [v for k, v in agg.items() if "conv_umma" in k]
n_umma, b_umma = sum(v[0] for v in umma), sum(v[2] for v in umma)
json.dump({
    "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py "
              "--steps 1 --warmup 1 --diffusion-steps 1 (every launch of two sampler calls of one diffusion step each)",
    "note": "per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
    "total_us": round(tt, 1), "conv_umma_launches": n_umma,
    "conv_umma_dram_bytes_per_launch_avg": b_umma / max(n_umma, 1),
    "kernels": {k: {"launches": v[0], "total_us": round(v[1], 1), "share": round(v[1] / tt, 4), "dram_MB_per_launch": round(v[2] / v[0] / 1e6, 2)}
                for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
}, open(f"{P}/r2_dram_traffic.json", "w"), indent=1)

# ---- full captures per shape class ----
raw = list(csv.reader(open(f"{R}/classes_raw.csv", errors="replace")))
hdr, units = raw[0], raw[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"] + [
    f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio"
    for k in ("no_instruction", "long_scoreboard", "wait", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "barrier", "not_selected")]
names = [l.split()[1] for l in open(f"{R}/classes.log") if l.startswith("class")]
caps = []
for i, r in enumerate(raw[2:]):
    e = {"class": names[i // 2] if i // 2 < len(names) else "?", "conv": "conv1" if i % 2 == 0 else "conv2 (+skip)",
         "kernel": r[hdr.index("Kernel Name")][:60]}
    for h, v, u in zip(hdr, r, units):
        if h in WANT:
            e[h] = f"{v} {u}".strip()
    caps.append(e)
json.dump({"source": "ncu --set full --clock-control none --profile-from-start off -k regex:conv_umma python tools/ncu_classes.py --only ... "
                     "(one ResBlock per shape class at batch 64: launch 2k = conv1, 2k+1 = conv2)", "captures": caps},
          open(f"{P}/r2_ncu_classes.json", "w"), indent=1)

# ---- stall tables of the captured launches with a SASS source page ----
def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0
with open(f"{P}/r2_ncu_stalls.txt", "w") as out:
    for i in (0, 2, 6, 8):
        path = f"{R}/source_{i}.csv.gz"
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(gzip.open(path, "rt")))
        name, h, data = rows[0][0], rows[1], rows[2:]
        ix = {k: j for j, k in enumerate(h)}
        tot_s = sum(f(r[ix["# Samples"]]) for r in data)
        out.write(f"== launch {i} ({names[i // 2]}, conv1): {name[:70]}\n   warp samples {int(tot_s)}; by stall reason: ")
        out.write(", ".join(f"{k[6:]} {100 * sum(f(r[ix[k]]) for r in data) / tot_s:.1f}%" for k in h if k.startswith("stall_")) + "\n")
        # role attribution by opcode neighbourhood: 2 KB code buckets labelled by their dominant opcodes
        base = int(data[0][ix["Address"]], 16)
        buckets = collections.OrderedDict()
        for r in data:
            b = (int(r[ix["Address"]], 16) - base) // 0x1000
            e = buckets.setdefault(b, [0.0, 0.0, collections.Counter()])
            e[0] += f(r[ix["# Samples"]]); e[1] += f(r[ix["Instructions Executed"]])
            op = r[ix["Source"]].strip().split()
            if op:
                e[2][(op[1] if op[0].startswith("@") and len(op) > 1 else op[0]).split(".")[0]] += f(r[ix["Instructions Executed"]])
        out.write("   4 KB code buckets with > 2 % of the samples (offset, % samples, warp instructions executed, dominant opcodes):\n")
        for b, (s, n, ops) in buckets.items():
            if s > 0.02 * tot_s:
                out.write(f"     +0x{b * 0x1000:05x}  {100 * s / tot_s:5.1f}%  {int(n):>11}  {' '.join(k for k, _ in ops.most_common(4))}\n")
        out.write("   top instructions by samples:\n")
        for r in sorted(data, key=lambda r: -f(r[ix["# Samples"]]))[:12]:
            st = {k[6:]: int(f(r[ix[k]])) for k in h if k.startswith("stall_") and f(r[ix[k]]) > 0.1 * f(r[ix["# Samples"]])}
            out.write(f"     {int(f(r[ix['# Samples']])):>6}  {r[ix['Source']].strip()[:70]:70s} {st}\n")
        out.write("\n")

# ---- SASS opcode histogram of the shipped library ----
sass = subprocess.run(["cuobjdump", "-sass", "vq_voice_swap_b200/libvqvs.so"], capture_output=True, text=True).stdout
ops = collections.Counter()
for l in sass.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(1)] += 1
with open(f"{P}/r2_sass_histogram.txt", "w") as out:
    out.write("cuobjdump -sass vq_voice_swap_b200/libvqvs.so: opcode counts (all kernels); Blackwell-specific first\n")
    key = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "UTMAPF", "FFMA2", "FMUL2", "FADD2", "ACQBULK", "UTCATOMSWS", "F2FP")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1]):
        if k.startswith(key):
            out.write(f"{v:>8}  {k}\n")
    out.write("-- everything, by count --\n")
    for k, v in ops.most_common(80):
        out.write(f"{v:>8}  {k}\n")
print("profiles/ refreshed from", R, "| conv_umma dram bytes per launch (launch-weighted over", n_umma, "launches):", round(b_umma / max(n_umma, 1)))
