"""Attribute an ncu SASS-level source page to CUDA source lines using nvdisasm -g line markers.
usage: ncu_by_line.py <report.ncu-rep> <cubin-disasm.txt> <kernel-substr> [launch-index]"""
import csv, re, subprocess, sys, collections
rep, dis, kern = sys.argv[1:4]
launch = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", launch, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
half = len(data) // 2 if len(data) > 2 and data[0][ix["Source"]] == data[len(data)//2][ix["Source"]] else len(data)
data = data[:half] if half != len(data) and all(data[i][ix["Source"]] == data[i+half][ix["Source"]] for i in range(0, half, max(1, half//50))) else data
# dedupe consecutive duplicate rows (ncu prints each instruction twice in this csv)
ded = []
for r in data:
    if ded and r[ix["Address"]] == ded[-1][ix["Address"]]:
        continue
    ded.append(r)
data = ded
# parse disasm: sequence of (line, instr) for the kernel
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
seq = []; cur = None
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text.") : break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
print("sass rows", len(data), "disasm instrs", len(seq))
def f(r, k):
    try: return float(r[ix[k]])
    except: return 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0])
n = min(len(data), len(seq))
for r, (loc, ins) in zip(data[:n], seq[:n]):
    key = loc[:2] if loc else ("?", 0)
    agg[key][0] += f(r, "Instructions Executed"); agg[key][1] += f(r, "# Samples")
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = {}
for (fn, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    if fn not in src:
        try: src[fn] = open(f"/root/repo/vq_voice_swap_b200/csrc/{fn}").read().splitlines()
        except Exception: src[fn] = []
    text = src[fn][ln - 1].strip()[:90] if 0 < ln <= len(src[fn]) else ""
    print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100 * v[0] / ti, 100 * v[1] / ts, fn, ln, text))
