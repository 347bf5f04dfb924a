#!/bin/bash
# One GPU-box pass producing what gets committed under profiles/ for round 2: tests, smoke, the three bench configs, the
# reference arm, small-batch latency, per-op profile, launch list + DRAM traffic of every launch of one UNet step, and a
# full ncu capture of eight conv shape classes.   usage: tools/gpu_round2.sh <tag>
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/gpu.txt
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 > $out/pytest_gpu.txt 2>&1; tail -2 $out/pytest_gpu.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; tail -2 $out/smoke.txt
timeout -s KILL 900 python bench.py > $out/bench_uncond.json 2> $out/bench_uncond.err; cut -c1-200 $out/bench_uncond.json
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-200 $out/bench_reference.json
for c in vqvae guided; do
  timeout -s KILL 900 python bench.py --config $c --steps 2 --warmup 2 --no-eager --no-cpu-baseline > $out/bench_$c.json 2> $out/bench_$c.err; cut -c1-160 $out/bench_$c.json
done
for pr in fast128 fast; do
  timeout -s KILL 600 python bench.py --precision $pr --steps 2 --warmup 3 --no-eager --no-cpu-baseline > $out/bench_uncond_$pr.json 2> $out/bench_$pr.err; cut -c1-120 $out/bench_uncond_$pr.json
done
for b in 1 4; do
  timeout -s KILL 600 python bench.py --batch $b --steps 3 --warmup 2 --no-eager --no-cpu-baseline > $out/bench_uncond_b$b.json 2> $out/bench_b$b.err
done
timeout -s KILL 300 python tools/op_profile.py > $out/op_profile.txt 2>&1
timeout -s KILL 300 python tools/op_profile.py --bc 32 --batch 32 > $out/op_profile_unet32_b32.txt 2>&1
# every launch of one sampler call with one diffusion step: duration + DRAM bytes (cold-cache, serialised: compare shares)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $out/launches.csv python bench.py --steps 1 --warmup 1 --diffusion-steps 1 --no-cpu-baseline --no-eager > $out/ncu_launches.log 2>&1
# full capture of eight shape classes (two launches each)
timeout -s KILL 1200 ncu --set full --clock-control none --profile-from-start off -k regex:conv_umma -o $out/classes -f \
    python tools/ncu_classes.py --only c64_l0,c128to64_l0,c64_down_l0,c128_l2,c256to128_l2,c128_up_l2,c256_l5,c512_l8 > $out/classes.log 2>&1
grep '^class' $out/classes.log
ncu -i $out/classes.ncu-rep --page raw --csv > $out/classes_raw.csv 2>/dev/null
for i in 0 2 6 8; do
  ncu -i $out/classes.ncu-rep --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | python tools/ncu_source_filter.py | gzip > $out/source_$i.csv.gz
done
rm -f $out/classes.ncu-rep
ls -la $out
