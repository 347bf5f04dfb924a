#!/bin/bash
# One GPU-box pass producing everything that gets committed under profiles/: tests, smoke, bench lines,
# ncu launch list and a full ncu capture of the dominant kernel.  usage: tools/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/gpu.txt
timeout -s KILL 600 python -m pytest tests -q -m gpu --timeout 120 > $out/pytest_gpu.txt 2>&1; tail -3 $out/pytest_gpu.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; tail -2 $out/smoke.txt
timeout -s KILL 900 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 600 $out/bench_reference.json
timeout -s KILL 300 python tools/op_profile.py > $out/op_profile.txt 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 64 --diffusion-steps 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 10 -c 3 -o $out/conv_umma \
    python bench.py --steps 1 --warmup 1 --batch 64 --diffusion-steps 1 --no-cpu-baseline > $out/ncu_full.log 2>&1
ls -la $out
