#!/bin/bash
run() { "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-eager --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks']['sm_mhz'], d['roofline']['whole_path_frac'])"; }
timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -2
echo "def(unstacked64+coll): $(run env)"
echo "stack64: $(run env VQVS_STACK64=1)"
echo "def: $(run env)"
echo "stack64: $(run env VQVS_STACK64=1)"
timeout 300 python tools/op_profile.py > gpurun_out/op_x7.txt 2>&1; sed -n 1,3p gpurun_out/op_x7.txt
