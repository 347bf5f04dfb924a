#!/bin/bash
run() { "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-eager --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks']['sm_mhz'], d['roofline']['whole_path_frac'])"; }
echo "def: $(run env)"
echo "old mt rule: $(run env VQVS_MT_RULE_R1=1)"
echo "def: $(run env)"
echo "old mt rule: $(run env VQVS_MT_RULE_R1=1)"
echo "stack64: $(run env VQVS_STACK64=1)"
