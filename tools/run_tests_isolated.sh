#!/bin/bash
# Run each GPU test id in its own process under a hard kill timeout (a hung kernel must not eat the GPU budget).
ids=$(python -m pytest tests/test_gpu_parity.py --collect-only -q -k "${1:-umma}" 2>/dev/null | grep "::")
for id in $ids; do
  timeout -s KILL ${2:-40} python -m pytest "$id" -q -x --timeout 30 > /tmp/one.log 2>&1
  rc=$?
  echo "rc=$rc $id $(tail -1 /tmp/one.log | cut -c1-80)"
done
