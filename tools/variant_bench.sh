#!/bin/bash
# usage: tools/variant_bench.sh lib1.so lib2.so ...   (runs cfg2 and cfg5 for each, prints value and row/col kernel ms)
for lib in default "$@"; do
  for wl in cfg2 cfg5; do
    if [ "$lib" = default ]; then unset ADMM_B200_LIB; else export ADMM_B200_LIB=$PWD/$lib; fi
    timeout 200 python bench.py --workload $wl --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']
print('%-45s %s value %.0f  rows %.2f cols %.2f  whole %.3f' % ('$lib', '$wl', d['value'], k['rows'], k['cols'], d['roofline']['whole_iteration']['frac']))"
  done
done
