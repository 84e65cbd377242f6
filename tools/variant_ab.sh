#!/bin/bash
# usage: tools/variant_ab.sh <workload> lib1.so ...   (value and per-kernel fractions for the default build and each variant)
wl=$1; shift
for rep in 1 2; do
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset ADMM_B200_LIB; else export ADMM_B200_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-cfg5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-28s %s value %.0f  ms/step %.3f  rows %.3f cols %.3f iter %.3f  mhz %s' % ('$lib', '$wl', d['value'], d['ms_per_step'], r.get('frac_rows') or 0, r.get('frac_cols') or 0, r.get('frac') or 0, d['clocks']['sm_mhz']))"
done
done
