#!/bin/bash
# value / e2e of bench.py for 1..4 split streams (ADMM_B200_SPLIT_STREAMS), cfg2 and the cfg5 shard
for wl in cfg2 cfg5; do
  for n in 1 2 3 4; do
    ADMM_B200_SPLIT_STREAMS=$n python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline --no-cfg5 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl streams=$n value %.0f e2e %.0f sm_mhz %s' % (d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))"
  done
done
