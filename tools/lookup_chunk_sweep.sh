#!/bin/bash
# in-situ lookup time (bench.py's roofline leg) for several dynamic-claim chunk sizes
for c in 1 2 3 4 6 8; do
  PVRAFT_LOOKUP_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c', round(d['roofline']['avg_launch_ms']*1e3,1),'us  frac', round(d['roofline']['frac'],3), ' ms/step', round(d['ms_per_step'],2))"
done
