#!/bin/bash
# quick bench lines: qb.sh wl [wl...]
for wl in "$@"; do
python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-e2e --also none 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read()
try:
    d=json.loads(l); print('$wl %.3e pairs/s %.2f ms'%(d['value'], d['ms_per_step']), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['roofline'].get('tensor_stats'), d.get('index_agreement',{}).get('identical_to_v0'), 'frac', round(d['roofline']['frac'],3), d['roofline']['bound'])
except Exception as e: print('$wl FAILED', l[-400:])
"
done
