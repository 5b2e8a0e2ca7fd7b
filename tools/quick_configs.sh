#!/bin/bash
cd "$(dirname "$0")/.."
run() { cfg=$1; shift; env "$@" python bench.py --config $cfg --no-cpu-baseline --no-other-configs --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); p=d['phases_ms']; print('config $cfg $*', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'post', round(p['postorder'],2), 'pre', round(p['preorder'],2), 'contract', round(p['contract'],3), 'frac', round(d['roofline']['frac'],3))"; }
for c in "$@"; do run $c A=1; done
