#!/bin/bash
cd "$(dirname "$0")/.."
run() { env "$@" python bench.py --config 5 --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); p=d['phases_ms']; print('$*', 'ms', round(d['ms_per_step'],3), 'post', round(p['postorder'],2), 'pre', round(p['preorder'],2), 'contract', round(p['contract'],3), 'frac', round(d['roofline']['frac'],3))"; }
run TTB2_GM61=8
run TTB2_GM61=16
run TTB2_GM61=16 TTB2_GM61F=16
run TTB2_GM61=16 TTB2_CHUNK_TARGET=4
run TTB2_GM61=16 TTB2_CHUNK_TARGET=8
run TTB2_GM61=16 TTB2_CHUNK_TARGET=32
