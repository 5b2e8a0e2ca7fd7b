#!/bin/bash
# A/B matrix for the 20-state warp-autonomous kernels (config 4); one line per setting
run() { env "$@" python tools/run_configs.py --only config4_aa_LG_like 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); t = d['timing']; p = t['phases_ms']
    print('%-60s total %.2f post %.2f pre %.2f ok=%s' % (' '.join(sys.argv[1:]), t['ms_per_eval'], p['postorder'], p['preorder'], d['parity']['ok']))
" "$@"; }
run TTB2_GM_LEGACY=1
run X=default
for v in 83 43 45; do run TTB2_GW_FWD=$v; done
for c in 8 32; do run TTB2_GW_FWD_CTAS=$c; done
for v in 42 44 83; do run TTB2_GW_BWD=$v; done
for c in 8 16; do run TTB2_CHUNK_TARGET=$c; run TTB2_CHUNK_TARGET=$c TTB2_GW_BWD=42; done
