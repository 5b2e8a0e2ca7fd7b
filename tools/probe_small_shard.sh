#!/bin/bash
# Small-shard fixed costs (the 8-GPU shard of the headline problem on one GPU): env-knob sweep + ncu launch list
cd "$(dirname "$0")/.."
run() { env "$@" python bench.py --patterns 12500 --no-cpu-baseline --no-other-configs --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d['gpu_launches']/30)"; }
run A=1
run TTB2_CHAIN_MIN_PATTERNS=0 TTB2_CHAIN_MAX=1
run TTB2_CHAIN_MIN_PATTERNS=0 TTB2_CHAIN_MAX=2
run TTB2_CHAIN_MIN_PATTERNS=0 TTB2_CHAIN_MAX=4
run TTB2_NO_PDL=1
run TTB2_GRAPH_MAX_UNITS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_12500.csv python bench.py --patterns 12500 --no-cpu-baseline --no-other-configs --steps 2 --warmup 3 --engine-flags 32 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_12500.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
# last evaluation only: take the final 60 launches
seq=[(r[ki][:60], float(r[vi].replace(',',''))) for r in rows[1:]]
agg=collections.OrderedDict()
for k,v in seq[-120:]:
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print('%-62s n=%3d total=%9.1f us avg=%7.2f us'%(k,n,t/1e3,t/1e3/n))
PY
