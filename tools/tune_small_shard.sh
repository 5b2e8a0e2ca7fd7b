for env in "X=1" "TTB2_GRAPH_MAX_UNITS=1e8" "TTB2_GRAPH_MAX_UNITS=0" "TTB2_CHUNK_TARGET=4" "TTB2_CHUNK_TARGET=16"; do
  for n in 12500 25000; do
  env $env python bench.py --patterns $n --steps 50 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); p=d['phases_ms']
print('$env', $n, 'ms %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:p[k] for k in ('pmatrix','postorder','root','preorder','contract')})
"
  done
done
