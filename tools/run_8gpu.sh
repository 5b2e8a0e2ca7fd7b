#!/bin/bash
# 8-GPU box: NCCL parity test, scaling of config 2 (pattern-sharded) and config 3 (draw-sharded)
cd "$(dirname "$0")/.."
python -m pytest tests/test_sharded_nccl_gpu.py -x -q 2>&1 | tail -3
for n in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"
  $TR bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/err_${n}.txt | grep '^{' > gpurun_out/r02_bench_${n}gpu.json || tail -20 gpurun_out/err_${n}.txt
  $TR bench.py --gpus $n --config 3 --steps 10 --warmup 3 2> gpurun_out/err3_${n}.txt | grep '^{' > gpurun_out/r02_bench_${n}gpu_config3.json || tail -20 gpurun_out/err3_${n}.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*gpu*.json")):
    try:
        d=json.load(open(f))
        print(f, "n", d["n_gpus"], "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "value %.3e"%d["value"], d["phases_ms"])
    except Exception as ex: print(f, "FAILED", ex)
PY
