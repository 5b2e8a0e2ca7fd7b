#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err || tail -20 gpurun_out/r02_bench_2gpu.err
$TR bench.py --gpus 2 --steps 30 --warmup 5 --patterns 25000 > gpurun_out/r02_bench_2gpu_p25000.json 2> gpurun_out/err2.txt || tail -20 gpurun_out/err2.txt
$TR bench.py --gpus 2 --config 3 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_config3.json 2> gpurun_out/err3.txt || tail -20 gpurun_out/err3.txt
python bench.py --patterns 12500 --no-cpu-baseline --no-other-configs --steps 30 --warmup 5 > gpurun_out/r02_shard_12500_b.json 2>/dev/null
python - <<'PY'
import json
for f in ("r02_bench_2gpu","r02_bench_2gpu_p25000","r02_bench_2gpu_config3","r02_shard_12500_b"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "value %.3e"%d["value"], "launches/step", d["gpu_launches"]/d["steps"], d["phases_ms"])
    except Exception as ex: print(f, "FAILED", ex)
PY
