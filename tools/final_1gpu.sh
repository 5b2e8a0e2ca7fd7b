#!/bin/bash
# Round-end evidence on one GPU: gpu tests, default bench line, ncu launch lists of configs 2/4/5
# (the reference arm, `bench.py --impl reference`, takes ~10 min and does not depend on the build:
# run it separately)
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 2>gpurun_out/err.txt | grep "^{" > gpurun_out/r02_bench_1gpu.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_config2_launches.csv python tools/profile_eval.py --config 2 --evals 2 > /dev/null 2>&1
for c in 4 5; do ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_config${c}_launches.csv python tools/profile_eval.py --config $c --evals 2 > /dev/null 2>&1; done
cp torchtree_b200/lib/libttb200.stamp gpurun_out/lib_stamp.txt
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
ls -la gpurun_out
