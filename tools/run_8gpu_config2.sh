cd /root/repo
for n in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"
  $TR bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/err_${n}.txt | grep '^{' > gpurun_out/r02_bench_${n}gpu.json || tail -20 gpurun_out/err_${n}.txt
done
