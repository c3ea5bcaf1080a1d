for w in cfg3 cfg4 cfg5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2j_scale_${w}_n8.json 2> gpurun_out/r2j_scale_${w}_n8.err
  echo "$w rc $?"
done
