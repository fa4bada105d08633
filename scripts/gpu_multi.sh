#!/bin/bash
# Multi-GPU bench session: bash scripts/gpu_multi.sh <N> [tag]   (under gpurun --gpus N)
N=${1:-2}
TAG=${2:-r01_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for wl in reddit_gws products_gs64 products_gs256; do
  for n in 1 $N; do
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 2>$OUT/${wl}_n$n.err | tail -1 | tee $OUT/${wl}_n$n.json
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --workload $wl --steps 20 --warmup 5 2>$OUT/${wl}_n$n.err | tail -1 | tee $OUT/${wl}_n$n.json
    fi
  done
done
ls -la $OUT
