#!/bin/bash
# Multi-GPU session: bash scripts/gpu_multi2.sh <N> [tag]   (under gpurun --gpus N)
# NCCL test of both exchange forms, then the bench at N GPUs with each form.
N=${1:-2}
TAG=${2:-r01c_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== two-GPU NCCL test"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_multi.txt
for wl in reddit_gws products_gs64; do
  for ex in pipeline allgather; do
    GEOT_B200_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $wl --steps 20 --warmup 5 2>$OUT/${wl}_n${N}_$ex.err | tail -1 | tee $OUT/${wl}_n${N}_$ex.json
  done
done
tail -5 $OUT/*.err
ls -la $OUT
