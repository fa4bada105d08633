#!/bin/bash
# Short multi-GPU session, most important first (the GPU budget left for it is a few minutes):
#   bash scripts/gpu_multi3.sh <N> [tag]      (under gpurun --gpus N)
# 1. the driver's own launch form of bench.py at N GPUs (default workload), 2. the two-GPU NCCL parity test,
# 3. products_gs64 with both exchange forms, 4. reddit with the all-gather form.  Every leg writes its own file as soon
# as it ends, so a call cut short still leaves what finished.
N=${1:-2}
TAG=${2:-r01d_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {   # run <file stem> <env assignments...> -- bench args
  local stem=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N "$@" > $OUT/$stem.json 2> $OUT/$stem.err
  echo "== $stem: exit $?"; tail -c 600 $OUT/$stem.json; tail -3 $OUT/$stem.err
}
date +%s > $OUT/t0
run reddit_gws_pipeline X=1 -- --steps 10 --warmup 3
echo "== two-GPU NCCL parity test"
timeout 240 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_multi.txt
run products_gs64_pipeline X=1 -- --workload products_gs64 --steps 10 --warmup 3
run products_gs64_allgather GEOT_B200_EXCHANGE=allgather -- --workload products_gs64 --steps 10 --warmup 3
run reddit_gws_allgather GEOT_B200_EXCHANGE=allgather -- --steps 10 --warmup 3
run products_gs64_replicated GEOT_B200_EXCHANGE=replicated -- --workload products_gs64 --steps 10 --warmup 3
run reddit_index_scatter X=1 -- --workload reddit_index_scatter --steps 5 --warmup 3
date +%s > $OUT/t1
ls -la $OUT
