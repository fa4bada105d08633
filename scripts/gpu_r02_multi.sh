#!/bin/bash
# Round-2 multi-GPU session (most important first; every leg writes its own file as soon as it ends):
#   gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_r02_multi.sh N [tag]'
# 1. NCCL parity tests (full exchange forms; needed-rows form + sharded GCN / GraphSAGE forward),
# 2. the driver's launch form of bench.py (default workload, default exchange),
# 3. products gs64 / gs256 with every exchange form (pipeline, needed, allgather, replicated),
# 4. Reddit gws with the needed-rows and all-gather forms, Reddit index_scatter (no exchange).
N=${1:-2}
TAG=${2:-r02_n$N}
# box time is charged N x: at N >= 4 only the essential legs run unless "full" is given as the third argument
MODE=${3:-$([ "$N" -ge 4 ] && echo short || echo full)}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {   # run <file stem> <env assignments...> -- bench args
  local stem=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N "$@" > $OUT/$stem.json 2> $OUT/$stem.err
  echo "== $stem: exit $?"; tail -c 700 $OUT/$stem.json; tail -3 $OUT/$stem.err
}
date +%s > $OUT/t0
echo "== NCCL parity tests"
GEOT_B200_TEST_EXPERIMENTS=1 timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_multi_needed.py -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_multi.txt
run reddit_gws_pipeline X=1 -- --steps 10 --warmup 3
for ex in pipeline needed push allgather replicated; do
  run products_gs64_$ex GEOT_B200_EXCHANGE=$ex -- --workload products_gs64 --steps 10 --warmup 3
done
run reddit_gws_push GEOT_B200_EXCHANGE=push -- --steps 10 --warmup 3
if [ "$MODE" = full ]; then
  for ex in pipeline needed push replicated; do
    run products_gs256_$ex GEOT_B200_EXCHANGE=$ex -- --workload products_gs256 --steps 10 --warmup 3
  done
  run reddit_gws_needed GEOT_B200_EXCHANGE=needed -- --steps 10 --warmup 3
  run reddit_gws_allgather GEOT_B200_EXCHANGE=allgather -- --steps 10 --warmup 3
  run reddit_gws_replicated GEOT_B200_EXCHANGE=replicated -- --steps 10 --warmup 3
  run reddit_index_scatter X=1 -- --workload reddit_index_scatter --steps 5 --warmup 3
fi
echo "== 3-layer GCN / GraphSAGE forward on the shards (configs[4])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/bench_model_multi.py > $OUT/model.jsonl 2> $OUT/model.err
cat $OUT/model.jsonl | cut -c1-900; tail -3 $OUT/model.err
date +%s > $OUT/t1
ls -la $OUT
