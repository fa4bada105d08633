#!/bin/bash
# One parametrised GPU session (replaces the per-session scripts of round 1).  Every leg writes its own file under
# gpurun_out/<tag>/ as soon as it ends.
#   gpurun [--gpus N] --timeout T -- 'bash scripts/gpu_session.sh <tag> <legs...>'
# legs: test (pytest -m gpu) | smoke | bench (default bench.py, own + reference arm) | benchN (torchrun bench at N = all
#       visible GPUs, own + reference arm) | exch (N > 1: products / reddit with every exchange form) | model (configs[4]
#       at N GPUs) | launches (ncu launch list of the default bench) | ncu:<workload> (one --set full capture) |
#       tune:<workload>[:chunks] | libs:<workload>:<lib,lib..>[:blocks] | ncut:<workload> | timeline:<workload>[:exchange] (N > 1: where the step goes, CUDA-graph replay) |
#       sweep:<workload>[:rounds[:ctas]] (N > 1: exchange rounds x push grid) | compare (the reference's own CUDA kernels beside ours, sorted and sorted=False) |
#       env:<NAME=VALUE> (exported for the legs that follow)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
TRUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for leg in "$@"; do
  echo "== leg $leg ($(date +%T))"
  case $leg in
    env:*) export "${leg#env:}";;
    test) timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt;;
    testmulti) timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -25 | tee $OUT/pytest_gpu_multi.txt;;
    benchNown) timeout 900 $TRUN bench.py --gpus $N 2>$OUT/bench_n$N.err | tail -1 > $OUT/bench_n$N.json
      tail -3 $OUT/bench_n$N.err; cut -c1-400 $OUT/bench_n$N.json;;
    smoke) timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/smoke.txt;;
    bench)
      timeout 900 python bench.py --impl reference 2>$OUT/bench_reference.err | tail -1 > $OUT/bench_reference.json
      timeout 900 python bench.py 2>$OUT/bench.err | tail -1 > $OUT/bench.json
      tail -3 $OUT/bench.err; cut -c1-600 $OUT/bench.json;;
    benchN)
      timeout 600 $TRUN bench.py --gpus $N --impl reference 2>$OUT/bench_n${N}_reference.err | tail -1 > $OUT/bench_n${N}_reference.json
      timeout 900 $TRUN bench.py --gpus $N 2>$OUT/bench_n$N.err | tail -1 > $OUT/bench_n$N.json
      tail -3 $OUT/bench_n$N.err; cut -c1-600 $OUT/bench_n$N.json;;
    exch|exch:*)
      WLS="reddit_gws products_gs64"; EXS="bucket push allgather replicated"
      if [ "$leg" != exch ]; then IFS=: read -r _ WLS EXS <<< "$leg"; WLS=${WLS//,/ }; EXS=${EXS//,/ }; fi
      for wl in $WLS; do for ex in $EXS; do
        GEOT_B200_BENCH_SECONDARY=0 GEOT_B200_EXCHANGE=$ex timeout 300 $TRUN bench.py --gpus $N --workload $wl --steps 10 --warmup 3 \
          2>$OUT/exch_${wl}_$ex.err | tail -1 > $OUT/exch_${wl}_$ex.json
        python - <<PY
import json
try:
    d = json.load(open("$OUT/exch_${wl}_$ex.json"))
    print("$wl $ex N=$N: %.4f ms/step, kernel %.4f ms, parity %s, e2e %.3f ms" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["parity"]["ok"], d["e2e"].get("ms_per_step", -1)))
except Exception as e:
    print("$wl $ex N=$N: FAILED", e)
PY
      done; done 2>&1 | tee $OUT/exch.txt;;
    model) timeout 600 $TRUN scripts/bench_model_multi.py > $OUT/model_n$N.jsonl 2> $OUT/model_n$N.err; cut -c1-900 $OUT/model_n$N.jsonl; tail -3 $OUT/model_n$N.err;;
    launches)
      GEOT_B200_BENCH_SECONDARY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
        python bench.py --steps 3 --warmup 3 > $OUT/bench_under_ncu.log 2>&1; grep -c geot $OUT/launches.csv;;
    ncu:*) IFS=: read -r _ wl cnt <<< "$leg"; cnt=${cnt:-1}      # cnt = main-kernel launches of one step (src blocks)
      GEOT_B200_BENCH_SECONDARY=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:segment_reduce_kernel -s $((3 * cnt)) -c $cnt -o $OUT/prof_$wl \
        python bench.py --workload $wl --steps 3 --warmup 3 > $OUT/prof_$wl.log 2>&1; ls -la $OUT/prof_$wl.ncu-rep;;
    compare) timeout 900 python scripts/compare_reference_cuda.py > $OUT/compare_reference.jsonl 2> $OUT/compare_reference.err
      timeout 300 python scripts/compare_reference_cuda.py unsorted > $OUT/compare_unsorted.jsonl 2>> $OUT/compare_reference.err
      cut -c1-400 $OUT/compare_reference.jsonl $OUT/compare_unsorted.jsonl; tail -3 $OUT/compare_reference.err;;
    timeline:*) IFS=: read -r _ wl ex <<< "$leg"
      timeout 300 $TRUN scripts/exchange_timeline.py $wl ${ex:-push} 2>$OUT/timeline_${wl}_${ex:-push}.err | grep -E "N=|rows pushed|rounds|graph|rror" | tee -a $OUT/timeline.txt
      tail -2 $OUT/timeline_${wl}_${ex:-push}.err;;
    sweep:*) IFS=: read -r _ wl rounds ctas <<< "$leg"
      timeout 300 $TRUN scripts/exchange_sweep.py $wl ${rounds:-1,2,3} ${ctas:-296,592} 2>$OUT/sweep_$wl.err | grep -E "N=|rror" | tee -a $OUT/sweep.txt
      tail -2 $OUT/sweep_$wl.err;;
    shards:*) IFS=: read -r _ wl parts <<< "$leg"
      timeout 300 python scripts/shard_probe.py $wl ${parts:-8} 2>&1 | grep -E "shard|rror" | tee -a $OUT/shard_probe.txt;;
    libs:*) IFS=: read -r _ wl libs blocks <<< "$leg"      # same-process A/B of tuning builds (Makefile VARIANT=...)
      timeout 300 python scripts/tune_libs.py $wl ${libs:-default} ${blocks:-1} 2>&1 | grep -E "lib=|rror" | tee -a $OUT/tune_libs.txt;;
    ncut:*) IFS=: read -r _ wl <<< "$leg"      # one --set full capture of the timed-loop kernel, driven by the light tune script
      timeout 240 ncu --set full --clock-control none --import-source on -k regex:segment_reduce_kernel -s 5 -c 1 -o $OUT/prof_$wl \
        python scripts/tune.py $wl > $OUT/prof_$wl.log 2>&1; ls -la $OUT/prof_$wl.ncu-rep;;
    tune:*) IFS=: read -r _ wl chunks <<< "$leg"
      timeout 300 python scripts/tune.py $wl ${chunks:-0} 2>&1 | grep -E "lib=|rror" | tee -a $OUT/tune.txt;;
    *) echo "unknown leg $leg";;
  esac
done
ls -la $OUT
