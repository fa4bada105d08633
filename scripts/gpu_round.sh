#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), ncu launch list + one full capture,
# reference-CUDA comparison, per-workload timings.
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench own arm"; timeout 600 python bench.py --steps 20 --warmup 5 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_reference.err | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 > $OUT/bench_under_ncu.log 2>&1
grep -c geot $OUT/launches.csv
echo "== ncu full captures of the main kernel (one launch each, after warm-up)"
for wl in reddit_gws products_gs64 reddit_index_scatter; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:segment_reduce_kernel -s 3 -c 1 -o $OUT/prof_$wl \
      python bench.py --workload $wl --steps 3 --warmup 3 > $OUT/prof_$wl.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sddmm_coo_kernel -s 2 -c 1 -o $OUT/prof_sddmm_reddit \
    python scripts/bench_next.py > $OUT/prof_sddmm.log 2>&1
echo "== per-workload timings"
for wl in reddit_gws products_gs64 products_gs256 proteins_gws256 arxiv_mh_spmm config1_index_scatter reddit_index_scatter; do
  timeout 300 python scripts/tune.py $wl 0 2>&1 | grep -E "lib=|Error|error" | tee -a $OUT/workloads.txt
done
echo "== reference CUDA kernels beside ours"
timeout 900 python scripts/compare_reference_cuda.py 2>$OUT/compare_reference.err | tee $OUT/compare_reference.jsonl
echo "== next rows"; timeout 1200 python scripts/bench_next.py 2>$OUT/bench_next.err | tee $OUT/bench_next.jsonl
ls -la $OUT
