#!/bin/bash
# Lean-ring session (1 GPU): parity tests, A/B of the ring variants per workload, ncu captures of the lean kernel.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_lean.sh [tag]'
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== ring variants per workload (GEOT_B200_RING: 2/3 first-generation ring, 35 lean depth 3, 39 lean depth 7, 0 registers)"
for wl in reddit_gws products_gs64 products_gs256 proteins_gws256 config1_index_scatter reddit_index_scatter arxiv_mh_spmm; do
  for ring in 2 3 35 39; do
    GEOT_B200_RING=$ring timeout 300 python scripts/tune.py $wl 0 2>&1 | grep -E "lib=|Error|error" | sed "s/^/ring=$ring /" | tee -a $OUT/ring_ab.txt
  done
done
echo "== ncu full captures of the lean kernel"
for wl in reddit_gws products_gs64; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:segment_reduce_kernel -s 4 -c 1 -o $OUT/prof_$wl \
      python scripts/tune.py $wl 0 > $OUT/prof_$wl.log 2>&1
done
ls -la $OUT
