#!/bin/bash
# Ring-depth experiment: the tuning library (f32 sum kernels with cp.async ring depths 1,2,3) against the direct path.
OUT=gpurun_out/${1:-ring}; mkdir -p $OUT
export GEOT_B200_LIB=$PWD/geot_b200/lib/libgeot_b200_ring.so
for wl in reddit_gws products_gs64 products_gs256 proteins_gws256 reddit_index_scatter config1_index_scatter; do
  for ring in 0 1 2 3; do
    GEOT_B200_RING=$ring timeout 300 python scripts/tune.py $wl 0 2>&1 | grep -E "lib=|Error|error" | sed "s/^/ring=$ring /" | tee -a $OUT/ring.txt
  done
done
unset GEOT_B200_LIB
echo "== parity with the ring (depth 2) in the production library"
GEOT_B200_RING=2 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_next.py -x -q 2>&1 | tail -5 | tee $OUT/pytest_ring2.txt
echo "== next rows"; timeout 1200 python scripts/bench_next.py 2>$OUT/bench_next.err | tee $OUT/bench_next.jsonl
