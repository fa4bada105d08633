#!/bin/bash
# Ring experiment 2: sub-batch size 8 vs 4, depths, after the two-batch-ahead index prefetch.
OUT=gpurun_out/${1:-ring2}; mkdir -p $OUT
for lib in ring8:0,1,2,3 ring4:2,3,4; do
  name=${lib%%:*}; rings=${lib##*:}
  export GEOT_B200_LIB=$PWD/geot_b200/lib/libgeot_b200_$name.so
  for wl in reddit_gws products_gs64 products_gs256 proteins_gws256 reddit_index_scatter config1_index_scatter; do
    for ring in ${rings//,/ }; do
      GEOT_B200_RING=$ring timeout 300 python scripts/tune.py $wl 0 2>&1 | grep -E "lib=|Error|error" | sed "s/^/ring=$ring /" | tee -a $OUT/ring.txt
    done
  done
done
