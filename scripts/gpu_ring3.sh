#!/bin/bash
# Ring experiment 3: TMA bulk-copy ring (ids 17..19 = depth 1..3) vs cp.async ring (id 2), sub-batch 4 and 8;
# column-slab sweep (GEOT_B200_LPR) for L2 residency.
OUT=gpurun_out/${1:-ring3}; mkdir -p $OUT
for lib in tma4:18,19,2 tma8:17,18,19; do
  name=${lib%%:*}; rings=${lib##*:}
  export GEOT_B200_LIB=$PWD/geot_b200/lib/libgeot_b200_$name.so
  for wl in reddit_gws products_gs64 products_gs256 proteins_gws256 reddit_index_scatter config1_index_scatter; do
    for ring in ${rings//,/ }; do
      GEOT_B200_RING=$ring timeout 300 python scripts/tune.py $wl 0 2>&1 | grep -E "lib=|Error|error" | sed "s/^/ring=$ring /" | tee -a $OUT/ring.txt
    done
  done
done
export GEOT_B200_LIB=$PWD/geot_b200/lib/libgeot_b200_tma4.so
for lpr in 16 8; do
  for ring in 2 18; do
    GEOT_B200_LPR=$lpr GEOT_B200_RING=$ring timeout 300 python scripts/tune.py reddit_gws 0 2>&1 | grep -E "lib=|Error|error" | sed "s/^/lprcap=$lpr ring=$ring /" | tee -a $OUT/ring.txt
  done
done
