#!/bin/bash
OUT=gpurun_out/tune; mkdir -p $OUT
echo "== tests"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
for wl in reddit_gws products_gs64 products_gs256 proteins_gws256 arxiv_mh_spmm config1_index_scatter reddit_index_scatter; do
  timeout 300 python scripts/tune.py $wl 0,64,128,256 2>&1 | grep -E "lib=|Error|error" | tee -a $OUT/tune_all.txt
done
