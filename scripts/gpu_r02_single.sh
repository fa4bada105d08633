#!/bin/bash
# Round-2 single-GPU session.  Build the tuning variant HERE first (it travels with the snapshot):
#   make -f geot_b200/csrc/Makefile VARIANT=_u4 EXTRA="-DGEOT_U0=4" -j8 lib
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r02_single.sh [tag]'
# Legs (each writes its own file as soon as it ends):
#  1. parity tests (the whole -m gpu suite: includes the exchange kernels, push_rows, compact host transport)
#  2. bench own arm with each host transport (e2e: 0 plain, 1 row pointers, 3 row pointers + int32 src ids)
#  3. bf16 max/min register-path kernels: U0 = 8 (spills) against the U0 = 4 variant (spill-free)
#  4. L2 capacity probe: Reddit gws at F = 32 / 64 / 128 (src 30 / 60 / 119 MB) with ncu hit rates
#  5. src blocking A/B (GEOT_B200_SRC_BLOCKS = 2 / 3 / 4)
#  6. ncu full capture of the lean depth-7 index_scatter kernel
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu (experiments included)"; GEOT_B200_TEST_EXPERIMENTS=1 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
for m in 0 1 3; do
  echo "== bench, host transport $m"
  GEOT_B200_HOST_COMPACT=$m timeout 600 python bench.py --steps 20 --warmup 5 2>$OUT/bench_compact$m.err | tail -1 | tee $OUT/bench_compact$m.json
done
echo "== ncu launch list of the default bench command (current build)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 > $OUT/bench_under_ncu.log 2>&1
grep -c geot $OUT/launches.csv
echo "== zero only the empty rows (GEOT_B200_ZERO_EMPTY) on the workload with gaps"
for z in 0 1; do
  GEOT_B200_ZERO_EMPTY=$z timeout 300 python scripts/tune.py arxiv_mh_spmm 0 2>&1 | grep -E "lib=|rror" | sed "s/^/zero_empty=$z /" | tee -a $OUT/zero_empty.txt
done
echo "== lean register path (GEOT_B200_RING=96) against the lean ring (35) per workload"
for wl in reddit_gws products_gs64 products_gs256 proteins_gws256 reddit_index_scatter; do
  for ring in 35 96; do
    GEOT_B200_RING=$ring timeout 300 python scripts/tune.py $wl 0 2>&1 | grep -E "lib=|rror" | sed "s/^/ring=$ring /" | tee -a $OUT/ring96_ab.txt
  done
done
echo "== chunk sweep on the two small (latency-bound) workloads"
for wl in arxiv_mh_spmm config1_index_scatter; do
  timeout 300 python scripts/tune.py $wl 0,16,32,64,128,256 2>&1 | grep -E "lib=|rror" | tee -a $OUT/chunk_sweep_small.txt
done
echo "== L2 capacity probe"
for F in 32 64 128; do
  timeout 600 ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none \
      -k regex:segment_reduce_kernel -s 4 -c 1 --csv --log-file $OUT/l2probe_F$F.csv python scripts/l2_probe.py $F > $OUT/l2probe_F$F.log 2>&1
done
echo "== src blocking (temporal L2 blocking of the src matrix) on the Reddit and proteins shapes"
for wl in reddit_gws proteins_gws256; do
  for b in 2 3 4; do
    GEOT_B200_SRC_BLOCKS=$b timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 2>$OUT/blocks_${wl}_$b.err | tail -1 > $OUT/blocks_${wl}_$b.json
    python -c "import json,sys; d=json.load(open('$OUT/blocks_${wl}_$b.json')); print('$wl blocks=$b', d['ms_per_step'], 'ms', d['value'], 'GB/s')" | tee -a $OUT/blocks.txt
  done
done
echo "== ncu full capture: reddit index_scatter (lean ring depth 7)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:segment_reduce_kernel -s 3 -c 1 -o $OUT/prof_reddit_index_scatter \
    python bench.py --workload reddit_index_scatter --steps 3 --warmup 3 > $OUT/prof_reddit_index_scatter.log 2>&1
ls -la $OUT
