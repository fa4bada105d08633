#!/bin/bash
TAG=${1:-r01_next}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== next rows"; timeout 1200 python scripts/bench_next.py 2>$OUT/bench_next.err | tee $OUT/bench_next.jsonl
tail -5 $OUT/bench_next.err
