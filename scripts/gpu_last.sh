#!/bin/bash
# Short sanity session for a nearly spent GPU budget: smoke() of the freshly built library, then the L2-hint A/B.
OUT=gpurun_out/${1:-r01e}
mkdir -p $OUT
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
timeout 80 python scripts/l2_persist_ab.py reddit_gws 2>&1 | tail -8 | tee $OUT/l2_persist_ab.txt
