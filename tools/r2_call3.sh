#!/usr/bin/env bash
# Round 2, GPU call 3: where does a panel's time go?  role timers (debug build), ring-depth sensitivity
set -u
mkdir -p gpurun_out
echo "--- role timers v0 (sets 1, tiles 1)"
BQP_LIB_SUFFIX=_d0 BQP_BUILD_DEFS="-DBQP_PANEL_DEBUG" timeout 200 python tools/iter_bench.py --instances 1 --iters 100 2>&1 | grep -E "^(P1|P2|UPD|_d)" | sort | uniq -c | sort -rn | head -12
echo "--- role timers v1 (sets 2, tiles 2)"
BQP_LIB_SUFFIX=_d1 BQP_BUILD_DEFS="-DBQP_PANEL_DEBUG -DBQP_P1_SETS=2 -DBQP_P1_TILES=2" timeout 200 python tools/iter_bench.py --instances 1 --iters 100 2>&1 | grep -E "^(P1|P2|UPD|_d)" | sort | uniq -c | sort -rn | head -12
echo "--- ring depth v0"
for s in 4 6 8; do BQP_PANEL_SLOTS=$s BQP_LIB_SUFFIX=_v0 BQP_BUILD_DEFS="-DBQP_P1_SETS=1 -DBQP_P1_TILES=1" timeout 200 python tools/iter_bench.py --instances 1 --iters 200 2>&1 | grep "^_v" ; done
