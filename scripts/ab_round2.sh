#!/bin/bash
# A/B candidates prepared at the end of round 1 without GPU time left.  Run on a GPU box:
#   bash scripts/ab_round2.sh
# Every candidate first has to pass the GPU test suite through WB_LIB (wrapped in `timeout`: a bulk-copy kernel whose byte
# count were wrong would wait on its mbarrier forever), then it is timed against the default build on c3.
set -e
cd "$(dirname "$0")/.."
scripts/build_variant.sh default
scripts/build_variant.sh staged -DWB_ATTRACT_STAGED=1
scripts/build_variant.sh pointhalf -DWB_POINT_HALF=1
scripts/build_variant.sh hitbatch -DWB_HIT_BATCH=1
scripts/build_variant.sh both -DWB_POINT_HALF=1 -DWB_HIT_BATCH=1
for v in staged pointhalf hitbatch both; do
  WB_LIB=$PWD/wembed_b200/lib/variants/libwb_$v.so timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
done
timeout 300 python scripts/gpu_ab.py c3 20 60 wembed_b200/lib/variants/libwb_default.so wembed_b200/lib/variants/libwb_staged.so wembed_b200/lib/variants/libwb_pointhalf.so wembed_b200/lib/variants/libwb_hitbatch.so wembed_b200/lib/variants/libwb_both.so
