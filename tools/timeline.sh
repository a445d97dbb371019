#!/bin/bash
# developer probe: per-phase cycle counters of the ordered smoothers (library built with -DAMGB200_TIMELINE in build_tl/)
export AMGB200_LIB=$PWD/build_tl/libamgb200_tl.so AMGB200_DEBUG_TIMING=1
for l in "$@"; do echo "=== level $l"; timeout 200 python tools/prof_level.py p3d 128 $l 2>&1 | grep -A3 "timeline\|dbg\|ms per" | tail -14; done
