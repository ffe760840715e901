#!/bin/bash
# times every ab_build/libdmfg_*.so (or the names given) with scripts/time_rollout.py; one line per build
cd "$(dirname "$0")/.."
libs="$@"
[ -z "$libs" ] && libs=$(ls ab_build/libdmfg_*.so)
for l in $libs; do
  DMFG_LIB_PATH=$l timeout 300 python scripts/time_rollout.py 2>&1 | tail -1
done
