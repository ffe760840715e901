#!/bin/bash
# times one reward update with every ab_build/libdmfg_*.so (and the main build first)
cd "$(dirname "$0")/.."
timeout 300 python scripts/time_irl_update.py 2>/dev/null | tail -1
for l in $(ls ab_build/libdmfg_*.so 2>/dev/null); do
  DMFG_LIB_PATH=$l timeout 300 python scripts/time_irl_update.py 2>/dev/null | tail -1
done
