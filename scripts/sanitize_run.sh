#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over scripts/sanitize_small.py -> gpurun_out/r2_compute_sanitizer.txt
cd "$(dirname "$0")/.."
out=gpurun_out/r2_compute_sanitizer.txt
: > $out
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> $out
  timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize_small.py 2>&1 | grep -v "^Inside" | tail -12 >> $out
done
cat $out
