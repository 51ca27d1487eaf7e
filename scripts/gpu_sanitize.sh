#!/bin/bash
# compute-sanitizer over scripts/sanitize_target.py, all four tools -> gpurun_out/r2_sanitizer.txt
out=gpurun_out/r2_sanitizer.txt
: > $out
for t in memcheck racecheck synccheck initcheck; do
  echo "## $t" >> $out
  timeout 900 compute-sanitizer --tool $t --print-limit 20 python scripts/sanitize_target.py 2>&1 | grep -v "^$" | tail -12 >> $out
done
echo "## racecheck, ring with stage reuse (scripts/sanitize_ring.py: 11 tiles through 4- and 2-stage rings, every CTA shape)" >> $out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_ring.py 2>&1 | grep -v "^$" | tail -8 >> $out
cat $out
