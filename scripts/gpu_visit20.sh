#!/bin/bash
set -u
mkdir -p gpurun_out
out=gpurun_out/r02_sanitizer_pencil.txt
echo "# compute-sanitizer on scripts/sanitize_pencil.py (pencil stage kernel 2-D/3-D Np=3..5, two species + sources, fused field phases, stand-alone field kernel, walls/outflow, partial patches, streamed host step)" > $out
timeout 200 python scripts/sanitize_pencil.py 2>&1 | tail -10
echo "## memcheck" >> $out
timeout 1200 compute-sanitizer --tool memcheck python scripts/sanitize_pencil.py 2>&1 | grep -v "^=========\s*$" | tail -25 >> $out
echo "## racecheck" >> $out
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_pencil.py 2>&1 | tail -25 >> $out
echo "## initcheck" >> $out
timeout 1200 compute-sanitizer --tool initcheck python scripts/sanitize_pencil.py 2>&1 | tail -15 >> $out
cat $out | cut -c1-200
