#!/bin/bash
# usage: tools/ab4.sh tag variant...   full grid and band 0 of 8
tag=$1; shift
out=gpurun_out/ab_$tag.txt; : > $out
for rep in 1 2; do
for v in "$@"; do
  for grid in 0.5 band8; do
  H9GPU_LIB=$PWD/variants/libh9gpu_$v.so python bench.py --grid $grid --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v', '$grid', 'ms', round(d['ms_per_step'], 3))" >> $out
  done
done; done
cat $out
