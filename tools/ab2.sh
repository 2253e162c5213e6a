#!/bin/bash
# usage: tools/ab2.sh tag "variant:H9_BLOCK" ...
tag=$1; shift
out=gpurun_out/ab_$tag.txt; : > $out
for rep in 1 2; do
for vb in "$@"; do
  v=${vb%%:*}; b=${vb##*:}
  H9_BLOCK=$b H9GPU_LIB=$PWD/variants/libh9gpu_$v.so python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$vb', 'ms', round(d['ms_per_step'], 3), 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'], 3))" >> $out
done; done
cat $out
