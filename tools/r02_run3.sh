set -x
for band in 4 0; do for blk in 64 4000; do
  H9_BENCH_BAND=$band H9_BLOCK=$blk ncu --set full --clock-control none --import-source on -k regex:days_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02_ncu_band${band}_blk${blk} python bench.py --grid band8 --days 30 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_ncu_band${band}_blk${blk}.log 2>&1
  tail -2 gpurun_out/r02_ncu_band${band}_blk${blk}.log
done; done
ls -la gpurun_out/*.ncu-rep
