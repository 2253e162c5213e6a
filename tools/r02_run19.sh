for blk in 64 4000; do
  H9_BENCH_BAND=0 H9_BLOCK=$blk ncu --set full --clock-control none --import-source on -k regex:days_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02_ncu_v8_band0_blk${blk} python bench.py --grid band8 --days 120 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_ncu_v8_band0_blk${blk}.log 2>&1
  tail -1 gpurun_out/r02_ncu_v8_band0_blk${blk}.log
done
H9_BENCH_BAND=4 H9_BLOCK=4000 ncu --set full --clock-control none --import-source on -k regex:days_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02_ncu_v8_band4_blk4000 python bench.py --grid band8 --days 120 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_ncu_v8_band4_blk4000.log 2>&1
