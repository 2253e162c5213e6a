python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_b24.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_ncu24a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:days_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02_ncu_final_fullgrid python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_ncu24b.log 2>&1
tail -1 gpurun_out/r02_ncu24b.log
