python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.txt 2>&1; tail -3 gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; tail -2 gpurun_out/final_smoke.txt
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
ncu --set full --clock-control none --import-source on -k regex:days_kernel -c 1 -o gpurun_out/exact_final python bench.py --math exact --days 20 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/exact_final_ncu.log 2>&1; tail -1 gpurun_out/exact_final_ncu.log
