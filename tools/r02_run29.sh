python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; echo "rc $?"; cat gpurun_out/r02_bench_final_n1.json | cut -c1-600
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_final_ref.json 2>> gpurun_out/r02_bench_final_n1.err; cat gpurun_out/r02_bench_final_ref.json | cut -c1-300
bash tools/bounds_check.sh 2>&1 | tail -3
