python -m pytest tests -m gpu -x -q > gpurun_out/final2_tests.txt 2>&1; tail -3 gpurun_out/final2_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.txt 2>&1; tail -2 gpurun_out/final2_smoke.txt
python bench.py --no-cpu > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; tail -c 400 gpurun_out/final2_bench.json
