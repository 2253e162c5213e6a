python tools/pair_budget.py --band 7 --spin 8 > gpurun_out/r02_pair_budget_band7.txt 2>&1; cat gpurun_out/r02_pair_budget_band7.txt
python tools/pair_budget.py --band 0 --spin 8 > gpurun_out/r02_pair_budget_band0.txt 2>&1; cat gpurun_out/r02_pair_budget_band0.txt
