python tools/cycle_budget.py --band 4 --cell 0 --days 60 > gpurun_out/r02_cycle_budget_band4.txt 2>&1
cat gpurun_out/r02_cycle_budget_band4.txt
python tools/cycle_budget.py --band 0 --cell 0 --days 60 > gpurun_out/r02_cycle_budget_band0.txt 2>&1
cat gpurun_out/r02_cycle_budget_band0.txt
python tools/cycle_budget.py --band 0 --cell 4000 --days 60 > gpurun_out/r02_cycle_budget_band0_c4000.txt 2>&1
cat gpurun_out/r02_cycle_budget_band0_c4000.txt
