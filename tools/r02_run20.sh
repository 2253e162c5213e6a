python tools/cycle_budget.py --band 0 --cell 0 --days 365 --spin 8 --coarse 2>&1 | grep -v Warning | grep -v stddev | grep -v "corr(" > gpurun_out/r02_cycle_budget_band0_equil.txt
cat gpurun_out/r02_cycle_budget_band0_equil.txt
