python tools/cycle_budget.py --band 4 --cell 0 --days 60 --coarse 2>&1 | grep -v Warning | grep -v stddev > gpurun_out/r02_cycle_budget_band4_v3e.txt
cat gpurun_out/r02_cycle_budget_band4_v3e.txt
