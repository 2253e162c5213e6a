python tools/cycle_budget.py --band 0 --cell 2976 --days 90 2>&1 | grep -v Warning | grep -v stddev | grep -v "corr(" | head -12
python tools/cycle_budget.py --band 0 --cell 2976 --days 90 --coarse 2>&1 | grep -v Warning | grep -v stddev | grep -v "corr(" | head -10
