for band in 6 0; do
python tools/cycle_budget.py --band $band --cell 0 --days 90 --coarse 2>&1 | grep -v Warning | grep -v stddev | grep -v "corr(" > gpurun_out/r02_cycle_budget_band${band}_v7.txt
cat gpurun_out/r02_cycle_budget_band${band}_v7.txt
done
bash tools/bounds_check.sh 2>&1 | tail -5
python -m pytest tests/test_gpu_poison.py tests/test_gpu_fullsize_oracle.py -q 2>&1 | tail -5
