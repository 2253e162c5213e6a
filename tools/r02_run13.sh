python -m pytest "tests/test_gpu_fullsize_oracle.py::test_half_degree_globe_10_days_vs_oracle" -q 2>&1 | grep -E "^E|passed|failed" | head -20
python tools/equilibrium_report.py 2>&1 | tail -60
bash tools/sanitize.sh 2>&1 | tail -40
