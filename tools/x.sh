python bench.py --math exact --days 60 --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('exact 60 days', d['ms_per_step'], 'ms ->', d['ms_per_step']*365/60, 'ms per year', d['value'])"
python -m pytest tests/test_gpu_vs_ref_bitwise.py tests/test_gpu_parity.py -q -k "exact or bitwise" 2>&1 | tail -2
