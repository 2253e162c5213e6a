for rep in 1 2; do for v in product rsssel; do
  if [ $v = product ]; then lib=$PWD/hybrid9_b200/libh9gpu.so; else lib=$PWD/variants/libh9gpu_$v.so; fi
  H9GPU_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v', 'ms', round(d['ms_per_step'], 3), 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'], 3))"
done; done
