#!/bin/bash
# Stand-in for compute-sanitizer memcheck (closed on this GPU pool): build the library with
# device-side assertions on every dynamic index (-DH9_BOUNDS_CHECK: shared-memory table rows,
# layer indices of the Drainage code, forcing offsets, cell indices) and run the GPU parity
# tests through it.  A violation traps the kernel (cudaErrorAssert) and fails the test.
set -e
python tools/build_variant.py bounds -DH9_BOUNDS_CHECK | tail -1
H9GPU_LIB=$PWD/variants/libh9gpu_bounds.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_pair.py \
  tests/test_gpu_api.py tests/test_ref_golden.py tests/test_gpu_poison.py tests/test_regrid.py -q 2>&1 | tail -4
