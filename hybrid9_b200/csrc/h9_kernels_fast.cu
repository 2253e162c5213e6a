/*
 * h9_kernels_fast.cu -- H9_MATH_FAST instantiation of the time-stepping
 * kernels: MUFU ex2/lg2/rcp based pow/exp/div and FMA contraction.  The
 * deviation from the exact mode is measured by tests/test_gpu_parity.py and
 * reported in DESIGN.md; it does not share the exact mode's tolerance gates.
 */
#include "h9_kernels.cuh"

namespace h9 {
H9_DEFINE_LAUNCHERS(fast, MathFast)
}
