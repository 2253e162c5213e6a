/*
 * h9_kernels_exact.cu -- H9_MATH_EXACT instantiation of the time-stepping
 * kernels.  This translation unit is compiled with -fmad=false -prec-div=true
 * so that every +,-,*,/ is the IEEE operation the reference's source order
 * implies; powf/expf/logf are CUDA's accurate versions.
 */
#include "h9_kernels.cuh"

namespace h9 {
H9_DEFINE_LAUNCHERS(exact, MathExact)
}
