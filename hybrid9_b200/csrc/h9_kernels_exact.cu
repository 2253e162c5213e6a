/*
 * h9_kernels_exact.cu -- H9_MATH_EXACT instantiation of the time-stepping
 * kernels.  This translation unit is compiled with -fmad=false -prec-div=true
 * so that every +,-,*,/ is the IEEE operation the reference's source order
 * implies; pow/exp/log are the portable correctly-rounded kernels of h9_physics.h
 * (the same code runs on the host in tests/twin and oracle/_ref/libh9ref_pk.so).
 */
#include "h9_kernels.cuh"

namespace h9 {
H9_DEFINE_LAUNCHERS(exact, MathExact)

const char* days_variant_exact(int nc, int block) {
  const bool capped = days_exact_capped(nc, block);
  const int bs = block >= 2000 ? 64 : block % 1000;
  if (bs == 32) return capped ? "h9::days_kernel<MathExact,32,16>" : "h9::days_kernel<MathExact,32,1>";
  if (bs == 128) return capped ? "h9::days_kernel<MathExact,128,4>" : "h9::days_kernel<MathExact,128,1>";
  return capped ? "h9::days_kernel<MathExact,64,8>" : "h9::days_kernel<MathExact,64,1>";
}

}
