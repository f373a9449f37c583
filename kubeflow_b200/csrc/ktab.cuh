// Kernel function k = f(d²) in FP64, branch-free, in registers.
// (A Taylor-table variant was measured first: two data-dependent 16-byte loads per element made the kernel L1-gather
//  bound — 104 ms vs 77 ms at cfg3 — so the table is gone; see profiles/README.md "r1 dead ends".)
//   exp(x), x <= 0:  n = rint(x·log2e) by the 1.5·2⁵² trick, f = x·log2e − n (two-term log2e), 2^f by a degree-11
//   polynomial in g = f·ln2 (|g| ≤ 0.3466, truncation 6e-15 relative), scaled by 2ⁿ through the exponent bits.
//   Results below 2⁻¹⁰²⁰ flush to 0 (irrelevant next to the 1e-5 contract).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ double kbo_exp_nonpos(double x) {
  const double t = x * 1.4426950408889634;
  const double u = t + 6755399441055744.0;          // 1.5·2^52: rounds t to nearest integer in the low word
  const int ni = __double2loint(u);
  const double nf = u - 6755399441055744.0;
  double f = fma(x, 1.4426950408889634, -nf);
  f = fma(x, 2.0355273740931033e-17, f);             // low part of log2(e)
  const double g = f * 0.6931471805599453;
  double p = 2.505210838544172e-08;                   // 1/11!
  p = fma(p, g, 2.755731922398589e-07);               // 1/10!
  p = fma(p, g, 2.7557319223985893e-06);              // 1/9!
  p = fma(p, g, 2.48015873015873e-05);                // 1/8!
  p = fma(p, g, 1.984126984126984e-04);               // 1/7!
  p = fma(p, g, 1.388888888888889e-03);               // 1/6!
  p = fma(p, g, 8.333333333333333e-03);               // 1/5!
  p = fma(p, g, 4.1666666666666664e-02);              // 1/4!
  p = fma(p, g, 1.6666666666666666e-01);              // 1/3!
  p = fma(p, g, 0.5);
  p = fma(p, g, 1.0);
  p = fma(p, g, 1.0);
  const int hi = __double2hiint(p) + (ni << 20);
  const double r = __hiloint2double(hi, __double2loint(p));
  return ni < -1020 ? 0.0 : r;
}

// branch-free sqrt for x >= 0: fp32 MUFU.RSQ seed (~2^-22), ONE Newton step for 1/sqrt in FP64 (→ ~1e-13), then the
// Newton correction on s = x·y itself, which squares the error again (≤ 1 ulp).
// x is floored at 1e-30 so the seed is finite; sqrt(0) comes out as 1e-15, i.e. k(0) = 1 − O(1e-30).
__device__ __forceinline__ double kbo_sqrt_nonneg(double x) {
  x = fmax(x, 1e-30);
  double y = (double)rsqrtf((float)x);
  y = y * fma(-0.5 * x, y * y, 1.5);
  const double s = x * y;
  return fma(fma(-s, s, x), 0.5 * y, s);
}

__device__ __forceinline__ double kbo_kernel_exact(double d2, int kind) {
  d2 = fmax(d2, 0.0);
  if (kind == 0 /*KBO_KERNEL_RBF*/) return kbo_exp_nonpos(-0.5 * d2);
  const double s = kbo_sqrt_nonneg(5.0 * d2);
  return fma(s, fma(s, 1.0 / 3.0, 1.0), 1.0) * kbo_exp_nonpos(-s);
}
