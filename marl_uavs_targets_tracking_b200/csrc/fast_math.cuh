// fast_math.cuh -- fp64 sine / cosine for headings (|h| <= pi plus one turn-rate step), shared by the step kernels.
//
// The reference evaluates math.cos / math.sin (glibc, < 1 ulp) on every heading once per step
// (src/agent/uav.py:92-95, src/agent/target.py:35-38).  Headings are wrapped to [-pi, pi), so the general
// argument reduction of the CUDA library routine (its slow path test, its call overhead: it is not inlined twice)
// is dead weight here.  This is the classical scheme: k = rint(h * 2/pi), Cody-Waite reduction with a two-part
// pi/2 (k <= 3: the first product is exact), the fdlibm kernel polynomials on [-pi/4, pi/4] with the reduction's
// tail, quadrant fix-up.  Max error < 1 ulp (tests/test_fast_math.py checks it against libm on the host build of
// this very file).  Compiles as plain C++ as well, so the host test exercises the same source.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD static inline
#endif

FM_HD double fm_hi_as_double(uint32_t hi) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)hi, 0);
#else
  uint64_t b = (uint64_t)hi << 32;
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
FM_HD uint32_t fm_hi_word(double x) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2hiint(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  return (uint32_t)(b >> 32);
#endif
}

// The coefficients sit in constant memory on the device: an fp64 immediate costs two extra instructions per use,
// a constant-bank operand none.
#define FM_COEFFS                                                                                                       \
  {6755399441055744.0, 6.36619772367581382433e-01, 1.57079632673412561417e+00, 6.07710050650619224932e-11,              \
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,    \
   -2.50507602534068634195e-08, 1.58969099521155010221e-10, 4.16666666666666019037e-02, -1.38888888888741095749e-03,    \
   2.48015872894767294178e-05, -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11}
#ifdef __CUDACC__
__constant__ double fm_dev_coeff[16] = FM_COEFFS;
#endif
static const double fm_host_coeff[16] = FM_COEFFS;
#ifdef __CUDA_ARCH__
#define FM_K(i) fm_dev_coeff[i]
#else
#define FM_K(i) fm_host_coeff[i]
#endif

// sin and cos of h for |h| < 4 (callers route anything else to the library routine)
FM_HD void fm_sincos_small(double h, double *sn, double *cs) {
  // k = rint(h * 2/pi) through the 1.5 * 2^52 shift; the low word of the shifted sum is k in two's complement
  const double SHIFT = FM_K(0);
  const double ks = fma(h, FM_K(1), SHIFT);
#ifdef __CUDA_ARCH__
  const int k = __double2loint(ks);
#else
  uint64_t kb;
  memcpy(&kb, &ks, 8);
  const int k = (int)(uint32_t)kb;
#endif
  const double kd = ks - SHIFT;
  const double r0 = fma(-kd, FM_K(2), h);  // first 33 bits of pi/2: the product is exact
  const double w = kd * FM_K(3);           // pi/2 - the above
  const double x = r0 - w;
  const double y = (r0 - x) - w;  // tail of the reduced argument
  const double z = x * x;
  // fdlibm __kernel_sin(x, y, 1)
  const double S1 = FM_K(4), S2 = FM_K(5), S3 = FM_K(6), S4 = FM_K(7), S5 = FM_K(8), S6 = FM_K(9);
  const double v = z * x;
  const double rs = fma(z, fma(z, fma(z, fma(z, S6, S5), S4), S3), S2);
  const double s = x - ((z * (0.5 * y - v * rs) - y) - v * S1);
  // fdlibm __kernel_cos(x, y)
  const double C1 = FM_K(10), C2 = FM_K(11), C3 = FM_K(12), C4 = FM_K(13), C5 = FM_K(14), C6 = FM_K(15);
  const double rc = z * fma(z, fma(z, fma(z, fma(z, fma(z, C6, C5), C4), C3), C2), C1);
  const uint32_t ix = fm_hi_word(x) & 0x7fffffffu;
  double qx = 0.0;                                  // |x| < 0.3: 1 - (z/2 - (z rc - x y))
  if (ix >= 0x3FD33333u) qx = (ix > 0x3fe90000u) ? 0.28125 : fm_hi_as_double(ix - 0x00200000u);  // x/4, low word cleared
  const double hz = 0.5 * z - qx;
  const double a = 1.0 - qx;
  const double c = a - (hz - (z * rc - x * y));
  // quadrant
  const double s_ = (k & 1) ? c : s, c_ = (k & 1) ? s : c;
  *sn = (k & 2) ? -s_ : s_;
  *cs = ((k + 1) & 2) ? -c_ : c_;
}

// ------------------------------------------------------------------------------------------------
// Table variant for |h| <= pi + a little (|h| < 3.3): h = k * pi/32 + r with |r| <= pi/64, sine and cosine of the
// 69 grid angles from a table, of r from two short series (r^9 / 9! < 5e-18), combined as
//   sin h = S_k + (S_k (cos r - 1) + C_k sin r),   cos h = C_k + (C_k (cos r - 1) - S_k sin r)
// so the table value enters unscaled and only a term <= 0.05 is computed: max error ~1 ulp (checked on the host by
// tests/test_fast_math.py).  26 fp64 instructions instead of ~70 for the polynomial-only routine above.
// ------------------------------------------------------------------------------------------------
#define FM_TAB_HALF 34                       // k = -34 .. 34 (|h| < 3.3 gives |k| <= 34)
#define FM_TAB_SIZE (2 * FM_TAB_HALF + 1)    // entries {sin, cos}

static inline void fm_fill_table(double *tab /* [FM_TAB_SIZE][2] */) {
  for (int k = -FM_TAB_HALF; k <= FM_TAB_HALF; k++) {
    // k * pi/32 evaluated exactly enough for a correctly rounded libm result: pi/32 = hi + lo
    const long double a = (long double)k * (3.14159265358979323846264338327950288L / 32.0L);
    tab[2 * (k + FM_TAB_HALF)] = (double)sinl(a);
    tab[2 * (k + FM_TAB_HALF) + 1] = (double)cosl(a);
  }
}

FM_HD void fm_sincos_tab(double h, const double *tab, double *sn, double *cs) {
  const double SHIFT = 6755399441055744.0;
  const double ks = fma(h, 1.01859163578813021189e+01 /* 32/pi */, SHIFT);
#ifdef __CUDA_ARCH__
  const int k = __double2loint(ks);
#else
  uint64_t kb;
  memcpy(&kb, &ks, 8);
  const int k = (int)(uint32_t)kb;
#endif
  const double kd = ks - SHIFT;
  // pi/32 = 9.81747704246810387019e-02: first 33 bits, then the rest (k <= 33: the first product is exact)
  const double r0 = fma(-kd, 9.81747704208828508854e-02, h);
  const double r = fma(-kd, 3.79818781656637015582e-12, r0);
  const double z = r * r;
  // sin r = r + r z (-1/6 + z (1/120 + z (-1/5040))),  cos r - 1 = z (-1/2 + z (1/24 + z (-1/720 + z / 40320)))
  const double ps = fma(z, fma(z, -1.98412698412698412526e-04, 8.33333333333333321769e-03), -1.66666666666666657415e-01);
  const double sr = fma(r * z, ps, r);
  const double cm = z * fma(z, fma(z, fma(z, 2.48015873015873015658e-05, -1.38888888888888894189e-03), 4.16666666666666643537e-02), -0.5);
#ifdef __CUDA_ARCH__
  const double2 sc = __ldg(reinterpret_cast<const double2 *>(tab) + (k + FM_TAB_HALF));  // one 128-bit load, L1 resident
  const double S = sc.x, C = sc.y;
#else
  const double S = tab[2 * (k + FM_TAB_HALF)], C = tab[2 * (k + FM_TAB_HALF) + 1];
#endif
  *sn = S + fma(S, cm, C * sr);
  *cs = C + fma(C, cm, -(S * sr));
}
