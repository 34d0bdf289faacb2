// crmath.cuh -- correctly rounded log / exp-near-zero for the RollingSharpe reward (rl4mm/rewards/RewardFunctions.py:10-22).
//
// get_sharpe takes np.diff(np.log(aum)) and np.exp(.) - 1 of AUMs that, at the reference's default cash of 1e12
// (rl4mm/helpers/main_helper.py:78), differ by 1e-10..1e-8 relative from step to step: ONE ulp of log(1e12) = 3.6e-15 is up to
// 1e-4 of such a return, so a libm that is merely "accurate to 1 ulp" (CUDA's log) changes the reward in the 4th digit.  numpy
// calls the platform's log (glibc: < 0.52 ulp, i.e. the correctly rounded value in all but a few percent of the arguments).  The
// device therefore computes log in double-double (~100 bits) and rounds ONCE: the correctly rounded result, which is what glibc
// returns except for arguments whose true logarithm lies within 0.02 ulp of a rounding boundary.
//   log x = k ln2 + log c + 2 atanh((m - c) / (m + c)),   x = 2^k m, m in [sqrt(1/2), sqrt(2)), c = round(32 m) / 32,
// |t| <= 0.0112, ten terms of the atanh series in double-double; ln2, log c and 1/(2n+1) are double-double constants (generated
// with 80-digit decimal arithmetic: tools/gen_crmath.py).  exp: Taylor to x^12 in double-double for |x| <= 2^-6, one rounding of 1 + p.
// Plain IEEE operations + fma only (compiled with --fmad=false / -ffp-contract=off), so host and device agree bit for bit;
// tests/test_abi_cpu.py checks 2e4 arguments against 60-digit decimal logarithms.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#if defined(__CUDACC__)
#define CRM_HD __host__ __device__ __forceinline__
#else
#define CRM_HD static inline
#endif

struct crm_dd { double hi, lo; };
CRM_HD crm_dd crm_two_sum(double a, double b) { crm_dd r; r.hi = a + b; const double bb = r.hi - a; r.lo = (a - (r.hi - bb)) + (b - bb); return r; }
CRM_HD crm_dd crm_quick_two_sum(double a, double b) { crm_dd r; r.hi = a + b; r.lo = b - (r.hi - a); return r; }
CRM_HD crm_dd crm_two_prod(double a, double b) { crm_dd r; r.hi = a * b; r.lo = fma(a, b, -r.hi); return r; }
CRM_HD crm_dd crm_add(crm_dd x, crm_dd y) {
  crm_dd s = crm_two_sum(x.hi, y.hi), t = crm_two_sum(x.lo, y.lo);
  s.lo += t.hi; s = crm_quick_two_sum(s.hi, s.lo);
  s.lo += t.lo; return crm_quick_two_sum(s.hi, s.lo);
}
CRM_HD crm_dd crm_mul(crm_dd x, crm_dd y) { crm_dd p = crm_two_prod(x.hi, y.hi); p.lo += x.hi * y.lo + x.lo * y.hi; return crm_quick_two_sum(p.hi, p.lo); }
CRM_HD crm_dd crm_mul_d(crm_dd x, double y) { crm_dd p = crm_two_prod(x.hi, y); p.lo += x.lo * y; return crm_quick_two_sum(p.hi, p.lo); }
CRM_HD crm_dd crm_div(crm_dd x, crm_dd y) {
  const double q1 = x.hi / y.hi;
  crm_dd p = crm_mul_d(y, q1); p.hi = -p.hi; p.lo = -p.lo;
  crm_dd r = crm_add(x, p);
  const double q2 = r.hi / y.hi;
  p = crm_mul_d(y, q2); p.hi = -p.hi; p.lo = -p.lo;
  r = crm_add(r, p);
  const double q3 = r.hi / y.hi;
  crm_dd q = crm_quick_two_sum(q1, q2);
  crm_dd t; t.hi = q3; t.lo = 0.0;
  return crm_add(q, t);
}
#define CRM_LN2_HI 0x1.62e42fefa39efp-1
#define CRM_LN2_LO 0x1.abc9e3b39803fp-56
// log(j / 32), j = 22 .. 46
#define CRM_LOGC_TABLE { \
  {-0x1.7fafa3bd8151cp-2, 0x1.219024acd3b77p-58}, \
  {-0x1.522ae0738a3d8p-2, 0x1.8f7e9b38a6979p-57}, \
  {-0x1.269621134db92p-2, -0x1.e0efadd9db02bp-56}, \
  {-0x1.f991c6cb3b379p-3, -0x1.f665066f980a2p-57}, \
  {-0x1.a93ed3c8ad9e3p-3, -0x1.bcafa9de97203p-57}, \
  {-0x1.5bf406b543db2p-3, 0x1.1f5b44c0df7e7p-61}, \
  {-0x1.1178e8227e47cp-3, 0x1.0e63a5f01c691p-58}, \
  {-0x1.9335e5d594989p-4, 0x1.478a85704ccb7p-58}, \
  {-0x1.08598b59e3a07p-4, 0x1.dd7009902bf32p-58}, \
  {-0x1.0415d89e74444p-5, -0x1.c05cf1d753622p-59}, \
  {0x0.0p+0, 0x0.0p+0}, \
  {0x1.f829b0e783300p-6, 0x1.33e3f04f1ef23p-60}, \
  {0x1.f0a30c01162a6p-5, 0x1.85f325c5bbacdp-59}, \
  {0x1.6f0d28ae56b4cp-4, -0x1.906d99184b992p-58}, \
  {0x1.e27076e2af2e6p-4, -0x1.61578001e0162p-60}, \
  {0x1.29552f81ff523p-3, 0x1.301771c407dbfp-57}, \
  {0x1.5ff3070a793d4p-3, -0x1.bc60efafc6f6ep-58}, \
  {0x1.9525a9cf456b4p-3, 0x1.d904c1d4e2e26p-57}, \
  {0x1.c8ff7c79a9a22p-3, -0x1.4f689f8434012p-57}, \
  {0x1.fb9186d5e3e2bp-3, -0x1.caaae64f21acbp-57}, \
  {0x1.1675cababa60ep-2, 0x1.ce63eab883717p-61}, \
  {0x1.2e8e2bae11d31p-2, -0x1.8f4cdb95ebdf9p-56}, \
  {0x1.4618bc21c5ec2p-2, 0x1.f42decdeccf1dp-56}, \
  {0x1.5d1bdbf5809cap-2, 0x1.4236383dc7fe1p-56}, \
  {0x1.739d7f6bbd007p-2, -0x1.8c76ceb014b04p-56}, \
}
// 1 / (2n + 1), n = 0 .. 9
#define CRM_ODD_INV_TABLE { \
  {0x1.0000000000000p+0, 0x0.0p+0}, \
  {0x1.5555555555555p-2, 0x1.5555555555555p-56}, \
  {0x1.999999999999ap-3, -0x1.999999999999ap-57}, \
  {0x1.2492492492492p-3, 0x1.2492492492492p-57}, \
  {0x1.c71c71c71c71cp-4, 0x1.c71c71c71c71cp-58}, \
  {0x1.745d1745d1746p-4, -0x1.745d1745d1746p-59}, \
  {0x1.3b13b13b13b14p-4, -0x1.3b13b13b13b14p-58}, \
  {0x1.1111111111111p-4, 0x1.1111111111111p-60}, \
  {0x1.e1e1e1e1e1e1ep-5, 0x1.e1e1e1e1e1e1ep-61}, \
  {0x1.af286bca1af28p-5, 0x1.af286bca1af28p-59}, \
}
// 1 / n!, n = 1 .. 12
#define CRM_INV_FACT_TABLE { \
  {0x1.0000000000000p+0, 0x0.0p+0}, \
  {0x1.0000000000000p-1, 0x0.0p+0}, \
  {0x1.5555555555555p-3, 0x1.5555555555555p-57}, \
  {0x1.5555555555555p-5, 0x1.5555555555555p-59}, \
  {0x1.1111111111111p-7, 0x1.1111111111111p-63}, \
  {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65}, \
  {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73}, \
  {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76}, \
  {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73}, \
  {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76}, \
  {0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80}, \
  {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83}, \
}

// the tables: __constant__ memory on the device (dynamic index j), plain statics on the host
static const crm_dd crm_logc_host[25] = CRM_LOGC_TABLE;
static const crm_dd crm_odd_inv_host[10] = CRM_ODD_INV_TABLE;
static const crm_dd crm_inv_fact_host[12] = CRM_INV_FACT_TABLE;
#if defined(__CUDACC__)
static __constant__ crm_dd crm_logc_dev[25] = CRM_LOGC_TABLE;
static __constant__ crm_dd crm_odd_inv_dev[10] = CRM_ODD_INV_TABLE;
static __constant__ crm_dd crm_inv_fact_dev[12] = CRM_INV_FACT_TABLE;
#endif
#if defined(__CUDA_ARCH__)
#define CRM_TAB(name) crm_##name##_dev
#else
#define CRM_TAB(name) crm_##name##_host
#endif

// correctly rounded natural logarithm of a positive, finite, normal double (anything else: the platform's log)
CRM_HD double cr_log(double x) {
  uint64_t bits; memcpy(&bits, &x, 8);
  const int ex = (int)((bits >> 52) & 0x7ff);
  if ((bits >> 63) || ex == 0 || ex == 0x7ff) return log(x);
  int k = ex - 1023;
  bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
  double m; memcpy(&m, &bits, 8);                                  // [1, 2)
  if (m >= 1.4142135623730951) { m *= 0.5; k += 1; }               // [sqrt(1/2), sqrt(2))
  const int j = (int)(m * 32.0 + 0.5);                             // 23 .. 45
  const double c = (double)j * 0.03125;
  const crm_dd* logc_tab = CRM_TAB(logc);
  const crm_dd* odd_inv = CRM_TAB(odd_inv);
  crm_dd num; num.hi = m - c; num.lo = 0.0;                        // exact (Sterbenz)
  const crm_dd t = crm_div(num, crm_two_sum(m, c));
  const crm_dd t2 = crm_mul(t, t);
  crm_dd p = odd_inv[9];
  for (int n = 8; n >= 0; n--) p = crm_add(crm_mul(p, t2), odd_inv[n]);
  crm_dd q = crm_mul(t, p);
  q.hi *= 2.0; q.lo *= 2.0;
  crm_dd ln2; ln2.hi = CRM_LN2_HI; ln2.lo = CRM_LN2_LO;
  crm_dd r = crm_add(crm_mul_d(ln2, (double)k), logc_tab[j - 22]);
  r = crm_add(r, q);
  return r.hi + r.lo;
}

// exp for the log-returns of get_sharpe: correctly rounded for |x| <= 2^-6 (where exp(x) - 1 would otherwise lose the return to
// the rounding of the platform's exp near 1), the platform's exp beyond
CRM_HD double cr_exp(double x) {
  if (!(fabs(x) <= 0.015625)) return exp(x);
  const crm_dd* inv_fact = CRM_TAB(inv_fact);
  crm_dd p = inv_fact[11];
  for (int n = 10; n >= 0; n--) p = crm_add(crm_mul_d(p, x), inv_fact[n]);
  p = crm_mul_d(p, x);                                             // e^x - 1
  crm_dd s = crm_two_sum(1.0, p.hi);
  return s.hi + (s.lo + p.lo);
}
