// ufm_pow.cuh -- x**y with the bits of the HOST's libm.
//
// The reference evaluates x**real through libm's pow (gfortran) and so does the CPU oracle.  CUDA's pow is within 2 ulp of
// it, which was the only source of GPU-vs-CPU differences in the dynamics -- and not a harmless one over a long run: the as-coded model
// has knife edges (a cell emptied by the out-flux limiter keeps a thickness of +-1e-17 m, and `Hi > 0` decides its masks), so one
// differing last bit in a velocity eventually flips a mask at one vertex and the trajectories part ways (first seen at step 3 of the
// 250 k-vertex bench workload, profiles/README.md).  The remedy is to compute pow exactly as the host does:
//
// glibc >= 2.28 (sysdeps/ieee754/dbl-64/e_pow.c, from ARM's optimized routines; here the FMA variant the ifunc resolver selects on every
// AVX2 machine) evaluates pow as exp(y * log(x)) with a 128-entry log table, a 128-entry exp table and two short polynomials, in plain
// IEEE double operations with fused multiply-adds at fixed places.  ufm_pow_main below is that instruction sequence, operation for
// operation; every IEEE operation gives the same bits on the GPU.  The tables are not copied into this repository: ufm_pow_host.cpp
// locates them in the running process's own libm.so.6, checks a host build of the same sequence against libm's pow on 200 000 random
// arguments, and only then uploads them.  If anything is off (another libm, a machine without FMA) the tables stay disabled and ufm_pow
// falls back to CUDA's pow: results are then within tolerance instead of bit-identical, and ufm_pow_mode() says so.
// tan (yield stress, grounding-line flux) is treated the same way on the range those routines use.
// Arguments off the main path (x <= 0, subnormal, non-finite; |y| < 2^-65 or >= 2^63; results near over/underflow) also go to CUDA's pow.
#pragma once
#include <stdint.h>
#include <math.h>

// x / d with the bits of the IEEE division, for a divisor whose correctly rounded reciprocal rd = 1.0 / d is already at hand (one real
// division per vertex instead of one per neighbour: the averages of map_Ac_to_Aa divide every term by the vertex degree).
// q0 = RN(x rd) is within 2 ulp of x/d; e = x - d q0 is exact in an fma; q0 + e rd differs from x/d by <= 2^-105 |x/d|, while for a divisor
// of at most 5 significant bits x/d is never closer than 2^-59 |x/d| to a rounding boundary -> RN(q0 + e rd) = RN(x/d).  Tiny and zero
// numerators (underflow in e) take the division itself, a zero is returned as it is (d > 0; zeros are common: ice-free cells).  Checked against the division on 10^9 numerators per
// divisor 1..17 when it was written and in tests/test_abi.py on every run (ufm_div_small_host is this function on the host).
#ifdef __CUDACC__
__host__ __device__ __forceinline__
#else
static inline
#endif
double ufm_div_small(const double x, const double d, const double rd)
{
  const double q0 = x * rd;
  double r = fma(fma(-q0, d, x), rd, q0);
  if (x == 0.0) r = x;                       // +-0 / d = +-0 (d > 0); the fma chain would turn -0 into +0
  else if (!(fabs(x) >= 1e-280)) r = x / d;  // underflow territory (and NaN): the division itself; never taken by model fields
  return r;
}

struct UfmPowTab {
  double ln2hi, ln2lo, A[7];
  double invln2N, shift, negln2hiN, negln2loN, C[4];
  int enabled, tan_enabled;
  double logt[128][4];            // invc, logc, logctail, -
  unsigned long long expt[128][2];  // tail, scale bits
  // tan on 0.07 <= |x| <= 0.78 (the friction angles of basal_yield_stress are 5..20 degrees = 0.087..0.349): glibc's s_tan.c there is
  // tan(x) = sign * (fi + pz*(fi+gi)/(gi-pz)), pz = z + z^3*(e0 + e1 z^2), z = |x| - xfg[i][0], (fi, gi) = (tan, cot) of the table point
  double tan_e0, tan_e1;
  double xfg[186][4];
};

#ifdef __CUDACC__
static __device__ UfmPowTab g_ufm_pow;   // one copy per translation unit; filled by ufm_powtab_upload() of that unit

__device__ __forceinline__ bool ufm_pow_main(const double x, const double y, double &out)
{
  const UfmPowTab &T = g_ufm_pow;
  const unsigned long long ix = (unsigned long long)__double_as_longlong(x), iy = (unsigned long long)__double_as_longlong(y);
  const unsigned topx = (unsigned)(ix >> 52), topy = (unsigned)(iy >> 52);
  if (topx - 1u > 0x7fdu) return false;
  if ((topy & 0x7ffu) - 0x3beu > 0x7fu) return false;
  const unsigned long long tmp = ix - 0x3fe6955500000000ull;
  const int i = (int)((tmp >> 45) & 127ull);
  const int k = (int)((long long)tmp >> 52);
  const double z = __longlong_as_double((long long)(ix - (tmp & (0xfffull << 52)))), kd = (double)k;
  const double2 e01 = *((const double2 *)&T.logt[i][0]);
  const double logctail = T.logt[i][2];
  const double r = fma(z, e01.x, -1.0);
  const double t1 = fma(kd, T.ln2hi, e01.y);
  const double t2 = t1 + r;
  const double lo1 = fma(kd, T.ln2lo, logctail);
  const double lo2 = t1 - t2 + r;
  const double ar = T.A[0] * r, ar2 = r * ar, ar3 = r * ar2;
  const double hi = t2 + ar2;
  const double lo3 = fma(ar, r, -ar2);
  const double lo4 = t2 - hi + ar2;
  const double p = fma(ar2, fma(ar2, fma(r, T.A[6], T.A[5]), fma(r, T.A[4], T.A[3])), fma(r, T.A[2], T.A[1]));
  const double lo = fma(ar3, p, lo1 + lo2 + lo3 + lo4);
  const double yl = hi + lo;
  const double tail = hi - yl + lo;
  const double ehi = y * yl;
  const double elo = fma(y, tail, fma(yl, y, -ehi));
  const unsigned abstop = (unsigned)(((unsigned long long)__double_as_longlong(ehi) >> 52) & 0x7ffull);
  if (abstop - 0x3c9u > 0x3eu) return false;
  const double zk = fma(ehi, T.invln2N, T.shift);
  const unsigned long long ki = (unsigned long long)__double_as_longlong(zk);
  const double kdd = zk - T.shift;
  double rr = fma(kdd, T.negln2loN, fma(kdd, T.negln2hiN, ehi));
  rr = elo + rr;
  const ulonglong2 et = *((const ulonglong2 *)&T.expt[ki & 127ull][0]);
  const double etail = __longlong_as_double((long long)et.x);
  const double scale = __longlong_as_double((long long)(et.y + (ki << 45)));
  const double r2 = rr * rr;
  const double q = fma(fma(rr, T.C[1], T.C[0]), r2, etail + rr);
  const double tmpv = fma(fma(rr, T.C[3], T.C[2]), r2 * r2, q);
  out = fma(tmpv, scale, scale);
  return true;
}
__device__ __forceinline__ double ufm_pow(const double x, const double y)
{
  double r;
  if (g_ufm_pow.enabled && ufm_pow_main(x, y, r)) return r;
  return pow(x, y);
}
__device__ __forceinline__ double ufm_tan(const double x)
{
  const UfmPowTab &T = g_ufm_pow;
  const double w = fabs(x);
  if (T.tan_enabled && w >= 0.07 && w <= 0.78) {
    const int i = (int)fma(w, 256.0, -15.5);
    const double2 xf = *((const double2 *)&T.xfg[i][0]);
    const double gi = T.xfg[i][2];
    const double z = w - xf.x, z2 = z * z;
    const double pz = fma(z * z2, fma(z2, T.tan_e1, T.tan_e0), z);
    const double t2 = ((xf.y + gi) * pz) / (gi - pz);
    return (x < 0.0 ? -1.0 : 1.0) * (xf.y + t2);
  }
  return tan(x);
}
// upload the tables into THIS translation unit's copy (current device)
static int ufm_powtab_upload_tu(const UfmPowTab *host) { return (int)cudaMemcpyToSymbol(g_ufm_pow, host, sizeof(UfmPowTab)); }
#endif
