// ufm_pow_host.cpp -- finds glibc's pow tables in the running process's libm and validates the re-implementation (see ufm_pow.cuh)
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ufm_pow.cuh"

static inline uint64_t asu(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double asd(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

// the same operation sequence as ufm_pow_main (device), on the host: fma() is exact by definition, the rest plain IEEE (this file is
// compiled with -ffp-contract=off like everything else)
static bool pow_main_host(const UfmPowTab &T, double x, double y, double *out)
{
  const uint64_t ix = asu(x), iy = asu(y);
  const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
  if (topx - 1u > 0x7fdu) return false;
  if ((topy & 0x7ffu) - 0x3beu > 0x7fu) return false;
  const uint64_t tmp = ix - 0x3fe6955500000000ull;
  const int i = (int)((tmp >> 45) & 127), k = (int)((int64_t)tmp >> 52);
  const double z = asd(ix - (tmp & (0xfffull << 52))), kd = (double)k;
  const double r = fma(z, T.logt[i][0], -1.0);
  const double t1 = fma(kd, T.ln2hi, T.logt[i][1]);
  const double t2 = t1 + r;
  const double lo1 = fma(kd, T.ln2lo, T.logt[i][2]);
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
  const uint32_t abstop = (uint32_t)((asu(ehi) >> 52) & 0x7ff);
  if (abstop - 0x3c9u > 0x3eu) return false;
  const double zk = fma(ehi, T.invln2N, T.shift);
  const uint64_t ki = asu(zk);
  const double kdd = zk - T.shift;
  double rr = fma(kdd, T.negln2loN, fma(kdd, T.negln2hiN, ehi));
  rr = elo + rr;
  const double etail = asd(T.expt[ki & 127][0]);
  const double scale = asd(T.expt[ki & 127][1] + (ki << 45));
  const double r2 = rr * rr;
  const double q = fma(fma(rr, T.C[1], T.C[0]), r2, etail + rr);
  const double tmpv = fma(fma(rr, T.C[3], T.C[2]), r2 * r2, q);
  *out = fma(tmpv, scale, scale);
  return true;
}

static double tan_mid_host(const UfmPowTab &T, double x)
{
  const double w = fabs(x);
  const int i = (int)fma(w, 256.0, -15.5);
  const double z = w - T.xfg[i][0], z2 = z * z;
  const double pz = fma(z * z2, fma(z2, T.tan_e1, T.tan_e0), z);
  const double fi = T.xfg[i][1], gi = T.xfg[i][2];
  const double t2 = ((fi + gi) * pz) / (gi - pz);
  return (x < 0.0 ? -1.0 : 1.0) * (fi + t2);
}

// glibc's tan on 0.07 <= |x| <= 0.78: the table xfg[186] = {x_i, tan x_i, cot x_i, -} is found by its structure (x_i within half a step of
// (i + 16) / 256), the two polynomial coefficients (e0 ~ 1/3, e1 ~ 2/15) among the doubles of libm close to those values: the pair that
// reproduces libm's tan on two million arguments bit for bit.  Leaves tan_enabled = 0 when nothing fits.
static void tantab_build(UfmPowTab *out, const std::vector<unsigned char> &b)
{
  const long n = (long)b.size();
  long pt = -1;
  for (long i = 0; i + 186 * 32 <= n && pt < 0; i += 8) {
    double r0[4];
    memcpy(r0, b.data() + i, 32);
    if (!(r0[0] > 15.5 / 256.0 && r0[0] < 16.5 / 256.0) || fabs(r0[1] - tan(r0[0])) > 1e-14) continue;
    bool ok = true;
    for (int k = 0; k < 186 && ok; k++) {
      double r[4];
      memcpy(r, b.data() + i + 32 * k, 32);
      ok = r[0] > (k + 15.5) / 256.0 && r[0] < (k + 16.5) / 256.0 && fabs(r[1] - tan(r[0])) <= 1e-14 && fabs(r[1] * r[2] - 1.0) <= 1e-14;
    }
    if (ok) pt = i;
  }
  if (pt < 0) return;
  memcpy(out->xfg, b.data() + pt, sizeof(out->xfg));
  std::vector<double> c0, c1;
  for (long i = 0; i + 8 <= n; i += 8) {
    double v;
    memcpy(&v, b.data() + i, 8);
    if (fabs(v - 1.0 / 3.0) < 1e-6 && c0.size() < 32 && std::find(c0.begin(), c0.end(), v) == c0.end()) c0.push_back(v);
    if (fabs(v - 2.0 / 15.0) < 1e-5 && c1.size() < 32 && std::find(c1.begin(), c1.end(), v) == c1.end()) c1.push_back(v);
  }
  for (double a : c0)
    for (double c : c1) {
      out->tan_e0 = a; out->tan_e1 = c;
      uint64_t s = 0x2545F4914F6CDD1Dull;
      auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); };
      bool ok = true;
      for (int t = 0; t < 2000000 && ok; t++) {
        const double x = (t & 1) ? (3.141592653589793 / 180.0) * (5.0 + 15.0 * rnd()) : (0.07 + 0.71 * rnd()) * ((t & 2) ? -1.0 : 1.0);
        ok = asu(tan_mid_host(*out, x)) == asu(tan(x));
      }
      if (ok) { out->tan_enabled = 1; return; }
    }
  out->tan_e0 = out->tan_e1 = 0.0;
}

// host entry point for tests (ctypes): the re-implementation's result, or libm's pow off the main path / when the tables are disabled
static UfmPowTab g_host_tab;
static int g_host_state = 0;   // 0 not tried, 1 enabled, -1 disabled

// Fills *out (enabled = 1 on success).  Returns 0 when the tables were found and the self-test passed, else a negative reason code.
extern "C" int ufm_powtab_build(UfmPowTab *out)
{
  memset(out, 0, sizeof(*out));
  if (getenv("UFM_POW_EXACT") && atoi(getenv("UFM_POW_EXACT")) == 0) return -1;
  Dl_info info;
  if (!dladdr((void *)(double (*)(double, double))&pow, &info) || !info.dli_fname) return -2;
  FILE *f = fopen(info.dli_fname, "rb");
  if (!f) return -3;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> b((size_t)n);
  const bool read_ok = fread(b.data(), 1, (size_t)n, f) == (size_t)n;
  fclose(f);
  if (!read_ok) return -3;
  // struct pow_log_data { ln2hi, ln2lo, poly[7], tab[128] = {invc, pad, logc, logctail} } and
  // struct exp_data { invln2N, shift, negln2hiN, negln2loN, poly[4], exp2_shift, exp2_poly[5], (exp10: 3 + 5), tab[2*128] }: found by their
  // leading constants (ln2 split for N = 128; 128 / ln2 and the rounding shift 0x1.8p52)
  const double a1[3] = {0x1.62e42fefa3800p-1, 0x1.ef35793c76730p-45, -0x1p-1};
  const double a2[2] = {0x1.71547652b82fep0 * 128, 0x1.8p52};
  long p1 = -1, p2 = -1;
  for (long i = 0; i + 8 * (9 + 4 * 128) <= n; i += 8) if (!memcmp(b.data() + i, a1, sizeof(a1))) { p1 = i; break; }
  for (long i = 0; i + 8 * (22 + 256) <= n; i += 8) if (!memcmp(b.data() + i, a2, sizeof(a2))) { p2 = i; break; }
  if (p1 < 0 || p2 < 0) return -4;
  double d[9 + 4 * 128], e[22];
  memcpy(d, b.data() + p1, sizeof(d));
  memcpy(e, b.data() + p2, sizeof(e));
  out->ln2hi = d[0]; out->ln2lo = d[1];
  for (int k = 0; k < 7; k++) out->A[k] = d[2 + k];
  for (int i = 0; i < 128; i++) { out->logt[i][0] = d[9 + 4 * i]; out->logt[i][1] = d[9 + 4 * i + 2]; out->logt[i][2] = d[9 + 4 * i + 3]; out->logt[i][3] = 0.0; }
  out->invln2N = e[0]; out->shift = e[1]; out->negln2hiN = e[2]; out->negln2loN = e[3];
  for (int k = 0; k < 4; k++) out->C[k] = e[4 + k];
  memcpy(out->expt, b.data() + p2 + 8 * 22, sizeof(out->expt));
  // self-test against the libm this process actually calls: arguments as the hot path produces them and a broad sweep
  uint64_t s = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); };
  const double ys[] = {3.0, 5.0, -1.0 / 3.0, -0.35, 4.0 / 3.0, 3.0 / 7.0, 1.0 / 3.0, 2.0, -2.0 / 3.0, 0.3};
  long on_path = 0;
  for (int t = 0; t < 200000; t++) {
    double x;
    switch (t & 3) {
      case 0: x = exp((rnd() - 0.5) * 80.0); break;
      case 1: x = 1e-12 + rnd() * 1e-3; break;
      case 2: x = rnd() * 1e7 + 1e-300; break;
      default: x = asd(asu(1.0) + (uint64_t)(rnd() * 4e15)); break;
    }
    const double y = (t % 3 == 0) ? ys[(t / 3) % 10] : (rnd() - 0.5) * 12.0;
    double mine;
    if (!pow_main_host(*out, x, y, &mine)) continue;
    on_path++;
    if (asu(mine) != asu(pow(x, y))) { memset(out, 0, sizeof(*out)); return -5; }
  }
  if (on_path < 100000) { memset(out, 0, sizeof(*out)); return -6; }
  out->enabled = 1;
  tantab_build(out, b);
  return 0;
}

// bit 0: this host's libm is the glibc whose pow the device re-states (tables found, 200 000 arguments bit-identical); bit 1: likewise tan.
// What ufm_pow_mode reports for a handle, without a device.
extern "C" int ufm_powtab_status(void)
{
  if (g_host_state == 0) g_host_state = ufm_powtab_build(&g_host_tab) == 0 ? 1 : -1;
  return g_host_state == 1 ? (g_host_tab.enabled ? 1 : 0) | (g_host_tab.tan_enabled ? 2 : 0) : 0;
}

extern "C" double ufm_div_small_host(double x, int n)
{
  const double d = (double)n;
  return ufm_div_small(x, d, 1.0 / d);
}

extern "C" double ufm_tan_host(double x)
{
  if (g_host_state == 0) g_host_state = ufm_powtab_build(&g_host_tab) == 0 ? 1 : -1;
  const double w = fabs(x);
  if (g_host_state == 1 && g_host_tab.tan_enabled && w >= 0.07 && w <= 0.78) return tan_mid_host(g_host_tab, x);
  return tan(x);
}

extern "C" double ufm_pow_host(double x, double y)
{
  if (g_host_state == 0) g_host_state = ufm_powtab_build(&g_host_tab) == 0 ? 1 : -1;
  double r;
  if (g_host_state == 1 && pow_main_host(g_host_tab, x, y, &r)) return r;
  return pow(x, y);
}
