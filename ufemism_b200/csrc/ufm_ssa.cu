// ufm_ssa.cu -- SSA stress balance on the combined AaAc mesh: hand-written sm_100a kernels.
//
// Replaces the bodies of solve_SSA / solve_SSA_linearised / SSA_effective_viscosity /
// SSA_sliding_term / basal_yield_stress / calculate_GL_flux (src/ice_dynamics_module.f90:408-949)
// and the AaAc operators they call (src/mesh_ArakawaC_module.f90:583-724,815-844).
//
// Arithmetic contract: this translation unit is compiled with -fmad=false, every expression is
// written in the reference's evaluation order, fp64 add/mul/div/sqrt are IEEE-exact on the GPU, so
// the SOR sweep, the Neumann pass and the linear-system setup reproduce the CPU restatement BIT FOR
// BIT; only pow() and tan() (CUDA libm <= 2 ulp vs glibc <= 1 ulp) can differ, by O(1e-16) relative.
//
// The hot loop is ONE persistent cooperative kernel per linear solve: 5 colour phases + 1 boundary
// phase per SOR iteration separated by grid-wide barriers, the max-residual reduced with warp
// shuffles + one atomicMax per CTA, the stop tests evaluated on the device.  No tensor cores: nothing
// here is a dense contraction (irregular 4..8-point stencils); the bound is HBM bandwidth.
#include <cooperative_groups.h>
#include <math.h>

#include <algorithm>
#include <climits>
#include <cstdio>

#include "ufm_internal.cuh"
#include "ufm_pow.cuh"

namespace cg = cooperative_groups;

// streaming (read-once) loads: evict-first so that (U,V) stays L2-resident across colour phases.
// UFM_LD_STREAM selects the cache policy at build time (tuning builds, tools/build_variant.py): 0 = ld.global.cs (LDG.EF),
// 1 = L1::no_allocate (LDG.NA), 2 = L1::no_allocate + L2 evict-first policy, 3 = ld.global.cg (L1 bypassed).
#ifndef UFM_LD_STREAM
#define UFM_LD_STREAM 0
#endif
#if UFM_LD_STREAM == 0
template <class T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
#elif UFM_LD_STREAM == 3
template <class T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcg(p); }
#elif UFM_LD_STREAM == 1
__device__ __forceinline__ int ld_stream(const int *p) { int v; asm("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ double ld_stream(const double *p) { double v; asm("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ double2 ld_stream(const double2 *p) { double2 v; asm("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }
#else
__device__ __forceinline__ unsigned long long ld_policy() { unsigned long long q; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(q)); return q; }
__device__ __forceinline__ int ld_stream(const int *p) { int v; asm("ld.global.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(ld_policy())); return v; }
__device__ __forceinline__ double ld_stream(const double *p) { double v; asm("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(ld_policy())); return v; }
__device__ __forceinline__ double2 ld_stream(const double2 *p) { double2 v; asm("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(ld_policy())); return v; }
#endif

// ---------------------------------------------------------------------------------------------
// Inter-GPU signalling over NVLink (vertex-partitioned runs).  Every rank owns a mailbox in its own
// HBM that its peers write with plain system-scope stores through CUDA-IPC mapped pointers; a rank only ever
// spins on its OWN memory.  peer_sync is an all-to-all epoch barrier executed by exactly one thread per GPU.
// A wait that exceeds SPIN_LIMIT polls (seconds) raises MAIL_ABORT instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
#define SPIN_LIMIT 40000000LL
// device-side control of one solve_SSA call (words of DevState::ctrl): lets a batch of outer iterations run back to back
// without a host round trip; every kernel of the batch first looks at SCTL_STOP
#define SCTL_BASE 96
#define SCTL_STOP 0       // 1: converged (RN < tol), aborted, or failed -> remaining kernels of the batch return at once
#define SCTL_NOUTER 1     // viscosity iterations started
#define SCTL_NINNER 2     // SOR iterations, total
#define SCTL_RESET 3      // velocities were reset once
#define SCTL_RC 4         // bit0 SOR hit max_inner (warning), bit1 unstable twice (abort), bit2 peer wait timed out
#define SCTL_RN 5         // last RN (double bits)
#define SCTL_MAXRES 6     // last max residual (double bits)
#define SCTL_NLAST 7      // SOR iterations of the last linear solve
__device__ __forceinline__ void peer_sync(const CommDev &cm)
{
  __threadfence_system();
  volatile unsigned long long *mine = cm.mail[cm.rank];
  const unsigned long long e = mine[MAIL_EPOCH] + 1ull;
  mine[MAIL_EPOCH] = e;
  for (int q = 0; q < cm.P; q++)
    if (q != cm.rank) *((volatile unsigned long long *)(cm.mail[q] + MAIL_FLAG + cm.rank)) = e;
  for (int q = 0; q < cm.P; q++) {
    if (q == cm.rank) continue;
    long long spins = 0;
    while (mine[MAIL_FLAG + q] < e) {
      if (++spins > SPIN_LIMIT) { mine[MAIL_ABORT] = 1ull; break; }
    }
  }
  __threadfence_system();
}
__global__ void k_peer_barrier(CommDev cm) { if (threadIdx.x == 0 && blockIdx.x == 0) peer_sync(cm); }

// Grid-wide barrier for the persistent kernels (all CTAs co-resident: cooperative launch).  The last CTA to arrive
// runs `last()` (the NVLink epoch exchange in partitioned runs) before it releases the others.
// HOOK variant: bar[32] = arrival count, bar[64] = generation (separate 128 B lines).
// MULTI variant (vertex-partitioned run): bar[32] = arrival count, bar[64] = generation (= low bits of the epoch).
// The last CTA to arrive releases the local CTAs, then runs pre() (e.g. publishes this rank's residual), issues ONE
// system-scope fence (cumulative over the peer pushes of all CTAs, which it has observed through the arrival atomics) and
// signals `epoch` into the mailbox of every peer.  Nobody waits for the peers here: rows that read peer-owned rows are
// swept last in a phase and wait_peers() is called just before them, so the NVLink latency hides behind interior work.
template <bool MULTI, class F>
__device__ __forceinline__ void grid_barrier(unsigned *bar, const int nblocks, const CommDev &cm, const unsigned long long epoch, const unsigned to, F &&pre)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!MULTI) {
      // single word: every CTA adds 1, CTA 0 adds 2^31 - (n-1), so bit 31 flips exactly when the last CTA arrives
      // and the pollers see the release as a side effect of that last atomic (one L2 round trip less)
      const unsigned add = blockIdx.x == 0 ? 0x80000000u - (unsigned)(nblocks - 1) : 1u;
      __threadfence();
      const unsigned old = atomicAdd(bar, add);
      while ((((old ^ *((volatile unsigned *)bar)) & 0x80000000u) == 0u)) { }
      __threadfence();
    } else {
      volatile unsigned *gen = bar + 64;
      __threadfence();
      if (atomicAdd(bar + 32, 1u) == (unsigned)nblocks - 1u) {
        *((volatile unsigned *)(bar + 32)) = 0u;
        __threadfence();
        *gen = (unsigned)epoch;
        pre();
        __threadfence_system();
        // `to`: the ranks that will wait for this epoch -- after a colour sweep only those that read rows of this rank (the strip's
        // neighbours); after the fifth colour (max residual) and after the Neumann pass (end of the solve) everybody
        for (int q = 0; q < cm.P; q++)
          if (q != cm.rank && ((to >> q) & 1u)) *((volatile unsigned long long *)(cm.mail[q] + MAIL_FLAG + cm.rank)) = epoch;
      } else {
        while (*gen != (unsigned)epoch) { }
      }
      __threadfence();
    }
  }
  __syncthreads();
}
// wait (on local memory) until every peer in `from` has signalled `epoch`; called by one lane, followed by a fence that also drops
// stale L1 lines of rows the peers have pushed
__device__ __forceinline__ void wait_peers(const CommDev &cm, const unsigned long long epoch, const unsigned from)
{
  volatile unsigned long long *mine = cm.mail[cm.rank];
  for (int q = 0; q < cm.P; q++) {
    if (q == cm.rank || !((from >> q) & 1u)) continue;
    long long spins = 0;
    while (mine[MAIL_FLAG + q] < epoch) {
      if (++spins > SPIN_LIMIT) { mine[MAIL_ABORT] = 1ull; break; }
    }
  }
  __threadfence();
}

__device__ __forceinline__ bool d_is_floating(double Hi, double Hb, double SL)
{
  return Hi < (SL - Hb) * UFM_SEAWATER_DENSITY / UFM_ICE_DENSITY;  // general_ice_model_data_module.f90:464-475
}

// ---------------------------------------------------------------------------------------------
// calculate_GL_flux, ice_dynamics_module.f90:848-949 (Coulomb_regularised).  One thread per Ac.
// ---------------------------------------------------------------------------------------------
__global__ void k_gl_flux(int nAc, const int4 *__restrict__ Aci, const unsigned *__restrict__ mbits_Ac, const unsigned *__restrict__ mbits,
                          const double *__restrict__ Hi, const double *__restrict__ Hb, const double *__restrict__ SL,
                          const double *__restrict__ phi_m, const int *__restrict__ aa2m, double A_flow,
                          const double *__restrict__ dHi_dx, const double *__restrict__ dHi_dy, const double *__restrict__ dSL_dx,
                          const double *__restrict__ dSL_dy, const double *__restrict__ dHb_dx, const double *__restrict__ dHb_dy,
                          const double *__restrict__ Dx_, const double *__restrict__ Dy_, double factor_Tsai_noA,
                          double *Qabs, double *Qp, double *Ux, double *Uy, const int *__restrict__ ac2m, double2 *UV,
                          const double *__restrict__ A_mean, double tsai_c1, double tsai_c2, double tsai_c3,
                          const unsigned char *__restrict__ own, int rank)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nAc || !UFM_OWNED(own, a, rank)) return;
  Qabs[a] = 0.0; Qp[a] = 0.0;
  if (!(mbits_Ac[a] & MB_GL)) return;
  int4 v = Aci[a];
  const double rr = UFM_SEAWATER_DENSITY / UFM_ICE_DENSITY;
  double TAFi = Hi[v.x] - ((SL[v.x] - Hb[v.x]) * rr);
  double TAFj = Hi[v.y] - ((SL[v.y] - Hb[v.y]) * rr);
  double lambda_GL = TAFi / (TAFi - TAFj);
  double Hi_GL = (Hi[v.x] * (1.0 - lambda_GL)) + (Hi[v.y] * lambda_GL);
  // ice%phi_fric_AaAc of the sheet-side vertex.  In a partitioned run that vertex may belong to the neighbour strip, whose yield-stress
  // pass this rank does not see: the angle is a function of the bed alone (basal_yield_stress :806-808), re-evaluated here with the same expression
  const int vg = (mbits[v.x] & MB_SHEET) ? v.x : v.y;
  double phi_fric_GL = own ? fmax(5.0, fmin(20.0, (5.0 + (20.0 - 5.0) * (1.0 + (Hb[vg] - 0.0) / (0.0 - (-1000.0)))))) : phi_m[aa2m[vg]];
  // factor_Tsai = 8 Q0 A (rho g)^n (1-rho_i/rho_w)^(n-1) / 4^n ; the A-independent part is evaluated on the host
  double factor_Tsai = A_flow * factor_Tsai_noA;
  if (A_mean) {  // temperature-dependent flow factor of the grounded side (:894-900); factor evaluated left to right as at :906-909
    const double A_GL = (mbits[v.x] & MB_SHEET) ? A_mean[v.x] : A_mean[v.y];
    factor_Tsai = 8.0 * 0.61 * A_GL * tsai_c1 * tsai_c2 / tsai_c3;
  }
  double q = factor_Tsai * ufm_pow(Hi_GL, UFM_N_FLOW + 2.0) / ufm_tan(phi_fric_GL * (UFM_PI / 180.0));
  Qabs[a] = q;
  double Fx = -(dHi_dx[a] - ((dSL_dx[a] - dHb_dx[a]) * rr));
  double Fy = -(dHi_dy[a] - ((dSL_dy[a] - dHb_dy[a]) * rr));
  double F = ufm_norm2_2(Fx, Fy);
  Fx = Fx / F; Fy = Fy / F;
  const double ux = q * Fx / Hi_GL, uy = q * Fy / Hi_GL;
  Ux[a] = ux; Uy[a] = uy;
  UV[ac2m[a]] = make_double2(ux, uy);  // the gather of :494-495 happens after calculate_GL_flux in the reference
  double Dx = Dx_[a], Dy = Dy_[a], D = ufm_norm2_2(Dx, Dy);
  Dx = Dx / D; Dy = Dy / D;
  Qp[a] = q * (Dx * Fx + Dy * Fy);
}

// ---------------------------------------------------------------------------------------------
// basal_yield_stress (:780-844) + gather of the Aa and Ac fields into AaAc order (:478-496).
// One thread per AaAc row.
// ---------------------------------------------------------------------------------------------
struct PrepArgs {
  int Mp;
  const int *src;
  const double *Hi, *Hb, *SL, *sx, *sy, *U, *V;              // Aa
  const double *Hi_Ac, *Hb_Ac, *SL_Ac, *sx_Ac, *sy_Ac, *Ux_Ac, *Uy_Ac;  // Ac
  const unsigned *mbits_Ac;
  int gl_fix;
  const double *A_mean, *A_mean_Ac;  // realistic flow factor (else NULL)
  double m_enh_ssa;
  double *Afac;
  double2 *UV, *rhsnum;
  double *tau_c, *phi, *Hm;
  unsigned char *mflag;
  const unsigned char *sowner; int rank;   // partitioned per-step kernels: rows of this rank's slices only (else NULL)
};
__global__ void k_ssa_prepare(PrepArgs a)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.Mp) return;
  if (a.sowner && a.sowner[p >> 5] != a.rank) return;
  int s = a.src[p];
  if (s == INT_MIN) return;
  double Hi, Hb, SL, sx, sy, U, V;
  unsigned char fl = 0;
  if (a.A_mean) {
    const double A = s >= 0 ? a.A_mean[s] : a.A_mean_Ac[~s];
    a.Afac[p] = ufm_pow(a.m_enh_ssa * 0.5 * A, -1.0 / UFM_N_FLOW);   // first factor of eta (ice_dynamics_module.f90:718)
  }
  if (s >= 0) { Hi = a.Hi[s]; Hb = a.Hb[s]; SL = a.SL[s]; sx = a.sx[s]; sy = a.sy[s]; U = a.U[s]; V = a.V[s]; }
  else {
    s = ~s;
    Hi = a.Hi_Ac[s]; Hb = a.Hb_Ac[s]; SL = a.SL_Ac[s]; sx = a.sx_Ac[s]; sy = a.sy_Ac[s]; U = a.Ux_Ac[s]; V = a.Uy_Ac[s];
    if (a.gl_fix && (a.mbits_Ac[s] & MB_GL)) fl |= 2;
  }
  const double pf1 = -1000.0, pf2 = 0.0, p_min = 5.0, p_max = 20.0;
  double lambda_p = fmax(0.0, fmin(1.0, (1.0 - (Hb - SL) / 1000.0)));
  double Hm = fmax(0.1, Hi);
  double pore_water_pressure = 0.96 * UFM_ICE_DENSITY * UFM_GRAV * Hm * lambda_p;
  double phi = fmax(p_min, fmin(p_max, (p_min + (p_max - p_min) * (1.0 + (Hb - pf2) / (pf2 - pf1)))));
  double tau_c = ufm_tan((UFM_PI / 180.0) * phi) * (UFM_ICE_DENSITY * UFM_GRAV * Hm - pore_water_pressure);
  if (!d_is_floating(Hi, Hb, SL)) fl |= 1;
  a.UV[p] = make_double2(U, V);
  a.rhsnum[p] = make_double2(UFM_ICE_DENSITY * UFM_GRAV * sx, UFM_ICE_DENSITY * UFM_GRAV * sy);
  a.tau_c[p] = tau_c; a.phi[p] = phi; a.Hm[p] = Hm; a.mflag[p] = fl;
}

// ---------------------------------------------------------------------------------------------
// SSA_effective_viscosity (:695-726) incl. get_mesh_derivatives_AaAc (mesh_ArakawaC_module.f90:583-618),
// N = eta*max(0.1,H), and the per-block partial sums of (N-Nprev)^2 and N^2 (:512-513).
// One warp per slice, one thread per row, grid-stride over slices.
// ---------------------------------------------------------------------------------------------
struct ViscArgs {
  CommDev cm;
  const int *rng;
  int n_slices;
  const long long *off;
  const unsigned char *deg;
  const int *idx;
  const double *nx, *ny, *nx0, *ny0, *Hm;
  const double2 *UV;
  double visc_A;  // (m_enh_ssa * 0.5 * A_flow)**(-1/n_flow), host libm (benchmark flow factor)
  const double *Afac;  // per-row factor (temperature-dependent flow factor), or NULL
  // fused SSA_sliding_term + RHS + centre coefficients (k_ssa_setup's work, done while eta is still in a register)
  int fuse_setup;
  const unsigned char *mflag;
  const double2 *rhsnum;
  const double *tau_c, *cU0, *cV0;
  double thr;
  double *S;
  double2 *RHS, *E;
  const unsigned long long *sctl;  // device-side solve control (SCTL_*), or NULL
  double *eta, *N;
  double2 *dU, *dV;
  double *partials;
};
template <int W>
__device__ __forceinline__ void visc_row(const ViscArgs &a, const long long o, const int lane, const int p, const int n,
                                         double &ux, double &uy, double &vx, double &vy)
{
  int j[W];
  double cx[W], cy[W];
#pragma unroll
  for (int c = 0; c < W; c++) {
    const long long e = o + (long long)c * 32 + lane;
    j[c] = ld_stream(a.idx + e); cx[c] = ld_stream(a.nx + e); cy[c] = ld_stream(a.ny + e);
  }
  const double2 u = a.UV[p];
  const double hx = ld_stream(a.nx0 + p), hy = ld_stream(a.ny0 + p);
  double2 nb[W];
#pragma unroll
  for (int c = 0; c < W; c++) nb[c] = a.UV[j[c]];
  ux = hx * u.x; uy = hy * u.x; vx = hx * u.y; vy = hy * u.y;
#pragma unroll
  for (int c = 0; c < W; c++)
    if (c < n) { ux = ux + cx[c] * nb[c].x; uy = uy + cy[c] * nb[c].x; vx = vx + cx[c] * nb[c].y; vy = vy + cy[c] * nb[c].y; }
}

// One warp per slice, slices of this rank's six block ranges taken grid-stride (no block-level synchronisation in the
// streaming loop).  The two sums are reduced per slice with a fixed xor-shuffle tree and stored at partials[slice] (and
// pushed to the peers of a partitioned run), so the final fixed-shape reduction over ALL slices gives the same bits for
// any number of GPUs and any grid size.
template <bool STORE_GRAD, int MINB>
__global__ void __launch_bounds__(256, MINB) k_ssa_viscosity(ViscArgs a)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  if (a.sctl && a.sctl[SCTL_STOP]) return;   // the solve has already converged / aborted (device-side control flow)
  for (int b = 0; b < 6; b++) {
    const int s0 = a.rng[(b * a.cm.P + a.cm.rank) * 3], s1 = a.rng[(b * a.cm.P + a.cm.rank) * 3 + 2];
    for (int s = s0 + wg; s < s1; s += nw) {
      const long long o = a.off[s];
      const int w = (int)((a.off[s + 1] - o) >> 5);
      const int p = s * 32 + lane;
      const int n = a.deg[p];
      double s_dn = 0.0, s_n = 0.0;
      if (n != UFM_DEG_PAD) {
        // row-local operands of the fused epilogue are requested up front so that they are in flight together with the
        // stencil streams (otherwise they would only be issued after the first pow())
        double hm = 0.0, nold = 0.0, tauc = 0.0, cu0 = 0.0, cv0 = 0.0;
        double2 rn = make_double2(0.0, 0.0);
        unsigned char mf = 0;
        if (!STORE_GRAD) {
          hm = a.Hm[p]; nold = a.N[p];
          if (a.fuse_setup) { tauc = ld_stream(a.tau_c + p); rn = ld_stream(a.rhsnum + p); cu0 = ld_stream(a.cU0 + p); cv0 = ld_stream(a.cV0 + p); mf = a.mflag[p]; }
        }
        double ux, uy, vx, vy;
        switch (w) {
          case 2: visc_row<2>(a, o, lane, p, n, ux, uy, vx, vy); break;
          case 3: visc_row<3>(a, o, lane, p, n, ux, uy, vx, vy); break;
          case 4: visc_row<4>(a, o, lane, p, n, ux, uy, vx, vy); break;
          case 5: visc_row<5>(a, o, lane, p, n, ux, uy, vx, vy); break;
          case 6: visc_row<6>(a, o, lane, p, n, ux, uy, vx, vy); break;
          case 7: visc_row<7>(a, o, lane, p, n, ux, uy, vx, vy); break;
          case 8: visc_row<8>(a, o, lane, p, n, ux, uy, vx, vy); break;
          default: {
            const double2 u = a.UV[p];
            ux = a.nx0[p] * u.x; uy = a.ny0[p] * u.x; vx = a.nx0[p] * u.y; vy = a.ny0[p] * u.y;
            for (int c = 0; c < w; c++) {
              if (c < n) {
                const long long e = o + (long long)c * 32 + lane;
                const double2 nb = a.UV[a.idx[e]];
                const double cx = a.nx[e], cy = a.ny[e];
                ux = ux + cx * nb.x; uy = uy + cy * nb.x;
                vx = vx + cx * nb.y; vy = vy + cy * nb.y;
              }
            }
          }
        }
        if (STORE_GRAD) { a.dU[p] = make_double2(ux, uy); a.dV[p] = make_double2(vx, vy); }
        else {
          const double epsilon_sq_0 = 1E-12;
          const double eta = (a.Afac ? a.Afac[p] : a.visc_A) * ufm_pow(ux * ux + vy * vy + ux * vy + 0.25 * ((uy + vx) * (uy + vx)) + epsilon_sq_0, (1.0 - UFM_N_FLOW) / (2.0 * UFM_N_FLOW));
          const double Nn = eta * hm;
          const double dn = Nn - nold;
          s_dn = dn * dn; s_n = Nn * Nn;
          a.eta[p] = eta; a.N[p] = Nn;
          if (a.fuse_setup) {   // identical expressions to k_ssa_setup
            const double delta_v = 1E-3, q_plastic = 0.30;
            const double2 u = a.UV[p];
            const double S = tauc * (ufm_pow(delta_v * delta_v + u.x * u.x + u.y * u.y, 0.5 * (q_plastic - 1.0))) / a.thr;
            a.RHS[p] = make_double2(rn.x / eta, rn.y / eta);
            double eu = cu0, ev = cv0;
            if (mf & 1) { const double t = S / (hm * eta); eu = eu - t; ev = ev - t; }
            a.S[p] = S;
            a.E[p] = make_double2(eu, ev);
          }
        }
      }
      if (STORE_GRAD) continue;
      for (int o2 = 16; o2 > 0; o2 >>= 1) { s_dn += __shfl_xor_sync(0xffffffffu, s_dn, o2); s_n += __shfl_xor_sync(0xffffffffu, s_n, o2); }
      if (lane == 0)
        for (int q = 0; q < a.cm.P; q++) { a.cm.partials[q][2 * s] = s_dn; a.cm.partials[q][2 * s + 1] = s_n; }
    }
  }
}
// fixed-shape tree over the per-block partials: deterministic for a given grid size
#define SUM_BLOCKS 64
__global__ void __launch_bounds__(256) k_sum_partials(int n, const double *partials, double *out2, unsigned long long *sctl, double RN_tol,
                                                      double *scratch, unsigned *ticket)
{
  // two-level tree of FIXED shape (depends on n only): SUM_BLOCKS CTAs reduce equal contiguous chunks, the last CTA to
  // finish adds the SUM_BLOCKS block sums in index order -> same bits for any arrival order, any rank, any GPU count
  __shared__ double sh[2][256];
  __shared__ bool last;
  if (sctl && sctl[SCTL_STOP]) return;
  const int len = (n + SUM_BLOCKS - 1) / SUM_BLOCKS, k0 = blockIdx.x * len, k1 = min(n, k0 + len);
  double t0 = 0.0, t1 = 0.0;
  for (int k = k0 + threadIdx.x; k < k1; k += 256) { const double2 v = ((const double2 *)partials)[k]; t0 += v.x; t1 += v.y; }
  sh[0][threadIdx.x] = t0; sh[1][threadIdx.x] = t1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    scratch[2 * blockIdx.x] = sh[0][0]; scratch[2 * blockIdx.x + 1] = sh[1][0];
    __threadfence();
    last = atomicAdd(ticket, 1u) == SUM_BLOCKS - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double s0 = 0.0, s1 = 0.0;
  for (int b = 0; b < SUM_BLOCKS; b++) { s0 += __ldcg(scratch + 2 * b); s1 += __ldcg(scratch + 2 * b + 1); }
  *ticket = 0u;
  out2[0] = s0; out2[1] = s1;
  if (sctl) {   // viscosity_iteration_i += 1 ; RN = SQRT(sum_DN_sq / sum_N_sq) ; IF (RN < C%SSA_RN_tol) EXIT  (:503-524)
    const double RN = sqrt(s0 / s1);
    sctl[SCTL_NOUTER] += 1ull;
    sctl[SCTL_RN] = (unsigned long long)__double_as_longlong(RN);
    if (RN < RN_tol) sctl[SCTL_STOP] = 1ull;
  }
}

// ---------------------------------------------------------------------------------------------
// SSA_sliding_term (:727-779) fused with the RHS / centre coefficients of solve_SSA_linearised (:581-596)
// ---------------------------------------------------------------------------------------------
struct SetupArgs {
  int Mp, P, rank;
  const unsigned char *sowner;
  const unsigned char *deg, *mflag;
  const double2 *UV, *rhsnum;
  const double *tau_c, *eta, *Hm, *cU0, *cV0;
  double thr;  // u_threshold**q_plastic, host libm
  double *S;
  double2 *RHS, *E;
};
__global__ void k_ssa_setup(SetupArgs a)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.Mp || a.deg[p] == UFM_DEG_PAD) return;
  if (a.P > 1 && a.sowner[p >> 5] != a.rank) return;
  const double delta_v = 1E-3, q_plastic = 0.30;
  const double2 u = a.UV[p];
  const double eta = a.eta[p];
  double S = a.tau_c[p] * (ufm_pow(delta_v * delta_v + u.x * u.x + u.y * u.y, 0.5 * (q_plastic - 1.0))) / a.thr;
  const double2 r = a.rhsnum[p];
  a.RHS[p] = make_double2(r.x / eta, r.y / eta);
  double eu = a.cU0[p], ev = a.cV0[p];
  if (a.mflag[p] & 1) {
    double t = S / (a.Hm[p] * eta);
    eu = eu - t; ev = ev - t;
  }
  a.S[p] = S;
  a.E[p] = make_double2(eu, ev);
}

// ---------------------------------------------------------------------------------------------
// The SOR loop of solve_SSA_linearised (:598-692): persistent cooperative kernel.
// ctrl[0..2]  rotating max-residual slots (bit pattern of a non-negative double; integer order = fp order)
// ctrl[8]     iterations executed     ctrl[9] bit0 did_reset, bit1 warning (hit max_inner)
// ctrl[10]    last max residual (bits)
// ---------------------------------------------------------------------------------------------
#ifndef SOR_BLOCK
#define SOR_BLOCK 1024
#endif
#ifndef UFM_VISC_MINB_DEFAULT
#define UFM_VISC_MINB_DEFAULT 4
#endif
#ifndef SOR_MIN_BLOCKS
#define SOR_MIN_BLOCKS 1
#endif
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct SorArgs {
  CommDev cm;
  const long long *off;
  const unsigned char *deg, *mflag, *xmask;
  const int *idx;
  const double *cU, *cV, *nxy, *nxy0, *nxysum;
  const double2 *E, *RHS;
  double2 *UV;
  const int *rng;   // slice range of block b, rank r at rng[(b*P + r)*2 + {0,1}]   (device memory: indexed dynamically)
  int bc_begin, bc_end;      // this rank's Neumann rows
  int corner_mask;           // bit k: corner k is owned by this rank
  const int *bc_pos, *bc_ptr, *bc_nbr;
  const int *corner;  // [8] corner_pos[4], corner_n[4]
  const int *corner_nbr, *corner_row;
  int Mp;
  int max_inner, force_iters;
  int chunk;      // 1: equal shares of a colour's slices per active warp (see colour_share); 0: plain grid-wide round robin
  int fuse_bc;    // 1 (single-GPU): the Neumann pass runs inside the fifth colour phase (see k_ssa_sor); 0: own phase
  int adj_end;    // fused Neumann pass: slices [begin of colour 5, adj_end) hold the colour-5 rows that the Neumann rows read
  unsigned long long *trace;   // tuning aid (env UFM_SOR_TRACE): per CTA and phase {start, first warp done, last warp done, barrier left} in ns
  int bar_rel;    // 1 (single GPU): grid barrier = release reduction + acquire poll against a locally tracked phase bit
  double omega, tol;
  unsigned long long *ctrl;
  unsigned long long *sctl;   // device-side solve control, or NULL when driven call by call
  // dataflow sweep (k_ssa_sor_df, single GPU): stage (c, k) = round k of colour c; st_K[c] rounds of st_act[c] slices each
  unsigned *stage_cnt;            // [n_stages * DF_CNT_STRIDE] CTAs that have passed the stage, monotone over the iterations of a launch
  const unsigned short *need;     // [n_slices] stage that must be complete before the slice gathers its neighbours (DF_NONE: none)
  const unsigned short *need_bc;  // [n_bc + 4] the same for the Neumann rows
  int n_stages, st_base[5], st_K[5], st_act[5];
};
#define DF_SPIN_LIMIT 4000000

__device__ __forceinline__ double2 bc_mean(const SorArgs &a, int row)
{
  double su = 0.0, sv = 0.0;
  const int b = a.bc_ptr[row], e = a.bc_ptr[row + 1];
  for (int k = b; k < e; k++) { const double2 q = __ldcg(a.UV + a.bc_nbr[k]); su = su + q.x; sv = sv + q.y; }
  const double nv = (double)(e - b);
  return make_double2(su / nv, sv / nv);
}

// store a freshly updated row and, in a partitioned run, push it over NVLink into the (U,V) array of every rank
// that reads it (same row index on every rank: the layout is replicated, only the work is partitioned)
template <bool MULTI>
__device__ __forceinline__ void store_row(const SorArgs &a, const int p, const double2 v)
{
  a.UV[p] = v;
  if (MULTI) {
    const unsigned xm = a.xmask[p];
    if (xm) {
      for (int q = 0; q < a.cm.P; q++) if ((xm >> q) & 1u) a.cm.uv[q][p] = v;   // made visible by the barrier's system fence
    }
  }
}

// Neumann row r of this rank: an edge row (mean of its non-edge neighbours) or, for r >= bc_end, one of the four corners
// (mean of all neighbours; edge neighbours at their NEW value, recomputed here from non-edge rows only)
template <bool MULTI>
__device__ __forceinline__ void neumann_row(const SorArgs &a, const int r)
{
  if (r < a.bc_end) { store_row<MULTI>(a, a.bc_pos[r], bc_mean(a, r)); return; }
  const int k = r - a.bc_end;
  if (!((a.corner_mask >> k) & 1)) return;
  const int n = a.corner[4 + k];
  double su = 0.0, sv = 0.0;
  for (int q = 0; q < n; q++) {
    const int row = a.corner_row[k * 16 + q];
    const double2 v = row >= 0 ? bc_mean(a, row) : __ldcg(a.UV + a.corner_nbr[k * 16 + q]);
    su = su + v.x; sv = sv + v.y;
  }
  store_row<MULTI>(a, a.corner[k], make_double2(su / (double)n, sv / (double)n));
}

// One SOR update of row p (ice_dynamics_module.f90:633-659).  W = slice width (compile time, so every index,
// coefficient and neighbour load of the row is issued before the first use: ~4W independent loads in flight per
// thread, which is what makes the sweep bandwidth- rather than latency-bound); n = row degree (<= W; the
// padding entries of a mixed-degree slice point at the home row with zero coefficients and are not accumulated).
template <int W, bool EXACT>
__device__ __forceinline__ double2 sor_row_value(const SorArgs &a, const long long o, const int lane, const int p, const int n, double &tmax)
{
  int j[W];
  double cu[W], cv[W], nx[W];
#pragma unroll
  for (int c = 0; c < W; c++) {
    const long long e = o + (long long)c * 32 + lane;
    j[c] = ld_stream(a.idx + e);
    cu[c] = ld_stream(a.cU + e);
    cv[c] = ld_stream(a.cV + e);
    if (EXACT) nx[c] = ld_stream(a.nxy + e);
  }
  const double2 u = a.UV[p];
  const double2 e2 = ld_stream(a.E + p), r2 = ld_stream(a.RHS + p);
  const double h = EXACT ? ld_stream(a.nxy0 + p) : ld_stream(a.nxysum + p);
  double2 nb[W];
#pragma unroll
  for (int c = 0; c < W; c++) nb[c] = a.UV[j[c]];
  double Uxy = u.x * h, Vxy = u.y * h;   // get_mesh_curvatures_vertex_AaAc as coded: home value times every coefficient
  if (EXACT) {
#pragma unroll
    for (int c = 0; c < W; c++) if (c < n) { Uxy = Uxy + u.x * nx[c]; Vxy = Vxy + u.y * nx[c]; }
  }
  double sumU = 0.0, sumV = 0.0;
#pragma unroll
  for (int c = 0; c < W; c++) if (c < n) { sumU = sumU + nb[c].x * cu[c]; sumV = sumV + nb[c].y * cv[c]; }
  const double LHSx = sumU + (3.0 * Vxy) + (e2.x * u.x);
  const double LHSy = sumV + (3.0 * Uxy) + (e2.y * u.y);
  const double resU = (LHSx - r2.x) / e2.x;
  const double resV = (LHSy - r2.y) / e2.y;
  tmax = fmax(tmax, fabs(resU));
  tmax = fmax(tmax, fabs(resV));
  return make_double2(u.x - a.omega * resU, u.y - a.omega * resV);
}
template <int W, bool EXACT, bool MULTI>
__device__ __forceinline__ double sor_row(const SorArgs &a, const long long o, const int lane, const int p, const int n, double tmax)
{
  store_row<MULTI>(a, p, sor_row_value<W, EXACT>(a, o, lane, p, n, tmax));
  return tmax;
}

// The slices [first, end) step `step` of colour c that the calling warp sweeps.  Single-GPU runs give every CTA one
// contiguous chunk of the colour block (equal to within one slice per SM, and spatially compact because rows are in
// Morton order inside a degree class, so neighbour gathers of one SM share L1 lines); its warps take the chunk round
// robin.  Partitioned runs keep the grid-wide round robin: their boundary slices sit at the end of the range and must be
// reached last by every warp.
// all CTAs co-resident (cooperative launch); `phase` is the caller's copy of bit 31 of the barrier word
__device__ __forceinline__ void grid_barrier_rel(unsigned *bar, const int nblocks, unsigned &phase)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned add = blockIdx.x == 0 ? 0x80000000u - (unsigned)(nblocks - 1) : 1u;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(add) : "memory");
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (((v ^ phase) & 0x80000000u) == 0u);
  }
  phase ^= 0x80000000u;
  __syncthreads();
}

template <bool MULTI>
__device__ __forceinline__ void colour_share(const SorArgs &a, const int *s_rng, const int c, const int wg, const int nw, int &first, int &end, int &step)
{
  const int b = s_rng[3 * c], e = s_rng[3 * c + 2];
  if (!MULTI && a.chunk) {
    // equal shares: k = ceil(slices / warps) slices for each of ceil(slices / k) warps, the other warps sit the phase out.
    // With the plain round robin most warps finish after floor(slices / warps) rounds and the last round runs with a
    // fraction of the loads in flight (measured: 5 us of a 37 us phase); here every active warp streams until the end.
    // The active warps are spread evenly over the CTAs and keep the CTA-contiguous numbering (L1 reuse of the gathers).
    const int n = e - b, k = (n + nw - 1) / nw;
    const int act = k > 0 ? (n + k - 1) / k : 0;
    const int lo = (int)(((long long)act * blockIdx.x) / gridDim.x), hi = (int)(((long long)act * (blockIdx.x + 1)) / gridDim.x);
    const int wib = (int)(threadIdx.x >> 5);
    first = wib < hi - lo ? b + lo + wib : e; end = e; step = act;
  } else {
    first = b + wg; end = e; step = nw;
  }
}

// ctrl[0..2]  rotating max-residual slots (bit pattern of a non-negative double; integer order = fp order)
// ctrl[8]     iterations executed     ctrl[9] bit0 did_reset, bit1 warning (hit max_inner), bit2 peer wait timed out
// ctrl[10]    last max residual (bits)   ctrl[32] grid barrier {count, generation}
template <bool EXACT, bool GLFIX, bool MULTI>
__global__ void __launch_bounds__(SOR_BLOCK, SOR_MIN_BLOCKS) k_ssa_sor(SorArgs a)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  const int nblocks = gridDim.x, P = MULTI ? a.cm.P : 1, rank = MULTI ? a.cm.rank : 0;
  unsigned *bar = (unsigned *)(a.ctrl + 32);
  volatile unsigned long long *mail = MULTI ? a.cm.mail[rank] : nullptr;
  __shared__ double sh[SOR_BLOCK / 32];
  __shared__ int s_rng[15];   // this rank's slice ranges [begin, boundary_begin, end) of the five colours (read once)
  __shared__ unsigned long long s_tr[2];
  if (a.sctl && a.sctl[SCTL_STOP]) return;   // uniform over the grid and over the ranks: nobody enters a barrier
  if (threadIdx.x < 15) s_rng[threadIdx.x] = a.rng[((threadIdx.x / 3) * P + rank) * 3 + (threadIdx.x % 3)];
  unsigned phase = *((volatile unsigned *)bar) & 0x80000000u;   // bit 31 cannot flip before this CTA's first arrival
  __syncthreads();
  int it = 0;
  bool done = false;
  unsigned flags = 0;
  double maxres = 0.0;
  const bool fused = !MULTI && a.fuse_bc;
  unsigned long long epoch = MULTI ? mail[MAIL_EPOCH] : 0ull;   // same on every rank: all ranks run the same barriers
  while (!done && it < a.max_inner) {
    it++;
    if (tid == 0) a.ctrl[(it + 1) % 3] = 0ull;
    double tmax = 0.0;
    for (int c = 0; c < 5; c++) {
      const int s_bnd = s_rng[3 * c + 1];
      int s_first, s_end, s_step;
      colour_share<MULTI>(a, s_rng, c, wg, nw, s_first, s_end, s_step);
      bool waited = !MULTI;
      const bool tr = a.trace && it == 4;
      if (tr && threadIdx.x == 0) {
        s_tr[0] = ~0ull; s_tr[1] = 0ull; a.trace[(blockIdx.x * 6 + c) * 4 + 0] = gtime();
        if (c == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); a.trace[(blockIdx.x * 6 + 5) * 4 + 0] = sm; }   // which SM this CTA sits on
      }
      if (tr) __syncthreads();
      for (int s = s_first; s < s_end; s += s_step) {
        if (MULTI && !waited && s >= s_bnd) {  // first boundary slice of this warp: the peers' previous phase must have landed
          // the previous epoch was a colour sweep (signalled to the neighbours) or, at c == 0, the Neumann pass (signalled to all)
          if (lane == 0) wait_peers(a.cm, epoch, a.cm.nbr);
          __syncwarp();
          waited = true;
        }
        const long long o = a.off[s];
        const int w = (int)((a.off[s + 1] - o) >> 5);
        const int p = s * 32 + lane;
        const int n = a.deg[p];
        bool act = n != UFM_DEG_PAD;
        if (GLFIX) { if (a.mflag[p] & 2) act = false; }
        if (act) switch (w) {  // warp-uniform
          case 3: tmax = sor_row<3, EXACT, MULTI>(a, o, lane, p, n, tmax); break;
          case 4: tmax = sor_row<4, EXACT, MULTI>(a, o, lane, p, n, tmax); break;
          case 5: tmax = sor_row<5, EXACT, MULTI>(a, o, lane, p, n, tmax); break;
          case 6: tmax = sor_row<6, EXACT, MULTI>(a, o, lane, p, n, tmax); break;
          case 7: tmax = sor_row<7, EXACT, MULTI>(a, o, lane, p, n, tmax); break;
          case 8: tmax = sor_row<8, EXACT, MULTI>(a, o, lane, p, n, tmax); break;
          default: {  // rare high-degree rows: generic path with the same accumulation order
            const double2 u = a.UV[p];
            const double h = EXACT ? a.nxy0[p] : a.nxysum[p];
            double sumU = 0.0, sumV = 0.0, Uxy = u.x * h, Vxy = u.y * h;
            for (int cc = 0; cc < w; cc++) {
              if (cc < n) {
                const long long e = o + (long long)cc * 32 + lane;
                const double2 nbv = a.UV[a.idx[e]];
                sumU = sumU + nbv.x * a.cU[e];
                sumV = sumV + nbv.y * a.cV[e];
                if (EXACT) { const double t = a.nxy[e]; Uxy = Uxy + u.x * t; Vxy = Vxy + u.y * t; }
              }
            }
            const double2 e2 = a.E[p], r2 = a.RHS[p];
            const double LHSx = sumU + (3.0 * Vxy) + (e2.x * u.x);
            const double LHSy = sumV + (3.0 * Uxy) + (e2.y * u.y);
            const double resU = (LHSx - r2.x) / e2.x;
            const double resV = (LHSy - r2.y) / e2.y;
            tmax = fmax(tmax, fabs(resU));
            tmax = fmax(tmax, fabs(resV));
            store_row<MULTI>(a, p, make_double2(u.x - a.omega * resU, u.y - a.omega * resV));
          }
        }
        if (!MULTI && fused && c == 4 && s < a.adj_end) {  // a slice the Neumann rows read: count it once its rows are visible
          __syncwarp();
          __threadfence();
          if (lane == 0) atomicAdd(a.ctrl + 12, 1ull);
        }
      }
      if (tr && lane == 0) { const unsigned long long t = gtime(); atomicMin(&s_tr[0], t); atomicMax(&s_tr[1], t); }
      if (!MULTI && fused && c == 4) {
        // apply_Neumann_boundary_AaAc inside the fifth colour phase: the colour-5 rows next to the domain edge are the first
        // slices of the colour block (upload order) and were counted above; every other row the pass reads belongs to
        // colours 1-4 and is final.  The wait is on work that was started first in this phase and never waits itself.
        const unsigned long long target = (unsigned long long)it * (unsigned long long)(a.adj_end - s_rng[12]);
        bool ready = false;
        // who takes the Neumann rows: with equal shares the warps that have no slice in this phase (they get here at once, so the pass
        // is over long before the sweep is; taken by the first CTAs' threads after their own slices it was the tail of the iteration:
        // 7 us between the last slice and the barrier, profiles/r02a_trace_bands64.json), otherwise everybody
        int r0 = a.bc_begin + tid, rs = nt;
        if (a.chunk) {
          const int n5 = s_rng[14] - s_rng[12], k5 = (n5 + nw - 1) / nw, act5 = k5 > 0 ? (n5 + k5 - 1) / k5 : 0;
          if ((nw - act5) * 32 >= a.bc_end + 4 - a.bc_begin) {
            const int lo5 = (int)(((long long)act5 * blockIdx.x) / gridDim.x), hi5 = (int)(((long long)act5 * (blockIdx.x + 1)) / gridDim.x);
            const int wib = (int)(threadIdx.x >> 5), nwb = (int)(blockDim.x >> 5);
            r0 = wib >= hi5 - lo5 ? a.bc_begin + (nwb * (int)blockIdx.x - lo5 + (wib - (hi5 - lo5))) * 32 + lane : a.bc_end + 4;
            rs = (nw - act5) * 32;
          }
        }
        for (int r = r0; r < a.bc_end + 4; r += rs) {
          if (!ready) { while (*((volatile unsigned long long *)(a.ctrl + 12)) < target) { } __threadfence(); ready = true; }
          neumann_row<MULTI>(a, r);
        }
      }
      if (c == 4) {  // publish this CTA's max residual before the barrier that precedes the stop test
        for (int o = 16; o > 0; o >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        if (lane == 0) sh[threadIdx.x >> 5] = tmax;
        __syncthreads();
        if (threadIdx.x == 0) {
          double m = 0.0;
          for (int k = 0; k < (int)(blockDim.x >> 5); k++) m = fmax(m, sh[k]);
          atomicMax(a.ctrl + (it % 3), (unsigned long long)__double_as_longlong(m));
        }
      }
      // end of colour phase: grid barrier; in a partitioned run the last CTA exchanges epochs with the peers
      // (and, after the fifth colour, this rank's max residual: the MPI_ALLREDUCE MAX of :673)
      ++epoch;
      if (!MULTI && a.bar_rel) grid_barrier_rel(bar, nblocks, phase);
      else grid_barrier<MULTI>(bar, nblocks, a.cm, epoch, c == 4 ? 0xffffffffu : a.cm.nbr, [&]() {
        if (MULTI && c == 4) {
          const unsigned long long r = *((volatile unsigned long long *)(a.ctrl + (it % 3)));
          for (int q = 0; q < P; q++) *((volatile unsigned long long *)(a.cm.mail[q] + MAIL_RESID + (it % 3) * UFM_MAX_RANKS + rank)) = r;
        }
      });
      if (tr && threadIdx.x == 0) { unsigned long long *q = a.trace + (blockIdx.x * 6 + c) * 4; q[1] = s_tr[0]; q[2] = s_tr[1]; q[3] = gtime(); }
    }
    // apply_Neumann_boundary_AaAc on U and V (mesh_ArakawaC_module.f90:660-724): edge rows from their
    // non-edge neighbours; the four corners from all neighbours, edge neighbours taken at their NEW value
    // (recomputed here from non-edge rows only, so the whole pass is one hazard-free phase).
    if (MULTI || !fused) {
      if (MULTI) {  // the Neumann rows may read peer-owned rows of the fifth colour
        if (threadIdx.x == 0) wait_peers(a.cm, epoch, 0xffffffffu);   // fifth-colour epoch: every rank's residual has arrived with it
        __syncthreads();
      }
      for (int r = a.bc_begin + tid; r < a.bc_end + 4; r += nt) neumann_row<MULTI>(a, r);
      ++epoch;
      if (!MULTI && a.bar_rel) grid_barrier_rel(bar, nblocks, phase);
      else grid_barrier<MULTI>(bar, nblocks, a.cm, epoch, 0xffffffffu, [&]() {});
    }
    if (MULTI) {
      // every rank's residual was published before it signalled the fifth-colour epoch, which wait_peers() above has seen
      unsigned long long r = 0ull;
      for (int q = 0; q < P; q++) { const unsigned long long v = mail[MAIL_RESID + (it % 3) * UFM_MAX_RANKS + q]; r = v > r ? v : r; }
      maxres = __longlong_as_double((long long)r);
      if (mail[MAIL_ABORT]) { flags |= 4; done = true; }
    } else {
      maxres = __longlong_as_double((long long)*((volatile unsigned long long *)(a.ctrl + (it % 3))));
    }
    if (!a.force_iters && !done) {
      if (maxres < a.tol) done = true;
      else if (maxres > 1E6) {
        for (int p = tid; p < a.Mp; p += nt) a.UV[p] = make_double2(0.0, 0.0);
        flags |= 1; done = true;
      } else if (it == a.max_inner) flags |= 2;
    }
  }
  if (MULTI) {  // leave only when every peer push of this solve has landed in our (U,V)
    if (threadIdx.x == 0) wait_peers(a.cm, epoch, 0xffffffffu);
    __syncthreads();
    if (tid == 0) mail[MAIL_EPOCH] = epoch;
  }
  if (tid == 0) {
    a.ctrl[8] = (unsigned long long)it; a.ctrl[9] = flags; a.ctrl[10] = (unsigned long long)__double_as_longlong(maxres);
    if (a.sctl) {   // bookkeeping of solve_SSA's outer loop (:530-540)
      a.sctl[SCTL_NINNER] += (unsigned long long)it; a.sctl[SCTL_NLAST] = (unsigned long long)it;
      a.sctl[SCTL_MAXRES] = (unsigned long long)__double_as_longlong(maxres);
      if (flags & 2) a.sctl[SCTL_RC] |= 1ull;
      if (flags & 4) { a.sctl[SCTL_RC] |= 4ull; a.sctl[SCTL_STOP] = 1ull; }
      if (flags & 1) {
        if (a.sctl[SCTL_RESET]) { a.sctl[SCTL_RC] |= 2ull; a.sctl[SCTL_STOP] = 1ull; }
        else a.sctl[SCTL_RESET] = 1ull;
      }
    }
  }
}

// =============================================================================================
// Dataflow SOR sweep (single GPU): the same row updates in an order that respects every read-after-write dependency of the
// colour-by-colour sweep, WITHOUT a grid barrier between the colours (one grid barrier per iteration remains: the stop test of
// ice_dynamics_module.f90:676-689 needs the global max residual before the next iteration may touch U).
//
// The result of a sweep does not depend on the global phase order, only on every row seeing its lower-coloured neighbours already
// updated and its higher-coloured ones not yet (tests/test_oracle.py::test_sor_sweep_is_a_dataflow_not_a_phase_order), so:
//   * a colour block of n slices is swept in K = ceil(n / warps) rounds of act = ceil(n / K) slices; round k of colour c is
//     "stage" st_base[c] + k.  Every warp walks through all stages in order (it has at most one slice per stage) and reports each
//     stage it has passed; per stage one shared-memory counter per CTA, and the last warp of a CTA adds 1 to the stage's global
//     counter with a gpu-scope release.  Stage g is complete when its counter has reached (iteration x CTAs); completeness is
//     monotone in g because every warp passes the stages in order.
//   * need[s] (k_sor_need, once per mesh) is the highest stage holding a lower-coloured neighbour of a row of slice s.  A warp issues
//     the read-only streams of its slice, then (only if that stage is not yet known to be complete) polls the counter, then gathers.
//     With the x-band row order (ufm_row_order_impl) need[s] lies about one round ahead of the slice's own position in the previous
//     colour, i.e. almost a whole colour phase in the past: nobody waits, and a warp that is done with colour c early simply goes on
//     with colour c+1, so the ramp-down of one colour overlaps the ramp-up of the next.
//   * the Neumann rows (apply_Neumann_boundary_AaAc) are "stage-less" readers with their own need_bc[r]; they are taken by the warps
//     that have no slice in the fifth colour (or by everybody when there are none), their index lists fetched before the wait.
// Deadlock-free: waits only ever target lower stages, all CTAs are co-resident (cooperative launch), and every poll loop gives up
// after DF_SPIN_LIMIT polls (flag bit 3 -> rc -9) instead of hanging the GPU.
// =============================================================================================
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
#ifdef DF_EXP_NO_ACQUIRE   // timing experiment only (NOT correct): no L1 invalidation
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) { return ld_relaxed_u32(p); }
#else
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
#endif

struct DfState {
  int it;                // SOR iteration of this launch, 1-based: a stage is complete when its counter has reached it x CTAs
  int known;             // highest stage this warp knows to be complete in this iteration (-1: none)
};
// make sure stage `nd` is complete (nd < 0: nothing to wait for).  Must be called by all 32 lanes with the same nd.
// Three levels: the warp's own `known`; the CTA's (*s_known = it << 16 | known + 1, so that it never needs a reset: a value from
// another iteration reads as "nothing known"); and the slow path, a real call (kept out of line for the registers): poll, then one
// acquire that also looks up to 32 stages ahead.  An acquire drops the SM's whole L1 (CCTL.IVALL) -- the reason for sharing what it
// learnt with the other 31 warps of the CTA: any L1 line they fetch from then on is at least that fresh.
__device__ __noinline__ int df_wait_slow(const unsigned *stage_cnt, unsigned long long *ctrl, const unsigned target, const int nd, const int n_stages)
{
  const int lane = threadIdx.x & 31;
  if (lane == 0) {
    int spins = 0;
    while (ld_relaxed_u32(stage_cnt + nd * DF_CNT_STRIDE) < target) {
      ++spins;
#ifdef DF_POLL_SLEEP_NS
      __nanosleep(DF_POLL_SLEEP_NS);   // back off: thousands of warps polling one counter would queue in front of the updates of that counter
#endif
      if ((spins & 1023) == 0 && *((volatile unsigned long long *)(ctrl + 13))) break;   // somebody else has given up: so do we
      if (spins > DF_SPIN_LIMIT) { *((volatile unsigned long long *)(ctrl + 13)) = 1ull; break; }
    }
  }
  __syncwarp();
  const int gi = min(nd + lane, n_stages - 1);
  const unsigned ok = __ballot_sync(0xffffffffu, ld_acquire_u32(stage_cnt + gi * DF_CNT_STRIDE) >= target);
  __syncwarp();
  const int run = ok == 0xffffffffu ? 32 : __ffs((int)~ok) - 1;
  return min(nd + max(run, 1) - 1, n_stages - 1);
}
__device__ __forceinline__ void df_signal(unsigned *sh_cnt, unsigned *cnt, const int g, const int lane, const unsigned n_warps);
__device__ __forceinline__ void df_wait(const SorArgs &a, DfState &d, volatile unsigned *s_known, unsigned *sh_cnt, int &pend, const int nd)
{
  if (nd <= d.known) return;
  const unsigned v = *s_known;
  if ((int)(v >> 16) == d.it) d.known = max(d.known, (int)(v & 0xFFFFu) - 1);
  if (nd <= d.known) return;
  // about to poll, possibly to block: whoever waits for this warp's last stage must not be kept waiting (deadlock otherwise)
  if (pend >= 0) { df_signal(sh_cnt, a.stage_cnt, pend, threadIdx.x & 31, blockDim.x >> 5); pend = -1; }
#ifdef DF_STATS   // tuning build: how often the slow path runs, whether it had to block, how far the look-ahead got
  const bool blocked = ld_relaxed_u32(a.stage_cnt + nd * DF_CNT_STRIDE) < (unsigned)d.it * gridDim.x;
  const unsigned long long t0 = gtime();
#endif
  d.known = df_wait_slow(a.stage_cnt, a.ctrl, (unsigned)d.it * gridDim.x, nd, a.n_stages);
#ifdef DF_STATS
  if (a.trace && (threadIdx.x & 31) == 0) {
    unsigned long long *q = a.trace + blockIdx.x * 8;
    atomicAdd(q + 0, 1ull); atomicAdd(q + 1, blocked ? 1ull : 0ull); atomicAdd(q + 2, gtime() - t0); atomicAdd(q + 3, (unsigned long long)(d.known - nd + 1));
  }
#endif
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned *)s_known, ((unsigned)d.it << 16) | (unsigned)(d.known + 1));
}
// the calling warp has passed stage g.  Its stores of that stage precede the call in program order; every warp makes its own stores
// visible device-wide (release at gpu scope: a fence that finds nothing outstanding when the call is deferred, see k_ssa_sor_df) before
// it counts itself in, and the last warp of the CTA publishes the stage.
__device__ __forceinline__ void df_signal(unsigned *sh_cnt, unsigned *cnt, const int g, const int lane, const unsigned n_warps)
{
  __syncwarp();
  if (lane == 0) {
    unsigned old;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sh_cnt + g);
#ifdef DF_EXP_NO_RELEASE   // timing experiment only (NOT correct): no fence before the count
    asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(sa), "r"(1u) : "memory");
#else
    asm volatile("atom.release.gpu.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(sa), "r"(1u) : "memory");
#endif
    if (old == n_warps - 1u) {   // last warp of this CTA: the other 31 released their stores at gpu scope before they counted themselves in
      asm volatile("fence.acq_rel.cta;" ::: "memory");
      sh_cnt[g] = 0u;            // next use: next iteration, after the grid barrier
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(cnt + g * DF_CNT_STRIDE), "r"(1u) : "memory");
    }
  }
}

// stage of the (swept) row at device position q, or -1 for rows outside the five colour blocks
__device__ __forceinline__ int df_stage_of(const int *rng15, const int *st_base, const int *st_act, const int q)
{
  const int sq = q >> 5;
#pragma unroll
  for (int c = 0; c < 5; c++)
    if (sq >= rng15[3 * c] && sq < rng15[3 * c + 2]) return st_base[c] + (sq - rng15[3 * c]) / st_act[c];
  return -1;
}
struct NeedArgs {
  int rng15[15], st_base[5], st_act[5];
  int n_slices;
  const long long *off; const unsigned char *deg; const int *idx;
  int n_bc; const int *bc_ptr, *bc_nbr, *corner, *corner_nbr, *corner_row;
  unsigned short *need, *need_bc;
  unsigned long long *trace;
};
// one warp per slice (colours 2..5), then one thread per Neumann row
__global__ void k_sor_need(NeedArgs a)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < a.n_slices; s += nw) {
    int c = -1;
    for (int k = 0; k < 5; k++) if (s >= a.rng15[3 * k] && s < a.rng15[3 * k + 2]) c = k;
    int nd = -1;
    if (c >= 1) {
      const long long o = a.off[s];
      const int p = s * 32 + lane, n = a.deg[p];
      if (n != UFM_DEG_PAD)
        for (int cc = 0; cc < n; cc++) {
          const int q = a.idx[o + (long long)cc * 32 + lane];
          const int g = df_stage_of(a.rng15, a.st_base, a.st_act, q);
          if (g >= 0 && g < a.st_base[c]) nd = max(nd, g);   // lower colours only; same colour never adjacent
        }
      for (int o2 = 16; o2 > 0; o2 >>= 1) nd = max(nd, __shfl_xor_sync(0xffffffffu, nd, o2));
    }
    if (lane == 0) a.need[s] = nd < 0 ? (unsigned short)DF_NONE : (unsigned short)nd;
#ifdef DF_STATS
    if (lane == 0 && a.trace && nd >= 0) atomicAdd(a.trace + 4096 * 12 + min(63, a.st_base[c] + (s - a.rng15[3 * c]) / a.st_act[c] - nd), 1ull);   // histogram of the slack in stages
#endif
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_bc + 4; r += gridDim.x * blockDim.x) {
    int nd = -1;
    if (r < a.n_bc) {
      for (int k = a.bc_ptr[r]; k < a.bc_ptr[r + 1]; k++) nd = max(nd, df_stage_of(a.rng15, a.st_base, a.st_act, a.bc_nbr[k]));
    } else {
      const int kc = r - a.n_bc, n = a.corner[4 + kc];
      for (int q = 0; q < n; q++) {
        const int row = a.corner_row[kc * 16 + q];
        if (row >= 0) { for (int k = a.bc_ptr[row]; k < a.bc_ptr[row + 1]; k++) nd = max(nd, df_stage_of(a.rng15, a.st_base, a.st_act, a.bc_nbr[k])); }
        else nd = max(nd, df_stage_of(a.rng15, a.st_base, a.st_act, a.corner_nbr[kc * 16 + q]));
      }
    }
    a.need_bc[r] = nd < 0 ? (unsigned short)DF_NONE : (unsigned short)nd;
  }
}

template <bool EXACT>
__device__ __forceinline__ double2 sor_row_generic(const SorArgs &a, const long long o, const int w, const int lane, const int p, const int n, double &tmax)
{
  // rare high-degree rows (slice width > 8): same accumulation order, loads not batched
  const double2 u = a.UV[p];
  const double h = EXACT ? a.nxy0[p] : a.nxysum[p];
  double sumU = 0.0, sumV = 0.0, Uxy = u.x * h, Vxy = u.y * h;
  for (int cc = 0; cc < w; cc++) {
    if (cc < n) {
      const long long e = o + (long long)cc * 32 + lane;
      const double2 nbv = a.UV[a.idx[e]];
      sumU = sumU + nbv.x * a.cU[e];
      sumV = sumV + nbv.y * a.cV[e];
      if (EXACT) { const double t = a.nxy[e]; Uxy = Uxy + u.x * t; Vxy = Vxy + u.y * t; }
    }
  }
  const double2 e2 = a.E[p], r2 = a.RHS[p];
  const double LHSx = sumU + (3.0 * Vxy) + (e2.x * u.x);
  const double LHSy = sumV + (3.0 * Uxy) + (e2.y * u.y);
  const double resU = (LHSx - r2.x) / e2.x;
  const double resV = (LHSy - r2.y) / e2.y;
  tmax = fmax(tmax, fabs(resU));
  tmax = fmax(tmax, fabs(resV));
  return make_double2(u.x - a.omega * resU, u.y - a.omega * resV);
}

// 32 Neumann rows (one per lane), r0 = first row: index lists first, then the wait, then one round of gathers
__device__ __forceinline__ void df_neumann_chunk(const SorArgs &a, DfState &d, volatile unsigned *s_known, unsigned *sh_cnt, const int r0)
{
  const int r = r0 + (int)(threadIdx.x & 31), r_end = a.bc_end + 4;
  const bool mine = r < r_end, edge = r < a.bc_end;
  int b = 0, e = 0, j[4] = {0, 0, 0, 0};
  int nd = -1;
  if (mine) { const unsigned v = a.need_bc[r]; nd = v == DF_NONE ? -1 : (int)v; }
  if (mine && edge) {
    b = a.bc_ptr[r]; e = a.bc_ptr[r + 1];
#pragma unroll
    for (int k = 0; k < 4; k++) if (b + k < e) j[k] = a.bc_nbr[b + k];
  }
  int ndw = nd;
  for (int o2 = 16; o2 > 0; o2 >>= 1) ndw = max(ndw, __shfl_xor_sync(0xffffffffu, ndw, o2));
  int none = -1;   // nothing pending here: the stage loop has flushed
  df_wait(a, d, s_known, sh_cnt, none, ndw);
  if (!mine) return;
  if (edge) {
    double su = 0.0, sv = 0.0;
    double2 q[4];
#pragma unroll
    for (int k = 0; k < 4; k++) if (b + k < e) q[k] = __ldcg(a.UV + j[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) if (b + k < e) { su = su + q[k].x; sv = sv + q[k].y; }
    for (int k = b + 4; k < e; k++) { const double2 t = __ldcg(a.UV + a.bc_nbr[k]); su = su + t.x; sv = sv + t.y; }
    const double nv = (double)(e - b);
    a.UV[a.bc_pos[r]] = make_double2(su / nv, sv / nv);
  } else neumann_row<false>(a, r);
}

template <bool EXACT, bool GLFIX>
__global__ void __launch_bounds__(SOR_BLOCK, SOR_MIN_BLOCKS) k_ssa_sor_df(SorArgs a)
{
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nblocks = gridDim.x;
  const unsigned n_warps = blockDim.x >> 5;
  unsigned *bar = (unsigned *)(a.ctrl + 32);
  __shared__ double sh[SOR_BLOCK / 32];
  // per stage: {first slice of this CTA, end of the colour block, warps of this CTA with a slice in the stage, -} -- a flat loop over
  // the stages that looks its bounds up here keeps the registers for the loads in flight
  extern __shared__ int4 s_stage[];
  unsigned *sh_cnt = (unsigned *)(s_stage + a.n_stages);
  __shared__ unsigned s_known;   // see df_wait
  if (a.sctl && a.sctl[SCTL_STOP]) return;   // uniform over the grid: nobody enters a barrier
  for (int g = threadIdx.x; g < a.n_stages; g += blockDim.x) {
    int c = 0;
    while (c < 4 && g >= a.st_base[c + 1]) c++;
    const int k = g - a.st_base[c], act = a.st_act[c];
    const int lo = (int)(((long long)act * blockIdx.x) / nblocks), hi = (int)(((long long)act * (blockIdx.x + 1)) / nblocks);
    s_stage[g] = make_int4(a.rng[3 * c] + k * act + lo, a.rng[3 * c + 2], hi - lo, 0);
    sh_cnt[g] = 0u;
  }
  if (threadIdx.x == 0) s_known = 0u;
  unsigned phase = *((volatile unsigned *)bar) & 0x80000000u;
  __syncthreads();
  bool done = false;
  unsigned flags = 0;
  DfState d;
  d.it = 0;
  // Neumann work: chunks of 32 rows for the warps without a slice in the fifth colour, if there are enough of them
  const int bc_chunks = (a.bc_end + 4 - a.bc_begin + 31) >> 5;
  while (!done && d.it < a.max_inner) {
    d.it++;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctrl[(d.it + 1) % 3] = 0ull;
    if (lane == 0) sh[wib] = 0.0;   // this warp's max residual (kept here, not in a register, across the slices)
    d.known = -1;
    // Signals are deferred: the stage of the slice just stored is reported in the middle of the NEXT slice, after that slice's loads
    // have been consumed and before its store -- the release fence then finds this warp's earlier stores long acknowledged and no load
    // outstanding, i.e. costs nothing; reported right after the store it would stall the warp for an L2 round trip per slice.
    int pend = -1;
#pragma unroll 1
    for (int g = 0; g < a.n_stages; g++) {
      const volatile int *tg = (const volatile int *)(s_stage + g);   // volatile: looked up per stage, not kept in registers
      const int s = tg[0] + wib;
      if (wib < tg[2] && s < tg[1]) {
        const long long o = a.off[s];
        const int w = (int)((a.off[s + 1] - o) >> 5);
        const int p = s * 32 + lane;
        const int n = a.deg[p];
        const unsigned ndu = a.need[s];   // fetched together with the slice header: no extra latency
        bool on = n != UFM_DEG_PAD;
        if (GLFIX) { if (a.mflag[p] & 2) on = false; }
        // the rows this slice reads must have been updated (warp-wide; almost always known already, see df_wait)
        df_wait(a, d, &s_known, sh_cnt, pend, ndu == DF_NONE ? -1 : (int)ndu);
        double tmax = 0.0;
        double2 nv = make_double2(0.0, 0.0);
        if (on) switch (w) {   // w is warp-uniform
          case 3: nv = sor_row_value<3, EXACT>(a, o, lane, p, n, tmax); break;
          case 4: nv = sor_row_value<4, EXACT>(a, o, lane, p, n, tmax); break;
          case 5: nv = sor_row_value<5, EXACT>(a, o, lane, p, n, tmax); break;
          case 6: nv = sor_row_value<6, EXACT>(a, o, lane, p, n, tmax); break;
          case 7: nv = sor_row_value<7, EXACT>(a, o, lane, p, n, tmax); break;
          case 8: nv = sor_row_value<8, EXACT>(a, o, lane, p, n, tmax); break;
          default: nv = sor_row_generic<EXACT>(a, o, w, lane, p, n, tmax);
        }
#ifndef DF_EXP_NO_REDUCE
        for (int o2 = 16; o2 > 0; o2 >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o2));   // (needs every load of the slice)
        if (lane == 0) { volatile double *q = sh + wib; if (tmax > *q) *q = tmax; }
#endif
        if (pend >= 0) df_signal(sh_cnt, a.stage_cnt, pend, lane, n_warps);
        if (on) a.UV[p] = nv;
        pend = g;
      } else {
        if (pend >= 0) df_signal(sh_cnt, a.stage_cnt, pend, lane, n_warps);
        df_signal(sh_cnt, a.stage_cnt, g, lane, n_warps);
        pend = -1;
      }
    }
    if (pend >= 0) df_signal(sh_cnt, a.stage_cnt, pend, lane, n_warps);
    // apply_Neumann_boundary_AaAc (mesh_ArakawaC_module.f90:660-724)
    {
      const int nw = nblocks * (int)n_warps;
      const int act5 = a.st_act[4];
      if (nw - act5 >= bc_chunks) {
        const int lo5 = (int)(((long long)act5 * blockIdx.x) / nblocks), hi5 = (int)(((long long)act5 * (blockIdx.x + 1)) / nblocks);
        if (wib >= hi5 - lo5) {
          const int i = (int)n_warps * (int)blockIdx.x - lo5 + (wib - (hi5 - lo5));   // index of this warp among the warps without colour-5 work
          if (i < bc_chunks) df_neumann_chunk(a, d, &s_known, sh_cnt, a.bc_begin + 32 * i);
        }
      } else {
        for (int i = blockIdx.x * (int)n_warps + wib; i < bc_chunks; i += nw) df_neumann_chunk(a, d, &s_known, sh_cnt, a.bc_begin + 32 * i);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = 0.0;
      for (int k = 0; k < (int)n_warps; k++) m = fmax(m, sh[k]);
      atomicMax(a.ctrl + (d.it % 3), (unsigned long long)__double_as_longlong(m));
    }
    grid_barrier_rel(bar, nblocks, phase);
    const double maxres = __longlong_as_double((long long)*((volatile unsigned long long *)(a.ctrl + (d.it % 3))));
    if (*((volatile unsigned long long *)(a.ctrl + 13))) { flags |= 8; done = true; }
    if (!a.force_iters && !done) {
      if (maxres < a.tol) done = true;
      else if (maxres > 1E6) {
        const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
        for (int p = tid; p < a.Mp; p += nt) a.UV[p] = make_double2(0.0, 0.0);
        flags |= 1; done = true;
      } else if (d.it == a.max_inner) flags |= 2;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long mr = d.it > 0 ? *((volatile unsigned long long *)(a.ctrl + (d.it % 3))) : 0ull;   // max residual of the last iteration
    a.ctrl[8] = (unsigned long long)d.it; a.ctrl[9] = flags; a.ctrl[10] = mr;
    if (a.sctl) {
      a.sctl[SCTL_NINNER] += (unsigned long long)d.it; a.sctl[SCTL_NLAST] = (unsigned long long)d.it;
      a.sctl[SCTL_MAXRES] = mr;
      if (flags & 2) a.sctl[SCTL_RC] |= 1ull;
      if (flags & 8) { a.sctl[SCTL_RC] |= 8ull; a.sctl[SCTL_STOP] = 1ull; }
      if (flags & 1) {
        if (a.sctl[SCTL_RESET]) { a.sctl[SCTL_RC] |= 2ull; a.sctl[SCTL_STOP] = 1ull; }
        else a.sctl[SCTL_RESET] = 1ull;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Halo exchange of the partitioned per-step kernels.  Storage is replicated, so an element has the same index on every
// GPU; what a rank computes for its own elements and a neighbour strip reads travels as: pack into the READER's exchange buffer
// (NVLink P2P stores; one region per parity and sender) -> peer barrier -> unpack from the own buffer into the own arrays.
// Two parities: a rank may pack exchange k+1 while a slower neighbour still unpacks exchange k.
// ---------------------------------------------------------------------------------------------
struct XArgs {
  CommDev cm;
  int parity, region, narr;
  const int *idx;
  int ptr[UFM_MAX_RANKS + 1];
  double *arr[4];
};
__global__ void __launch_bounds__(256) k_xpack(XArgs a)
{
  const int P = a.cm.P, me = a.cm.rank;
  for (int q = 0; q < P; q++) {
    const int b = a.ptr[q], n = a.ptr[q + 1] - b;
    if (q == me || n == 0) continue;
    double *dst = a.cm.xbuf[q] + ((size_t)a.parity * P + me) * a.region;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * a.narr; i += gridDim.x * blockDim.x) {
      const int k = i / n, j = i - k * n;
      dst[(size_t)k * n + j] = a.arr[k][a.idx[b + j]];
    }
  }
  __threadfence_system();
}
__global__ void __launch_bounds__(256) k_xunpack(XArgs a)
{
  const int P = a.cm.P, me = a.cm.rank;
  for (int q = 0; q < P; q++) {
    const int b = a.ptr[q], n = a.ptr[q + 1] - b;
    if (q == me || n == 0) continue;
    const double *src = a.cm.xbuf[me] + ((size_t)a.parity * P + q) * a.region;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * a.narr; i += gridDim.x * blockDim.x) {
      const int k = i / n, j = i - k * n;
      a.arr[k][a.idx[b + j]] = __ldcv(src + (size_t)k * n + j);   // written by a peer: never from a stale cache line
    }
  }
}
// all-reduce of up to 4 words over the ranks: op 0 = minimum (order-preserving keys of doubles), 1 = sum.  One thread.
__global__ void k_peer_allreduce(CommDev cm, int parity, unsigned long long *vals, int n, int op)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int q = 0; q < cm.P; q++)
    for (int k = 0; k < n; k++) *((volatile unsigned long long *)(cm.mail[q] + MAIL_RED + (parity * UFM_MAX_RANKS + cm.rank) * 4 + k)) = vals[k];
  peer_sync(cm);
  volatile unsigned long long *mine = cm.mail[cm.rank];
  for (int k = 0; k < n; k++) {
    unsigned long long r = mine[MAIL_RED + (parity * UFM_MAX_RANKS) * 4 + k];
    for (int q = 1; q < cm.P; q++) {
      const unsigned long long v = mine[MAIL_RED + (parity * UFM_MAX_RANKS + q) * 4 + k];
      r = op == 0 ? (v < r ? v : r) : r + v;
    }
    vals[k] = r;
  }
}
// rows of this rank that a peer reads (xmask): their (U,V) as set up by the gather / the grounding-line flux, before the first viscosity pass
__global__ void __launch_bounds__(256) k_push_uv_halo(CommDev cm, int Mp, const unsigned char *xmask, const unsigned char *sowner, const double2 *UV)
{
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Mp; p += gridDim.x * blockDim.x) {
    const unsigned xm = xmask[p];
    if (!xm || sowner[p >> 5] != cm.rank) continue;
    const double2 v = UV[p];
    for (int q = 0; q < cm.P; q++) if ((xm >> q) & 1u) cm.uv[q][p] = v;
  }
  __threadfence_system();
}

// all-gather of the final (U,V): every rank pushes its own rows to all peers (partitioned runs only)
__global__ void k_push_uv(CommDev cm, int n_slices, const unsigned char *sowner, const unsigned char *deg, const double2 *UV)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < n_slices; s += nw) {
    if (sowner[s] != cm.rank) continue;
    const int p = s * 32 + lane;
    if (deg[p] == UFM_DEG_PAD) continue;
    const double2 v = UV[p];
    for (int q = 0; q < cm.P; q++) if (q != cm.rank) cm.uv[q][p] = v;
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// scatter AaAc -> Aa / Ac and rotate_xy_to_po (mesh_ArakawaC_module.f90:815-844)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ssa_finish(int nV, int nAc, const int *aa2m, const int *ac2m, const double2 *UV, const double *Dx_, const double *Dy_,
                                                    double *U, double *V, double *Ux, double *Uy, double *Up, double *Uo, const double *rmin, unsigned long long *cfl_key,
                                                    const unsigned char *__restrict__ own_aa, const unsigned char *__restrict__ own_ac, int rank)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double mS = 1000.0;
  if (i < nV && UFM_OWNED(own_aa, i, rank)) {
    const double2 q = UV[aa2m[i]];
    U[i] = q.x; V[i] = q.y;
    mS = rmin[i] / (fabs(q.x) + fabs(q.y));   // epilogue: this vertex's SSA critical time step (see k_cfl)
  }
  if (i < nAc && UFM_OWNED(own_ac, i, rank)) {
    const double2 q = UV[ac2m[i]];
    const double Dx = Dx_[i], Dy = Dy_[i], D = sqrt(Dx * Dx + Dy * Dy);
    Ux[i] = q.x; Uy[i] = q.y;
    Up[i] = q.x * Dx / D + q.y * Dy / D;
    Uo[i] = q.y * Dx / D - q.x * Dy / D;
  }
  block_min_to_key(mS, cfl_key);
}

__global__ void k_zero_d(size_t n, double *p) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.0; }

__global__ void k_sum_mask_sheet(int nV, const unsigned *mbits, unsigned long long *out, const unsigned char *__restrict__ own, int rank)
{
  unsigned long long c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nV; i += gridDim.x * blockDim.x) c += (UFM_OWNED(own, i, rank) && (mbits[i] & MB_SHEET)) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// =============================================================================================
// host launchers
// =============================================================================================
int ufm_ssa_powtab_init(const UfmPowTab *t) { return ufm_powtab_upload_tu(t); }

int ufm_halo_exchange(ufm_handle *h, int kind, int narr, double *const *arrays)
{
  DevMesh &m = h->mesh;
  if (!m.part_step) return 0;
  if (!h->comm_connected) return ufm_set_error(-6, "partitioned mesh but ufm_comm_connect has not been called");
  if (narr < 1 || narr > 4) return ufm_set_error(-2, "ufm_halo_exchange: 1..4 arrays");
  XArgs a;
  a.cm = h->comm; a.parity = h->xparity; a.region = m.x_region; a.narr = narr;
  for (int k = 0; k < 4; k++) a.arr[k] = k < narr ? arrays[k] : nullptr;
  a.idx = kind == 0 ? m.xa_s_idx : m.xc_s_idx;
  memcpy(a.ptr, kind == 0 ? m.xa_s_ptr : m.xc_s_ptr, sizeof(a.ptr));
  k_xpack<<<h->num_sms, 256, 0, h->stream>>>(a);
  k_peer_barrier<<<1, 32, 0, h->stream>>>(h->comm);
  a.idx = kind == 0 ? m.xa_r_idx : m.xc_r_idx;
  memcpy(a.ptr, kind == 0 ? m.xa_r_ptr : m.xc_r_ptr, sizeof(a.ptr));
  k_xunpack<<<h->num_sms, 256, 0, h->stream>>>(a);
  h->cnt.kernel_launches += 3;
  h->xparity ^= 1;
  return ufm_cuda_check(cudaGetLastError(), "halo exchange");
}

int ufm_peer_allreduce(ufm_handle *h, unsigned long long *vals_dev, int n, int op)
{
  if (h->mesh.P <= 1) return 0;
  if (!h->comm_connected) return ufm_set_error(-6, "partitioned mesh but ufm_comm_connect has not been called");
  if (n < 1 || n > 4) return ufm_set_error(-2, "ufm_peer_allreduce: 1..4 words");
  k_peer_allreduce<<<1, 32, 0, h->stream>>>(h->comm, h->xparity, vals_dev, n, op);
  h->cnt.kernel_launches++;
  h->xparity ^= 1;
  return ufm_cuda_check(cudaGetLastError(), "k_peer_allreduce");
}

int ufm_push_uv_halo(ufm_handle *h)
{
  DevMesh &m = h->mesh;
  if (m.P <= 1) return 0;
  if (!h->comm_connected) return ufm_set_error(-6, "partitioned mesh but ufm_comm_connect has not been called");
  k_push_uv_halo<<<h->num_sms * 2, 256, 0, h->stream>>>(h->comm, m.Mp, m.m_xmask, m.m_sowner, h->st.UV);
  k_peer_barrier<<<1, 32, 0, h->stream>>>(h->comm);
  h->cnt.kernel_launches += 2;
  return ufm_cuda_check(cudaGetLastError(), "k_push_uv_halo");
}

static inline int grid_for(int n, int b) { return (n + b - 1) / b; }

int ufm_k_sum_mask_sheet(ufm_handle *h, long long *out)
{
  DevState &s = h->st;
  UFM_CUDA(cudaMemsetAsync(s.ctrl + 16, 0, sizeof(unsigned long long), h->stream));
  k_sum_mask_sheet<<<h->num_sms * 2, 256, 0, h->stream>>>(h->mesh.nV, s.mbits, s.ctrl + 16, h->mesh.part_step ? h->mesh.own_aa : nullptr, h->mesh.rank);
  h->cnt.kernel_launches++;
  if (h->mesh.part_step) { int rc_x = ufm_peer_allreduce(h, s.ctrl + 16, 1, 1); if (rc_x) return rc_x; }   // SUM( ice%mask_sheet) over the ranks
  unsigned long long v = 0;
  UFM_CUDA(cudaMemcpyAsync(&v, s.ctrl + 16, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  *out = (long long)v;
  return 0;
}

int ufm_k_ssa_zero(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  UFM_CUDA(cudaMemsetAsync(s.UV, 0, sizeof(double2) * (size_t)m.Mp, h->stream));
  UFM_CUDA(cudaMemsetAsync(s.U_SSA, 0, sizeof(double) * (size_t)m.nVp, h->stream));
  UFM_CUDA(cudaMemsetAsync(s.V_SSA, 0, sizeof(double) * (size_t)m.nVp, h->stream));
  for (int k = 0; k < 4; k++) UFM_CUDA(cudaMemsetAsync(s.U_SSA_Ac[k], 0, sizeof(double) * (size_t)m.nAcp, h->stream));
  h->cfl_ok[1] = true;   // zero velocities: the SSA critical time step is the reference's initial 1000 yr
  return ufm_cfl_key_reset(h, 2);
}

int ufm_k_ssa_prepare(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  PrepArgs a;
  a.Mp = m.Mp; a.src = m.m_src;
  a.Hi = s.Hi; a.Hb = s.Hb; a.SL = s.SL; a.sx = s.dHs_dx_shelf; a.sy = s.dHs_dy_shelf; a.U = s.U_SSA; a.V = s.V_SSA;
  a.Hi_Ac = s.Hi_Ac; a.Hb_Ac = s.Hb_Ac; a.SL_Ac = s.SL_Ac; a.sx_Ac = s.dHs_dx_shelf_Ac; a.sy_Ac = s.dHs_dy_shelf_Ac;
  a.Ux_Ac = s.U_SSA_Ac[0]; a.Uy_Ac = s.U_SSA_Ac[1]; a.mbits_Ac = s.mbits_Ac; a.gl_fix = h->P.use_analytical_GL_flux;
  a.A_mean = s.realistic_A ? s.A_mean : nullptr; a.A_mean_Ac = s.A_mean_Ac; a.m_enh_ssa = h->P.m_enh_ssa; a.Afac = s.Afac;
  a.UV = s.UV; a.rhsnum = s.rhsnum; a.tau_c = s.tau_c; a.phi = s.phi; a.Hm = s.Hm; a.mflag = s.mflag;
  a.sowner = m.part_step ? m.m_sowner : nullptr; a.rank = m.rank;
  k_ssa_prepare<<<grid_for(m.Mp, 256), 256, 0, h->stream>>>(a);
  h->cnt.kernel_launches++;
  if (h->P.use_analytical_GL_flux) {
    // factor_Tsai = 8 Q0 A (rho g)^n (1-rho_i/rho_w)^(n-1) / 4^n, evaluated left to right as at :906-909 (host libm)
    const double Q0 = 0.61;
    double f = 8.0 * Q0 * s.A_flow_const * pow(UFM_ICE_DENSITY * UFM_GRAV, UFM_N_FLOW) *
               pow(1.0 - (UFM_ICE_DENSITY / UFM_SEAWATER_DENSITY), UFM_N_FLOW - 1.0) / pow(4.0, UFM_N_FLOW);
    k_gl_flux<<<grid_for(m.nAc, 256), 256, 0, h->stream>>>(m.nAc, m.ac_Aci, s.mbits_Ac, s.mbits, s.Hi, s.Hb, s.SL, s.phi, m.aa2m, 1.0,
                                                           s.dHi_Ac[0], s.dHi_Ac[1], s.dSL_Ac[0], s.dSL_Ac[1], s.dHb_Ac[0], s.dHb_Ac[1],
                                                           m.ac_Dx, m.ac_Dy, f, s.Qabs_GL_Ac, s.Qp_GL_Ac, s.U_SSA_Ac[0], s.U_SSA_Ac[1],
                                                           m.ac2m, s.UV, s.realistic_A ? s.A_mean : nullptr, pow(UFM_ICE_DENSITY * UFM_GRAV, UFM_N_FLOW),
                                                           pow(1.0 - (UFM_ICE_DENSITY / UFM_SEAWATER_DENSITY), UFM_N_FLOW - 1.0), pow(4.0, UFM_N_FLOW),
                                                           m.part_step ? m.own_ac : nullptr, m.rank);
    h->cnt.kernel_launches++;
  }
  if (m.part_step) {   // the start values of the rows the neighbour strips read (viscosity, first sweep)
    int rc_x = ufm_push_uv_halo(h);
    if (rc_x) return rc_x;
  }
  return ufm_cuda_check(cudaGetLastError(), "k_ssa_prepare");
}

static void fill_visc_args(ufm_handle *h, ViscArgs &a, bool fuse, bool device_ctl)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  a.cm = h->comm; a.rng = m.rng_dev;
  a.n_slices = m.m.n_slices; a.off = m.m.off; a.deg = m.m.deg; a.idx = m.m_idx; a.nx = m.m_nx; a.ny = m.m_ny; a.nx0 = m.m_nx0; a.ny0 = m.m_ny0;
  a.Hm = s.Hm; a.UV = s.UV;
  a.visc_A = pow(h->P.m_enh_ssa * 0.5 * s.A_flow_const, -1.0 / UFM_N_FLOW);
  a.Afac = s.realistic_A ? s.Afac : nullptr;
  a.eta = s.eta; a.N = s.N; a.dU = s.dU; a.dV = s.dV; a.partials = s.partials;
  a.fuse_setup = fuse ? 1 : 0; a.mflag = s.mflag; a.rhsnum = s.rhsnum; a.tau_c = s.tau_c; a.cU0 = m.m_cU0; a.cV0 = m.m_cV0;
  a.thr = pow(100.0, 0.30); a.S = s.S; a.RHS = s.RHS; a.E = s.E;
  a.sctl = device_ctl ? s.ctrl + SCTL_BASE : nullptr;
}

// one viscosity evaluation (+ fused sliding term / linear-system setup) and the RN sums; no host synchronisation
static int enqueue_viscosity(ufm_handle *h, bool fuse, bool device_ctl)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  if (m.P > 1 && !h->comm_connected) return ufm_set_error(-6, "partitioned mesh but ufm_comm_connect has not been called");
  ViscArgs a;
  fill_visc_args(h, a, fuse, device_ctl);
  {
    // resident CTAs per SM: 2 (about 120 registers, every load of a row in flight at once), 3 or 4 (64 registers, twice the warps)
    if (h->visc_minb < 0) { const char *e = getenv("UFM_VISC_MINB"); h->visc_minb = e ? atoi(e) : UFM_VISC_MINB_DEFAULT; }
    const int minb = h->visc_minb;
    if (minb == 6) k_ssa_viscosity<false, 6><<<h->num_sms * 6, 256, 0, h->stream>>>(a);
    else if (minb == 5) k_ssa_viscosity<false, 5><<<h->num_sms * 5, 256, 0, h->stream>>>(a);
    else if (minb == 4) k_ssa_viscosity<false, 4><<<h->num_sms * 4, 256, 0, h->stream>>>(a);
    else if (minb == 3) k_ssa_viscosity<false, 3><<<h->num_sms * 3, 256, 0, h->stream>>>(a);
    else if (minb == 2) k_ssa_viscosity<false, 2><<<h->num_sms * 2, 256, 0, h->stream>>>(a);
    else k_ssa_viscosity<false, 1><<<h->num_sms * 8, 256, 0, h->stream>>>(a);
  }
  if (m.P > 1) { k_peer_barrier<<<1, 32, 0, h->stream>>>(h->comm); h->cnt.kernel_launches++; }
  k_sum_partials<<<SUM_BLOCKS, 256, 0, h->stream>>>(m.m.n_slices, s.partials, s.scal, device_ctl ? s.ctrl + SCTL_BASE : nullptr, h->P.SSA_RN_tol,
                                                    s.red_scratch, (unsigned *)(s.ctrl + CTRL_RN_TICKET));
  h->cnt.kernel_launches += 2;
  return ufm_cuda_check(cudaGetLastError(), "k_ssa_viscosity");
}

int ufm_k_ssa_viscosity(ufm_handle *h, double sums2[2])
{
  DevState &s = h->st;
  int rc = enqueue_viscosity(h, true, false);
  if (rc) return rc;
  UFM_CUDA(cudaMemcpyAsync(s.scal_h, s.scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  sums2[0] = s.scal_h[0]; sums2[1] = s.scal_h[1];
  return 0;
}

// dU_SSA_dx_AaAc ... dV_SSA_dy_AaAc are pure diagnostics in the reference (written at ice_dynamics_module.f90:712-713, read
// nowhere else); the device path does not store them per viscosity iteration but recomputes them from the current
// U,V when the host asks for them.
int ufm_k_ssa_gradients(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  ViscArgs a;
  // diagnostic gradients on ALL rows (each rank's (U,V) is complete after ufm_ssa_finish): single-rank view of the ranges
  fill_visc_args(h, a, false, false);
  a.cm.P = 1; a.cm.rank = 0; a.rng = m.rng_all_dev; a.visc_A = 0.0; a.Afac = nullptr; a.eta = s.eta; a.N = s.N; a.dU = s.dU; a.dV = s.dV; a.partials = s.partials;
  int grid = h->num_sms * 8;
  k_ssa_viscosity<true, 1><<<grid, 256, 0, h->stream>>>(a);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_ssa_gradients");
}

int ufm_k_ssa_sliding_setup(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  SetupArgs a;
  a.Mp = m.Mp; a.P = m.P; a.rank = m.rank; a.sowner = m.m_sowner; a.deg = m.m.deg; a.mflag = s.mflag; a.UV = s.UV; a.rhsnum = s.rhsnum; a.tau_c = s.tau_c; a.eta = s.eta; a.Hm = s.Hm;
  a.cU0 = m.m_cU0; a.cV0 = m.m_cV0; a.thr = pow(100.0, 0.30); a.S = s.S; a.RHS = s.RHS; a.E = s.E;
  k_ssa_setup<<<grid_for(m.Mp, 256), 256, 0, h->stream>>>(a);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_ssa_setup");
}

typedef void (*sor_kernel_t)(SorArgs);
static sor_kernel_t pick_sor(const ufm_handle *h)
{
  const bool ex = h->P.exact_xy != 0, gl = h->P.use_analytical_GL_flux != 0, mu = h->mesh.P > 1;
  if (h->sor_df && !mu) {
    if (ex) return gl ? k_ssa_sor_df<true, true> : k_ssa_sor_df<true, false>;
    return gl ? k_ssa_sor_df<false, true> : k_ssa_sor_df<false, false>;
  }
  if (mu) {
    if (ex) return gl ? k_ssa_sor<true, true, true> : k_ssa_sor<true, false, true>;
    return gl ? k_ssa_sor<false, true, true> : k_ssa_sor<false, false, true>;
  }
  if (ex) return gl ? k_ssa_sor<true, true, false> : k_ssa_sor<true, false, false>;
  return gl ? k_ssa_sor<false, true, false> : k_ssa_sor<false, false, false>;
}

int ufm_sor_configure(ufm_handle *h)
{
  int per_sm = 0;
  // tuning switches, re-read on every configure so that one process can compare variants (tools/sor_probe.py)
  { const char *e = getenv("UFM_SOR_CHUNK"); h->sor_chunk = e ? atoi(e) : UFM_SOR_CHUNK_DEFAULT; }
  if (getenv("UFM_SOR_TRACE") && !h->sor_trace) { UFM_CUDA(cudaMalloc((void **)&h->sor_trace, 4096 * 24 * sizeof(unsigned long long))); UFM_CUDA(cudaMemset(h->sor_trace, 0, 4096 * 24 * sizeof(unsigned long long))); }
  { const char *e = getenv("UFM_SOR_BAR"); h->sor_bar = e ? atoi(e) : UFM_SOR_BAR_DEFAULT; }
  { const char *e = getenv("UFM_SOR_FUSE_BC"); h->sor_fuse_bc = e ? atoi(e) : UFM_SOR_FUSE_BC_DEFAULT; }
  h->sor_block = SOR_BLOCK;
  h->sor_smem = 0;
  { const char *e = getenv("UFM_SOR_DATAFLOW"); h->sor_df = (h->mesh.df_layout && h->mesh.P == 1 && e && atoi(e) != 0) ? 1 : 0; }
  // one CTA of 1024 threads per SM; UFM_SOR_GRID (tests only) shrinks the grid so that small meshes are swept in several rounds per colour
  int grid_target = h->num_sms;
  { const char *e = getenv("UFM_SOR_GRID"); if (e && atoi(e) > 0 && atoi(e) < grid_target) grid_target = atoi(e); }
  if (h->sor_df) {
    // the stage table depends only on the mesh and the grid (the occupancy query below confirms that the grid is resident)
    DevMesh &m = h->mesh;
    const int nw = grid_target * (h->sor_block / 32);
    int base = 0;
    for (int c = 0; c < 5; c++) {
      const int n = m.rng[c][0][2] - m.rng[c][0][0];
      const int K = n > 0 ? (n + nw - 1) / nw : 1;
      m.df_K[c] = K; m.df_act[c] = n > 0 ? (n + K - 1) / K : 1; m.df_base[c] = base;
      base += K;
    }
    m.df_n_stages = base;
    if (base > DF_MAX_STAGES) h->sor_df = 0;   // meshes far beyond one GPU's memory: barrier kernel
    else h->sor_smem = (size_t)base * (sizeof(int4) + sizeof(unsigned));
  }
  // The streaming kernels keep ~230 B of loads per thread in flight and reuse gathered (U,V) lines: they want the unified
  // L1 / shared-memory array as L1 (a 120 KB shared-memory carve-out costs the sweep 30 %).  UFM_L1_CARVEOUT
  // (percent of shared memory, -1 = leave the driver's default) exists for A/B measurements.
  {
    // once per kernel variant and handle (the attribute is per device): ufm_sor_configure runs before every piecewise SOR call
    const void *fn = (const void *)pick_sor(h);
    if (h->carveout_done_for != fn) {
      const char *e = getenv("UFM_L1_CARVEOUT");
      const int pct = e ? atoi(e) : 0;
      if (pct >= 0) {
        UFM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        UFM_CUDA(cudaFuncSetAttribute((const void *)k_ssa_viscosity<false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
      }
      h->carveout_done_for = fn;
    }
  }
  UFM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pick_sor(h), h->sor_block, h->sor_smem));
  if (per_sm < 1) return ufm_set_error(-3, "SOR kernel cannot be made resident");
  h->sor_grid = getenv("UFM_SOR_GRID") ? std::min(per_sm * h->num_sms, grid_target) : per_sm * h->num_sms;
  if (h->sor_df) {
    // need[] of the dataflow sweep for this grid (once per mesh and grid)
    DevMesh &m = h->mesh;
    const int nw = h->sor_grid * (h->sor_block / 32);
    if (h->sor_grid != grid_target) return ufm_set_error(-3, "dataflow SOR sweep: %d CTAs resident, expected %d", h->sor_grid, grid_target);
    if (m.df_ready_for != nw) {
      NeedArgs na;
      for (int c = 0; c < 5; c++) {
        for (int q = 0; q < 3; q++) na.rng15[3 * c + q] = m.rng[c][0][q];
        na.st_base[c] = m.df_base[c]; na.st_act[c] = m.df_act[c];
      }
      na.n_slices = m.m.n_slices; na.off = m.m.off; na.deg = m.m.deg; na.idx = m.m_idx;
      na.n_bc = m.n_bc; na.bc_ptr = m.bc_ptr; na.bc_nbr = m.bc_nbr; na.corner = m.corner_dev; na.corner_nbr = m.corner_nbr; na.corner_row = m.corner_row;
      na.need = m.df_need; na.need_bc = m.df_need_bc; na.trace = h->sor_trace;
      k_sor_need<<<h->num_sms * 8, 256, 0, h->stream>>>(na);
      UFM_CUDA(cudaGetLastError());
      h->cnt.kernel_launches++;
      m.df_ready_for = nw;
    }
  }
  return 0;
}

static int enqueue_sor(ufm_handle *h, int max_inner, int force_iters, bool device_ctl, cudaEvent_t e0, cudaEvent_t e1)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  if (m.P > 1 && !h->comm_connected) return ufm_set_error(-6, "partitioned mesh but ufm_comm_connect has not been called");
  SorArgs a;
  a.cm = h->comm; a.xmask = m.m_xmask; a.rng = m.rng_dev;
  a.bc_begin = m.bc_rng[m.rank]; a.bc_end = m.bc_rng[m.rank + 1];
  a.corner_mask = 0;
  for (int k = 0; k < 4; k++) if (m.corner_owner[k] == m.rank) a.corner_mask |= 1 << k;
  a.off = m.m.off; a.deg = m.m.deg; a.mflag = s.mflag; a.idx = m.m_idx; a.cU = m.m_cU; a.cV = m.m_cV; a.nxy = m.m_nxy; a.nxy0 = m.m_nxy0;
  a.nxysum = m.m_nxysum; a.E = s.E; a.RHS = s.RHS; a.UV = s.UV;
  a.corner = m.corner_dev;
  a.bc_pos = m.bc_pos; a.bc_ptr = m.bc_ptr; a.bc_nbr = m.bc_nbr;
  a.corner_nbr = m.corner_nbr; a.corner_row = m.corner_row;
  a.chunk = h->sor_chunk; a.trace = h->sor_trace; a.bar_rel = h->sor_bar;
  a.fuse_bc = (h->sor_fuse_bc && m.P == 1 && !m.df_layout) ? 1 : 0; a.adj_end = m.adj5_end;
  a.stage_cnt = m.df_stage_cnt; a.need = m.df_need; a.need_bc = m.df_need_bc; a.n_stages = m.df_n_stages;
  for (int c = 0; c < 5; c++) { a.st_base[c] = m.df_base[c]; a.st_K[c] = m.df_K[c]; a.st_act[c] = m.df_act[c]; }
  if (h->sor_df) UFM_CUDA(cudaMemsetAsync(m.df_stage_cnt, 0, sizeof(unsigned) * DF_CNT_STRIDE * (size_t)std::max(1, m.df_n_stages), h->stream));
  a.Mp = m.Mp; a.max_inner = max_inner; a.force_iters = force_iters; a.omega = h->P.SSA_SOR_omega; a.tol = h->P.SSA_max_residual_UV;
  a.ctrl = s.ctrl; a.sctl = device_ctl ? s.ctrl + SCTL_BASE : nullptr;
  UFM_CUDA(cudaMemsetAsync(s.ctrl, 0, 16 * sizeof(unsigned long long), h->stream));
  void *args[] = {&a};
  UFM_CUDA(cudaEventRecord(e0, h->stream));
  UFM_CUDA(cudaLaunchCooperativeKernel((void *)pick_sor(h), dim3(h->sor_grid), dim3(h->sor_block), args, h->sor_smem, h->stream));
  UFM_CUDA(cudaEventRecord(e1, h->stream));
  h->cnt.kernel_launches++; h->cnt.sor_launches++;
  return 0;
}

int ufm_k_ssa_sor(ufm_handle *h, int max_inner, int force_iters, ufm_ssa_stats *st)
{
  DevState &s = h->st;
  int rc = ufm_sor_configure(h);
  if (rc) return rc;
  if ((rc = enqueue_sor(h, max_inner, force_iters, false, h->ev0, h->ev1))) return rc;
  unsigned long long *res = (unsigned long long *)(s.scal_h + 8);
  UFM_CUDA(cudaMemcpyAsync(res, s.ctrl + 8, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  UFM_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->cnt.sor_ms += ms; h->cnt.sor_iterations += (long long)res[0];
  if (st) {
    st->n_inner_last = (int)res[0];
    st->did_reset = (int)(res[1] & 1);
    st->rc = (res[1] & 2) ? 1 : 0;
    if (res[1] & 4) return ufm_set_error(-7, "SOR: wait for a peer GPU timed out (partitioned run)");
    if (res[1] & 8) return ufm_set_error(-9, "SOR: a dependency wait of the dataflow sweep timed out");
    double r;
    memcpy(&r, &res[2], sizeof(r));
    st->last_max_residual = r;
  }
  return 0;
}

// The outer (viscosity) loop of solve_SSA (:501-543) with device-side control flow: UFM_OUTER_BATCH outer iterations are
// enqueued back to back (viscosity+setup, RN reduction, SOR) and every kernel returns at once when SCTL_STOP is set, so
// the host synchronises once per batch instead of twice per outer iteration.
#define UFM_OUTER_BATCH 10
int ufm_k_ssa_outer_loop(ufm_handle *h, ufm_ssa_stats *st)
{
  DevState &s = h->st;
  int rc = ufm_sor_configure(h);
  if (rc) return rc;
  if (!h->ev_pool[0]) for (int k = 0; k < 2 * UFM_OUTER_BATCH; k++) UFM_CUDA(cudaEventCreate(&h->ev_pool[k]));
  UFM_CUDA(cudaMemsetAsync(s.ctrl + SCTL_BASE, 0, 8 * sizeof(unsigned long long), h->stream));
  unsigned long long *c = (unsigned long long *)(s.scal_h + 24);
  const int max_outer = h->P.SSA_max_outer_loops;
  int launched = 0;
  long long inner_before = 0;
  bool stop = false;
  while (!stop && launched < max_outer) {
    const int nb = (max_outer - launched) < UFM_OUTER_BATCH ? (max_outer - launched) : UFM_OUTER_BATCH;
    for (int k = 0; k < nb; k++) {
      if ((rc = enqueue_viscosity(h, true, true))) return rc;
      if ((rc = enqueue_sor(h, h->P.SSA_max_inner_loops, 0, true, h->ev_pool[2 * k], h->ev_pool[2 * k + 1]))) return rc;
    }
    launched += nb;
    UFM_CUDA(cudaMemcpyAsync(c, s.ctrl + SCTL_BASE, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    UFM_CUDA(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < nb; k++) { float ms = 0.f; UFM_CUDA(cudaEventElapsedTime(&ms, h->ev_pool[2 * k], h->ev_pool[2 * k + 1])); h->cnt.sor_ms += ms; }
    h->cnt.sor_iterations += (long long)c[SCTL_NINNER] - inner_before;
    inner_before = (long long)c[SCTL_NINNER];
    stop = c[SCTL_STOP] != 0;
  }
  double d;
  st->n_outer = (int)c[SCTL_NOUTER]; st->n_inner_total = (int)c[SCTL_NINNER]; st->n_inner_last = (int)c[SCTL_NLAST];
  st->did_reset = (int)c[SCTL_RESET];
  memcpy(&d, &c[SCTL_RN], sizeof(d)); st->last_RN = d;
  memcpy(&d, &c[SCTL_MAXRES], sizeof(d)); st->last_max_residual = d;
  st->rc = (c[SCTL_RC] & 2) ? -1 : ((c[SCTL_RC] & 1) ? 1 : 0);
  if (c[SCTL_RC] & 4) return ufm_set_error(-7, "SOR: wait for a peer GPU timed out (partitioned run)");
  if (c[SCTL_RC] & 8) return ufm_set_error(-9, "SOR: a dependency wait of the dataflow sweep timed out");
  return 0;
}

int ufm_k_ssa_finish(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  int n = m.nV > m.nAc ? m.nV : m.nAc;
  if (m.P > 1 && !m.part_step) {   // replicated per-step kernels: every rank scatters every row, so it needs every row
    if (!h->comm_connected) return ufm_set_error(-6, "partitioned mesh but ufm_comm_connect has not been called");
    k_push_uv<<<h->num_sms * 4, 256, 0, h->stream>>>(h->comm, m.m.n_slices, m.m_sowner, m.m.deg, s.UV);
    k_peer_barrier<<<1, 32, 0, h->stream>>>(h->comm);
    h->cnt.kernel_launches += 2;
  }
  int rc_ = ufm_cfl_key_reset(h, 2);
  if (rc_) return rc_;
  k_ssa_finish<<<grid_for(n, 256), 256, 0, h->stream>>>(m.nV, m.nAc, m.aa2m, m.ac2m, s.UV, m.ac_Dx, m.ac_Dy, s.U_SSA, s.V_SSA,
                                                        s.U_SSA_Ac[0], s.U_SSA_Ac[1], s.U_SSA_Ac[2], s.U_SSA_Ac[3], m.aa_rmin, s.ctrl + CTRL_CFL_KEYS + 1,
                                                        m.part_step ? m.own_aa : nullptr, m.part_step ? m.own_ac : nullptr, m.rank);
  h->cfl_ok[1] = true;
  h->cnt.kernel_launches++;
  if (m.part_step) {   // the next thickness update of the neighbour strips reads the parallel velocity on the staggered vertices it shares with us
    double *arr[1] = {s.U_SSA_Ac[2]};
    int rc_x = ufm_halo_exchange(h, 1, 1, arr);
    if (rc_x) return rc_x;
  }
  return ufm_cuda_check(cudaGetLastError(), "k_ssa_finish");
}
