// ufm_geom.cu -- per-model-step kernels: update_general_ice_model_data, solve_SIA,
// calculate_ice_thickness_change, the CFL minima, and the permuted field copies.
//
// Same arithmetic contract as ufm_ssa.cu (-fmad=false, reference evaluation order): everything
// here is add/mul/div/sqrt/compare and reproduces the CPU restatement bit for bit, except the one
// pow(x, 3.0) in the SIA diffusivity (CUDA libm <= 2 ulp).
//
// All kernels are streaming / gather kernels bounded by HBM bandwidth; rows are Morton-ordered so the
// gathers of a warp fall into a few neighbouring 128 B lines.
#include <math.h>

#include <climits>
#include <cstring>

#include "ufm_internal.cuh"
#include "ufm_pow.cuh"

int ufm_geom_powtab_init(const UfmPowTab *t) { return ufm_powtab_upload_tu(t); }
static inline int grid_for(long long n, int b) { return (int)((n + b - 1) / b); }

__device__ __forceinline__ bool g_is_floating(double Hi, double Hb, double SL)
{
  return Hi < (SL - Hb) * UFM_SEAWATER_DENSITY / UFM_ICE_DENSITY;
}
__device__ __forceinline__ double g_Hs(double Hi, double Hb, double SL)
{
  // general_ice_model_data_module.f90:46-52
  return Hi + fmax(SL - UFM_ICE_DENSITY / UFM_SEAWATER_DENSITY * Hi, Hb);
}
__device__ __forceinline__ unsigned g_bits1(double Hi, double Hb, double SL)
{
  // determine_masks stage 1, :129-160 ; mask codes :108-116
  unsigned b, code = 0;
  const bool ocean = g_is_floating(Hi, Hb, SL), ice = Hi > 0.0;
  b = ocean ? MB_OCEAN : MB_LAND;
  if (ocean) code = 1;
  if (ice) b |= MB_ICE;
  if (ice && !ocean) { b |= MB_SHEET; code = 3; }
  if (ice && ocean) { b |= MB_SHELF; code = 4; }
  return b | (code << MB_CODE_SHIFT);
}

// Device-driven region loop (ufm_api.cu, run_model_device): the per-step kernels are enqueued many steps ahead; whether a step still
// runs, and which of its actions are due, is decided on the device (k_step_control) and read here.  gate == NULL: always run.
#define UFM_GATE(g) do { if ((g) && !*((const volatile int *)(g))) return; } while (0)

// ---- Aa stage 1: Hs, dHs_dt, primary mask bits ----
__global__ void k_geom_aa1(int nV, const double *__restrict__ Hi, const double *__restrict__ Hb, const double *__restrict__ SL,
                           const double *__restrict__ dHb_dt, const double *__restrict__ dHi_dt, double *__restrict__ Hs,
                           double *__restrict__ dHs_dt, unsigned *__restrict__ mbits, const int *gate, const unsigned char *__restrict__ act)
{
  UFM_GATE(gate);
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  if (act && !act[v]) return;   // partitioned: own vertices and the ones an owned Aa / Ac element reads
  const double hi = Hi[v], hb = Hb[v], sl = SL[v];
  Hs[v] = g_Hs(hi, hb, sl);
  dHs_dt[v] = dHb_dt[v] + dHi_dt[v];
  mbits[v] = g_bits1(hi, hb, sl);
}

// ---- Aa stage 2: gradients of Hi and Hs (get_mesh_derivatives, mesh_derivatives_module.f90:315-371),
//      neighbour-dependent masks (:163-226), shelf slopes (:70-79).  One warp per slice. ----
struct GeomAa2Args {
  int n_slices;
  const long long *off;
  const unsigned char *deg;
  const int *C;
  const double *Nx, *Ny, *Nx0, *Ny0;
  const double *Hi, *Hs;
  unsigned *mbits;
  double *dHi_dx, *dHi_dy, *dHs_dx, *dHs_dy, *sx, *sy;
  const int *gate;
  const unsigned char *own; int rank;
};
__global__ void __launch_bounds__(256) k_geom_aa2(GeomAa2Args a)
{
  UFM_GATE(a.gate);
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < a.n_slices; s += nw) {
    const long long o = a.off[s];
    const int w = (int)((a.off[s + 1] - o) >> 5);
    const int v = s * 32 + lane;
    const int n = a.deg[v];
    if (n == UFM_DEG_PAD || !UFM_OWNED(a.own, v, a.rank)) continue;
    const double hi = a.Hi[v], hs = a.Hs[v];
    const unsigned own = a.mbits[v];
    double ix = a.Nx0[v] * hi, iy = a.Ny0[v] * hi, sx = a.Nx0[v] * hs, sy = a.Ny0[v] * hs;
    unsigned any_ocean = 0, any_noice = 0, any_shelf = 0;
    for (int c = 0; c < w; c++) {
      if (c < n) {
        const long long e = o + (long long)c * 32 + lane;
        const int j = __ldcs(a.C + e);
        const double cx = __ldcs(a.Nx + e), cy = __ldcs(a.Ny + e);
        const double hj = a.Hi[j], sj = a.Hs[j];
        const unsigned bj = a.mbits[j];
        ix = ix + cx * hj; iy = iy + cy * hj;
        sx = sx + cx * sj; sy = sy + cy * sj;
        any_ocean |= bj & MB_OCEAN; any_noice |= (~bj) & MB_ICE; any_shelf |= bj & MB_SHELF;
      }
    }
    unsigned b = own & 0x3F, code = (own >> MB_CODE_SHIFT) & 0xF;
    if ((own & MB_LAND) && any_ocean) { b |= MB_COAST; code = 5; }
    if ((own & MB_ICE) && any_noice) { b |= MB_MARGIN; code = 6; }
    if ((own & MB_SHEET) && any_shelf) { b |= MB_GL; code = 7; }
    if ((own & MB_ICE) && any_ocean) { b |= MB_CF; code = 8; }
    a.mbits[v] = b | (code << MB_CODE_SHIFT);
    a.dHi_dx[v] = ix; a.dHi_dy[v] = iy; a.dHs_dx[v] = sx; a.dHs_dy[v] = sy;
    if (!(own & MB_OCEAN)) { a.sx[v] = sx; a.sy[v] = sy; }
    else {
      a.sx[v] = (1.0 - UFM_ICE_DENSITY / UFM_SEAWATER_DENSITY) * ix;
      a.sy[v] = (1.0 - UFM_ICE_DENSITY / UFM_SEAWATER_DENSITY) * iy;
    }
  }
}

// ---- Ac: map_Aa_to_Ac x3, Hs_Ac, masks on Ac (:232-294), get_mesh_derivatives_Ac x4
//      (mesh_ArakawaC_module.f90:542-581) fused: coefficients and indices are read once for 16 outputs ----
struct GeomAcArgs {
  int nAc;
  const int4 *Aci;
  const double *Nx[4], *Ny[4], *No[4], *Np;
  const double *Hi, *Hb, *SL, *Hs;
  const unsigned *mbits;
  double *Hi_Ac, *Hb_Ac, *SL_Ac, *Hs_Ac;
  double *dHi[4], *dHb[4], *dHs[4], *dSL[4];
  double *sx, *sy;
  unsigned *mbits_Ac;
  const int *gate;
  const unsigned char *own; int rank;
};
// (compiled for 5, 6, 8 resident CTAs per SM instead of the 4 the compiler picks by itself: 0.280, 0.351, 0.442 ms per update_general against 0.252 -- spills)
__global__ void __launch_bounds__(256) k_geom_ac(GeomAcArgs a)
{
  UFM_GATE(a.gate);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nAc || !UFM_OWNED(a.own, i, a.rank)) return;
  const int4 v = a.Aci[i];
  const int vv[4] = {v.x, v.y, v.z, v.w};
  double hi[4], hb[4], sl[4], hs[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { hi[k] = a.Hi[vv[k]]; hb[k] = a.Hb[vv[k]]; sl[k] = a.SL[vv[k]]; hs[k] = a.Hs[vv[k]]; }
  const double hi_ac = (hi[0] + hi[1]) / 2.0, hb_ac = (hb[0] + hb[1]) / 2.0, sl_ac = (sl[0] + sl[1]) / 2.0;
  a.Hi_Ac[i] = hi_ac; a.Hb_Ac[i] = hb_ac; a.SL_Ac[i] = sl_ac;
  a.Hs_Ac[i] = g_Hs(hi_ac, hb_ac, sl_ac);
  double d[4][3];  // [field][x,y,o]
#pragma unroll
  for (int f = 0; f < 4; f++) { d[f][0] = 0.0; d[f][1] = 0.0; d[f][2] = 0.0; }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double nx = __ldcs(a.Nx[k] + i), ny = __ldcs(a.Ny[k] + i), no = __ldcs(a.No[k] + i);
    d[0][0] = d[0][0] + nx * hi[k]; d[0][1] = d[0][1] + ny * hi[k]; d[0][2] = d[0][2] + no * hi[k];
    d[1][0] = d[1][0] + nx * hb[k]; d[1][1] = d[1][1] + ny * hb[k]; d[1][2] = d[1][2] + no * hb[k];
    d[2][0] = d[2][0] + nx * hs[k]; d[2][1] = d[2][1] + ny * hs[k]; d[2][2] = d[2][2] + no * hs[k];
    d[3][0] = d[3][0] + nx * sl[k]; d[3][1] = d[3][1] + ny * sl[k]; d[3][2] = d[3][2] + no * sl[k];
  }
  const double np = __ldcs(a.Np + i);
  a.dHi[0][i] = d[0][0]; a.dHi[1][i] = d[0][1]; a.dHi[2][i] = np * (hi[1] - hi[0]); a.dHi[3][i] = d[0][2];
  a.dHb[0][i] = d[1][0]; a.dHb[1][i] = d[1][1]; a.dHb[2][i] = np * (hb[1] - hb[0]); a.dHb[3][i] = d[1][2];
  a.dHs[0][i] = d[2][0]; a.dHs[1][i] = d[2][1]; a.dHs[2][i] = np * (hs[1] - hs[0]); a.dHs[3][i] = d[2][2];
  a.dSL[0][i] = d[3][0]; a.dSL[1][i] = d[3][1]; a.dSL[2][i] = np * (sl[1] - sl[0]); a.dSL[3][i] = d[3][2];
  unsigned b = g_bits1(hi_ac, hb_ac, sl_ac), code = (b >> MB_CODE_SHIFT) & 0xF;
  b &= 0x3F;
  const unsigned bi = a.mbits[v.x], bj = a.mbits[v.y];
  if (((bi & MB_LAND) && (bj & MB_OCEAN)) || ((bj & MB_LAND) && (bi & MB_OCEAN))) { b |= MB_COAST; code = 5; }
  if (((bi & MB_ICE) && !(bj & MB_ICE)) || ((bj & MB_ICE) && !(bi & MB_ICE))) { b |= MB_MARGIN; code = 6; }
  if (((bi & MB_SHEET) && (bj & MB_SHELF)) || ((bj & MB_SHEET) && (bi & MB_SHELF))) { b |= MB_GL; code = 7; }
  if (((bi & MB_ICE) && !(bj & MB_SHELF) && (bj & MB_OCEAN)) || ((bj & MB_ICE) && !(bi & MB_SHELF) && (bi & MB_OCEAN))) { b |= MB_CF; code = 8; }
  a.mbits_Ac[i] = b | (code << MB_CODE_SHIFT);
  if (!(b & MB_OCEAN)) { a.sx[i] = d[2][0]; a.sy[i] = d[2][1]; }
  else {
    a.sx[i] = (1.0 - UFM_ICE_DENSITY / UFM_SEAWATER_DENSITY) * d[0][0];
    a.sy[i] = (1.0 - UFM_ICE_DENSITY / UFM_SEAWATER_DENSITY) * d[0][1];
  }
}

// ---- ice_physical_properties, temperature-dependent branch (general_ice_model_data_module.f90:372-462): Arrhenius flow
//      factor per layer; the (.,nZ) arrays A_flow / A_flow_Ac / Ti_Ac are never materialised, only their vertical means ----
struct ZetaConst { int nZ; double dz[UFM_MAX_NZ]; double z3[UFM_MAX_NZ]; };
__device__ __forceinline__ double g_arrhenius(double Ti)
{
  const double A_low_temp = 1.14E-05, A_high_temp = 5.47E+10, Q_low_temp = 6.0E+04, Q_high_temp = 13.9E+04, R_gas = 8.314;
  return (Ti < 263.15) ? A_low_temp * exp(-Q_low_temp / (R_gas * Ti)) : A_high_temp * exp(-Q_high_temp / (R_gas * Ti));
}
// mode 0: Aa vertices; mode 1: Ac vertices with Ti_Ac = (Ti(vi)+Ti(vj))/2 (map_Aa_to_Ac_3D, mesh_ArakawaC_module.f90:748-769)
template <int MODE>
__global__ void __launch_bounds__(256) k_flow_mean(int n, int nVp, ZetaConst Z, const int4 *__restrict__ Aci, const double *__restrict__ Ti,
                                                   const unsigned *__restrict__ mbits, double *__restrict__ A_mean)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int vi = i, vj = i;
  if (MODE == 1) { const int4 v = Aci[i]; vi = v.x; vj = v.y; }
  auto T = [&](int k) { return MODE == 0 ? Ti[(size_t)k * nVp + vi] : (Ti[(size_t)k * nVp + vi] + Ti[(size_t)k * nVp + vj]) / 2.0; };
  double out;
  if (mbits[i] & MB_SHEET) {
    double prev = g_arrhenius(T(0)), avg = 0.0;
    for (int k = 1; k < Z.nZ; k++) { const double cur = g_arrhenius(T(k)); avg = avg + 0.5 * (cur + prev) * Z.dz[k]; prev = cur; }
    out = avg;
  } else {
    out = g_arrhenius((T(0) + UFM_SMT) / 2.0);
  }
  A_mean[i] = out;
}

// dt_D_2D of one staggered vertex (UFEMISM_main_model.f90:752; the kind-less 1E-09 is a single-precision literal)
__device__ __forceinline__ double cfl_dt_D(const double dist2, const double D) { return dist2 / (-6.0 * UFM_PI * (D - (double)1E-09f)); }

// solve_SIA with a per-layer flow factor: f(k) = m_enh_sia * A_flow_Ac(aci,k) * zeta(k)**n, integrated in registers
__global__ void __launch_bounds__(256) k_sia_ac_T(int nAc, int nVp, ZetaConst Z, double m_enh_sia, const int4 *__restrict__ Aci, const double *__restrict__ Ti,
                                                  const unsigned *__restrict__ mbits_Ac, const double *__restrict__ Hi_Ac, const double *__restrict__ hx,
                                                  const double *__restrict__ hy, const double *__restrict__ hp, const double *__restrict__ ho,
                                                  double *__restrict__ D_SIA_Ac, double *__restrict__ Ux, double *__restrict__ Uy, double *__restrict__ Up, double *__restrict__ Uo,
                                                  const double *__restrict__ dist2, unsigned long long *cfl_key, const unsigned char *__restrict__ own, int rank)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < nAc && UFM_OWNED(own, i, rank);
  double D = 0.0, ux = 0.0, uy = 0.0, up = 0.0, uo = 0.0;
  if (in && (mbits_Ac[i] & MB_SHEET)) {
    const double D_uv_3D_cutoff = -1E5;
    const int4 v = Aci[i];
    const double H = Hi_Ac[i], sp = hp[i], so = ho[i];
    const double D_0 = ufm_pow(UFM_ICE_DENSITY * UFM_GRAV * H, UFM_N_FLOW) * (sp * sp + so * so);
    const double twoH = 2.0 * H;
    double I[UFM_MAX_NZ];
    I[Z.nZ - 1] = 0.0;
    double fk1 = m_enh_sia * g_arrhenius((Ti[(size_t)(Z.nZ - 1) * nVp + v.x] + Ti[(size_t)(Z.nZ - 1) * nVp + v.y]) / 2.0) * Z.z3[Z.nZ - 1];
    for (int k = Z.nZ - 2; k >= 0; k--) {   // vertical_integrate, zeta_module.f90:81-84
      const double fk = m_enh_sia * g_arrhenius((Ti[(size_t)k * nVp + v.x] + Ti[(size_t)k * nVp + v.y]) / 2.0) * Z.z3[k];
      I[k] = I[k + 1] - 0.5 * (fk1 + fk) * Z.dz[k + 1];
      fk1 = fk;
    }
    double prev = D_0 * (twoH * I[0]);
    if (prev < D_uv_3D_cutoff) prev = D_uv_3D_cutoff;
    double avg = 0.0;
    for (int k = 1; k < Z.nZ; k++) {
      double cur = D_0 * (twoH * I[k]);
      if (cur < D_uv_3D_cutoff) cur = D_uv_3D_cutoff;
      avg = avg + 0.5 * (cur + prev) * Z.dz[k];
      prev = cur;
    }
    D = H * avg; ux = avg * hx[i]; uy = avg * hy[i]; up = avg * sp; uo = avg * so;
  }
  if (in) { D_SIA_Ac[i] = D; Ux[i] = ux; Uy[i] = uy; Up[i] = up; Uo[i] = uo; }
  block_min_to_key(in ? cfl_dt_D(dist2[i], D) : 1000.0, cfl_key);
}

// ---- solve_SIA on Ac (ice_dynamics_module.f90:240-306); the 3-D diffusivity profile stays in registers.
//      With the benchmark (vertically constant) flow factor the integral of m_enh*A*zeta^n is the same for
//      every column and is evaluated once on the host exactly as vertical_integrate does (zeta_module.f90:58-85). ----
struct SiaConst { int nZ; double I[UFM_MAX_NZ]; double dz[UFM_MAX_NZ]; };
__global__ void __launch_bounds__(256) k_sia_ac(int nAc, SiaConst K, const unsigned *__restrict__ mbits_Ac, const double *__restrict__ Hi_Ac,
                                                const double *__restrict__ hx, const double *__restrict__ hy, const double *__restrict__ hp,
                                                const double *__restrict__ ho, double *__restrict__ D_SIA_Ac, double *__restrict__ Ux,
                                                double *__restrict__ Uy, double *__restrict__ Up, double *__restrict__ Uo,
                                                const double *__restrict__ dist2, unsigned long long *cfl_key, const int *gate,
                                                const unsigned char *__restrict__ own, int rank)
{
  UFM_GATE(gate);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < nAc && UFM_OWNED(own, i, rank);
  double D = 0.0, ux = 0.0, uy = 0.0, up = 0.0, uo = 0.0;
  if (in && (mbits_Ac[i] & MB_SHEET)) {
    const double D_uv_3D_cutoff = -1E5;
    const double H = Hi_Ac[i], sp = hp[i], so = ho[i];
    // (rho g H)**n_flow * (hp**2 + ho**2)**((n_flow-1)/2)   [x**1.0 == x]
    const double D_0 = ufm_pow(UFM_ICE_DENSITY * UFM_GRAV * H, UFM_N_FLOW) * (sp * sp + so * so);
    const double twoH = 2.0 * H;
    double prev = D_0 * (twoH * K.I[0]);
    if (prev < D_uv_3D_cutoff) prev = D_uv_3D_cutoff;
    double avg = 0.0;
    for (int k = 1; k < K.nZ; k++) {
      double cur = D_0 * (twoH * K.I[k]);
      if (cur < D_uv_3D_cutoff) cur = D_uv_3D_cutoff;
      avg = avg + 0.5 * (cur + prev) * K.dz[k];  // vertical_average, zeta_module.f90:53-56
      prev = cur;
    }
    D = H * avg; ux = avg * hx[i]; uy = avg * hy[i]; up = avg * sp; uo = avg * so;
  }
  if (in) { D_SIA_Ac[i] = D; Ux[i] = ux; Uy[i] = uy; Up[i] = up; Uo[i] = uo; }
  // epilogue: this vertex's diffusive critical time step, reduced over the CTA into the cached minimum (see ufm_k_cfl)
  block_min_to_key(in ? cfl_dt_D(dist2[i], D) : 1000.0, cfl_key);
}

// ---- solve_SIA_3D, U_3D / V_3D half (ice_dynamics_module.f90:317-367) + apply_Neumann_boundary_3D
//      (mesh_derivatives_module.f90:538-597).  Layout of the (nV,nZ) arrays: k-major, [k*nVp + v]. ----
template <bool REALISTIC>
__global__ void __launch_bounds__(256) k_sia3d_uv(int nV, int nVp, SiaConst K, ZetaConst Z, double m_enh_sia, const double *__restrict__ Ti,
                                                  const unsigned *__restrict__ mbits, const double *__restrict__ Hi, const double *__restrict__ hx,
                                                  const double *__restrict__ hy, const double *__restrict__ U_SSA, const double *__restrict__ V_SSA,
                                                  double *__restrict__ U3, double *__restrict__ V3, const int *gate)
{
  UFM_GATE(gate);
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const double us = U_SSA[v], vs = V_SSA[v], H = Hi[v];
  const bool plain = (mbits[v] & MB_SHELF) || (H == 0.0);
  double I[UFM_MAX_NZ];
  double D_0 = 0.0, dx = 0.0, dy = 0.0;
  if (!plain) {
    dx = hx[v]; dy = hy[v];
    D_0 = ufm_pow(UFM_ICE_DENSITY * UFM_GRAV * H, UFM_N_FLOW) * (dx * dx + dy * dy);
    if (REALISTIC) {
      I[Z.nZ - 1] = 0.0;
      double fk1 = m_enh_sia * g_arrhenius(Ti[(size_t)(Z.nZ - 1) * nVp + v]) * Z.z3[Z.nZ - 1];
      for (int k = Z.nZ - 2; k >= 0; k--) {
        const double fk = m_enh_sia * g_arrhenius(Ti[(size_t)k * nVp + v]) * Z.z3[k];
        I[k] = I[k + 1] - 0.5 * (fk1 + fk) * Z.dz[k + 1];
        fk1 = fk;
      }
    }
  }
  const double twoH = 2.0 * H;
  for (int k = 0; k < K.nZ; k++) {
    double u = us, w = vs;
    if (!plain) {
      const double D = fmax(D_0 * (twoH * (REALISTIC ? I[k] : K.I[k])), -1E5);
      u = D * dx + us; w = D * dy + vs;
    }
    U3[(size_t)k * nVp + v] = u; V3[(size_t)k * nVp + v] = w;
  }
}
// stage 0: domain-edge vertices except the corners, from their non-edge neighbours; stage 1: the four corners (reference
// vertices 1..4) from all neighbours at their new values
__global__ void __launch_bounds__(256) k_neumann_3d(int stage, int n_slices, int nVp, int nZ, const long long *__restrict__ off,
                                                    const unsigned char *__restrict__ deg, const unsigned char *__restrict__ edge,
                                                    const int *__restrict__ dev2ref, const int *__restrict__ C, double *U3, double *V3, const int *gate = nullptr)
{
  UFM_GATE(gate);
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < n_slices; s += nw) {
    const int v = s * 32 + lane;
    const int n = deg[v];
    if (n == UFM_DEG_PAD || edge[v] == 0) continue;
    const bool corner = dev2ref[v] < 4;
    if (corner != (stage == 1)) continue;
    const long long o = off[s];
    for (int k = 0; k < nZ; k++) {
      double su = 0.0, sv = 0.0;
      int nvals = 0;
      for (int c = 0; c < n; c++) {
        const int j = C[o + (long long)c * 32 + lane];
        if (stage == 0 && edge[j] > 0) continue;
        nvals++;
        su = su + U3[(size_t)k * nVp + j]; sv = sv + V3[(size_t)k * nVp + j];
      }
      U3[(size_t)k * nVp + v] = su / (double)nvals; V3[(size_t)k * nVp + v] = sv / (double)nvals;
    }
  }
}

// ---- map_Ac_to_Aa x3 (mesh_ArakawaC_module.f90:770-791), diagnostic U_SIA, V_SIA, D_SIA ----
struct SiaAaArgs {
  int n_slices;
  const long long *off;
  const unsigned char *deg;
  const int *iAci;
  const double *Ux, *Uy, *D;
  double *U_SIA, *V_SIA, *D_SIA;
  const int *gate;
  const unsigned char *own; int rank;
};
// resident CTAs per SM the kernel is compiled for and connections per batch: measured at 1 M vertices (solve_SIA total, tools/probes_r02/
// r02_probe_minb.sh): (1, 4) 0.100-0.102 ms, (4, 4) 0.093, (5, 4) 0.095, (4, 2) 0.093, (6, 2) 0.090, (1, 8) 0.112 -- latency bound: warps beat batch size
#ifndef UFM_SIA_AA_MINB
#define UFM_SIA_AA_MINB 6
#endif
#ifndef UFM_SIA_AA_CHUNK
#define UFM_SIA_AA_CHUNK 2
#endif
__global__ void __launch_bounds__(256, UFM_SIA_AA_MINB) k_sia_aa(SiaAaArgs a)
{
  UFM_GATE(a.gate);
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < a.n_slices; s += nw) {
    const long long o = a.off[s];
    const int w = (int)((a.off[s + 1] - o) >> 5);
    const int v = s * 32 + lane;
    const int n = a.deg[v];
    if (n == UFM_DEG_PAD || !UFM_OWNED(a.own, v, a.rank)) continue;
    const double dn = (double)n, rn = 1.0 / dn;
    double u = 0.0, vv = 0.0, dd = 0.0;
    // a batch of connections at a time: all index loads, then all gathers, then the sums in connection order (columns past the
    // vertex's own degree read entry 0 and are not added: unconditional loads are what lets the compiler put them in flight together).
    // Measured (1 M vertices, ncu): one connection at a time with a division per term 81-92 us; batches of four 52 us; the three fields
    // split over blockIdx.y (three times the warps, a third of the gathers each) 60 us.
    for (int c0 = 0; c0 < w; c0 += UFM_SIA_AA_CHUNK) {
      int ac[UFM_SIA_AA_CHUNK];
      double x[UFM_SIA_AA_CHUNK], y[UFM_SIA_AA_CHUNK], z[UFM_SIA_AA_CHUNK];
#pragma unroll
      for (int k = 0; k < UFM_SIA_AA_CHUNK; k++) {
        const int c = c0 + k < w ? c0 + k : w - 1;
        const int ia = a.iAci[o + (long long)c * 32 + lane] & 0x7fffffff;
        ac[k] = c0 + k < n ? ia : 0;
      }
#pragma unroll
      for (int k = 0; k < UFM_SIA_AA_CHUNK; k++) { x[k] = a.Ux[ac[k]]; y[k] = a.Uy[ac[k]]; z[k] = a.D[ac[k]]; }
#pragma unroll
      for (int k = 0; k < UFM_SIA_AA_CHUNK; k++) {
        if (c0 + k < n) {
          // term / nC(vi), term by term as the reference does; ufm_div_small = the same bits without 3 n divisions per vertex
          u = u + ufm_div_small(x[k], dn, rn); vv = vv + ufm_div_small(y[k], dn, rn); dd = dd + ufm_div_small(z[k], dn, rn);
        }
      }
    }
    a.U_SIA[v] = u; a.V_SIA[v] = vv; a.D_SIA[v] = dd;
  }
}

// ---- calculate_ice_thickness_change (ice_dynamics_module.f90:31-237) as two gather passes.
//      The limiter only rescales OUT-fluxes by a factor of the source vertex (:115-166), so the result is
//      order independent: pass 1 = factor per vertex, pass 2 = flux sum with the neighbours' factors. ----
struct ThkArgs {
  int n_slices;
  const long long *off;
  const unsigned char *deg, *edge;
  const int *C, *iAci;
  const double *A, *Cw, *UpSIA, *UpSSA, *Hi, *SMB, *BMB;
  const int *noice;
  double dt;
  const double *dt_dev;              // device-driven loop: the time step lives on the device (else NULL)
  const int *gate;
  const unsigned char *own; int rank;
  int clamp_edge;                    // 0 only for 'SSA_icestream': no boundary condition on the thickness (:189-206)
  double *factor, *smb;              // pass 1 out / pass 2 in
  const double *flux;                // edge pass out (EDGE variants): h_upwind * Upar * Cw per Ac vertex
  double *Hi_new, *dHi_dt;           // pass 2 out
};
// Edge pass (UFM_THK_EDGE=1): the flux of a connection is the same number seen from either end (:66-113 computes it once per Ac vertex
// and stores it twice with opposite signs), so one thread per Ac vertex evaluates h_upwind * Upar * Cw -- the first two of the three
// multiplications, dt follows in the vertex passes -- and the vertex passes gather one value per connection instead of four.
__global__ void __launch_bounds__(256) k_thk_flux(int nAc, const int4 *__restrict__ Aci, const double *__restrict__ UpSIA, const double *__restrict__ UpSSA,
                                                  const double *__restrict__ Cw, const double *__restrict__ Hi, double *__restrict__ flux,
                                                  const int *gate, const unsigned char *__restrict__ own_aa, int rank)
{
  UFM_GATE(gate);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nAc) return;
  const int4 v = Aci[i];
  if (own_aa && own_aa[v.x] != rank && own_aa[v.y] != rank) return;   // partitioned: the connections of the vertices this rank updates
  const double Upar = UpSIA[i] + UpSSA[i];
  const double h_up = (Upar > 0.0) ? Hi[v.x] : Hi[v.y];
  flux[i] = h_up * Upar * Cw[i];
}
// Four connections of a vertex at a time: their index loads, then all their gathers, then the arithmetic in connection order.  Columns
// past the vertex's own degree load entry 0 / the vertex itself and are not added: unconditional loads from valid addresses are what
// lets the compiler put a chunk's gathers in flight together (one memory round trip per chunk instead of one per connection).
template <bool EDGE>
__device__ __forceinline__ void thk_chunk(const ThkArgs &a, const long long o, const int lane, const int w, const int n, const int c0, const int v, const double hv,
                                          double en[4], int j[4])
{
  int ia[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int c = c0 + k < w ? c0 + k : w - 1;
    const long long e = o + (long long)c * 32 + lane;
    const int t = a.iAci[e], tj = a.C[e];
    ia[k] = c0 + k < n ? t : 0; j[k] = c0 + k < n ? tj : v;
  }
  if (EDGE) {
    double f[4];
#pragma unroll
    for (int k = 0; k < 4; k++) f[k] = a.flux[ia[k] & 0x7fffffff];
    asm volatile("" ::: "memory");
#pragma unroll
    for (int k = 0; k < 4; k++) { const double dVi = f[k] * a.dt; en[k] = ia[k] < 0 ? -dVi : dVi; }
  } else {
    double u1[4], u2[4], cw[4], hj[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int ac = ia[k] & 0x7fffffff; u1[k] = a.UpSIA[ac]; u2[k] = a.UpSSA[ac]; cw[k] = a.Cw[ac]; hj[k] = a.Hi[j[k]]; }
    asm volatile("" ::: "memory");   // keep the sixteen gathers ahead of the arithmetic (the scheduler otherwise sinks them entry by entry to save registers)
#pragma unroll
    for (int k = 0; k < 4; k++) {   // = thk_entry
      const bool first = ia[k] < 0;
      const double Upar = u1[k] + u2[k];
      const double h_up = (Upar > 0.0) ? (first ? hv : hj[k]) : (first ? hj[k] : hv);
      const double dVi = h_up * Upar * cw[k] * a.dt;
      en[k] = first ? -dVi : dVi;
    }
  }
}
// resident CTAs per SM: thickness update total 0.098-0.099 ms at 1, 0.097 at 5, 0.094 at 6 (same probe)
#ifndef UFM_THK_MINB
#define UFM_THK_MINB 6
#endif
template <int PASS, bool EDGE>
__global__ void __launch_bounds__(256, UFM_THK_MINB) k_thk(ThkArgs a)
{
  UFM_GATE(a.gate);
  if (a.dt_dev) a.dt = *a.dt_dev;
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < a.n_slices; s += nw) {
    const long long o = a.off[s];
    const int w = (int)((a.off[s + 1] - o) >> 5);
    const int v = s * 32 + lane;
    const int n = a.deg[v];
    if (n == UFM_DEG_PAD || !UFM_OWNED(a.own, v, a.rank)) continue;
    const double hv = a.Hi[v], Av = a.A[v];
    if (PASS == 1) {
      double Vi_out = 0.0;
      for (int c0 = 0; c0 < w; c0 += 4) {
        double en[4];
        int j[4];
        thk_chunk<EDGE>(a, o, lane, w, n, c0, v, hv, en, j);
#pragma unroll
        for (int k = 0; k < 4; k++) if (c0 + k < n && !(en[k] > 0.0)) Vi_out = Vi_out - en[k];
      }
      double Vi_SMB = (a.SMB[v] + a.BMB[v]) * Av * a.dt;
      const double Vi_available = Av * hv;
      double rescale_factor = 1.0;
      if (-Vi_SMB >= Vi_available) { Vi_SMB = -Vi_available; rescale_factor = 0.0; }
      if (Vi_out > Vi_available + Vi_SMB) rescale_factor = (Vi_available + Vi_SMB) / Vi_out;
      a.factor[v] = rescale_factor; a.smb[v] = Vi_SMB;
    } else {
      const double fv = a.factor[v];
      double dVi = 0.0;
      for (int c0 = 0; c0 < w; c0 += 4) {
        double en[4], fj[4];
        int j[4];
        thk_chunk<EDGE>(a, o, lane, w, n, c0, v, hv, en, j);
        // the source vertex's out-flux factor, for the fluxes that come in: a second, smaller batch of gathers
#pragma unroll
        for (int k = 0; k < 4; k++) fj[k] = (c0 + k < n && en[k] > 0.0) ? a.factor[j[k]] : 1.0;
        asm volatile("" ::: "memory");
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (c0 + k < n) {
            double e = en[k];
            if (e < 0.0) { if (fv < 1.0) e = e * fv; }
            else if (e > 0.0) { if (fj[k] < 1.0) e = -((-e) * fj[k]); }
            dVi = dVi + e;
          }
        }
      }
      double dh = (dVi + a.smb[v]) / (Av * a.dt);
      if (a.dt == 0.0) dh = 0.0;
      double hn = hv + (dh * a.dt);
      if (a.clamp_edge && a.edge[v] > 0) hn = 0.0;
      if (a.noice[v] == 1) hn = 0.0;
      a.dHi_dt[v] = dh; a.Hi_new[v] = hn;
    }
  }
}

// ---- conservative remapping, application only (mesh_mapping_module.f90:3964-3983, 4010-4043) ----
// gradients of an Aa field on the old mesh (get_mesh_derivatives), written in REFERENCE order next to the field itself
struct StashArgs {
  int n_slices;
  const long long *off;
  const unsigned char *deg;
  const int *C, *dev2ref;
  const double *Nx, *Ny, *Nx0, *Ny0, *f;
  double *d, *ddx, *ddy;
};
__global__ void __launch_bounds__(256) k_remap_stash(StashArgs a)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < a.n_slices; s += nw) {
    const long long o = a.off[s];
    const int v = s * 32 + lane;
    const int n = a.deg[v];
    if (n == UFM_DEG_PAD) continue;
    const double fv = a.f[v];
    double gx = a.Nx0[v] * fv, gy = a.Ny0[v] * fv;
    for (int c = 0; c < n; c++) {
      const long long e = o + (long long)c * 32 + lane;
      const double fj = a.f[a.C[e]];
      gx = gx + a.Nx[e] * fj; gy = gy + a.Ny[e] * fj;
    }
    const int r = a.dev2ref[v];
    a.d[r] = fv; a.ddx[r] = gx; a.ddy[r] = gy;
  }
}
// one thread per destination vertex, entries accumulated in list order exactly as the reference loop does
__global__ void __launch_bounds__(256) k_remap_apply(int nV_dst, int order, const int *__restrict__ vli1, const int *__restrict__ vli2, const int *__restrict__ vi,
                                                     const double *__restrict__ w0, const double *__restrict__ w1x, const double *__restrict__ w1y,
                                                     const double *__restrict__ d, const double *__restrict__ ddx, const double *__restrict__ ddy,
                                                     const int *__restrict__ ref2dev, double *out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nV_dst) return;
  double acc = 0.0;
  for (int l = vli1[i]; l <= vli2[i]; l++) {
    const int sv = vi[l - 1] - 1;
    if (order == 1) acc = acc + (d[sv] * w0[l - 1]);
    else acc = acc + (d[sv] * w0[l - 1]) + (ddx[sv] * w1x[l - 1]) + (ddy[sv] * w1y[l - 1]);
  }
  out[ref2dev[i]] = acc;
}

// ---- critical time steps (UFEMISM_main_model.f90:747-768) ----
// `which`: bit 0 dt_D_2D, bit 1 dt_V_2D_SSA, bit 2 dt_V_3D_SIA -- only the minima whose cached value is stale are recomputed
__global__ void __launch_bounds__(256) k_cfl(int which, int nV, int nVp, int nAc, int nZ, const double *__restrict__ dist2, const double *__restrict__ D_SIA_Ac,
                                             const double *__restrict__ U, const double *__restrict__ V, const double *__restrict__ rmin,
                                             const double *__restrict__ sqrtApi, const double *__restrict__ U3, const double *__restrict__ V3, unsigned long long *keys,
                                             const int *gate, const unsigned char *__restrict__ own_aa, const unsigned char *__restrict__ own_ac, int rank)
{
  UFM_GATE(gate);
  double mD = 1000.0, mS = 1000.0, m3 = 1000.0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ((which & 1) && i < nAc && UFM_OWNED(own_ac, i, rank)) mD = cfl_dt_D(dist2[i], D_SIA_Ac[i]);
  if (i < nV && UFM_OWNED(own_aa, i, rank)) {
    // the reference takes dist / (|U| + |V|) at both ends of every connection and SQRT(A/pi) / (|U| + |V|) per vertex (:754-762); IEEE
    // division is monotone in its numerator, so per vertex that is the smallest numerator (rmin, fixed per mesh) over the same sum
    if (which & 2) mS = rmin[i] / (fabs(U[i]) + fabs(V[i]));
    if (which & 4) {
      const double r = sqrtApi[i];
      for (int k = 0; k < nZ; k++) m3 = fmin(r / (fabs(U3[(size_t)k * nVp + i]) + fabs(V3[(size_t)k * nVp + i])), m3);
    }
  }
  if (which & 1) block_min_to_key(mD, keys + 0);
  if (which & 2) { __syncthreads(); block_min_to_key(mS, keys + 1); }
  if (which & 4) { __syncthreads(); block_min_to_key(m3, keys + 2); }
}
__global__ void k_cfl_key_reset(int which, unsigned long long *keys)
{
  if (threadIdx.x < 3 && ((which >> threadIdx.x) & 1)) keys[threadIdx.x] = ord_key(1000.0);
}

// ---- permuted copies between reference order (device staging) and device order ----
// dev element (row r) lives at dev[r*stride + comp]; ref element i at ref[i]
__global__ void k_perm_d(int n, const int *__restrict__ r2d, double *dev, int stride, int comp, double *ref, int to_device)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t p = (size_t)r2d[i] * stride + comp;
  if (to_device) dev[p] = ref[i]; else ref[i] = dev[p];
}
__global__ void k_perm_i(int n, const int *__restrict__ r2d, int *dev, int *ref, int to_device)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (to_device) dev[r2d[i]] = ref[i]; else ref[i] = dev[r2d[i]];
}
// mode 0: (bits & arg) != 0 ; mode 1: the ice%mask code
__global__ void k_perm_mask(int n, const int *__restrict__ r2d, const unsigned *__restrict__ bits, unsigned arg, int mode, int *ref)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned b = bits[r2d[i]];
  ref[i] = mode ? (int)((b >> MB_CODE_SHIFT) & 0xF) : ((b & arg) ? 1 : 0);
}
__global__ void k_perm_3d(int n, int nZ, int nVp, const int *__restrict__ r2d, double *dev, double *ref, int to_device)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < nZ; k++) {
    size_t p = (size_t)k * nVp + r2d[i], q = (size_t)k * n + i;
    if (to_device) dev[p] = ref[q]; else ref[q] = dev[p];
  }
}

// =============================================================================================
// host launchers
// =============================================================================================
int ufm_perm_double(ufm_handle *h, int n, const int *r2d, double *dev, int stride, int comp, double *ref_dev, int to_device)
{
  k_perm_d<<<grid_for(n, 256), 256, 0, h->stream>>>(n, r2d, dev, stride, comp, ref_dev, to_device);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_perm_d");
}
int ufm_perm_int(ufm_handle *h, int n, const int *r2d, int *dev, int *ref_dev, int to_device)
{
  k_perm_i<<<grid_for(n, 256), 256, 0, h->stream>>>(n, r2d, dev, ref_dev, to_device);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_perm_i");
}
int ufm_perm_mask(ufm_handle *h, int n, const int *r2d, const unsigned *bits, unsigned arg, int mode, int *ref_dev)
{
  k_perm_mask<<<grid_for(n, 256), 256, 0, h->stream>>>(n, r2d, bits, arg, mode, ref_dev);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_perm_mask");
}
int ufm_perm_3d(ufm_handle *h, int n, int nZ, int nVp, const int *r2d, double *dev, double *ref_dev, int to_device)
{
  k_perm_3d<<<grid_for(n, 256), 256, 0, h->stream>>>(n, nZ, nVp, r2d, dev, ref_dev, to_device);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_perm_3d");
}

int ufm_k_geom(ufm_handle *h, double time)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  // ice_physical_properties, benchmark branches (general_ice_model_data_module.f90:321-368)
  const int b = h->P.benchmark;
  double A_flow = 1.0E-16;
  if (b == UFM_BM_MISMIP_MOD || b == UFM_BM_SSA_ICESTREAM) {
    if (time < 25000.0) A_flow = 1.0E-16; else if (time < 50000.0) A_flow = 1.0E-17; else if (time < 75000.0) A_flow = 1.0E-16;
  }
  s.A_flow_const = A_flow;
  k_geom_aa1<<<grid_for(m.nV, 256), 256, 0, h->stream>>>(m.nV, s.Hi, s.Hb, s.SL, s.dHb_dt, s.dHi_dt, s.Hs, s.dHs_dt, s.mbits, h->gate[0], m.part_step ? m.act_aa1 : nullptr);
  GeomAa2Args a2;
  a2.n_slices = m.aa.n_slices; a2.off = m.aa.off; a2.deg = m.aa.deg; a2.C = m.aa_C; a2.Nx = m.aa_Nx; a2.Ny = m.aa_Ny; a2.Nx0 = m.aa_Nx0; a2.Ny0 = m.aa_Ny0;
  a2.Hi = s.Hi; a2.Hs = s.Hs; a2.mbits = s.mbits; a2.dHi_dx = s.dHi_dx; a2.dHi_dy = s.dHi_dy; a2.dHs_dx = s.dHs_dx; a2.dHs_dy = s.dHs_dy;
  a2.sx = s.dHs_dx_shelf; a2.sy = s.dHs_dy_shelf; a2.gate = h->gate[0]; a2.own = m.part_step ? m.own_aa : nullptr; a2.rank = m.rank;
  int g2 = grid_for((long long)m.aa.n_slices * 32, 256);
  k_geom_aa2<<<g2, 256, 0, h->stream>>>(a2);
  GeomAcArgs ac;
  ac.nAc = m.nAc; ac.Aci = m.ac_Aci; ac.Np = m.ac_Np;
  for (int k = 0; k < 4; k++) { ac.Nx[k] = m.ac_Nx[k]; ac.Ny[k] = m.ac_Ny[k]; ac.No[k] = m.ac_No[k]; ac.dHi[k] = s.dHi_Ac[k]; ac.dHb[k] = s.dHb_Ac[k]; ac.dHs[k] = s.dHs_Ac[k]; ac.dSL[k] = s.dSL_Ac[k]; }
  ac.Hi = s.Hi; ac.Hb = s.Hb; ac.SL = s.SL; ac.Hs = s.Hs; ac.mbits = s.mbits;
  ac.Hi_Ac = s.Hi_Ac; ac.Hb_Ac = s.Hb_Ac; ac.SL_Ac = s.SL_Ac; ac.Hs_Ac = s.Hs_Ac; ac.sx = s.dHs_dx_shelf_Ac; ac.sy = s.dHs_dy_shelf_Ac; ac.mbits_Ac = s.mbits_Ac; ac.gate = h->gate[0]; ac.own = m.part_step ? m.own_ac : nullptr; ac.rank = m.rank;
  k_geom_ac<<<grid_for(m.nAc, 256), 256, 0, h->stream>>>(ac);
  h->cnt.kernel_launches += 3;
  if (s.realistic_A) {
    ZetaConst Z;
    Z.nZ = h->P.nZ; Z.dz[0] = 0.0;
    for (int k = 1; k < Z.nZ; k++) Z.dz[k] = h->P.zeta[k] - h->P.zeta[k - 1];
    for (int k = 0; k < Z.nZ; k++) Z.z3[k] = h->zeta3[k];
    k_flow_mean<0><<<grid_for(m.nV, 256), 256, 0, h->stream>>>(m.nV, m.nVp, Z, m.ac_Aci, s.Ti, s.mbits, s.A_mean);
    k_flow_mean<1><<<grid_for(m.nAc, 256), 256, 0, h->stream>>>(m.nAc, m.nVp, Z, m.ac_Aci, s.Ti, s.mbits_Ac, s.A_mean_Ac);
    h->cnt.kernel_launches += 2;
  }
  return ufm_cuda_check(cudaGetLastError(), "k_geom");
}

int ufm_k_sia(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  SiaConst K;
  K.nZ = h->P.nZ;
  // vertical_integrate( C%m_enh_sia * A_flow_Ac(aci,:) * C%zeta**n_flow ), zeta_module.f90:81-84
  double f[UFM_MAX_NZ];
  for (int k = 0; k < K.nZ; k++) f[k] = h->P.m_enh_sia * s.A_flow_const * h->zeta3[k];
  K.I[K.nZ - 1] = 0.0;
  for (int k = K.nZ - 1; k >= 1; k--) K.I[k - 1] = K.I[k] - 0.5 * (f[k] + f[k - 1]) * (h->P.zeta[k] - h->P.zeta[k - 1]);
  K.dz[0] = 0.0;
  for (int k = 1; k < K.nZ; k++) K.dz[k] = h->P.zeta[k] - h->P.zeta[k - 1];
  // the diffusive critical time step is re-reduced by the kernel below (device-driven loop: k_step_control resets the key, and only
  // when the solve is due)
  if (!h->gate[1]) { int rc_ = ufm_cfl_key_reset(h, 1); if (rc_) return rc_; }
  if (s.realistic_A) {
    ZetaConst Z;
    Z.nZ = h->P.nZ; Z.dz[0] = 0.0;
    for (int k = 1; k < Z.nZ; k++) Z.dz[k] = h->P.zeta[k] - h->P.zeta[k - 1];
    for (int k = 0; k < Z.nZ; k++) Z.z3[k] = h->zeta3[k];
    k_sia_ac_T<<<grid_for(m.nAc, 256), 256, 0, h->stream>>>(m.nAc, m.nVp, Z, h->P.m_enh_sia, m.ac_Aci, s.Ti, s.mbits_Ac, s.Hi_Ac, s.dHs_Ac[0], s.dHs_Ac[1],
                                                            s.dHs_Ac[2], s.dHs_Ac[3], s.D_SIA_Ac, s.U_SIA_Ac[0], s.U_SIA_Ac[1], s.U_SIA_Ac[2], s.U_SIA_Ac[3],
                                                            m.ac_dist2, s.ctrl + CTRL_CFL_KEYS + 0, m.part_step ? m.own_ac : nullptr, m.rank);
  } else
  k_sia_ac<<<grid_for(m.nAc, 256), 256, 0, h->stream>>>(m.nAc, K, s.mbits_Ac, s.Hi_Ac, s.dHs_Ac[0], s.dHs_Ac[1], s.dHs_Ac[2], s.dHs_Ac[3],
                                                       s.D_SIA_Ac, s.U_SIA_Ac[0], s.U_SIA_Ac[1], s.U_SIA_Ac[2], s.U_SIA_Ac[3],
                                                       m.ac_dist2, s.ctrl + CTRL_CFL_KEYS + 0, h->gate[1], m.part_step ? m.own_ac : nullptr, m.rank);
  h->cfl_ok[0] = true;
  SiaAaArgs a;
  a.n_slices = m.aa.n_slices; a.off = m.aa.off; a.deg = m.aa.deg; a.iAci = m.aa_iAci; a.Ux = s.U_SIA_Ac[0]; a.Uy = s.U_SIA_Ac[1]; a.D = s.D_SIA_Ac;
  a.U_SIA = s.U_SIA; a.V_SIA = s.V_SIA; a.D_SIA = s.D_SIA; a.gate = h->gate[1]; a.own = m.part_step ? m.own_aa : nullptr; a.rank = m.rank;
  if (m.part_step) {   // the diagnostic map to the vertices and the next thickness update read staggered vertices of the neighbour strips
    double *arr[4] = {s.U_SIA_Ac[2], s.U_SIA_Ac[0], s.U_SIA_Ac[1], s.D_SIA_Ac};
    int rc_x = ufm_halo_exchange(h, 1, 4, arr);
    if (rc_x) return rc_x;
  }
  k_sia_aa<<<grid_for((long long)m.aa.n_slices * 32, 256), 256, 0, h->stream>>>(a);
  h->cnt.kernel_launches += 2;
  return ufm_cuda_check(cudaGetLastError(), "k_sia");
}

int ufm_k_sia3d(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  SiaConst K;
  ZetaConst Z;
  K.nZ = Z.nZ = h->P.nZ;
  double f[UFM_MAX_NZ];
  for (int k = 0; k < K.nZ; k++) f[k] = h->P.m_enh_sia * s.A_flow_const * h->zeta3[k];
  K.I[K.nZ - 1] = 0.0;
  for (int k = K.nZ - 1; k >= 1; k--) K.I[k - 1] = K.I[k] - 0.5 * (f[k] + f[k - 1]) * (h->P.zeta[k] - h->P.zeta[k - 1]);
  K.dz[0] = Z.dz[0] = 0.0;
  for (int k = 1; k < K.nZ; k++) K.dz[k] = Z.dz[k] = h->P.zeta[k] - h->P.zeta[k - 1];
  for (int k = 0; k < K.nZ; k++) Z.z3[k] = h->zeta3[k];
  if (s.realistic_A)
    k_sia3d_uv<true><<<grid_for(m.nV, 256), 256, 0, h->stream>>>(m.nV, m.nVp, K, Z, h->P.m_enh_sia, s.Ti, s.mbits, s.Hi, s.dHs_dx, s.dHs_dy, s.U_SSA, s.V_SSA, s.U_3D, s.V_3D, h->gate[3]);
  else
    k_sia3d_uv<false><<<grid_for(m.nV, 256), 256, 0, h->stream>>>(m.nV, m.nVp, K, Z, h->P.m_enh_sia, s.Ti, s.mbits, s.Hi, s.dHs_dx, s.dHs_dy, s.U_SSA, s.V_SSA, s.U_3D, s.V_3D, h->gate[3]);
  int g = grid_for((long long)m.aa.n_slices * 32, 256);
  k_neumann_3d<<<g, 256, 0, h->stream>>>(0, m.aa.n_slices, m.nVp, h->P.nZ, m.aa.off, m.aa.deg, m.aa_edge, m.aa_dev2ref, m.aa_C, s.U_3D, s.V_3D, h->gate[3]);
  k_neumann_3d<<<g, 256, 0, h->stream>>>(1, m.aa.n_slices, m.nVp, h->P.nZ, m.aa.off, m.aa.deg, m.aa_edge, m.aa_dev2ref, m.aa_C, s.U_3D, s.V_3D, h->gate[3]);
  h->cnt.kernel_launches += 3;
  h->cfl_ok[2] = false;   // reduced again from the final (U,V)_3D by the next ufm_cfl (once per thermodynamics step)
  return ufm_cuda_check(cudaGetLastError(), "k_sia3d");
}

// ---- run_SMB_model, benchmark branches (src/SMB_module.f90:55-97): EISMINT_SMB (:172-238), Bueler_solution_MB (:240-283),
//      the constants of Halfar / MISMIP_mod, mesh_generation_test.  Everything that does not depend on the vertex is
//      evaluated on the host with the host libm, as the reference does once per call. ----
struct SmbArgs { int mode; double E, S_b, M_max, H0f1, f2, R0, lam_tp_spy, value; };
__global__ void __launch_bounds__(256) k_smb_benchmark(int nV, SmbArgs a, const double2 *__restrict__ xy, double *__restrict__ SMB_year, const int *gate,
                                                       const unsigned char *__restrict__ own, int rank)
{
  UFM_GATE(gate);
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV || !UFM_OWNED(own, v, rank)) return;
  const double2 p = xy[v];
  double out;
  if (a.mode == 0) out = a.value;
  else if (a.mode == 1) out = fmin(a.M_max, a.S_b * (a.E - ufm_norm2_2(p.x, p.y)));   // dist = NORM2( mesh%V( vi,:))
  else if (a.mode == 2) {
    const double f3 = sqrt((p.x * p.x) + (p.y * p.y)) / a.R0;                     // x**2._dp: pow( x, 2) is exactly x*x
    const double f4 = fmax(0.0, 1.0 - ufm_pow(a.f2 * f3, 4.0 / 3.0));
    const double H = a.H0f1 * ufm_pow(f4, 3.0 / 7.0);
    out = a.lam_tp_spy * H * UFM_SEC_PER_YEAR;
  } else {
    const double R = ufm_norm2_2(p.x, p.y);
    out = R < 250000.0 ? 0.3 : fmax(-2.0, 0.3 - (R - 250000.0) / 200000.0);
  }
  SMB_year[v] = out;
}
int ufm_k_smb_benchmark(ufm_handle *h, double time, double H0, double R0, double lambda)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  const int b = h->P.benchmark;
  SmbArgs a;
  memset(&a, 0, sizeof(a));
  if (b >= UFM_BM_EISMINT_1 && b <= UFM_BM_EISMINT_6) {
    a.mode = 1; a.E = 450000.0; a.S_b = 0.01 / 1000.0; a.M_max = 0.5;
    if (b == UFM_BM_EISMINT_2) { if (!(time < 0.0)) a.E = 450000.0 + 100000.0 * sin(2.0 * UFM_PI * time / 20000.0); }
    else if (b == UFM_BM_EISMINT_3) { if (!(time < 0.0)) a.E = 450000.0 + 100000.0 * sin(2.0 * UFM_PI * time / 40000.0); }
    else if (b == UFM_BM_EISMINT_4) { a.M_max = 0.3; a.E = 999000.0; }
    else if (b == UFM_BM_EISMINT_5) { a.E = 999000.0; a.M_max = time < 0.0 ? 0.3 : 0.3 + 0.2 * sin(2.0 * UFM_PI * time / 20000.0); }
    else if (b == UFM_BM_EISMINT_6) { a.E = 999000.0; a.M_max = time < 0.0 ? 0.3 : 0.3 + 0.2 * sin(2.0 * UFM_PI * time / 40000.0); }
  } else if (b == UFM_BM_HALFAR) { a.mode = 0; a.value = 0.0; }
  else if (b == UFM_BM_MISMIP_MOD || b == UFM_BM_SSA_ICESTREAM) { a.mode = 0; a.value = 0.3; }
  else if (b == UFM_BM_MESH_GENERATION_TEST) a.mode = 3;
  else if (b == UFM_BM_BUELER) {
    const double A_flow = 1E-16, rho = 910.0, g = 9.81, n = 3.0;
    const double alpha = (2.0 - (n + 1.0) * lambda) / ((5.0 * n) + 3.0);
    const double beta = (1.0 + ((2.0 * n) + 1.0) * lambda) / ((5.0 * n) + 3.0);
    const double Gamma = 2.0 / 5.0 * (A_flow / UFM_SEC_PER_YEAR) * pow(rho * g, n);
    double f1 = ((2.0 * n) + 1) / (n + 1.0);
    double f2 = (pow(R0, n + 1.0)) / (pow(H0, (2.0 * n) + 1.0));
    const double t0 = (beta / Gamma) * (pow(f1, n)) * f2;
    const double tp = time * UFM_SEC_PER_YEAR;
    f1 = pow(tp / t0, -alpha);
    f2 = pow(tp / t0, -beta);
    a.mode = 2; a.H0f1 = H0 * f1; a.f2 = f2; a.R0 = R0; a.lam_tp_spy = lambda / tp;
  } else return ufm_set_error(-4, "no closed-form SMB for benchmark %d: SMB_year comes from the host", b);
  k_smb_benchmark<<<grid_for(m.nV, 256), 256, 0, h->stream>>>(m.nV, a, m.aa_xy, s.SMB_year, h->gate[2], m.part_step ? m.own_aa : nullptr, m.rank);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_smb_benchmark");
}

// apply_Neumann_boundary_3D on two (nV,nZ) fields at once (or one, passed twice)
int ufm_k_neumann3d_pair(ufm_handle *h, double *A3, double *B3)
{
  DevMesh &m = h->mesh;
  const int g = grid_for((long long)m.aa.n_slices * 32, 256);
  k_neumann_3d<<<g, 256, 0, h->stream>>>(0, m.aa.n_slices, m.nVp, h->P.nZ, m.aa.off, m.aa.deg, m.aa_edge, m.aa_dev2ref, m.aa_C, A3, B3);
  k_neumann_3d<<<g, 256, 0, h->stream>>>(1, m.aa.n_slices, m.nVp, h->P.nZ, m.aa.off, m.aa.deg, m.aa_edge, m.aa_dev2ref, m.aa_C, A3, B3);
  h->cnt.kernel_launches += 2;
  return ufm_cuda_check(cudaGetLastError(), "k_neumann_3d");
}

int ufm_k_remap_stash(ufm_handle *h, int slot, double *field_dev)
{
  DevMesh &m = h->mesh;
  ufm_handle::Stash &st = h->stash[slot];
  if (st.d) { cudaFree(st.d); cudaFree(st.ddx); cudaFree(st.ddy); st.d = st.ddx = st.ddy = nullptr; }
  st.n = m.nV;
  UFM_CUDA(cudaMalloc((void **)&st.d, sizeof(double) * (size_t)m.nV));
  UFM_CUDA(cudaMalloc((void **)&st.ddx, sizeof(double) * (size_t)m.nV));
  UFM_CUDA(cudaMalloc((void **)&st.ddy, sizeof(double) * (size_t)m.nV));
  StashArgs a;
  a.n_slices = m.aa.n_slices; a.off = m.aa.off; a.deg = m.aa.deg; a.C = m.aa_C; a.dev2ref = m.aa_dev2ref; a.Nx = m.aa_Nx; a.Ny = m.aa_Ny;
  a.Nx0 = m.aa_Nx0; a.Ny0 = m.aa_Ny0; a.f = field_dev; a.d = st.d; a.ddx = st.ddx; a.ddy = st.ddy;
  k_remap_stash<<<grid_for((long long)m.aa.n_slices * 32, 256), 256, 0, h->stream>>>(a);
  h->cnt.kernel_launches++;
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  return ufm_cuda_check(cudaGetLastError(), "k_remap_stash");
}

int ufm_k_remap_apply(ufm_handle *h, int slot, const ufm_remap_cons *map, int order, double *field_dev)
{
  DevMesh &m = h->mesh;
  ufm_handle::Stash &st = h->stash[slot];
  int *vli1 = nullptr, *vli2 = nullptr, *vi = nullptr;
  double *w0 = nullptr, *w1x = nullptr, *w1y = nullptr;
  const size_t nd = (size_t)map->nV_dst, nt = (size_t)(map->n_tot > 0 ? map->n_tot : 1);
  UFM_CUDA(cudaMalloc((void **)&vli1, 4 * nd)); UFM_CUDA(cudaMalloc((void **)&vli2, 4 * nd)); UFM_CUDA(cudaMalloc((void **)&vi, 4 * nt));
  UFM_CUDA(cudaMalloc((void **)&w0, 8 * nt)); UFM_CUDA(cudaMalloc((void **)&w1x, 8 * nt)); UFM_CUDA(cudaMalloc((void **)&w1y, 8 * nt));
  UFM_CUDA(cudaMemcpyAsync(vli1, map->vli1, 4 * nd, cudaMemcpyHostToDevice, h->stream));
  UFM_CUDA(cudaMemcpyAsync(vli2, map->vli2, 4 * nd, cudaMemcpyHostToDevice, h->stream));
  UFM_CUDA(cudaMemcpyAsync(vi, map->vi, 4 * (size_t)map->n_tot, cudaMemcpyHostToDevice, h->stream));
  UFM_CUDA(cudaMemcpyAsync(w0, map->w0, 8 * (size_t)map->n_tot, cudaMemcpyHostToDevice, h->stream));
  if (order == 2) {
    UFM_CUDA(cudaMemcpyAsync(w1x, map->w1x, 8 * (size_t)map->n_tot, cudaMemcpyHostToDevice, h->stream));
    UFM_CUDA(cudaMemcpyAsync(w1y, map->w1y, 8 * (size_t)map->n_tot, cudaMemcpyHostToDevice, h->stream));
  }
  k_remap_apply<<<grid_for(m.nV, 256), 256, 0, h->stream>>>(m.nV, order, vli1, vli2, vi, w0, w1x, w1y, st.d, st.ddx, st.ddy, m.aa_ref2dev, field_dev);
  h->cnt.kernel_launches++;
  h->cnt.h2d_bytes += 8.0 * nd + (4.0 + 8.0 * (order == 2 ? 3 : 1)) * (double)map->n_tot;
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(vli1); cudaFree(vli2); cudaFree(vi); cudaFree(w0); cudaFree(w1x); cudaFree(w1y);
  return ufm_cuda_check(cudaGetLastError(), "k_remap_apply");
}

int ufm_k_thickness(ufm_handle *h, double dt)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  ThkArgs a;
  a.n_slices = m.aa.n_slices; a.off = m.aa.off; a.deg = m.aa.deg; a.edge = m.aa_edge; a.C = m.aa_C; a.iAci = m.aa_iAci; a.A = m.aa_A; a.Cw = m.ac_Cw;
  a.UpSIA = s.U_SIA_Ac[2]; a.UpSSA = s.U_SSA_Ac[2]; a.Hi = s.Hi; a.SMB = s.SMB_year; a.BMB = s.BMB; a.noice = s.mask_noice; a.dt = dt;
  a.factor = s.thk_factor; a.smb = s.thk_smb; a.Hi_new = s.Hi_alt; a.dHi_dt = s.dHi_dt;
  a.dt_dev = h->dt_dev; a.gate = h->gate[0]; a.own = m.part_step ? m.own_aa : nullptr; a.rank = m.rank;
  a.clamp_edge = h->P.benchmark != UFM_BM_SSA_ICESTREAM;
  int g = grid_for((long long)m.aa.n_slices * 32, 256);
  const char *edge_env = getenv("UFM_THK_EDGE");      // read per call: A/B runs and tests switch it inside one process
  const bool edge = !edge_env || atoi(edge_env) != 0;
  a.flux = s.thk_flux;
  if (edge) {
    k_thk_flux<<<grid_for(m.nAc, 256), 256, 0, h->stream>>>(m.nAc, m.ac_Aci, a.UpSIA, a.UpSSA, a.Cw, a.Hi, s.thk_flux, a.gate, a.own, a.rank);
    h->cnt.kernel_launches++;
    k_thk<1, true><<<g, 256, 0, h->stream>>>(a);
  } else k_thk<1, false><<<g, 256, 0, h->stream>>>(a);
  if (m.part_step) {   // the out-flux factors of the neighbour strips' vertices scale the fluxes that come in from them
    double *arr[1] = {s.thk_factor};
    int rc_x = ufm_halo_exchange(h, 0, 1, arr);
    if (rc_x) return rc_x;
  }
  if (edge) k_thk<2, true><<<g, 256, 0, h->stream>>>(a);
  else k_thk<2, false><<<g, 256, 0, h->stream>>>(a);
  h->cnt.kernel_launches += 2;
  // Hi_prev = Hi ; Hi = new  (pointer swap: the old buffer IS Hi_prev)
  double *t = s.Hi; s.Hi = s.Hi_alt; s.Hi_alt = t;
  if (m.part_step) {   // the new thickness of the vertices the neighbour strips read (geometry, next thickness update)
    double *arr[1] = {s.Hi};
    int rc_x = ufm_halo_exchange(h, 0, 1, arr);
    if (rc_x) return rc_x;
  }
  return ufm_cuda_check(cudaGetLastError(), "k_thk");
}

int ufm_cfl_key_reset(ufm_handle *h, int which)
{
  k_cfl_key_reset<<<1, 32, 0, h->stream>>>(which, h->st.ctrl + CTRL_CFL_KEYS);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_cfl_key_reset");
}

// the 3-D critical time step alone, re-reduced from (U,V)_3D into its key; no host synchronisation (device-driven loop, gate: thermodynamics due)
int ufm_k_cfl3d_enqueue(ufm_handle *h)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  k_cfl<<<grid_for(m.nV, 256), 256, 0, h->stream>>>(4, m.nV, m.nVp, m.nAc, h->P.nZ, m.ac_dist2, s.D_SIA_Ac, s.U_SSA, s.V_SSA,
                                                    m.aa_rmin, m.aa_sqrtApi, s.U_3D, s.V_3D, s.ctrl + CTRL_CFL_KEYS, h->gate[3], nullptr, nullptr, 0);
  h->cnt.kernel_launches++;
  return ufm_cuda_check(cudaGetLastError(), "k_cfl (3-D)");
}

int ufm_k_cfl(ufm_handle *h, double out3[3])
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  unsigned long long *keys = s.ctrl + CTRL_CFL_KEYS;
  int which = 0;
  for (int k = 0; k < 3; k++) if (!h->cfl_ok[k]) which |= 1 << k;
  if (getenv("UFM_CFL_RECOMPUTE")) which = 7;   // A/B and tests: ignore the cached minima
  if (which) {
    int rc = ufm_cfl_key_reset(h, which);
    if (rc) return rc;
    const int n = (which & 1) ? (m.nAc > m.nV ? m.nAc : m.nV) : m.nV;   // thread i: staggered vertex i (bit 0) and vertex i (bits 1, 2)
    k_cfl<<<grid_for(n, 256), 256, 0, h->stream>>>(which, m.nV, m.nVp, m.nAc, h->P.nZ, m.ac_dist2, s.D_SIA_Ac, s.U_SSA, s.V_SSA,
                                                   m.aa_rmin, m.aa_sqrtApi, s.U_3D, s.V_3D, keys, nullptr, m.part_step ? m.own_aa : nullptr,
                                                   m.part_step ? m.own_ac : nullptr, m.rank);
    h->cnt.kernel_launches++;
    h->cfl_ok[0] = h->cfl_ok[1] = h->cfl_ok[2] = true;
  }
  unsigned long long *res = (unsigned long long *)(s.scal_h + 16);
  if (m.part_step) {
    // every rank holds the minima over ITS elements (they stay cached as such): minimum over the ranks on a copy
    // (the MPI_ALLREDUCE MIN of UFEMISM_main_model.f90:764-766)
    unsigned long long *tmp = s.ctrl + CTRL_CFL_TMP;
    UFM_CUDA(cudaMemcpyAsync(tmp, keys, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream));
    int rc = ufm_peer_allreduce(h, tmp, 3, 0);
    if (rc) return rc;
    keys = tmp;
  }
  UFM_CUDA(cudaMemcpyAsync(res, keys, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  const double dt_correction_factor = 0.9;
  for (int k = 0; k < 3; k++) out3[k] = ord_unkey(res[k]) * dt_correction_factor;
  return 0;
}
