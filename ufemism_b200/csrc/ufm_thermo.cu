// ufm_thermo.cu -- thermodynamics on the device: the vertical velocity half of solve_SIA_3D
// (src/ice_dynamics_module.f90:369-405) and update_ice_temperature (src/thermodynamics_module.f90:23-202) with everything
// it calls: bottom_frictional_heating (:281-311), calculate_zeta_derivatives (src/zeta_module.f90:86-112, folded into the
// column kernel: the (nV,nZ) Jacobians are never materialised), get_upwind_derivative_vertex_3D
// (src/mesh_derivatives_module.f90:435-483), tridiagonal_solve = LAPACK DGTSV (:313-358) and the Robin-solution safety
// net (:204-279).  Ki, Cpi and Ti_pmp of ice_physical_properties (src/general_ice_model_data_module.f90:336-343,388-404)
// are recomputed per layer from Ti / Hi, which have not changed since update_general_ice_model_data stored them
// (run_model order, src/UFEMISM_main_model.f90:115-169).
//
// One thread owns one ice column (15 layers in registers / local memory), one warp one 32-row slice of the Aa sliced ELL;
// (nV,nZ) arrays are k-major so a warp reads 32 consecutive doubles per layer.  Every expression keeps the reference's
// evaluation order and the file is compiled with -fmad=false: with benchmark ice properties and no sliding the result is
// bit-identical to the CPU restatement; pow (frictional heating), exp (Ki) and erf (Robin) differ from glibc by <= 2 ulp.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ufm_internal.cuh"
#include "ufm_pow.cuh"

#ifndef UFM_HEAT_MINB_DEFAULT
#define UFM_HEAT_MINB_DEFAULT 6
#endif
#define UFM_T0 273.16
#define UFM_CC 8.7E-04

struct ThermoConst {
  int nZ, realistic;
  double zeta[UFM_MAX_NZ];
  double a_zeta[UFM_MAX_NZ], b_zeta[UFM_MAX_NZ], c_zeta[UFM_MAX_NZ], a_zz[UFM_MAX_NZ], b_zz[UFM_MAX_NZ], c_zz[UFM_MAX_NZ];  // index k-1 for k = 2..nZ-1
  double dt_thermo, ten_q;   // u_threshold**q_plastic
};

struct ThermoArgs {
  int n_slices, nVp;
  const long long *off;
  const unsigned char *deg, *edge;
  const int *C, *iTri;
  const double *Nx, *Ny, *Nx0, *Ny0, *R;
  const double2 *xy;
  const TriRec *tri;
  const unsigned *mbits;
  const double *Hi, *dHs_dx, *dHs_dy, *dHi_dx, *dHi_dy, *dHb_dt, *dHs_dt, *dHi_dt, *U_SSA, *V_SSA, *tau_c, *GHF, *T2m, *SMB_year;
  const int *aa2m;
  const double *U3, *V3;
  double *W3, *Ti, *Ti_new, *fric;
  unsigned long long *status;   // [0] columns replaced by the Robin solution, [1] bit0 DGTSV info /= 0, bit1 no upwind triangle
};

// ---- solve_SIA_3D, vertical velocity (ice_dynamics_module.f90:369-403); the Neumann pass on W_3D follows the heat kernel ----
__global__ void __launch_bounds__(256) k_thermo_w3d(ThermoArgs a, ThermoConst K)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nZ = K.nZ;
  const size_t ld = (size_t)a.nVp;
  for (int s = wg; s < a.n_slices; s += nw) {
    const long long o = a.off[s];
    const int w = (int)((a.off[s + 1] - o) >> 5);
    const int v = s * 32 + lane;
    const int n = a.deg[v];
    if (n == UFM_DEG_PAD) continue;
    if (a.edge[v] > 0 || !(a.mbits[v] & MB_SHEET)) {
      for (int k = 0; k < nZ; k++) a.W3[k * ld + v] = 0.0;
      continue;
    }
    // dU/dx and dV/dy on every layer: home coefficient first, then the neighbours in C order (get_mesh_derivatives_vertex_3D)
    double gx[UFM_MAX_NZ], gy[UFM_MAX_NZ];
    {
      const double nx0 = a.Nx0[v], ny0 = a.Ny0[v];
      for (int k = 0; k < nZ; k++) { gx[k] = nx0 * a.U3[k * ld + v]; gy[k] = ny0 * a.V3[k * ld + v]; }
    }
    for (int c = 0; c < w; c++) {
      if (c < n) {
        const long long e = o + (long long)c * 32 + lane;
        const int j = a.C[e];
        const double nx = a.Nx[e], ny = a.Ny[e];
        for (int k = 0; k < nZ; k++) { gx[k] = gx[k] + nx * a.U3[k * ld + j]; gy[k] = gy[k] + ny * a.V3[k * ld + j]; }
      }
    }
    const double Hi = a.Hi[v], hsx = a.dHs_dx[v], hsy = a.dHs_dy[v], hix = a.dHi_dx[v], hiy = a.dHi_dy[v];
    const double dHb_dx = hsx - hix, dHb_dy = hsy - hiy;
    double Wk1 = a.dHb_dt[v] + a.U3[(nZ - 1) * ld + v] * dHb_dx + a.V3[(nZ - 1) * ld + v] * dHb_dy;
    a.W3[(nZ - 1) * ld + v] = Wk1;
    double Uk1 = a.U3[(nZ - 1) * ld + v], Vk1 = a.V3[(nZ - 1) * ld + v];
    const double Hm = fmax(0.1, Hi);
    for (int k = nZ - 2; k >= 0; k--) {
      const double Uk = a.U3[k * ld + v], Vk = a.V3[k * ld + v];
      const double zk = K.zeta[k], zk1 = K.zeta[k + 1];
      const double w1 = (gx[k] + gx[k + 1]) / 2.0;
      const double w2 = (gy[k] + gy[k + 1]) / 2.0;
      const double w3 = ((hsx - 0.5 * (zk1 + zk) * hix) / Hm) * ((Uk1 - Uk) / (zk1 - zk));
      const double w4 = ((hsy - 0.5 * (zk1 + zk) * hiy) / Hm) * ((Vk1 - Vk) / (zk1 - zk));
      const double Wk = Wk1 - Hi * (w1 + w2 + w3 + w4) * (zk1 - zk);
      a.W3[k * ld + v] = Wk;
      Wk1 = Wk; Uk1 = Uk; Vk1 = Vk;
    }
  }
}

// LAPACK DGTSV, one right-hand side (netlib reference algorithm): Gaussian elimination with partial pivoting
__device__ __forceinline__ int d_dgtsv(const int n, double *dl, double *d, double *du, double *b)
{
  for (int i = 0; i < n - 2; i++) {
    if (fabs(d[i]) >= fabs(dl[i])) {
      if (d[i] != 0.0) {
        const double fact = dl[i] / d[i];
        d[i + 1] = d[i + 1] - fact * du[i];
        b[i + 1] = b[i + 1] - fact * b[i];
      } else return i + 1;
      dl[i] = 0.0;
    } else {
      const double fact = d[i] / dl[i];
      d[i] = dl[i];
      double temp = d[i + 1];
      d[i + 1] = du[i] - fact * temp;
      dl[i] = du[i + 1];
      du[i + 1] = -fact * dl[i];
      du[i] = temp;
      temp = b[i];
      b[i] = b[i + 1];
      b[i + 1] = temp - fact * b[i + 1];
    }
  }
  if (n > 1) {
    const int i = n - 2;
    if (fabs(d[i]) >= fabs(dl[i])) {
      if (d[i] != 0.0) {
        const double fact = dl[i] / d[i];
        d[i + 1] = d[i + 1] - fact * du[i];
        b[i + 1] = b[i + 1] - fact * b[i];
      } else return i + 1;
    } else {
      const double fact = d[i] / dl[i];
      d[i] = dl[i];
      double temp = d[i + 1];
      d[i + 1] = du[i] - fact * temp;
      du[i] = temp;
      temp = b[i];
      b[i] = b[i + 1];
      b[i + 1] = temp - fact * b[i + 1];
    }
  }
  if (d[n - 1] == 0.0) return n;
  b[n - 1] = b[n - 1] / d[n - 1];
  if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
  for (int i = n - 3; i >= 0; i--) b[i] = (b[i] - du[i] * b[i + 1] - dl[i] * b[i + 2]) / d[i];
  return 0;
}

__device__ __forceinline__ double d_surface_temperature(const ThermoArgs &a, const int v)
{
  double sT = 0.0;
  for (int mo = 0; mo < 12; mo++) sT = sT + a.T2m[(size_t)mo * a.nVp + v];
  return fmin(UFM_T0, sT / 12.0);
}

// ---- the heat equation, one implicit step per column (thermodynamics_module.f90:66-172) ----
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_thermo_heat(ThermoArgs a, ThermoConst K)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nZ = K.nZ;
  const size_t ld = (size_t)a.nVp;
  for (int s = wg; s < a.n_slices; s += nw) {
    const long long o = a.off[s];
    const int w = (int)((a.off[s + 1] - o) >> 5);
    const int v = s * 32 + lane;
    const int n = a.deg[v];
    if (n == UFM_DEG_PAD) continue;
    const double Ts = d_surface_temperature(a, v);   // ice%Ti(vi,1) = MIN(T0, SUM(T2m(vi,:)) / 12)
    const unsigned mb = a.mbits[v];
    const bool sheet = (mb & MB_SHEET) != 0;
    // bottom_frictional_heating
    double fh = 0.0;
    if (sheet) {
      const double delta_v = 1E-3, q_plastic = 0.30;
      const double u = a.U_SSA[v], vv = a.V_SSA[v];
      const double beta_base = a.tau_c[a.aa2m[v]] * (ufm_pow(delta_v * delta_v + u * u + vv * vv, 0.5 * (q_plastic - 1.0))) / K.ten_q;
      fh = beta_base * (u * u + vv * vv);
    }
    a.fric[v] = fh;
    if (a.edge[v] > 0) continue;   // filled by the Neumann pass
    if (!(mb & MB_ICE)) {
      for (int k = 0; k < nZ; k++) a.Ti_new[k * ld + v] = Ts;
      continue;
    }
    const double Hi = a.Hi[v], hsx = a.dHs_dx[v], hsy = a.dHs_dy[v], hix = a.dHi_dx[v], hiy = a.dHi_dy[v];
    const double hst = a.dHs_dt[v], hit = a.dHi_dt[v];
    const double inverse_Hi = 1.0 / fmax(0.1, Hi);
    const double dzeta_dz = -inverse_Hi;
    const double2 pv = a.xy[v];
    const double Rv = a.R[v];
    double dl[UFM_MAX_NZ], dd[UFM_MAX_NZ], du[UFM_MAX_NZ], x[UFM_MAX_NZ];
    // surface boundary condition
    dd[0] = 1.0; du[0] = 0.0; x[0] = Ts;
    for (int k = 1; k <= nZ - 2; k++) {   // reference layers 2 .. NZ-1
      const double Uk = a.U3[k * ld + v], Vk = a.V3[k * ld + v];
      const double Tk = a.Ti[k * ld + v];
      // get_upwind_derivative_vertex_3D
      double dTi_dx = 0.0, dTi_dy = 0.0;
      if (fabs(Uk) < 1E-10 && fabs(Vk) < 1E-10) {
        dTi_dx = a.Nx0[v] * Tk; dTi_dy = a.Ny0[v] * Tk;
        for (int c = 0; c < w; c++) {
          if (c < n) {
            const long long e = o + (long long)c * 32 + lane;
            const double t = a.Ti[k * ld + a.C[e]];
            dTi_dx = dTi_dx + a.Nx[e] * t; dTi_dy = dTi_dy + a.Ny[e] * t;
          }
        }
      } else {
        const double den = 4.0 * sqrt(Uk * Uk + Vk * Vk);
        const double px = pv.x - Uk * Rv / den, py = pv.y - Vk * Rv / den;
        int tup = -1;
        for (int c = 0; c < w && tup < 0; c++) {
          if (c < n) {
            const int ti = a.iTri[o + (long long)c * 32 + lane];
            if (ti < 0) break;
            const TriRec &q = a.tri[ti];
            const double tol = 1E-8;
            const double as_x = px - q.ax, as_y = py - q.ay;
            const double s1 = ((q.bx - q.ax) * as_y - (q.by - q.ay) * as_x);
            const double s2 = ((q.cx - q.ax) * as_y - (q.cy - q.ay) * as_x);
            const double s3 = ((q.cx - q.bx) * (py - q.by) - (q.cy - q.by) * (px - q.bx));
            if (s1 > -tol && s2 < tol && s3 > -tol) tup = ti;
          }
        }
        if (tup < 0) atomicOr(a.status + 1, 2ull);
        else {
          const TriRec &q = a.tri[tup];
          const double t0 = a.Ti[k * ld + q.v[0]], t1 = a.Ti[k * ld + q.v[1]], t2 = a.Ti[k * ld + q.v[2]];
          dTi_dx = q.nx[0] * t0 + q.nx[1] * t1 + q.nx[2] * t2;
          dTi_dy = q.ny[0] * t0 + q.ny[1] * t1 + q.ny[2] * t2;
        }
      }
      const double Cpi = K.realistic ? 2115.3 + 7.79293 * (Tk - UFM_T0) : 2009.0;
      const double Ki = K.realistic ? 3.101E+08 * exp(-0.0057 * Tk) : 2.1 * UFM_SEC_PER_YEAR;
      double internal_heating = 0.0;
      if (sheet) {
        const double Um = a.U3[(k - 1) * ld + v], Up = a.U3[(k + 1) * ld + v], Vm = a.V3[(k - 1) * ld + v], Vp = a.V3[(k + 1) * ld + v];
        internal_heating = ((-UFM_GRAV * K.zeta[k]) / Cpi) * ((K.a_zeta[k] * Um + K.b_zeta[k] * Uk + K.c_zeta[k] * Up) * hsx +
                                                             (K.a_zeta[k] * Vm + K.b_zeta[k] * Vk + K.c_zeta[k] * Vp) * hsy);
      }
      const double dzeta_dt = inverse_Hi * (hst - K.zeta[k] * hit);
      const double dzeta_dx = inverse_Hi * (hsx - K.zeta[k] * hix);
      const double dzeta_dy = inverse_Hi * (hsy - K.zeta[k] * hiy);
      const double f1 = (Ki * (dzeta_dz * dzeta_dz)) / (UFM_ICE_DENSITY * Cpi);
      const double f2 = dzeta_dt + dzeta_dx * Uk + dzeta_dy * Vk + dzeta_dz * a.W3[k * ld + v];
      const double f3 = internal_heating + (Uk * dTi_dx + Vk * dTi_dy) - Tk / K.dt_thermo;
      dl[k - 1] = f1 * K.a_zz[k] - f2 * K.a_zeta[k];                       // alpha(k) = ldiag(k-1)
      dd[k] = f1 * K.b_zz[k] - f2 * K.b_zeta[k] - 1.0 / K.dt_thermo;      // beta(k)
      du[k] = f1 * K.c_zz[k] - f2 * K.c_zeta[k];                          // gamma(k)
      x[k] = f3;                                                          // delta(k)
    }
    const double T_bot = a.Ti[(nZ - 1) * ld + v];
    const double Ki_bot = K.realistic ? 3.101E+08 * exp(-0.0057 * T_bot) : 2.1 * UFM_SEC_PER_YEAR;
    const double pmp_bot = K.realistic ? UFM_T0 - UFM_CC * Hi * K.zeta[nZ - 1] : UFM_T0 - (K.zeta[nZ - 1] * Hi * 8.7E-04);
    const double bottom_flux = (K.zeta[nZ - 1] - K.zeta[nZ - 2]) * (a.GHF[v] + fh) / (dzeta_dz * Ki_bot);
    if ((mb & MB_SHELF) || (mb & MB_GL)) { dl[nZ - 2] = 0.0; dd[nZ - 1] = 1.0; x[nZ - 1] = UFM_SMT; }
    else {
      dl[nZ - 2] = 1.0; dd[nZ - 1] = -1.0; x[nZ - 1] = bottom_flux;
      if (T_bot >= pmp_bot) { dl[nZ - 2] = 0.0; dd[nZ - 1] = 1.0; x[nZ - 1] = pmp_bot; }
    }
    const double T_above_bot = a.Ti[(nZ - 2) * ld + v];
    if (d_dgtsv(nZ, dl, dd, du, x)) atomicOr(a.status + 1, 1ull);
    for (int k = 0; k < nZ - 1; k++) {
      const double pmp = K.realistic ? UFM_T0 - UFM_CC * Hi * K.zeta[k] : UFM_T0 - (K.zeta[k] * Hi * 8.7E-04);
      a.Ti_new[k * ld + v] = fmin(x[k], pmp);
    }
    double tb = x[nZ - 1];
    if (tb >= pmp_bot) tb = fmin(pmp_bot, T_above_bot - bottom_flux);
    a.Ti_new[(nZ - 1) * ld + v] = tb;
  }
}

// ---- Ti = Ti_new, then the safety net: columns colder than 150 K anywhere get the Robin solution (:174-200, :204-279) ----
__global__ void __launch_bounds__(256) k_thermo_finish(int nV, ThermoArgs a, ThermoConst K)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const int nZ = K.nZ;
  const size_t ld = (size_t)a.nVp;
  double mn = 1e300;
  for (int k = 0; k < nZ; k++) { const double t = a.Ti_new[k * ld + v]; a.Ti[k * ld + v] = t; mn = fmin(mn, t); }
  if (!(mn < 150.0)) return;
  atomicAdd(a.status, 1ull);
  const double kappa_0_ice_conductivity = 9.828, kappa_e_ice_conductivity = 0.0057, c_0_specific_heat = 2127.5, Claus_Clap_gradient = 8.7E-04;
  const double thermal_conductivity_robin = kappa_0_ice_conductivity * UFM_SEC_PER_YEAR * exp(-kappa_e_ice_conductivity * UFM_T0);
  const double thermal_diffusivity_robin = thermal_conductivity_robin / (UFM_ICE_DENSITY * c_0_specific_heat);
  const double bottom_temperature_gradient_robin = -a.GHF[v] / thermal_conductivity_robin;
  const double Ts = d_surface_temperature(a, v);
  const double Hi = a.Hi[v];
  const unsigned mb = a.mbits[v];
  for (int k = 0; k < nZ; k++) {
    double t;
    if (mb & MB_SHEET) {
      const double smb = a.SMB_year[v];
      if (smb > 0.0) {
        const double thermal_length_scale = sqrt(2.0 * thermal_diffusivity_robin * Hi / smb);
        const double distance_above_bed = (1.0 - K.zeta[k]) * Hi;
        const double erf1 = erf(distance_above_bed / thermal_length_scale);
        const double erf2 = erf(Hi / thermal_length_scale);
        t = Ts + sqrt(UFM_PI) / 2.0 * thermal_length_scale * bottom_temperature_gradient_robin * (erf1 - erf2);
      } else t = Ts + ((UFM_T0 - Claus_Clap_gradient * Hi) - Ts) * K.zeta[k];
    } else if (mb & MB_SHELF) t = Ts + K.zeta[k] * (UFM_SMT - Ts);
    else t = Ts;
    a.Ti[k * ld + v] = fmin(t, UFM_T0 - Claus_Clap_gradient * Hi * K.zeta[k]);
  }
}

// =============================================================================================
int ufm_thermo_powtab_init(const UfmPowTab *t) { return ufm_powtab_upload_tu(t); }
static inline int grid_for(long long n, int b) { return (int)((n + b - 1) / b); }

static bool thermo_skipped(const ufm_handle *h)
{
  const int b = h->P.benchmark;   // thermodynamics_module.f90:44-64
  return b == UFM_BM_MISMIP_MOD || b == UFM_BM_MESH_GENERATION_TEST || b == UFM_BM_HALFAR || b == UFM_BM_BUELER || b == UFM_BM_SSA_ICESTREAM;
}

static int thermo_setup(ufm_handle *h, ThermoArgs &a, ThermoConst &K)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  if (!m.has_tri) return ufm_set_error(-2, "thermodynamics needs a mesh uploaded with Tri / niTri / iTri / R / NxTri / NyTri");
  if (h->P.nZ < 3) return ufm_set_error(-2, "thermodynamics needs nZ >= 3");
  if (!(h->P.dt_thermo > 0.0)) return ufm_set_error(-2, "thermodynamics needs dt_thermo > 0 (ufm_params)");
  memset(&K, 0, sizeof(K));
  const int nZ = K.nZ = h->P.nZ;
  K.realistic = s.realistic_A ? 1 : 0;
  K.dt_thermo = h->P.dt_thermo;
  K.ten_q = pow(100.0, 0.30);   // u_threshold**q_plastic, host libm as the reference
  for (int k = 0; k < nZ; k++) K.zeta[k] = h->P.zeta[k];
  // initialize_zeta_discretization (zeta_module.f90:113-173); reference index k = 2..NZ-1 stored at k-1
  for (int k = 2; k <= nZ - 1; k++) {
    const double a_k = h->P.zeta[k - 1] - h->P.zeta[k - 2], b_k = h->P.zeta[k] - h->P.zeta[k - 1];
    K.a_zeta[k - 1] = -b_k / (a_k * (a_k + b_k));
    K.b_zeta[k - 1] = (b_k - a_k) / (a_k * b_k);
    K.c_zeta[k - 1] = a_k / (b_k * (a_k + b_k));
    K.a_zz[k - 1] = 2.0 / (a_k * (a_k + b_k));
    K.b_zz[k - 1] = -2.0 / (a_k * b_k);
    K.c_zz[k - 1] = 2.0 / (b_k * (a_k + b_k));
  }
  a.n_slices = m.aa.n_slices; a.nVp = m.nVp; a.off = m.aa.off; a.deg = m.aa.deg; a.edge = m.aa_edge; a.C = m.aa_C; a.iTri = m.aa_iTri;
  a.Nx = m.aa_Nx; a.Ny = m.aa_Ny; a.Nx0 = m.aa_Nx0; a.Ny0 = m.aa_Ny0; a.R = m.aa_R; a.xy = m.aa_xy; a.tri = m.tri; a.mbits = s.mbits;
  a.Hi = s.Hi; a.dHs_dx = s.dHs_dx; a.dHs_dy = s.dHs_dy; a.dHi_dx = s.dHi_dx; a.dHi_dy = s.dHi_dy; a.dHb_dt = s.dHb_dt; a.dHs_dt = s.dHs_dt; a.dHi_dt = s.dHi_dt;
  a.U_SSA = s.U_SSA; a.V_SSA = s.V_SSA; a.tau_c = s.tau_c; a.GHF = s.GHF; a.T2m = s.T2m; a.SMB_year = s.SMB_year; a.aa2m = m.aa2m;
  a.U3 = s.U_3D; a.V3 = s.V_3D; a.W3 = s.W_3D; a.Ti = s.Ti; a.Ti_new = s.Ti_new; a.fric = s.fric_heat;
  a.status = s.ctrl + CTRL_THERMO_STATUS;
  return 0;
}

int ufm_k_thermo_w3d(ufm_handle *h)
{
  ThermoArgs a; ThermoConst K;
  int rc = thermo_setup(h, a, K);
  if (rc) return rc;
  k_thermo_w3d<<<grid_for((long long)a.n_slices * 32, 256), 256, 0, h->stream>>>(a, K);
  h->cnt.kernel_launches++;
  if ((rc = ufm_cuda_check(cudaGetLastError(), "k_thermo_w3d"))) return rc;
  return ufm_k_neumann3d_pair(h, h->st.W_3D, h->st.W_3D);
}

int ufm_k_thermo_heat(ufm_handle *h, ufm_thermo_stats *st)
{
  DevMesh &m = h->mesh; DevState &s = h->st;
  ThermoArgs a; ThermoConst K;
  int rc = thermo_setup(h, a, K);
  if (rc) return rc;
  UFM_CUDA(cudaMemsetAsync(a.status, 0, 2 * sizeof(unsigned long long), h->stream));
  {
    static int minb = -1;   // resident CTAs per SM asked of ptxas (env UFM_HEAT_MINB): trades registers for warps
    if (minb < 0) { const char *e = getenv("UFM_HEAT_MINB"); minb = e ? atoi(e) : UFM_HEAT_MINB_DEFAULT; }
    const int g = grid_for((long long)a.n_slices * 32, 128);
    if (minb >= 12) k_thermo_heat<12><<<g, 128, 0, h->stream>>>(a, K);
    else if (minb >= 8) k_thermo_heat<8><<<g, 128, 0, h->stream>>>(a, K);
    else if (minb >= 6) k_thermo_heat<6><<<g, 128, 0, h->stream>>>(a, K);
    else k_thermo_heat<1><<<g, 128, 0, h->stream>>>(a, K);
  }
  h->cnt.kernel_launches++;
  if ((rc = ufm_cuda_check(cudaGetLastError(), "k_thermo_heat"))) return rc;
  if ((rc = ufm_k_neumann3d_pair(h, s.Ti_new, s.Ti_new))) return rc;
  k_thermo_finish<<<grid_for(m.nV, 256), 256, 0, h->stream>>>(m.nV, a, K);
  h->cnt.kernel_launches++;
  unsigned long long *res = (unsigned long long *)(s.scal_h + 40);
  UFM_CUDA(cudaMemcpyAsync(res, a.status, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  st->n_unstable = (int)res[0];
  st->rc = 0;
  if (res[1] & 2ull) { st->rc = -10; return ufm_set_error(-10, "update_ice_temperature: couldnt find upwind triangle (mesh_derivatives_module.f90:473-477)"); }
  if (res[1] & 1ull) { st->rc = -9; return ufm_set_error(-9, "update_ice_temperature: DGTSV problem with tridiagonal system (thermodynamics_module.f90:350-354)"); }
  // IF (n_unstable > CEILING(REAL(mesh%nV) / 100._dp)) -- REAL() is single precision
  if (st->n_unstable > (int)ceil((double)(float)m.nV / 100.0)) {
    st->rc = -8;
    return ufm_set_error(-8, "thermodynamics: heat equation solver unstable for more than 1%% of vertices (%d of %d)", st->n_unstable, m.nV);
  }
  return 0;
}

extern "C" {

#define NEED_MESH_T(h) do { if (!(h)) return ufm_set_error(-2, "NULL handle"); if (!(h)->has_mesh) return ufm_set_error(-2, "no mesh resident"); \
  int rc0__ = ufm_cuda_check(cudaSetDevice((h)->device), "cudaSetDevice"); if (rc0__) return rc0__; } while (0)

int ufm_thermo_w3d(ufm_handle *h) { NEED_MESH_T(h); return ufm_k_thermo_w3d(h); }

int ufm_thermo_heat(ufm_handle *h, ufm_thermo_stats *st)
{
  NEED_MESH_T(h);
  ufm_thermo_stats tmp;
  if (!st) st = &tmp;
  st->n_unstable = 0; st->rc = 0;
  return ufm_k_thermo_heat(h, st);
}

int ufm_update_ice_temperature(ufm_handle *h, ufm_thermo_stats *st)
{
  NEED_MESH_T(h);
  ufm_thermo_stats tmp;
  if (!st) st = &tmp;
  st->n_unstable = 0; st->rc = 0;
  if (thermo_skipped(h)) return 0;
  int rc = ufm_k_sia3d(h);   // U_3D, V_3D (+ their Neumann passes)
  if (rc) return rc;
  if ((rc = ufm_k_thermo_w3d(h))) return rc;
  return ufm_k_thermo_heat(h, st);
}

}  // extern "C"
