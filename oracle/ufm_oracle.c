/*
 * ufm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See ufm_oracle.h ("PARITY PIN").
 *
 * Statement-by-statement CPU restatement of the UFEMISM v1.1.1 ice-dynamics hot path.
 * "Ranks" of the reference's MPI shared-memory SPMD model are OpenMP threads here: every
 * routine body is executed by all threads of one parallel region, each on the index range
 * partition_list() would give that rank, and CALL sync is `#pragma omp barrier`.
 * With nthreads = 1 this is the serial reference semantics.
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off (no -ffast-math): see oracle/Makefile.
 */
#include "ufm_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* src/parameters_module.f90:9-21 */
/* NORM2([x, y]) as gfortran evaluates it: the intrinsic calls libgfortran's _gfortran_norm2_r8 (m4/norm2.m4), a scaled sum of
 * squares -- not the C library's hypotenuse function, from which it differs in the last bit for about one argument pair in three.
 * Found by running the reference's own source through oracle/f90py.py (tests/test_reference_source.py). */
static double ora_norm2_2(double x, double y)
{
  double result = 0.0, scale = 1.0;
  const double v[2] = {x, y};
  for (int k = 0; k < 2; k++) {
    if (v[k] != 0.0) {
      const double absX = fabs(v[k]);
      if (scale < absX) { const double val = scale / absX; result = 1.0 + result * val * val; scale = absX; }
      else { const double val = absX / scale; result += val * val; }
    }
  }
  return scale * sqrt(result);
}

static const double pi = 3.141592653589793;
static const double sec_per_year = 31556943.36;
static const double grav = 9.81;
static const double n_flow = 3.0;
static const double ice_density = 910.0;
static const double seawater_density = 1028.0;
static const double SMT = 271.15;
static const double T0 = 273.16;
static const double CC = 8.7E-04;

#define A2(a, i, j, ld) (a)[((size_t)((j) - 1)) * (size_t)(ld) + (size_t)((i) - 1)]
#define A1(a, i) (a)[(size_t)(i) - 1]
#define SYNC _Pragma("omp barrier")

void ora_config_defaults(ora_config *c)
{
  /* src/configuration_module.f90:37,124-126,169-184 (m_enh as in the benchmark configs: 1.0) */
  static const double z[15] = {0.00, 0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 0.925, 0.95, 0.975, 0.99, 1.00};
  memset(c, 0, sizeof(*c));
  c->nZ = 15;
  for (int k = 0; k < 15; k++) c->zeta[k] = z[k];
  c->m_enh_sia = 1.0; c->m_enh_ssa = 1.0;
  c->use_analytical_GL_flux = 0;
  c->SSA_RN_tol = 1E-5; c->SSA_max_outer_loops = 50; c->SSA_max_residual_UV = 2.5;
  c->SSA_SOR_omega = 1.2; c->SSA_max_inner_loops = 10000;
  c->dt_max = 10.0;
  c->benchmark = ORA_BM_HALFAR;
  c->nthreads = 1;
  c->dt_thermo = 10.0;
}

/* partition_list, src/mesh_help_functions_module.f90:1475-1496 (i = 0-based rank) */
void ora_partition_list(int ntot, int i, int n, int *i1, int *i2)
{
  if (ntot > n * 2) {
    /* i1 = MAX(1, FLOOR(REAL(ntot*i/n)) + 1), i2 = MIN(ntot, FLOOR(REAL(ntot*(i+1)/n))), src/mesh_help_functions_module.f90:1483-1484:
     * integer division, then REAL() WITHOUT a kind, i.e. single precision.  Above 2**24 = 16 777 216 the quotient is rounded to 24
     * bits, so the reference's ranges shift and can even miss the last element (ntot = 16 777 217, 8 ranks: nobody owns it) -- a
     * latent defect of the reference for lists longer than 16.7 M, found by running its source through oracle/f90py.py.  Restated
     * as coded; every list of the BASELINE configurations (<= 16.0 M combined-mesh vertices) is below that length. */
    long a = (long)floorf((float)(((long)ntot * i) / n)) + 1, b = (long)floorf((float)(((long)ntot * (i + 1)) / n));
    *i1 = (int)(a < 1 ? 1 : a);
    *i2 = (int)(b > ntot ? ntot : b);
  } else {
    if (i == 0) { *i1 = 1; *i2 = ntot; } else { *i1 = 1; *i2 = 0; }
  }
}

typedef struct { int i, n, v1, v2, ac1, ac2, a1, a2, cv1[5], cv2[5]; } rank_t;
static rank_t rank_of(const ora_mesh *m)
{
  rank_t r;
  r.i = omp_get_thread_num(); r.n = omp_get_num_threads();
  ora_partition_list(m->nV, r.i, r.n, &r.v1, &r.v2);
  ora_partition_list(m->nAc, r.i, r.n, &r.ac1, &r.ac2);
  ora_partition_list(m->nVAaAc, r.i, r.n, &r.a1, &r.a2);
  for (int c = 0; c < 5; c++) ora_partition_list(m->colour_nV[c], r.i, r.n, &r.cv1[c], &r.cv2[c]);
  return r;
}

/* is_floating, src/general_ice_model_data_module.f90:464-475 */
static inline int is_floating(double Hi, double Hb, double SL) { return Hi < (SL - Hb) * seawater_density / ice_density; }

static inline int is_benchmark_simple(int b)
{ return (b >= ORA_BM_EISMINT_1 && b <= ORA_BM_EISMINT_6) || b == ORA_BM_HALFAR || b == ORA_BM_BUELER; }

/* ============================================================================================
 * zeta_module: vertical_average (:33-57), vertical_integrate (:58-85)
 * ============================================================================================ */
static double vertical_average(const ora_config *c, const double *f)
{
  double average_f = 0.0;
  for (int k = 1; k <= c->nZ - 1; k++) average_f = average_f + 0.5 * (f[k] + f[k - 1]) * (c->zeta[k] - c->zeta[k - 1]);
  return average_f;
}
static void vertical_integrate(const ora_config *c, const double *f, double *int_f)
{
  int_f[c->nZ - 1] = 0.0;
  for (int k = c->nZ - 1; k >= 1; k--) int_f[k - 1] = int_f[k] - 0.5 * (f[k] + f[k - 1]) * (c->zeta[k] - c->zeta[k - 1]);
}

/* ============================================================================================
 * mesh_derivatives_module / mesh_ArakawaC_module stencils
 * ============================================================================================ */
/* get_mesh_derivatives_vertex, src/mesh_derivatives_module.f90:352-371 */
static inline void get_mesh_derivatives_vertex(const ora_mesh *m, const double *d, double *ddx, double *ddy, int vi)
{
  int nV = m->nV, n = A1(m->nC, vi);
  *ddx = A2(m->Nx, vi, n + 1, nV) * A1(d, vi);
  *ddy = A2(m->Ny, vi, n + 1, nV) * A1(d, vi);
  for (int ci = 1; ci <= n; ci++) {
    *ddx = *ddx + A2(m->Nx, vi, ci, nV) * A1(d, A2(m->C, vi, ci, nV));
    *ddy = *ddy + A2(m->Ny, vi, ci, nV) * A1(d, A2(m->C, vi, ci, nV));
  }
}
/* get_mesh_derivatives, :315-330 (rank body) */
static void get_mesh_derivatives_r(const ora_mesh *m, const rank_t *r, const double *d, double *ddx, double *ddy)
{
  for (int vi = r->v1; vi <= r->v2; vi++) get_mesh_derivatives_vertex(m, d, &A1(ddx, vi), &A1(ddy, vi), vi);
  SYNC
}
/* get_mesh_derivatives_vertex_Ac, src/mesh_ArakawaC_module.f90:557-581 */
static inline void get_mesh_derivatives_vertex_Ac(const ora_mesh *m, const double *d_Aa, double *ddx, double *ddy, double *ddp, double *ddo, int aci)
{
  int nAc = m->nAc;
  double x = 0.0, y = 0.0, o = 0.0;
  for (int n = 1; n <= 4; n++) {
    double dv = A1(d_Aa, A2(m->Aci, aci, n, nAc));
    x = x + A2(m->Nx_Ac, aci, n, nAc) * dv;
    y = y + A2(m->Ny_Ac, aci, n, nAc) * dv;
    o = o + A2(m->No_Ac, aci, n, nAc) * dv;
  }
  *ddx = x; *ddy = y; *ddo = o;
  *ddp = A1(m->Np_Ac, aci) * (A1(d_Aa, A2(m->Aci, aci, 2, nAc)) - A1(d_Aa, A2(m->Aci, aci, 1, nAc)));
}
static void get_mesh_derivatives_Ac_r(const ora_mesh *m, const rank_t *r, const double *d_Aa, double *ddx, double *ddy, double *ddp, double *ddo)
{
  for (int aci = r->ac1; aci <= r->ac2; aci++)
    get_mesh_derivatives_vertex_Ac(m, d_Aa, &A1(ddx, aci), &A1(ddy, aci), &A1(ddp, aci), &A1(ddo, aci), aci);
  SYNC
}
/* get_mesh_derivatives_vertex_AaAc, src/mesh_ArakawaC_module.f90:599-618 */
static inline void get_mesh_derivatives_vertex_AaAc(const ora_mesh *m, const double *d, double *ddx, double *ddy, int ai)
{
  int M = m->nVAaAc, n = A1(m->nCAaAc, ai);
  *ddx = A2(m->Nx_AaAc, ai, n + 1, M) * A1(d, ai);
  *ddy = A2(m->Ny_AaAc, ai, n + 1, M) * A1(d, ai);
  for (int ci = 1; ci <= n; ci++) {
    *ddx = *ddx + A2(m->Nx_AaAc, ai, ci, M) * A1(d, A2(m->CAaAc, ai, ci, M));
    *ddy = *ddy + A2(m->Ny_AaAc, ai, ci, M) * A1(d, A2(m->CAaAc, ai, ci, M));
  }
}
/* get_mesh_curvatures_vertex_AaAc, src/mesh_ArakawaC_module.f90:636-658 -- AS CODED: every
 * coefficient multiplies the HOME value d(ai); `ac` is assigned and never used.
 * Only ddxy is consumed by the caller; ddxx/ddyy have no side effects and are not evaluated. */
static inline double get_mesh_curvature_xy_vertex_AaAc(const ora_mesh *m, const double *d, int ai)
{
  int M = m->nVAaAc, n = A1(m->nCAaAc, ai);
  double ddxy = A1(d, ai) * A2(m->Nxy_AaAc, ai, n + 1, M);
  for (int ci = 1; ci <= n; ci++) ddxy = ddxy + A1(d, ai) * A2(m->Nxy_AaAc, ai, ci, M);
  return ddxy;
}
static inline int is_edge_AaAc(const ora_mesh *m, int ai)
{ return ai <= m->nV ? A1(m->edge_index, ai) > 0 : A1(m->edge_index_Ac, ai - m->nV) > 0; }

/* apply_Neumann_boundary_AaAc, src/mesh_ArakawaC_module.f90:660-724 (rank body) */
static void apply_Neumann_boundary_AaAc_r(const ora_mesh *m, const rank_t *r, double *d)
{
  int M = m->nVAaAc, W = m->nC_mem;
  double vals[64];
  for (int ai = (r->a1 > 5 ? r->a1 : 5); ai <= r->a2; ai++) {
    if (!is_edge_AaAc(m, ai)) continue;
    for (int k = 0; k < W; k++) vals[k] = 0.0;
    int nvals = 0;
    for (int ci = 1; ci <= A1(m->nCAaAc, ai); ci++) {
      int ac = A2(m->CAaAc, ai, ci, M);
      if (is_edge_AaAc(m, ac)) continue;
      nvals = nvals + 1;
      vals[nvals - 1] = A1(d, ac);
    }
    double s = 0.0;
    for (int k = 0; k < W; k++) s = s + vals[k];
    A1(d, ai) = s / nvals;
  }
  SYNC
  if (r->i == 0) {
    for (int ai = 1; ai <= 4; ai++) {
      for (int k = 0; k < W; k++) vals[k] = 0.0;
      int nvals = 0;
      for (int ci = 1; ci <= A1(m->nCAaAc, ai); ci++) {
        int ac = A2(m->CAaAc, ai, ci, M);
        nvals = nvals + 1;
        vals[nvals - 1] = A1(d, ac);
      }
      double s = 0.0;
      for (int k = 0; k < W; k++) s = s + vals[k];
      A1(d, ai) = s / nvals;
    }
  }
  SYNC
}
/* map_Aa_to_Ac, src/mesh_ArakawaC_module.f90:726-747 */
static void map_Aa_to_Ac_r(const ora_mesh *m, const rank_t *r, const double *d_Aa, double *d_Ac)
{
  for (int aci = r->ac1; aci <= r->ac2; aci++) {
    int vi = A2(m->Aci, aci, 1, m->nAc), vj = A2(m->Aci, aci, 2, m->nAc);
    A1(d_Ac, aci) = (A1(d_Aa, vi) + A1(d_Aa, vj)) / 2.0;
  }
  SYNC
}
/* map_Aa_to_Ac_3D, :748-769 */
static void map_Aa_to_Ac_3D_r(const ora_mesh *m, const rank_t *r, int nZ, const double *d_Aa, double *d_Ac)
{
  for (int aci = r->ac1; aci <= r->ac2; aci++) {
    int vi = A2(m->Aci, aci, 1, m->nAc), vj = A2(m->Aci, aci, 2, m->nAc);
    for (int k = 1; k <= nZ; k++) A2(d_Ac, aci, k, m->nAc) = (A2(d_Aa, vi, k, m->nV) + A2(d_Aa, vj, k, m->nV)) / 2.0;
  }
  SYNC
}
/* map_Ac_to_Aa, :770-791 */
static void map_Ac_to_Aa_r(const ora_mesh *m, const rank_t *r, const double *d_Ac, double *d_Aa)
{
  for (int vi = r->v1; vi <= r->v2; vi++) A1(d_Aa, vi) = 0.0;
  for (int vi = r->v1; vi <= r->v2; vi++)
    for (int ci = 1; ci <= A1(m->nC, vi); ci++) {
      int aci = A2(m->iAci, vi, ci, m->nV);
      A1(d_Aa, vi) = A1(d_Aa, vi) + A1(d_Ac, aci) / A1(m->nC, vi);
    }
  SYNC
}
/* rotate_xy_to_po, :815-844 */
static void rotate_xy_to_po_r(const ora_mesh *m, const rank_t *r, const double *dx_, const double *dy_, double *dp_, double *do_)
{
  for (int ci = r->ac1; ci <= r->ac2; ci++) {
    int vi = A2(m->Aci, ci, 1, m->nAc), vj = A2(m->Aci, ci, 2, m->nAc);
    double Dx = A2(m->V, vj, 1, m->nV) - A2(m->V, vi, 1, m->nV);
    double Dy = A2(m->V, vj, 2, m->nV) - A2(m->V, vi, 2, m->nV);
    double D = sqrt(Dx * Dx + Dy * Dy);
    A1(dp_, ci) = A1(dx_, ci) * Dx / D + A1(dy_, ci) * Dy / D;
    A1(do_, ci) = A1(dy_, ci) * Dx / D - A1(dx_, ci) * Dy / D;
  }
  SYNC
}

/* ============================================================================================
 * calculate_ice_thickness_change, src/ice_dynamics_module.f90:31-237
 * ============================================================================================ */
void ora_calculate_ice_thickness_change(const ora_mesh *m, ora_ice *ice, const ora_config *c, double dt)
{
  int nV = m->nV, nAc = m->nAc, W = m->nC_mem;
  double *Vi_SMB = (double *)calloc((size_t)nV, sizeof(double));
#pragma omp parallel num_threads(c->nthreads)
  {
    rank_t r = rank_of(m);
    for (int ci = 1; ci <= W; ci++)
      for (int vi = r.v1; vi <= r.v2; vi++) A2(ice->dVi_in, vi, ci, nV) = 0.0;
    SYNC
    /* ice fluxes across all Aa vertex connections, :66-105 */
    for (int aci = r.ac1; aci <= r.ac2; aci++) {
      int vi = A2(m->Aci, aci, 1, nAc), vj = A2(m->Aci, aci, 2, nAc), ci = 0, cj = 0;
      for (int cii = 1; cii <= A1(m->nC, vi); cii++) if (A2(m->C, vi, cii, nV) == vj) { ci = cii; break; }
      for (int cji = 1; cji <= A1(m->nC, vj); cji++) if (A2(m->C, vj, cji, nV) == vi) { cj = cji; break; }
      double Upar = A1(ice->Up_SIA_Ac, aci) + A1(ice->Up_SSA_Ac, aci), dVi;
      if (Upar > 0.0) dVi = A1(ice->Hi, vi) * Upar * A2(m->Cw, vi, ci, nV) * dt;
      else            dVi = A1(ice->Hi, vj) * Upar * A2(m->Cw, vi, ci, nV) * dt;
      A2(ice->dVi_in, vi, ci, nV) = -dVi;
      A2(ice->dVi_in, vj, cj, nV) = dVi;
    }
    SYNC
    /* :112 */
    for (int vi = r.v1; vi <= r.v2; vi++) Vi_SMB[vi - 1] = (A1(ice->SMB_year, vi) + A1(ice->BMB, vi)) * A1(m->A, vi) * dt;
    SYNC
    /* out-flux limiter, :115-166 */
    for (int vi = r.v1; vi <= r.v2; vi++) {
      double Vi_available = A1(m->A, vi) * A1(ice->Hi, vi);
      double Vi_in = 0.0, Vi_out = 0.0;
      for (int ci = 1; ci <= A1(m->nC, vi); ci++) {
        if (A2(ice->dVi_in, vi, ci, nV) > 0.0) Vi_in = Vi_in + A2(ice->dVi_in, vi, ci, nV);
        else Vi_out = Vi_out - A2(ice->dVi_in, vi, ci, nV);
      }
      (void)Vi_in;
      double rescale_factor = 1.0;
      if (-Vi_SMB[vi - 1] >= Vi_available) { Vi_SMB[vi - 1] = -Vi_available; rescale_factor = 0.0; }
      if (Vi_out > Vi_available + Vi_SMB[vi - 1]) rescale_factor = (Vi_available + Vi_SMB[vi - 1]) / Vi_out;
      if (rescale_factor < 1.0) {
        for (int ci = 1; ci <= A1(m->nC, vi); ci++) {
          int vj = A2(m->C, vi, ci, nV);
          if (A2(ice->dVi_in, vi, ci, nV) < 0.0) {
            A2(ice->dVi_in, vi, ci, nV) = A2(ice->dVi_in, vi, ci, nV) * rescale_factor;
            for (int cji = 1; cji <= A1(m->nC, vj); cji++)
              if (A2(m->C, vj, cji, nV) == vi) { A2(ice->dVi_in, vj, cji, nV) = -A2(ice->dVi_in, vi, ci, nV); break; }
          }
        }
      }
    }
    SYNC
    /* :172-187 */
    for (int vi = r.v1; vi <= r.v2; vi++) {
      double dVi = 0.0;
      for (int ci = 1; ci <= W; ci++) dVi = dVi + A2(ice->dVi_in, vi, ci, nV);
      A1(ice->dHi_dt, vi) = (dVi + Vi_SMB[vi - 1]) / (A1(m->A, vi) * dt);
    }
    SYNC
    if (dt == 0.0) for (int vi = r.v1; vi <= r.v2; vi++) A1(ice->dHi_dt, vi) = 0.0;
    SYNC
    for (int vi = r.v1; vi <= r.v2; vi++) {
      A1(ice->Hi_prev, vi) = A1(ice->Hi, vi);
      A1(ice->Hi, vi) = A1(ice->Hi, vi) + (A1(ice->dHi_dt, vi) * dt);
    }
    /* boundary conditions, :189-228 (identical in the benchmark and the realistic branch; 'SSA_icestream' has none, :206) */
    if (c->benchmark != ORA_BM_SSA_ICESTREAM)
      for (int vi = r.v1; vi <= r.v2; vi++) if (A1(m->edge_index, vi) > 0) A1(ice->Hi, vi) = 0.0;
    SYNC
    for (int vi = r.v1; vi <= r.v2; vi++) if (A1(ice->mask_noice, vi) == 1) A1(ice->Hi, vi) = 0.0;
    SYNC
  }
  free(Vi_SMB);
}

/* ============================================================================================
 * general_ice_model_data_module
 * ============================================================================================ */
/* determine_masks, src/general_ice_model_data_module.f90:97-297 (rank body) */
static void determine_masks_r(const ora_mesh *m, const rank_t *r, ora_ice *ice)
{
  enum { type_land = 0, type_ocean = 1, type_lake = 2, type_sheet = 3, type_shelf = 4, type_coast = 5, type_margin = 6,
         type_groundingline = 7, type_calvingfront = 8 };
  int nV = m->nV, nAc = m->nAc;
  for (int vi = r->v1; vi <= r->v2; vi++) {
    A1(ice->mask_land, vi) = 1; A1(ice->mask_ocean, vi) = 0; A1(ice->mask_lake, vi) = 0; A1(ice->mask_ice, vi) = 0;
    A1(ice->mask_sheet, vi) = 0; A1(ice->mask_shelf, vi) = 0; A1(ice->mask_coast, vi) = 0; A1(ice->mask_margin, vi) = 0;
    A1(ice->mask_gl, vi) = 0; A1(ice->mask_cf, vi) = 0; A1(ice->mask, vi) = type_land;
  }
  for (int aci = r->ac1; aci <= r->ac2; aci++) {
    A1(ice->mask_land_Ac, aci) = 1; A1(ice->mask_ocean_Ac, aci) = 0; A1(ice->mask_lake_Ac, aci) = 0; A1(ice->mask_ice_Ac, aci) = 0;
    A1(ice->mask_sheet_Ac, aci) = 0; A1(ice->mask_shelf_Ac, aci) = 0; A1(ice->mask_coast_Ac, aci) = 0; A1(ice->mask_margin_Ac, aci) = 0;
    A1(ice->mask_gl_Ac, aci) = 0; A1(ice->mask_cf_Ac, aci) = 0; A1(ice->mask_Ac, aci) = type_land;
  }
  SYNC
  for (int vi = r->v1; vi <= r->v2; vi++) {
    if (is_floating(A1(ice->Hi, vi), A1(ice->Hb, vi), A1(ice->SL, vi))) { A1(ice->mask_ocean, vi) = 1; A1(ice->mask_land, vi) = 0; A1(ice->mask, vi) = type_ocean; }
    if (A1(ice->Hi, vi) > 0.0) A1(ice->mask_ice, vi) = 1;
    if (A1(ice->mask_ice, vi) == 1 && A1(ice->mask_land, vi) == 1) { A1(ice->mask_sheet, vi) = 1; A1(ice->mask, vi) = type_sheet; }
    if (A1(ice->mask_ice, vi) == 1 && A1(ice->mask_ocean, vi) == 1) { A1(ice->mask_shelf, vi) = 1; A1(ice->mask, vi) = type_shelf; }
  }
  SYNC
  for (int vi = r->v1; vi <= r->v2; vi++) {
    int n = A1(m->nC, vi);
    if (A1(ice->mask_land, vi) == 1)
      for (int ci = 1; ci <= n; ci++) if (A1(ice->mask_ocean, A2(m->C, vi, ci, nV)) == 1) { A1(ice->mask, vi) = type_coast; A1(ice->mask_coast, vi) = 1; }
    if (A1(ice->mask_ice, vi) == 1)
      for (int ci = 1; ci <= n; ci++) if (A1(ice->mask_ice, A2(m->C, vi, ci, nV)) == 0) { A1(ice->mask, vi) = type_margin; A1(ice->mask_margin, vi) = 1; }
    if (A1(ice->mask_sheet, vi) == 1)
      for (int ci = 1; ci <= n; ci++) if (A1(ice->mask_shelf, A2(m->C, vi, ci, nV)) == 1) { A1(ice->mask, vi) = type_groundingline; A1(ice->mask_gl, vi) = 1; }
    if (A1(ice->mask_ice, vi) == 1)
      for (int ci = 1; ci <= n; ci++) if (A1(ice->mask_ocean, A2(m->C, vi, ci, nV)) == 1) { A1(ice->mask, vi) = type_calvingfront; A1(ice->mask_cf, vi) = 1; }
  }
  SYNC
  for (int aci = r->ac1; aci <= r->ac2; aci++) {
    if (is_floating(A1(ice->Hi_Ac, aci), A1(ice->Hb_Ac, aci), A1(ice->SL_Ac, aci))) { A1(ice->mask_ocean_Ac, aci) = 1; A1(ice->mask_land_Ac, aci) = 0; A1(ice->mask_Ac, aci) = type_ocean; }
    if (A1(ice->Hi_Ac, aci) > 0.0) A1(ice->mask_ice_Ac, aci) = 1;
    if (A1(ice->mask_ice_Ac, aci) == 1 && A1(ice->mask_land_Ac, aci) == 1) { A1(ice->mask_sheet_Ac, aci) = 1; A1(ice->mask_Ac, aci) = type_sheet; }
    if (A1(ice->mask_ice_Ac, aci) == 1 && A1(ice->mask_ocean_Ac, aci) == 1) { A1(ice->mask_shelf_Ac, aci) = 1; A1(ice->mask_Ac, aci) = type_shelf; }
  }
  SYNC
  for (int aci = r->ac1; aci <= r->ac2; aci++) {
    int vi = A2(m->Aci, aci, 1, nAc), vj = A2(m->Aci, aci, 2, nAc);
    if ((A1(ice->mask_land, vi) == 1 && A1(ice->mask_ocean, vj) == 1) || (A1(ice->mask_land, vj) == 1 && A1(ice->mask_ocean, vi) == 1)) { A1(ice->mask_Ac, aci) = type_coast; A1(ice->mask_coast_Ac, aci) = 1; }
    if ((A1(ice->mask_ice, vi) == 1 && A1(ice->mask_ice, vj) == 0) || (A1(ice->mask_ice, vj) == 1 && A1(ice->mask_ice, vi) == 0)) { A1(ice->mask_Ac, aci) = type_margin; A1(ice->mask_margin_Ac, aci) = 1; }
    if ((A1(ice->mask_sheet, vi) == 1 && A1(ice->mask_shelf, vj) == 1) || (A1(ice->mask_sheet, vj) == 1 && A1(ice->mask_shelf, vi) == 1)) { A1(ice->mask_Ac, aci) = type_groundingline; A1(ice->mask_gl_Ac, aci) = 1; }
    if ((A1(ice->mask_ice, vi) == 1 && A1(ice->mask_shelf, vj) == 0 && A1(ice->mask_ocean, vj) == 1) ||
        (A1(ice->mask_ice, vj) == 1 && A1(ice->mask_shelf, vi) == 0 && A1(ice->mask_ocean, vi) == 1)) { A1(ice->mask_Ac, aci) = type_calvingfront; A1(ice->mask_cf_Ac, aci) = 1; }
  }
  SYNC
}

/* ice_physical_properties, src/general_ice_model_data_module.f90:298-462 (rank body) */
static void ice_physical_properties_r(const ora_mesh *m, const rank_t *r, ora_ice *ice, const ora_config *c, double time)
{
  const double A_low_temp = 1.14E-05, A_high_temp = 5.47E+10, Q_low_temp = 6.0E+04, Q_high_temp = 13.9E+04, R_gas = 8.314;
  int nV = m->nV, nAc = m->nAc, nZ = c->nZ;
  if (c->benchmark != ORA_BM_NONE) {
    double A_flow;
    if (is_benchmark_simple(c->benchmark) || c->benchmark == ORA_BM_MESH_GENERATION_TEST) A_flow = 1.0E-16;
    else { /* MISMIP_mod (and our synthetic SSA_icestream), :351-368 */
      A_flow = 1.0E-16;
      if (time < 25000.0) A_flow = 1.0E-16; else if (time < 50000.0) A_flow = 1.0E-17; else if (time < 75000.0) A_flow = 1.0E-16;
    }
    for (int k = 1; k <= nZ; k++) {
      for (int vi = r->v1; vi <= r->v2; vi++) A2(ice->A_flow, vi, k, nV) = A_flow;
      for (int aci = r->ac1; aci <= r->ac2; aci++) A2(ice->A_flow_Ac, aci, k, nAc) = A_flow;
    }
    for (int vi = r->v1; vi <= r->v2; vi++) A1(ice->A_flow_mean, vi) = A_flow;
    for (int aci = r->ac1; aci <= r->ac2; aci++) A1(ice->A_flow_mean_Ac, aci) = A_flow;
    if (is_benchmark_simple(c->benchmark) || c->benchmark == ORA_BM_MESH_GENERATION_TEST) { /* :336-343 */
      for (int k = 1; k <= nZ; k++)
        for (int vi = r->v1; vi <= r->v2; vi++) {
          A2(ice->Ki, vi, k, nV) = 2.1 * sec_per_year;
          A2(ice->Cpi, vi, k, nV) = 2009.0;
          A2(ice->Ti_pmp, vi, k, nV) = T0 - (c->zeta[k - 1] * A1(ice->Hi, vi) * 8.7E-04);
        }
    }
    SYNC
    return;
  }
  double prof[32];
  for (int vi = r->v1; vi <= r->v2; vi++) {
    for (int k = 1; k <= nZ; k++) {
      double Ti = A2(ice->Ti, vi, k, nV);
      A2(ice->Ti_pmp, vi, k, nV) = T0 - CC * A1(ice->Hi, vi) * c->zeta[k - 1];   /* :388 */
      if (Ti < 263.15) A2(ice->A_flow, vi, k, nV) = A_low_temp * exp(-Q_low_temp / (R_gas * Ti));
      else             A2(ice->A_flow, vi, k, nV) = A_high_temp * exp(-Q_high_temp / (R_gas * Ti));
      A2(ice->Cpi, vi, k, nV) = 2115.3 + 7.79293 * (Ti - T0);                     /* :401 */
      A2(ice->Ki, vi, k, nV) = 3.101E+08 * exp(-0.0057 * Ti);                     /* :404 */
    }
    if (A1(ice->mask_sheet, vi) == 1) {
      for (int k = 1; k <= nZ; k++) prof[k - 1] = A2(ice->A_flow, vi, k, nV);
      A1(ice->A_flow_mean, vi) = vertical_average(c, prof);
    } else {
      double Ti_mean = (A2(ice->Ti, vi, 1, nV) + SMT) / 2.0;
      if (Ti_mean < 263.15) A1(ice->A_flow_mean, vi) = A_low_temp * exp(-Q_low_temp / (R_gas * Ti_mean));
      else                  A1(ice->A_flow_mean, vi) = A_high_temp * exp(-Q_high_temp / (R_gas * Ti_mean));
    }
  }
  SYNC
  for (int ci = r->ac1; ci <= r->ac2; ci++) {
    for (int k = 1; k <= nZ; k++) {
      double Ti = A2(ice->Ti_Ac, ci, k, nAc);
      if (Ti < 263.15) A2(ice->A_flow_Ac, ci, k, nAc) = A_low_temp * exp(-Q_low_temp / (R_gas * Ti));
      else             A2(ice->A_flow_Ac, ci, k, nAc) = A_high_temp * exp(-Q_high_temp / (R_gas * Ti));
    }
    if (A1(ice->mask_sheet_Ac, ci) == 1) {
      for (int k = 1; k <= nZ; k++) prof[k - 1] = A2(ice->A_flow_Ac, ci, k, nAc);
      A1(ice->A_flow_mean_Ac, ci) = vertical_average(c, prof);
    } else {
      double Ti_mean = (A2(ice->Ti_Ac, ci, 1, nAc) + SMT) / 2.0;
      if (Ti_mean < 263.15) A1(ice->A_flow_mean_Ac, ci) = A_low_temp * exp(-Q_low_temp / (R_gas * Ti_mean));
      else                  A1(ice->A_flow_mean_Ac, ci) = A_high_temp * exp(-Q_high_temp / (R_gas * Ti_mean));
    }
  }
  SYNC
}

/* update_general_ice_model_data, src/general_ice_model_data_module.f90:23-96 */
void ora_update_general_ice_model_data(const ora_mesh *m, ora_ice *ice, const ora_config *c, double time)
{
#pragma omp parallel num_threads(c->nthreads)
  {
    rank_t r = rank_of(m);
    map_Aa_to_Ac_r(m, &r, ice->Hi, ice->Hi_Ac);
    map_Aa_to_Ac_r(m, &r, ice->Hb, ice->Hb_Ac);
    map_Aa_to_Ac_r(m, &r, ice->SL, ice->SL_Ac);
    map_Aa_to_Ac_3D_r(m, &r, c->nZ, ice->Ti, ice->Ti_Ac);
    for (int vi = r.v1; vi <= r.v2; vi++)
      A1(ice->Hs, vi) = A1(ice->Hi, vi) + fmax(A1(ice->SL, vi) - ice_density / seawater_density * A1(ice->Hi, vi), A1(ice->Hb, vi));
    SYNC
    for (int aci = r.ac1; aci <= r.ac2; aci++)
      A1(ice->Hs_Ac, aci) = A1(ice->Hi_Ac, aci) + fmax(A1(ice->SL_Ac, aci) - ice_density / seawater_density * A1(ice->Hi_Ac, aci), A1(ice->Hb_Ac, aci));
    for (int vi = r.v1; vi <= r.v2; vi++) A1(ice->dHs_dt, vi) = A1(ice->dHb_dt, vi) + A1(ice->dHi_dt, vi);
    determine_masks_r(m, &r, ice);
    get_mesh_derivatives_r(m, &r, ice->Hi, ice->dHi_dx, ice->dHi_dy);
    get_mesh_derivatives_r(m, &r, ice->Hs, ice->dHs_dx, ice->dHs_dy);
    get_mesh_derivatives_Ac_r(m, &r, ice->Hi, ice->dHi_dx_Ac, ice->dHi_dy_Ac, ice->dHi_dp_Ac, ice->dHi_do_Ac);
    get_mesh_derivatives_Ac_r(m, &r, ice->Hb, ice->dHb_dx_Ac, ice->dHb_dy_Ac, ice->dHb_dp_Ac, ice->dHb_do_Ac);
    get_mesh_derivatives_Ac_r(m, &r, ice->Hs, ice->dHs_dx_Ac, ice->dHs_dy_Ac, ice->dHs_dp_Ac, ice->dHs_do_Ac);
    get_mesh_derivatives_Ac_r(m, &r, ice->SL, ice->dSL_dx_Ac, ice->dSL_dy_Ac, ice->dSL_dp_Ac, ice->dSL_do_Ac);
    for (int vi = r.v1; vi <= r.v2; vi++) {
      if (A1(ice->mask_ocean, vi) == 0) { A1(ice->dHs_dx_shelf, vi) = A1(ice->dHs_dx, vi); A1(ice->dHs_dy_shelf, vi) = A1(ice->dHs_dy, vi); }
      else {
        A1(ice->dHs_dx_shelf, vi) = (1.0 - ice_density / seawater_density) * A1(ice->dHi_dx, vi);
        A1(ice->dHs_dy_shelf, vi) = (1.0 - ice_density / seawater_density) * A1(ice->dHi_dy, vi);
      }
    }
    for (int aci = r.ac1; aci <= r.ac2; aci++) {
      if (A1(ice->mask_ocean_Ac, aci) == 0) { A1(ice->dHs_dx_shelf_Ac, aci) = A1(ice->dHs_dx_Ac, aci); A1(ice->dHs_dy_shelf_Ac, aci) = A1(ice->dHs_dy_Ac, aci); }
      else {
        A1(ice->dHs_dx_shelf_Ac, aci) = (1.0 - ice_density / seawater_density) * A1(ice->dHi_dx_Ac, aci);
        A1(ice->dHs_dy_shelf_Ac, aci) = (1.0 - ice_density / seawater_density) * A1(ice->dHi_dy_Ac, aci);
      }
    }
    SYNC
    ice_physical_properties_r(m, &r, ice, c, time);
  }
}

/* ============================================================================================
 * solve_SIA, src/ice_dynamics_module.f90:240-314
 * ============================================================================================ */
void ora_solve_SIA(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
  const double D_uv_3D_cutoff = -1E5;
  int nAc = m->nAc, nZ = c->nZ;
#pragma omp parallel num_threads(c->nthreads)
  {
    rank_t r = rank_of(m);
    double f[32], D_deformation[32], prof[32];
    for (int k = 1; k <= nZ; k++) for (int aci = r.ac1; aci <= r.ac2; aci++) A2(ice->D_SIA_3D_Ac, aci, k, nAc) = 0.0;
    for (int aci = r.ac1; aci <= r.ac2; aci++) {
      A1(ice->D_SIA_Ac, aci) = 0.0; A1(ice->Ux_SIA_Ac, aci) = 0.0; A1(ice->Uy_SIA_Ac, aci) = 0.0;
      A1(ice->Up_SIA_Ac, aci) = 0.0; A1(ice->Uo_SIA_Ac, aci) = 0.0;
    }
    for (int vi = r.v1; vi <= r.v2; vi++) { A1(ice->D_SIA, vi) = 0.0; A1(ice->U_SIA, vi) = 0.0; A1(ice->V_SIA, vi) = 0.0; }
    SYNC
    for (int aci = r.ac1; aci <= r.ac2; aci++) {
      if (A1(ice->mask_sheet_Ac, aci) == 1) {
        double dp_ = A1(ice->dHs_dp_Ac, aci), do_ = A1(ice->dHs_do_Ac, aci);
        double D_0 = pow(ice_density * grav * A1(ice->Hi_Ac, aci), n_flow) * pow((dp_ * dp_ + do_ * do_), (n_flow - 1.0) / 2.0);
        for (int k = 1; k <= nZ; k++) f[k - 1] = c->m_enh_sia * A2(ice->A_flow_Ac, aci, k, nAc) * pow(c->zeta[k - 1], n_flow);
        vertical_integrate(c, f, D_deformation);
        for (int k = 1; k <= nZ; k++) D_deformation[k - 1] = 2.0 * A1(ice->Hi_Ac, aci) * D_deformation[k - 1];
        for (int k = 1; k <= nZ; k++) A2(ice->D_SIA_3D_Ac, aci, k, nAc) = D_0 * D_deformation[k - 1];
      }
      for (int k = 1; k <= nZ; k++)
        if (A2(ice->D_SIA_3D_Ac, aci, k, nAc) < D_uv_3D_cutoff) A2(ice->D_SIA_3D_Ac, aci, k, nAc) = D_uv_3D_cutoff;
    }
    SYNC
    for (int aci = r.ac1; aci <= r.ac2; aci++) {
      if (A1(ice->mask_sheet_Ac, aci) == 1) {
        for (int k = 1; k <= nZ; k++) prof[k - 1] = A2(ice->D_SIA_3D_Ac, aci, k, nAc);
        double D_uv_2D = vertical_average(c, prof);
        A1(ice->D_SIA_Ac, aci) = A1(ice->Hi_Ac, aci) * D_uv_2D;
        A1(ice->Ux_SIA_Ac, aci) = D_uv_2D * A1(ice->dHs_dx_Ac, aci);
        A1(ice->Uy_SIA_Ac, aci) = D_uv_2D * A1(ice->dHs_dy_Ac, aci);
        A1(ice->Up_SIA_Ac, aci) = D_uv_2D * A1(ice->dHs_dp_Ac, aci);
        A1(ice->Uo_SIA_Ac, aci) = D_uv_2D * A1(ice->dHs_do_Ac, aci);
      }
    }
    SYNC
    map_Ac_to_Aa_r(m, &r, ice->Ux_SIA_Ac, ice->U_SIA);
    map_Ac_to_Aa_r(m, &r, ice->Uy_SIA_Ac, ice->V_SIA);
    map_Ac_to_Aa_r(m, &r, ice->D_SIA_Ac, ice->D_SIA);
  }
}

/* apply_Neumann_boundary_3D, src/mesh_derivatives_module.f90:538-597 (rank body) */
static void apply_Neumann_boundary_3D_r(const ora_mesh *m, const rank_t *r, double *d, int nz)
{
  int nV = m->nV, W = m->nC_mem;
  double vals[64];
  for (int vi = (r->v1 > 5 ? r->v1 : 5); vi <= r->v2; vi++) {
    if (A1(m->edge_index, vi) == 0) continue;
    for (int k = 1; k <= nz; k++) {
      for (int q = 0; q < W; q++) vals[q] = 0.0;
      int nvals = 0;
      for (int ci = 1; ci <= A1(m->nC, vi); ci++) {
        int vc = A2(m->C, vi, ci, nV);
        if (A1(m->edge_index, vc) > 0) continue;
        nvals = nvals + 1;
        vals[nvals - 1] = A2(d, vc, k, nV);
      }
      double s = 0.0;
      for (int q = 0; q < W; q++) s = s + vals[q];
      A2(d, vi, k, nV) = s / nvals;
    }
  }
  SYNC
  if (r->i == 0) {
    for (int vi = 1; vi <= 4; vi++)
      for (int k = 1; k <= nz; k++) {
        for (int q = 0; q < W; q++) vals[q] = 0.0;
        int nvals = 0;
        for (int ci = 1; ci <= A1(m->nC, vi); ci++) { nvals = nvals + 1; vals[nvals - 1] = A2(d, A2(m->C, vi, ci, nV), k, nV); }
        double s = 0.0;
        for (int q = 0; q < W; q++) s = s + vals[q];
        A2(d, vi, k, nV) = s / nvals;
      }
  }
  SYNC
}

/* solve_SIA_3D, src/ice_dynamics_module.f90:317-367: the U_3D / V_3D half (the vertical velocity W_3D, :369-403, only feeds
 * thermodynamics, out of scope).  U_3D / V_3D set the third critical time step (src/UFEMISM_main_model.f90:764-767). */
void ora_solve_SIA_3D_UV(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
  const double D_uv_3D_cutoff = -1E5;
  int nV = m->nV, nZ = c->nZ;
#pragma omp parallel num_threads(c->nthreads)
  {
    rank_t r = rank_of(m);
    double f[32], D_deformation[32];
    for (int k = 1; k <= nZ; k++) for (int vi = r.v1; vi <= r.v2; vi++) { A2(ice->U_3D, vi, k, nV) = 0.0; A2(ice->V_3D, vi, k, nV) = 0.0; }
    SYNC
    for (int vi = r.v1; vi <= r.v2; vi++) {
      for (int k = 1; k <= nZ; k++) { A2(ice->U_3D, vi, k, nV) = A1(ice->U_SSA, vi); A2(ice->V_3D, vi, k, nV) = A1(ice->V_SSA, vi); }
      if (A1(ice->mask_shelf, vi) == 1) continue;
      if (A1(ice->Hi, vi) == 0.0) continue;
      double dx = A1(ice->dHs_dx, vi), dy = A1(ice->dHs_dy, vi);
      double D_0 = pow(ice_density * grav * A1(ice->Hi, vi), n_flow) * pow((dx * dx + dy * dy), (n_flow - 1.0) / 2.0);
      for (int k = 1; k <= nZ; k++) f[k - 1] = c->m_enh_sia * A2(ice->A_flow, vi, k, nV) * pow(c->zeta[k - 1], n_flow);
      vertical_integrate(c, f, D_deformation);
      for (int k = 1; k <= nZ; k++) D_deformation[k - 1] = 2.0 * A1(ice->Hi, vi) * D_deformation[k - 1];
      for (int k = 1; k <= nZ; k++) {
        double D_SIA_3D = fmax(D_0 * D_deformation[k - 1], D_uv_3D_cutoff);
        A2(ice->U_3D, vi, k, nV) = D_SIA_3D * dx + A1(ice->U_SSA, vi);
        A2(ice->V_3D, vi, k, nV) = D_SIA_3D * dy + A1(ice->V_SSA, vi);
      }
    }
    SYNC
    apply_Neumann_boundary_3D_r(m, &r, ice->U_3D, nZ);
    apply_Neumann_boundary_3D_r(m, &r, ice->V_3D, nZ);
  }
}

/* ============================================================================================
 * Thermodynamics: solve_SIA_3D W half, update_ice_temperature and what it calls
 * ============================================================================================ */
/* get_mesh_derivatives_vertex_3D, src/mesh_derivatives_module.f90:372-391 */
static inline void get_mesh_derivatives_vertex_3D(const ora_mesh *m, const double *d, double *ddx, double *ddy, int vi, int k)
{
  int nV = m->nV, n = A1(m->nC, vi);
  double x = A2(m->Nx, vi, n + 1, nV) * A2(d, vi, k, nV);
  double y = A2(m->Ny, vi, n + 1, nV) * A2(d, vi, k, nV);
  for (int ci = 1; ci <= n; ci++) {
    int vc = A2(m->C, vi, ci, nV);
    x = x + A2(m->Nx, vi, ci, nV) * A2(d, vc, k, nV);
    y = y + A2(m->Ny, vi, ci, nV) * A2(d, vc, k, nV);
  }
  *ddx = x; *ddy = y;
}

/* is_in_triangle, src/mesh_help_functions_module.f90:648-672 */
static inline int is_in_triangle(const double *pa, const double *pb, const double *pc, const double *p)
{
  const double tol = 1E-8;
  double as_x = p[0] - pa[0], as_y = p[1] - pa[1];
  double s1 = ((pb[0] - pa[0]) * as_y - (pb[1] - pa[1]) * as_x);
  double s2 = ((pc[0] - pa[0]) * as_y - (pc[1] - pa[1]) * as_x);
  double s3 = ((pc[0] - pb[0]) * (p[1] - pb[1]) - (pc[1] - pb[1]) * (p[0] - pb[0]));
  return (s1 > -tol && s2 < tol && s3 > -tol);
}

/* get_upwind_derivative_vertex_3D, src/mesh_derivatives_module.f90:435-483; returns 0 when no upwind triangle is found */
static int get_upwind_derivative_vertex_3D(const ora_mesh *m, const double *U, const double *V, const double *d, int vi, int k, double *ddx, double *ddy)
{
  int nV = m->nV, nTri = m->nTri;
  double u = A2(U, vi, k, nV), v = A2(V, vi, k, nV);
  *ddx = 0.0; *ddy = 0.0;
  if (fabs(u) < 1E-10 && fabs(v) < 1E-10) { get_mesh_derivatives_vertex_3D(m, d, ddx, ddy, vi, k); return 1; }
  double den = 4.0 * sqrt(u * u + v * v);
  double W[2] = {u * A1(m->R, vi) / den, v * A1(m->R, vi) / den};
  double p[2] = {A2(m->V, vi, 1, nV) - W[0], A2(m->V, vi, 2, nV) - W[1]};
  int tup = 0;
  for (int iti = 1; iti <= A1(m->niTri, vi); iti++) {
    int ti = A2(m->iTri, vi, iti, nV);
    double pa[2], pb[2], pc[2];
    int a = A2(m->Tri, ti, 1, nTri), b = A2(m->Tri, ti, 2, nTri), cc = A2(m->Tri, ti, 3, nTri);
    pa[0] = A2(m->V, a, 1, nV); pa[1] = A2(m->V, a, 2, nV);
    pb[0] = A2(m->V, b, 1, nV); pb[1] = A2(m->V, b, 2, nV);
    pc[0] = A2(m->V, cc, 1, nV); pc[1] = A2(m->V, cc, 2, nV);
    if (is_in_triangle(pa, pb, pc, p)) { tup = ti; break; }
  }
  if (tup == 0) return 0;
  int a = A2(m->Tri, tup, 1, nTri), b = A2(m->Tri, tup, 2, nTri), cc = A2(m->Tri, tup, 3, nTri);
  *ddx = A2(m->NxTri, tup, 1, nTri) * A2(d, a, k, nV) + A2(m->NxTri, tup, 2, nTri) * A2(d, b, k, nV) + A2(m->NxTri, tup, 3, nTri) * A2(d, cc, k, nV);
  *ddy = A2(m->NyTri, tup, 1, nTri) * A2(d, a, k, nV) + A2(m->NyTri, tup, 2, nTri) * A2(d, b, k, nV) + A2(m->NyTri, tup, 3, nTri) * A2(d, cc, k, nV);
  return 1;
}

/* solve_SIA_3D, src/ice_dynamics_module.f90:317-405: U_3D / V_3D as ora_solve_SIA_3D_UV, then the vertical velocity (:369-403) */
void ora_solve_SIA_3D(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
  int nV = m->nV, nZ = c->nZ;
  ora_solve_SIA_3D_UV(m, ice, c);   /* :339-367 incl. W_3D = 0 (below) and the two Neumann passes */
#pragma omp parallel num_threads(c->nthreads)
  {
    rank_t r = rank_of(m);
    for (int k = 1; k <= nZ; k++) for (int vi = r.v1; vi <= r.v2; vi++) A2(ice->W_3D, vi, k, nV) = 0.0;
    SYNC
    for (int vi = r.v1; vi <= r.v2; vi++) {
      if (A1(m->edge_index, vi) > 0) continue;
      if (A1(ice->mask_sheet, vi) == 0) continue;
      double dHb_dx = A1(ice->dHs_dx, vi) - A1(ice->dHi_dx, vi);
      double dHb_dy = A1(ice->dHs_dy, vi) - A1(ice->dHi_dy, vi);
      A2(ice->W_3D, vi, nZ, nV) = A1(ice->dHb_dt, vi) + A2(ice->U_3D, vi, nZ, nV) * dHb_dx + A2(ice->V_3D, vi, nZ, nV) * dHb_dy;
      for (int k = nZ - 1; k >= 1; k--) {
        double dUdx_k, dUdy_k, dUdx_kp1, dUdy_kp1, dVdx_k, dVdy_k, dVdx_kp1, dVdy_kp1;
        get_mesh_derivatives_vertex_3D(m, ice->U_3D, &dUdx_k, &dUdy_k, vi, k);
        get_mesh_derivatives_vertex_3D(m, ice->U_3D, &dUdx_kp1, &dUdy_kp1, vi, k + 1);
        get_mesh_derivatives_vertex_3D(m, ice->V_3D, &dVdx_k, &dVdy_k, vi, k);
        get_mesh_derivatives_vertex_3D(m, ice->V_3D, &dVdx_kp1, &dVdy_kp1, vi, k + 1);
        double zk = c->zeta[k - 1], zk1 = c->zeta[k];
        double w1 = (dUdx_k + dUdx_kp1) / 2.0;
        double w2 = (dVdy_k + dVdy_kp1) / 2.0;
        double w3 = ((A1(ice->dHs_dx, vi) - 0.5 * (zk1 + zk) * A1(ice->dHi_dx, vi)) / fmax(0.1, A1(ice->Hi, vi))) *
                    ((A2(ice->U_3D, vi, k + 1, nV) - A2(ice->U_3D, vi, k, nV)) / (zk1 - zk));
        double w4 = ((A1(ice->dHs_dy, vi) - 0.5 * (zk1 + zk) * A1(ice->dHi_dy, vi)) / fmax(0.1, A1(ice->Hi, vi))) *
                    ((A2(ice->V_3D, vi, k + 1, nV) - A2(ice->V_3D, vi, k, nV)) / (zk1 - zk));
        A2(ice->W_3D, vi, k, nV) = A2(ice->W_3D, vi, k + 1, nV) - A1(ice->Hi, vi) * (w1 + w2 + w3 + w4) * (zk1 - zk);
      }
    }
    SYNC
    apply_Neumann_boundary_3D_r(m, &r, ice->W_3D, nZ);
  }
}

/* LAPACK DGTSV, NRHS = 1 (netlib reference implementation; the reference links the system LAPACK,
 * src/thermodynamics_module.f90:339,348): Gaussian elimination with partial pivoting on a tridiagonal system. */
void ora_dgtsv(int n, double *dl, double *d, double *du, double *b, int *info)
{
  *info = 0;
  if (n == 0) return;
#define DL(i) dl[(i) - 1]
#define D(i) d[(i) - 1]
#define DU(i) du[(i) - 1]
#define B(i) b[(i) - 1]
  for (int i = 1; i <= n - 2; i++) {
    if (fabs(D(i)) >= fabs(DL(i))) {
      if (D(i) != 0.0) {
        double fact = DL(i) / D(i);
        D(i + 1) = D(i + 1) - fact * DU(i);
        B(i + 1) = B(i + 1) - fact * B(i);
      } else { *info = i; return; }
      DL(i) = 0.0;
    } else {
      double fact = D(i) / DL(i);
      D(i) = DL(i);
      double temp = D(i + 1);
      D(i + 1) = DU(i) - fact * temp;
      DL(i) = DU(i + 1);
      DU(i + 1) = -fact * DL(i);
      DU(i) = temp;
      temp = B(i);
      B(i) = B(i + 1);
      B(i + 1) = temp - fact * B(i + 1);
    }
  }
  if (n > 1) {
    int i = n - 1;
    if (fabs(D(i)) >= fabs(DL(i))) {
      if (D(i) != 0.0) {
        double fact = DL(i) / D(i);
        D(i + 1) = D(i + 1) - fact * DU(i);
        B(i + 1) = B(i + 1) - fact * B(i);
      } else { *info = i; return; }
    } else {
      double fact = D(i) / DL(i);
      D(i) = DL(i);
      double temp = D(i + 1);
      D(i + 1) = DU(i) - fact * temp;
      DU(i) = temp;
      temp = B(i);
      B(i) = B(i + 1);
      B(i + 1) = temp - fact * B(i + 1);
    }
  }
  if (D(n) == 0.0) { *info = n; return; }
  B(n) = B(n) / D(n);
  if (n > 1) B(n - 1) = (B(n - 1) - DU(n - 1) * B(n)) / D(n - 1);
  for (int i = n - 2; i >= 1; i--) B(i) = (B(i) - DU(i) * B(i + 1) - DL(i) * B(i + 2)) / D(i);
#undef DL
#undef D
#undef DU
#undef B
}

/* replace_Ti_with_robin_solution, src/thermodynamics_module.f90:204-279 */
void ora_replace_Ti_with_robin_solution(const ora_mesh *m, ora_ice *ice, const ora_config *c, int vi)
{
  const double kappa_0_ice_conductivity = 9.828, kappa_e_ice_conductivity = 0.0057, c_0_specific_heat = 2127.5, Claus_Clap_gradient = 8.7E-04;
  int nV = m->nV, nZ = c->nZ;
  double thermal_conductivity_robin = kappa_0_ice_conductivity * sec_per_year * exp(-kappa_e_ice_conductivity * T0);
  double thermal_diffusivity_robin = thermal_conductivity_robin / (ice_density * c_0_specific_heat);
  double bottom_temperature_gradient_robin = -A1(ice->GHF, vi) / thermal_conductivity_robin;
  double sT = 0.0;
  for (int mo = 1; mo <= 12; mo++) sT = sT + A2(ice->T2m, vi, mo, nV);
  double Ts = fmin(T0, sT / 12.0);
  double Hi = A1(ice->Hi, vi);
  if (A1(ice->mask_sheet, vi) == 1) {
    if (A1(ice->SMB_year, vi) > 0.0) {
      double thermal_length_scale = sqrt(2.0 * thermal_diffusivity_robin * Hi / A1(ice->SMB_year, vi));
      for (int k = 1; k <= nZ; k++) {
        double distance_above_bed = (1.0 - c->zeta[k - 1]) * Hi;
        double erf1 = erf(distance_above_bed / thermal_length_scale);
        double erf2 = erf(Hi / thermal_length_scale);
        A2(ice->Ti, vi, k, nV) = Ts + sqrt(pi) / 2.0 * thermal_length_scale * bottom_temperature_gradient_robin * (erf1 - erf2);
      }
    } else {
      for (int k = 1; k <= nZ; k++) A2(ice->Ti, vi, k, nV) = Ts + ((T0 - Claus_Clap_gradient * Hi) - Ts) * c->zeta[k - 1];
    }
  } else if (A1(ice->mask_shelf, vi) == 1) {
    for (int k = 1; k <= nZ; k++) A2(ice->Ti, vi, k, nV) = Ts + c->zeta[k - 1] * (SMT - Ts);
  } else {
    for (int k = 1; k <= nZ; k++) A2(ice->Ti, vi, k, nV) = Ts;
  }
  for (int k = 1; k <= nZ; k++) A2(ice->Ti, vi, k, nV) = fmin(A2(ice->Ti, vi, k, nV), T0 - Claus_Clap_gradient * Hi * c->zeta[k - 1]);
}

/* update_ice_temperature, src/thermodynamics_module.f90:23-202, with bottom_frictional_heating (:281-311),
 * calculate_zeta_derivatives (src/zeta_module.f90:86-112) and the coefficients of initialize_zeta_discretization (:113-173) */
int ora_update_ice_temperature(const ora_mesh *m, ora_ice *ice, const ora_config *c, int *n_unstable_out)
{
  int nV = m->nV, nZ = c->nZ;
  *n_unstable_out = 0;
  if (c->benchmark == ORA_BM_MISMIP_MOD || c->benchmark == ORA_BM_MESH_GENERATION_TEST || c->benchmark == ORA_BM_HALFAR ||
      c->benchmark == ORA_BM_BUELER || c->benchmark == ORA_BM_SSA_ICESTREAM) return 0;   /* :44-64 */
  /* initialize_zeta_discretization */
  double a_k[33], b_k[33], a_zeta[33], b_zeta[33], c_zeta[33], a_zetazeta[33], b_zetazeta[33], c_zetazeta[33];
  for (int k = 2; k <= nZ; k++) a_k[k] = c->zeta[k - 1] - c->zeta[k - 2];
  for (int k = 1; k <= nZ - 1; k++) b_k[k] = c->zeta[k] - c->zeta[k - 1];
  for (int k = 2; k <= nZ - 1; k++) {
    a_zeta[k] = -b_k[k] / (a_k[k] * (a_k[k] + b_k[k]));
    b_zeta[k] = (b_k[k] - a_k[k]) / (a_k[k] * b_k[k]);
    c_zeta[k] = a_k[k] / (b_k[k] * (a_k[k] + b_k[k]));
    a_zetazeta[k] = 2.0 / (a_k[k] * (a_k[k] + b_k[k]));
    b_zetazeta[k] = -2.0 / (a_k[k] * b_k[k]);
    c_zetazeta[k] = 2.0 / (b_k[k] * (a_k[k] + b_k[k]));
  }
  ora_solve_SIA_3D(m, ice, c);
  int rc = 0, n_unstable = 0;
#pragma omp parallel num_threads(c->nthreads) reduction(+ : n_unstable)
  {
    rank_t r = rank_of(m);
    for (int k = 1; k <= nZ; k++) for (int vi = r.v1; vi <= r.v2; vi++) A2(ice->Ti_new, vi, k, nV) = 0.0;
    SYNC
    /* bottom_frictional_heating */
    {
      const double delta_v = 1E-3, q_plastic = 0.30, u_threshold = 100.0;
      for (int vi = r.v1; vi <= r.v2; vi++) A1(ice->frictional_heating, vi) = 0.0;
      SYNC
      for (int vi = r.v1; vi <= r.v2; vi++) {
        if (A1(ice->mask_sheet, vi) == 1) {
          double u = A1(ice->U_SSA, vi), v = A1(ice->V_SSA, vi);
          double beta_base = A1(ice->tau_c_AaAc, vi) * (pow(delta_v * delta_v + u * u + v * v, 0.5 * (q_plastic - 1.0))) / pow(u_threshold, q_plastic);
          A1(ice->frictional_heating, vi) = beta_base * (u * u + v * v);
        }
      }
      SYNC
    }
    /* calculate_zeta_derivatives */
    for (int vi = r.v1; vi <= r.v2; vi++) {
      double inverse_Hi = 1.0 / fmax(0.1, A1(ice->Hi, vi));
      A1(ice->dzeta_dz, vi) = -inverse_Hi;
      for (int k = 1; k <= nZ; k++) {
        A2(ice->dzeta_dt, vi, k, nV) = inverse_Hi * (A1(ice->dHs_dt, vi) - c->zeta[k - 1] * A1(ice->dHi_dt, vi));
        A2(ice->dzeta_dx, vi, k, nV) = inverse_Hi * (A1(ice->dHs_dx, vi) - c->zeta[k - 1] * A1(ice->dHi_dx, vi));
        A2(ice->dzeta_dy, vi, k, nV) = inverse_Hi * (A1(ice->dHs_dy, vi) - c->zeta[k - 1] * A1(ice->dHi_dy, vi));
      }
    }
    SYNC
    /* surface temperature = annual mean 2 m air temperature (:73-77) */
    for (int vi = r.v1; vi <= r.v2; vi++) {
      double sT = 0.0;
      for (int mo = 1; mo <= 12; mo++) sT = sT + A2(ice->T2m, vi, mo, nV);
      A2(ice->Ti, vi, 1, nV) = fmin(T0, sT / 12.0);
    }
    SYNC
    /* heat equation, one column per vertex (:80-172) */
    double alpha[33], beta[33], gamma[33], delta[33], dl[33], dd[33], du[33], x[33];
    for (int vi = r.v1; vi <= r.v2; vi++) {
      if (A1(m->edge_index, vi) > 0) continue;
      if (A1(ice->mask_ice, vi) == 0) {
        for (int k = 1; k <= nZ; k++) A2(ice->Ti_new, vi, k, nV) = A2(ice->Ti, vi, 1, nV);
        continue;
      }
      beta[1] = 1.0; gamma[1] = 0.0; delta[1] = A2(ice->Ti, vi, 1, nV);
      for (int k = 2; k <= nZ - 1; k++) {
        double dTi_dx, dTi_dy, internal_heating;
        if (!get_upwind_derivative_vertex_3D(m, ice->U_3D, ice->V_3D, ice->Ti, vi, k, &dTi_dx, &dTi_dy)) {
#pragma omp atomic write
          rc = -3;
          dTi_dx = 0.0; dTi_dy = 0.0;
        }
        double Uk = A2(ice->U_3D, vi, k, nV), Vk = A2(ice->V_3D, vi, k, nV);
        if (A1(ice->mask_sheet, vi) == 1) {
          internal_heating = ((-grav * c->zeta[k - 1]) / A2(ice->Cpi, vi, k, nV)) * (
              (a_zeta[k] * A2(ice->U_3D, vi, k - 1, nV) + b_zeta[k] * Uk + c_zeta[k] * A2(ice->U_3D, vi, k + 1, nV)) * A1(ice->dHs_dx, vi) +
              (a_zeta[k] * A2(ice->V_3D, vi, k - 1, nV) + b_zeta[k] * Vk + c_zeta[k] * A2(ice->V_3D, vi, k + 1, nV)) * A1(ice->dHs_dy, vi));
        } else internal_heating = 0.0;
        double dz = A1(ice->dzeta_dz, vi);
        double f1 = (A2(ice->Ki, vi, k, nV) * (dz * dz)) / (ice_density * A2(ice->Cpi, vi, k, nV));
        double f2 = A2(ice->dzeta_dt, vi, k, nV) + A2(ice->dzeta_dx, vi, k, nV) * Uk + A2(ice->dzeta_dy, vi, k, nV) * Vk + dz * A2(ice->W_3D, vi, k, nV);
        double f3 = internal_heating + (Uk * dTi_dx + Vk * dTi_dy) - A2(ice->Ti, vi, k, nV) / c->dt_thermo;
        alpha[k] = f1 * a_zetazeta[k] - f2 * a_zeta[k];
        beta[k] = f1 * b_zetazeta[k] - f2 * b_zeta[k] - 1.0 / c->dt_thermo;
        gamma[k] = f1 * c_zetazeta[k] - f2 * c_zeta[k];
        delta[k] = f3;
      }
      double bottom_flux = (c->zeta[nZ - 1] - c->zeta[nZ - 2]) * (A1(ice->GHF, vi) + A1(ice->frictional_heating, vi)) / (A1(ice->dzeta_dz, vi) * A2(ice->Ki, vi, nZ, nV));
      if (A1(ice->mask_shelf, vi) == 1 || A1(ice->mask_gl, vi) == 1) {
        alpha[nZ] = 0.0; beta[nZ] = 1.0; delta[nZ] = SMT;
      } else {
        alpha[nZ] = 1.0; beta[nZ] = -1.0; delta[nZ] = bottom_flux;
        if (A2(ice->Ti, vi, nZ, nV) >= A2(ice->Ti_pmp, vi, nZ, nV)) { alpha[nZ] = 0.0; beta[nZ] = 1.0; delta[nZ] = A2(ice->Ti_pmp, vi, nZ, nV); }
      }
      /* tridiagonal_solve( alpha, beta, gamma, delta): ldiag = alpha(2:NZ), diag = beta, udiag = gamma(1:NZ-1) */
      for (int k = 1; k <= nZ; k++) { dd[k - 1] = beta[k]; x[k - 1] = delta[k]; }
      for (int k = 1; k <= nZ - 1; k++) { dl[k - 1] = alpha[k + 1]; du[k - 1] = gamma[k]; }
      int info;
      ora_dgtsv(nZ, dl, dd, du, x, &info);
      if (info != 0) {
#pragma omp atomic write
        rc = -2;
      }
      for (int k = 1; k <= nZ; k++) A2(ice->Ti_new, vi, k, nV) = x[k - 1];
      for (int k = 1; k <= nZ - 1; k++) A2(ice->Ti_new, vi, k, nV) = fmin(A2(ice->Ti_new, vi, k, nV), A2(ice->Ti_pmp, vi, k, nV));
      if (A2(ice->Ti_new, vi, nZ, nV) >= A2(ice->Ti_pmp, vi, nZ, nV))
        A2(ice->Ti_new, vi, nZ, nV) = fmin(A2(ice->Ti_pmp, vi, nZ, nV), A2(ice->Ti, vi, nZ - 1, nV) - bottom_flux);
    }
    SYNC
    apply_Neumann_boundary_3D_r(m, &r, ice->Ti_new, nZ);
    SYNC
    for (int k = 1; k <= nZ; k++) for (int vi = r.v1; vi <= r.v2; vi++) A2(ice->Ti, vi, k, nV) = A2(ice->Ti_new, vi, k, nV);
    SYNC
    /* safety net (:181-200) */
    for (int vi = r.v1; vi <= r.v2; vi++) {
      double mn = A2(ice->Ti, vi, 1, nV);
      for (int k = 2; k <= nZ; k++) mn = fmin(mn, A2(ice->Ti, vi, k, nV));
      if (mn < 150.0) { ora_replace_Ti_with_robin_solution(m, ice, c, vi); n_unstable = n_unstable + 1; }
    }
  }
  *n_unstable_out = n_unstable;
  if (rc) return rc;
  if (n_unstable > (int)ceil((double)(float)m->nV / 100.0)) return -1;
  return 0;
}

/* ============================================================================================
 * SSA
 * ============================================================================================ */
/* basal_yield_stress, src/ice_dynamics_module.f90:780-844 (rank body) */
static void basal_yield_stress_r(const ora_mesh *m, const rank_t *r, ora_ice *ice)
{
  const double pf1 = -1000.0, pf2 = 0.0, p_min = 5.0, p_max = 20.0;
  for (int vi = r->v1; vi <= r->v2; vi++) {
    int ai = vi;
    double lambda_p = fmax(0.0, fmin(1.0, (1.0 - (A1(ice->Hb, vi) - A1(ice->SL, vi)) / 1000.0)));
    double pore_water_pressure = 0.96 * ice_density * grav * fmax(0.1, A1(ice->Hi, vi)) * lambda_p;
    A1(ice->phi_fric_AaAc, ai) = fmax(p_min, fmin(p_max, (p_min + (p_max - p_min) * (1.0 + (A1(ice->Hb, vi) - pf2) / (pf2 - pf1)))));
    A1(ice->tau_c_AaAc, ai) = tan((pi / 180.0) * A1(ice->phi_fric_AaAc, ai)) * (ice_density * grav * fmax(0.1, A1(ice->Hi, vi)) - pore_water_pressure);
  }
  SYNC
  for (int aci = r->ac1; aci <= r->ac2; aci++) {
    int ai = aci + m->nV;
    double lambda_p = fmax(0.0, fmin(1.0, (1.0 - (A1(ice->Hb_Ac, aci) - A1(ice->SL_Ac, aci)) / 1000.0)));
    double pore_water_pressure = 0.96 * ice_density * grav * fmax(0.1, A1(ice->Hi_Ac, aci)) * lambda_p;
    A1(ice->phi_fric_AaAc, ai) = fmax(p_min, fmin(p_max, (p_min + (p_max - p_min) * (1.0 + (A1(ice->Hb_Ac, aci) - pf2) / (pf2 - pf1)))));
    A1(ice->tau_c_AaAc, ai) = tan((pi / 180.0) * A1(ice->phi_fric_AaAc, ai)) * (ice_density * grav * fmax(0.1, A1(ice->Hi_Ac, aci)) - pore_water_pressure);
  }
  SYNC
}

/* calculate_GL_flux, src/ice_dynamics_module.f90:848-949 (Coulomb_regularised branch; every rank
 * loops over ALL Ac vertices in the reference, :878 -- idempotent, so one rank does it here) */
static void calculate_GL_flux_r(const ora_mesh *m, const rank_t *r, ora_ice *ice)
{
  const double Q0 = 0.61;
  int nAc = m->nAc, nV = m->nV;
  for (int aci = r->ac1; aci <= r->ac2; aci++) { A1(ice->Qabs_GL_Ac, aci) = 0.0; A1(ice->Qp_GL_Ac, aci) = 0.0; }
  SYNC
  if (r->i == 0) {
    for (int aci = 1; aci <= nAc; aci++) {
      if (A1(ice->mask_gl_Ac, aci) == 0) continue;
      int vi = A2(m->Aci, aci, 1, nAc), vj = A2(m->Aci, aci, 2, nAc);
      double TAFi = A1(ice->Hi, vi) - ((A1(ice->SL, vi) - A1(ice->Hb, vi)) * (seawater_density / ice_density));
      double TAFj = A1(ice->Hi, vj) - ((A1(ice->SL, vj) - A1(ice->Hb, vj)) * (seawater_density / ice_density));
      double lambda_GL = TAFi / (TAFi - TAFj);
      double Hi_GL = (A1(ice->Hi, vi) * (1.0 - lambda_GL)) + (A1(ice->Hi, vj) * lambda_GL);
      double phi_fric_GL, A_flow_GL;
      if (A1(ice->mask_sheet, vi) == 1) { phi_fric_GL = A1(ice->phi_fric_AaAc, vi); A_flow_GL = A1(ice->A_flow_mean, vi); }
      else { phi_fric_GL = A1(ice->phi_fric_AaAc, vj); A_flow_GL = A1(ice->A_flow_mean, vj); }
      double factor_Tsai = (8.0 * Q0 * A_flow_GL * pow(ice_density * grav, n_flow) *
                            pow(1.0 - (ice_density / seawater_density), n_flow - 1.0) / pow(4.0, n_flow));
      A1(ice->Qabs_GL_Ac, aci) = factor_Tsai * pow(Hi_GL, n_flow + 2.0) / tan(phi_fric_GL * (pi / 180.0));
      double Fx = -(A1(ice->dHi_dx_Ac, aci) - ((A1(ice->dSL_dx_Ac, aci) - A1(ice->dHb_dx_Ac, aci)) * (seawater_density / ice_density)));
      double Fy = -(A1(ice->dHi_dy_Ac, aci) - ((A1(ice->dSL_dy_Ac, aci) - A1(ice->dHb_dy_Ac, aci)) * (seawater_density / ice_density)));
      double F = ora_norm2_2(Fx, Fy); /* NORM2 */
      Fx = Fx / F; Fy = Fy / F;
      A1(ice->Ux_SSA_Ac, aci) = A1(ice->Qabs_GL_Ac, aci) * Fx / Hi_GL;
      A1(ice->Uy_SSA_Ac, aci) = A1(ice->Qabs_GL_Ac, aci) * Fy / Hi_GL;
      double Dx = A2(m->V, vj, 1, nV) - A2(m->V, vi, 1, nV), Dy = A2(m->V, vj, 2, nV) - A2(m->V, vi, 2, nV);
      double D = ora_norm2_2(Dx, Dy);
      Dx = Dx / D; Dy = Dy / D;
      A1(ice->Qp_GL_Ac, aci) = A1(ice->Qabs_GL_Ac, aci) * (Dx * Fx + Dy * Fy);
    }
  }
  SYNC
}

/* gather Aa + Ac fields into the AaAc arrays, src/ice_dynamics_module.f90:478-496 */
static void SSA_gather_AaAc_r(const ora_mesh *m, const rank_t *r, ora_ice *ice)
{
  int nV = m->nV;
  for (int vi = r->v1; vi <= r->v2; vi++) {
    A1(ice->Hi_AaAc, vi) = A1(ice->Hi, vi); A1(ice->Hb_AaAc, vi) = A1(ice->Hb, vi); A1(ice->SL_AaAc, vi) = A1(ice->SL, vi);
    A1(ice->dHs_dx_shelf_AaAc, vi) = A1(ice->dHs_dx_shelf, vi); A1(ice->dHs_dy_shelf_AaAc, vi) = A1(ice->dHs_dy_shelf, vi);
    A1(ice->A_flow_mean_AaAc, vi) = A1(ice->A_flow_mean, vi);
    A1(ice->U_SSA_AaAc, vi) = A1(ice->U_SSA, vi); A1(ice->V_SSA_AaAc, vi) = A1(ice->V_SSA, vi);
  }
  for (int aci = r->ac1; aci <= r->ac2; aci++) {
    int ai = nV + aci;
    A1(ice->Hi_AaAc, ai) = A1(ice->Hi_Ac, aci); A1(ice->Hb_AaAc, ai) = A1(ice->Hb_Ac, aci); A1(ice->SL_AaAc, ai) = A1(ice->SL_Ac, aci);
    A1(ice->dHs_dx_shelf_AaAc, ai) = A1(ice->dHs_dx_shelf_Ac, aci); A1(ice->dHs_dy_shelf_AaAc, ai) = A1(ice->dHs_dy_shelf_Ac, aci);
    A1(ice->A_flow_mean_AaAc, ai) = A1(ice->A_flow_mean_Ac, aci);
    A1(ice->U_SSA_AaAc, ai) = A1(ice->Ux_SSA_Ac, aci); A1(ice->V_SSA_AaAc, ai) = A1(ice->Uy_SSA_Ac, aci);
  }
  SYNC
}

/* SSA_effective_viscosity, src/ice_dynamics_module.f90:695-726 (rank body) */
static void SSA_effective_viscosity_r(const ora_mesh *m, const rank_t *r, ora_ice *ice, const ora_config *c)
{
  const double epsilon_sq_0 = 1E-12;
  for (int ai = r->a1; ai <= r->a2; ai++) get_mesh_derivatives_vertex_AaAc(m, ice->U_SSA_AaAc, &A1(ice->dU_SSA_dx_AaAc, ai), &A1(ice->dU_SSA_dy_AaAc, ai), ai);
  SYNC
  for (int ai = r->a1; ai <= r->a2; ai++) get_mesh_derivatives_vertex_AaAc(m, ice->V_SSA_AaAc, &A1(ice->dV_SSA_dx_AaAc, ai), &A1(ice->dV_SSA_dy_AaAc, ai), ai);
  SYNC
  for (int ai = r->a1; ai <= r->a2; ai++) {
    double ux = A1(ice->dU_SSA_dx_AaAc, ai), uy = A1(ice->dU_SSA_dy_AaAc, ai), vx = A1(ice->dV_SSA_dx_AaAc, ai), vy = A1(ice->dV_SSA_dy_AaAc, ai);
    A1(ice->eta_AaAc, ai) = pow(c->m_enh_ssa * 0.5 * A1(ice->A_flow_mean_AaAc, ai), -1.0 / n_flow) *
                            pow(ux * ux + vy * vy + ux * vy + 0.25 * ((uy + vx) * (uy + vx)) + epsilon_sq_0, (1.0 - n_flow) / (2.0 * n_flow));
    A1(ice->N_AaAc, ai) = A1(ice->eta_AaAc, ai) * fmax(0.1, A1(ice->Hi_AaAc, ai));
  }
  SYNC
}

/* SSA_sliding_term, src/ice_dynamics_module.f90:727-779 (Coulomb_regularised) */
static void SSA_sliding_term_r(const ora_mesh *m, const rank_t *r, ora_ice *ice)
{
  const double delta_v = 1E-3, q_plastic = 0.30, u_threshold = 100.0;
  (void)m;
  for (int ai = r->a1; ai <= r->a2; ai++) {
    double U = A1(ice->U_SSA_AaAc, ai), V = A1(ice->V_SSA_AaAc, ai);
    A1(ice->S_AaAc, ai) = A1(ice->tau_c_AaAc, ai) * (pow(delta_v * delta_v + U * U + V * V, 0.5 * (q_plastic - 1.0))) / (pow(u_threshold, q_plastic));
  }
  SYNC
}

/* solve_SSA_linearised, src/ice_dynamics_module.f90:558-694 (rank body).
 * shared[]: per-rank max residuals (the MPI_ALLREDUCE MAX of :673). */
static void solve_SSA_linearised_r(const ora_mesh *m, const rank_t *r, ora_ice *ice, const ora_config *c, int max_inner, int force_iters,
                                   double *shared, int *n_inner_out, double *max_res_out, int *did_reset_out, int *warn_out)
{
  int M = m->nVAaAc, nV = m->nV;
  for (int ai = r->a1; ai <= r->a2; ai++) {
    A1(ice->RHSx_AaAc, ai) = ice_density * grav * A1(ice->dHs_dx_shelf_AaAc, ai) / A1(ice->eta_AaAc, ai);
    A1(ice->RHSy_AaAc, ai) = ice_density * grav * A1(ice->dHs_dy_shelf_AaAc, ai) / A1(ice->eta_AaAc, ai);
  }
  for (int ai = r->a1; ai <= r->a2; ai++) {
    int n = A1(m->nCAaAc, ai);
    if (!is_floating(A1(ice->Hi_AaAc, ai), A1(ice->Hb_AaAc, ai), A1(ice->SL_AaAc, ai))) {
      A1(ice->eu_i_AaAc, ai) = (4.0 * A2(m->Nxx_AaAc, ai, n + 1, M) + A2(m->Nyy_AaAc, ai, n + 1, M)) - A1(ice->S_AaAc, ai) / (fmax(0.1, A1(ice->Hi_AaAc, ai)) * A1(ice->eta_AaAc, ai));
      A1(ice->ev_i_AaAc, ai) = (4.0 * A2(m->Nyy_AaAc, ai, n + 1, M) + A2(m->Nxx_AaAc, ai, n + 1, M)) - A1(ice->S_AaAc, ai) / (fmax(0.1, A1(ice->Hi_AaAc, ai)) * A1(ice->eta_AaAc, ai));
    } else {
      A1(ice->eu_i_AaAc, ai) = (4.0 * A2(m->Nxx_AaAc, ai, n + 1, M) + A2(m->Nyy_AaAc, ai, n + 1, M));
      A1(ice->ev_i_AaAc, ai) = (4.0 * A2(m->Nyy_AaAc, ai, n + 1, M) + A2(m->Nxx_AaAc, ai, n + 1, M));
    }
  }
  SYNC
  int has_converged = 0, inner_loop_i = 0, did_reset = 0, warn = 0;
  double max_residual_UV = 0.0;
  while (!has_converged && inner_loop_i < max_inner) {
    inner_loop_i = inner_loop_i + 1;
    max_residual_UV = 0.0;
    for (int fci = 1; fci <= 5; fci++) {
      for (int fcvi = r->cv1[fci - 1]; fcvi <= r->cv2[fci - 1]; fcvi++) {
        int ai = A2(m->colour_vi, fcvi, fci, M);
        if (is_edge_AaAc(m, ai)) continue;
        if (c->use_analytical_GL_flux && ai > nV && A1(ice->mask_gl_Ac, ai - nV) == 1) continue;
        double Uxyi = get_mesh_curvature_xy_vertex_AaAc(m, ice->U_SSA_AaAc, ai);
        double Vxyi = get_mesh_curvature_xy_vertex_AaAc(m, ice->V_SSA_AaAc, ai);
        double sumUc = 0.0, sumVc = 0.0;
        for (int ci = 1; ci <= A1(m->nCAaAc, ai); ci++) {
          int ac = A2(m->CAaAc, ai, ci, M);
          sumUc = sumUc + A1(ice->U_SSA_AaAc, ac) * (4.0 * A2(m->Nxx_AaAc, ai, ci, M) + A2(m->Nyy_AaAc, ai, ci, M));
          sumVc = sumVc + A1(ice->V_SSA_AaAc, ac) * (4.0 * A2(m->Nyy_AaAc, ai, ci, M) + A2(m->Nxx_AaAc, ai, ci, M));
        }
        A1(ice->LHSx_AaAc, ai) = sumUc + (3.0 * Vxyi) + (A1(ice->eu_i_AaAc, ai) * A1(ice->U_SSA_AaAc, ai));
        A1(ice->LHSy_AaAc, ai) = sumVc + (3.0 * Uxyi) + (A1(ice->ev_i_AaAc, ai) * A1(ice->V_SSA_AaAc, ai));
        A1(ice->resU_AaAc, ai) = (A1(ice->LHSx_AaAc, ai) - A1(ice->RHSx_AaAc, ai)) / A1(ice->eu_i_AaAc, ai);
        A1(ice->resV_AaAc, ai) = (A1(ice->LHSy_AaAc, ai) - A1(ice->RHSy_AaAc, ai)) / A1(ice->ev_i_AaAc, ai);
        max_residual_UV = fmax(max_residual_UV, fabs(A1(ice->resU_AaAc, ai)));
        max_residual_UV = fmax(max_residual_UV, fabs(A1(ice->resV_AaAc, ai)));
        A1(ice->U_SSA_AaAc, ai) = A1(ice->U_SSA_AaAc, ai) - c->SSA_SOR_omega * A1(ice->resU_AaAc, ai);
        A1(ice->V_SSA_AaAc, ai) = A1(ice->V_SSA_AaAc, ai) - c->SSA_SOR_omega * A1(ice->resV_AaAc, ai);
      }
      SYNC
    }
    apply_Neumann_boundary_AaAc_r(m, r, ice->U_SSA_AaAc);
    apply_Neumann_boundary_AaAc_r(m, r, ice->V_SSA_AaAc);
    /* MPI_ALLREDUCE MAX */
    shared[r->i] = max_residual_UV;
    SYNC
    for (int q = 0; q < r->n; q++) max_residual_UV = fmax(max_residual_UV, shared[q]);
    SYNC
    if (force_iters) continue;
    if (max_residual_UV < c->SSA_max_residual_UV) { did_reset = 0; has_converged = 1; }
    else if (max_residual_UV > 1E6) {
      for (int ai = r->a1; ai <= r->a2; ai++) { A1(ice->U_SSA_AaAc, ai) = 0.0; A1(ice->V_SSA_AaAc, ai) = 0.0; }
      did_reset = 1; has_converged = 1;
    } else if (inner_loop_i == max_inner) { warn = 1; }
  }
  SYNC
  /* every rank holds the same values (the reference's loop variables are rank-local copies too) */
  *n_inner_out = inner_loop_i; *max_res_out = max_residual_UV; *did_reset_out = did_reset; *warn_out = warn;
}

/* solve_SSA, src/ice_dynamics_module.f90:408-557 */
int ora_solve_SSA(const ora_mesh *m, ora_ice *ice, const ora_config *c, ora_ssa_stats *st)
{
  int nV = m->nV;
  memset(st, 0, sizeof(*st));
  int set_zero = 0;
  if (c->benchmark != ORA_BM_NONE) {
    if (is_benchmark_simple(c->benchmark)) set_zero = 1;
    else if (c->benchmark == ORA_BM_MISMIP_MOD || c->benchmark == ORA_BM_MESH_GENERATION_TEST || c->benchmark == ORA_BM_SSA_ICESTREAM) { }
    else { st->rc = -2; return -2; }
  }
  long sum_sheet = 0;
  for (int vi = 1; vi <= nV; vi++) sum_sheet += A1(ice->mask_sheet, vi);
  if (sum_sheet == 0) set_zero = 1;
  if (set_zero) {
    for (int ai = 1; ai <= m->nVAaAc; ai++) { A1(ice->U_SSA_AaAc, ai) = 0.0; A1(ice->V_SSA_AaAc, ai) = 0.0; }
    for (int vi = 1; vi <= nV; vi++) { A1(ice->U_SSA, vi) = 0.0; A1(ice->V_SSA, vi) = 0.0; }
    for (int aci = 1; aci <= m->nAc; aci++) { A1(ice->Ux_SSA_Ac, aci) = 0.0; A1(ice->Uy_SSA_Ac, aci) = 0.0; A1(ice->Up_SSA_Ac, aci) = 0.0; A1(ice->Uo_SSA_Ac, aci) = 0.0; }
    return 0;
  }
  int nth = c->nthreads;
  double *shared = (double *)calloc((size_t)nth * 2 + 2, sizeof(double));
  int rc = 0;
#pragma omp parallel num_threads(nth)
  {
    rank_t r = rank_of(m);
    basal_yield_stress_r(m, &r, ice);
    if (c->use_analytical_GL_flux) calculate_GL_flux_r(m, &r, ice);
    SSA_gather_AaAc_r(m, &r, ice);
    int has_converged = 0, viscosity_iteration_i = 0, did_reset_before = 0, did_reset_now = 0, abort_ = 0;
    int n_inner = 0, warn = 0, n_inner_total = 0, any_warn = 0;
    double max_res = 0.0, RN = 0.0;
    while (!has_converged && viscosity_iteration_i < c->SSA_max_outer_loops && !abort_) {
      viscosity_iteration_i = viscosity_iteration_i + 1;
      for (int ai = r.a1; ai <= r.a2; ai++) A1(ice->N_AaAc_prev, ai) = A1(ice->N_AaAc, ai);
      SYNC
      SSA_effective_viscosity_r(m, &r, ice, c);
      double sum_DN_sq = 0.0, sum_N_sq = 0.0;
      for (int ai = r.a1; ai <= r.a2; ai++) {
        double dN = A1(ice->N_AaAc, ai) - A1(ice->N_AaAc_prev, ai);
        sum_DN_sq = sum_DN_sq + dN * dN;
      }
      for (int ai = r.a1; ai <= r.a2; ai++) sum_N_sq = sum_N_sq + A1(ice->N_AaAc, ai) * A1(ice->N_AaAc, ai);
      shared[2 * r.i] = sum_DN_sq; shared[2 * r.i + 1] = sum_N_sq;
      SYNC
      sum_DN_sq = 0.0; sum_N_sq = 0.0;
      for (int q = 0; q < r.n; q++) { sum_DN_sq += shared[2 * q]; sum_N_sq += shared[2 * q + 1]; }
      SYNC
      RN = sqrt(sum_DN_sq / sum_N_sq);
      if (RN < c->SSA_RN_tol) { has_converged = 1; break; }
      SSA_sliding_term_r(m, &r, ice);
      solve_SSA_linearised_r(m, &r, ice, c, c->SSA_max_inner_loops, 0, shared, &n_inner, &max_res, &did_reset_now, &warn);
      SYNC
      n_inner_total += n_inner; any_warn |= warn;
      if (did_reset_now) { if (!did_reset_before) did_reset_before = 1; else abort_ = 1; }
    }
    for (int vi = r.v1; vi <= r.v2; vi++) { A1(ice->U_SSA, vi) = A1(ice->U_SSA_AaAc, vi); A1(ice->V_SSA, vi) = A1(ice->V_SSA_AaAc, vi); }
    for (int aci = r.ac1; aci <= r.ac2; aci++) { A1(ice->Ux_SSA_Ac, aci) = A1(ice->U_SSA_AaAc, nV + aci); A1(ice->Uy_SSA_Ac, aci) = A1(ice->V_SSA_AaAc, nV + aci); }
    SYNC
    rotate_xy_to_po_r(m, &r, ice->Ux_SSA_Ac, ice->Uy_SSA_Ac, ice->Up_SSA_Ac, ice->Uo_SSA_Ac);
    if (r.i == 0) {
      st->n_outer = viscosity_iteration_i; st->n_inner_total = n_inner_total; st->n_inner_last = n_inner;
      st->did_reset = did_reset_before; st->last_max_residual = max_res; st->last_RN = RN;
      st->rc = abort_ ? -1 : (any_warn ? 1 : 0);
      rc = st->rc;
    }
  }
  free(shared);
  return rc;
}

/* ---- pieces exposed for kernel-level parity tests ---- */
void ora_basal_yield_stress(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); basal_yield_stress_r(m, &r, ice); }
}
void ora_calculate_GL_flux(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); calculate_GL_flux_r(m, &r, ice); }
}
void ora_SSA_gather_AaAc(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); SSA_gather_AaAc_r(m, &r, ice); }
}
void ora_SSA_effective_viscosity(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); SSA_effective_viscosity_r(m, &r, ice, c); }
}
void ora_SSA_sliding_term(const ora_mesh *m, ora_ice *ice, const ora_config *c)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); SSA_sliding_term_r(m, &r, ice); }
}
int ora_solve_SSA_linearised(const ora_mesh *m, ora_ice *ice, const ora_config *c, int max_inner_override, int force_iters,
                             int *n_inner, double *max_residual, int *did_reset)
{
  int nth = c->nthreads, warn = 0;
  double *shared = (double *)calloc((size_t)nth * 2 + 2, sizeof(double));
  int max_inner = max_inner_override > 0 ? max_inner_override : c->SSA_max_inner_loops;
#pragma omp parallel num_threads(nth)
  {
    rank_t r = rank_of(m);
    solve_SSA_linearised_r(m, &r, ice, c, max_inner, force_iters, shared, n_inner, max_residual, did_reset, &warn);
  }
  free(shared);
  return warn;
}
void ora_apply_Neumann_boundary_AaAc(const ora_mesh *m, const ora_config *c, double *d_AaAc)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); apply_Neumann_boundary_AaAc_r(m, &r, d_AaAc); }
}
void ora_get_mesh_derivatives(const ora_mesh *m, const ora_config *c, const double *d, double *ddx, double *ddy)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); get_mesh_derivatives_r(m, &r, d, ddx, ddy); }
}
void ora_map_Ac_to_Aa(const ora_mesh *m, const ora_config *c, const double *d_Ac, double *d_Aa)
{
#pragma omp parallel num_threads(c->nthreads)
  { rank_t r = rank_of(m); map_Ac_to_Aa_r(m, &r, d_Ac, d_Aa); }
}

/* remap_cons_1st_order_2D / remap_cons_2nd_order_2D, src/mesh_mapping_module.f90:3964-3983, 4010-4043 (application only;
 * ddx_src, ddy_src = get_mesh_derivatives of d_src on the source mesh, computed by the caller with ora_get_mesh_derivatives) */
void ora_remap_cons_2D(int order, int nV_dst, const int *vli1, const int *vli2, const int *vi, const double *w0, const double *w1x, const double *w1y,
                       const double *d_src, const double *ddx_src, const double *ddy_src, double *d_dst)
{
  for (int vi_dst = 1; vi_dst <= nV_dst; vi_dst++) {
    A1(d_dst, vi_dst) = 0.0;
    for (int vli = A1(vli1, vi_dst); vli <= A1(vli2, vi_dst); vli++) {
      if (order == 1) A1(d_dst, vi_dst) = A1(d_dst, vi_dst) + (A1(d_src, A1(vi, vli)) * A1(w0, vli));
      else A1(d_dst, vi_dst) = A1(d_dst, vi_dst) + (A1(d_src, A1(vi, vli)) * A1(w0, vli)) + (A1(ddx_src, A1(vi, vli)) * A1(w1x, vli)) + (A1(ddy_src, A1(vi, vli)) * A1(w1y, vli));
    }
  }
}

/* ============================================================================================
 * determine_timesteps_and_actions, critical time steps, src/UFEMISM_main_model.f90:738-778.
 * out3 = {dt_D_2D_min, dt_V_2D_SSA_min, dt_V_3D_SIA_min}, each already times 0.9.
 * NOTE `1E-09` at :752 is a default-REAL (single precision) literal.
 * ============================================================================================ */
void ora_determine_timesteps(const ora_mesh *m, const ora_ice *ice, const ora_config *c, double out3[3])
{
  const double dt_correction_factor = 0.9;
  int nth = c->nthreads, nV = m->nV, nAc = m->nAc;
  double *sh = (double *)malloc(sizeof(double) * 3 * (size_t)nth);
#pragma omp parallel num_threads(nth)
  {
    rank_t r = rank_of(m);
    double dt_D_2D_min = 1000.0, dt_V_3D_SIA_min = 1000.0, dt_V_2D_SSA_min = 1000.0;
    for (int ci = r.ac1; ci <= r.ac2; ci++) {
      int vi = A2(m->Aci, ci, 1, nAc), vj = A2(m->Aci, ci, 2, nAc);
      double dx = A2(m->V, vj, 1, nV) - A2(m->V, vi, 1, nV), dy = A2(m->V, vj, 2, nV) - A2(m->V, vi, 2, nV);
      double dist = sqrt(dx * dx + dy * dy);
      double dt_D_2D = (dist * dist) / (-6.0 * pi * (A1(ice->D_SIA_Ac, ci) - (double)1E-09f));
      dt_D_2D_min = fmin(dt_D_2D, dt_D_2D_min);
      double dt_V_2D_SSA = dist / (fabs(A1(ice->U_SSA, vi)) + fabs(A1(ice->V_SSA, vi)));
      dt_V_2D_SSA_min = fmin(dt_V_2D_SSA, dt_V_2D_SSA_min);
      dt_V_2D_SSA = dist / (fabs(A1(ice->U_SSA, vj)) + fabs(A1(ice->V_SSA, vj)));
      dt_V_2D_SSA_min = fmin(dt_V_2D_SSA, dt_V_2D_SSA_min);
    }
    for (int vi = r.v1; vi <= r.v2; vi++) {
      double dt_V_2D_SSA = sqrt(A1(m->A, vi) / pi) / (fabs(A1(ice->U_SSA, vi)) + fabs(A1(ice->V_SSA, vi)));
      dt_V_2D_SSA_min = fmin(dt_V_2D_SSA, dt_V_2D_SSA_min);
      for (int k = 1; k <= c->nZ; k++) {
        double dt_V_3D_SIA = sqrt(A1(m->A, vi) / pi) / (fabs(A2(ice->U_3D, vi, k, nV)) + fabs(A2(ice->V_3D, vi, k, nV)));
        dt_V_3D_SIA_min = fmin(dt_V_3D_SIA, dt_V_3D_SIA_min);
      }
    }
    sh[3 * r.i] = dt_D_2D_min; sh[3 * r.i + 1] = dt_V_2D_SSA_min; sh[3 * r.i + 2] = dt_V_3D_SIA_min;
  }
  double a = 1000.0, b = 1000.0, d = 1000.0;
  for (int q = 0; q < nth; q++) { a = fmin(a, sh[3 * q]); b = fmin(b, sh[3 * q + 1]); d = fmin(d, sh[3 * q + 2]); }
  free(sh);
  out3[0] = a * dt_correction_factor; out3[1] = b * dt_correction_factor; out3[2] = d * dt_correction_factor;
}

/* ============================================================================================
 * Analytic solutions and benchmark SMB (host side in the reference as well)
 * ============================================================================================ */
/* Halfar_solution, src/reference_fields_module.f90:707-745 */
double ora_Halfar_solution(double H0, double R0, double x, double y, double t)
{
  double A_flow = 1E-16, rho = 910.0, g = 9.81;
  double Gamma = (2.0 / 5.0) * (A_flow / sec_per_year) * pow(rho * g, 3.0);
  double t0 = 1.0 / (18.0 * Gamma) * pow(7.0 / 4.0, 3.0) * (pow(R0, 4.0)) / (pow(H0, 7.0));
  double tp = (t * sec_per_year) + t0;
  double r = sqrt(pow(x, 2.0) + pow(y, 2.0));
  double f1 = pow(t0 / tp, 1.0 / 9.0), f2 = pow(t0 / tp, 1.0 / 18.0), f3 = (r / R0);
  return H0 * f1 * pow(fmax(0.0, (1.0 - pow(f2 * f3, 4.0 / 3.0))), 3.0 / 7.0);
}
/* Bueler_solution, :747-793 */
static double bueler_H(double H0, double R0, double lambda, double x, double y, double t, double *tp_out)
{
  double A_flow = 1E-16, rho = 910.0, g = 9.81, n = 3.0;
  double alpha = (2.0 - (n + 1.0) * lambda) / ((5.0 * n) + 3.0);
  double beta = (1.0 + ((2.0 * n) + 1.0) * lambda) / ((5.0 * n) + 3.0);
  double Gamma = 2.0 / 5.0 * (A_flow / sec_per_year) * pow(rho * g, n);
  double f1 = ((2.0 * n) + 1) / (n + 1.0);
  double f2 = (pow(R0, n + 1.0)) / (pow(H0, (2.0 * n) + 1.0));
  double t0 = (beta / Gamma) * (pow(f1, n)) * f2;
  double tp = t * sec_per_year;
  f1 = pow(tp / t0, -alpha);
  f2 = pow(tp / t0, -beta);
  double f3 = sqrt((pow(x, 2.0)) + (pow(y, 2.0))) / R0;
  double f4 = fmax(0.0, 1.0 - pow(f2 * f3, (n + 1.0) / n));
  if (tp_out) *tp_out = tp;
  return H0 * f1 * pow(f4, n / ((2.0 * n) + 1.0));
}
double ora_Bueler_solution(double H0, double R0, double lambda, double x, double y, double t) { return bueler_H(H0, R0, lambda, x, y, t, 0); }
/* Bueler_solution_MB, src/SMB_module.f90:240-283 */
double ora_Bueler_solution_MB(double H0, double R0, double lambda, double x, double y, double t)
{
  double tp, H = bueler_H(H0, R0, lambda, x, y, t, &tp);
  return (lambda / tp) * H * sec_per_year;
}
/* run_SMB_model benchmark branches, src/SMB_module.f90:55-97, EISMINT_SMB :172-238 */
void ora_run_SMB_benchmark(const ora_mesh *m, ora_ice *ice, const ora_config *c, double time, double H0, double R0, double lambda)
{
  int nV = m->nV, b = c->benchmark;
  if (b >= ORA_BM_EISMINT_1 && b <= ORA_BM_EISMINT_6) {
    double E = 450000.0, S_b = 0.01 / 1000.0, M_max = 0.5;
    if (b == ORA_BM_EISMINT_2) { if (!(time < 0.0)) E = 450000.0 + 100000.0 * sin(2.0 * pi * time / 20000.0); }
    else if (b == ORA_BM_EISMINT_3) { if (!(time < 0.0)) E = 450000.0 + 100000.0 * sin(2.0 * pi * time / 40000.0); }
    else if (b == ORA_BM_EISMINT_4) { M_max = 0.3; E = 999000.0; }
    else if (b == ORA_BM_EISMINT_5) { if (time < 0.0) { M_max = 0.3; E = 999000.0; } else { M_max = 0.3 + 0.2 * sin(2.0 * pi * time / 20000.0); E = 999000.0; } }
    else if (b == ORA_BM_EISMINT_6) { if (time < 0.0) { M_max = 0.3; E = 999000.0; } else { M_max = 0.3 + 0.2 * sin(2.0 * pi * time / 40000.0); E = 999000.0; } }
    for (int vi = 1; vi <= nV; vi++) {
      double dist = ora_norm2_2(A2(m->V, vi, 1, nV), A2(m->V, vi, 2, nV));
      A1(ice->SMB_year, vi) = fmin(M_max, S_b * (E - dist));
    }
  } else if (b == ORA_BM_HALFAR) {
    for (int vi = 1; vi <= nV; vi++) A1(ice->SMB_year, vi) = 0.0;
  } else if (b == ORA_BM_BUELER) {
    for (int vi = 1; vi <= nV; vi++) A1(ice->SMB_year, vi) = ora_Bueler_solution_MB(H0, R0, lambda, A2(m->V, vi, 1, nV), A2(m->V, vi, 2, nV), time);
  } else if (b == ORA_BM_MISMIP_MOD || b == ORA_BM_SSA_ICESTREAM) {
    for (int vi = 1; vi <= nV; vi++) A1(ice->SMB_year, vi) = 0.3;
  } else if (b == ORA_BM_MESH_GENERATION_TEST) {
    for (int vi = 1; vi <= nV; vi++) {
      double R = ora_norm2_2(A2(m->V, vi, 1, nV), A2(m->V, vi, 2, nV));
      if (R < 250000.0) A1(ice->SMB_year, vi) = 0.3; else A1(ice->SMB_year, vi) = fmax(-2.0, 0.3 - (R - 250000.0) / 200000.0);
    }
  }
}

/* ============================================================================================
 * Region time loop for benchmark physics: run_model (src/UFEMISM_main_model.f90:37-214) with
 * determine_timesteps_and_actions (:708-843).  Timers that only pace the loop in the benchmark
 * experiments (thermodynamics, climate, BMB, ELRA, output) are kept as timers; their physics is
 * a no-op for the dynamics (BMB = 0: src/BMB_module.f90:51-69; ELRA: bedrock_ELRA_module:35-52).
 * ============================================================================================ */
void ora_region_init(ora_region *r, double start_time)
{
  /* initialise_model, src/UFEMISM_main_model.f90:352-390 */
  memset(r, 0, sizeof(*r));
  r->time = start_time;
  for (int k = 0; k < ORA_NT; k++) { r->t0[k] = start_time; r->t1[k] = start_time; r->do_[k] = 1; }
  r->dtc[ORA_T_SIA] = 0; r->dtc[ORA_T_SSA] = 0; r->dtc[ORA_T_THERMO] = 10.0; r->dtc[ORA_T_CLIMATE] = 10.0;
  r->dtc[ORA_T_SMB] = 10.0; r->dtc[ORA_T_BMB] = 10.0; r->dtc[ORA_T_ELRA] = 100.0; r->dtc[ORA_T_OUTPUT] = 5000.0;
  r->t1[ORA_T_THERMO] = start_time + r->dtc[ORA_T_THERMO];
  r->do_[ORA_T_THERMO] = 0;
  r->dt = 0.0; r->dt_prev = 1000.0;
  r->H0 = 5000.0; r->R0 = 300000.0; r->lambda = 5.0;
}

int ora_run_model(const ora_mesh *m, ora_ice *ice, const ora_config *c, ora_region *r, double t_end, long max_steps)
{
  long steps = 0;
  /* the thermodynamics timer runs on C%dt_thermo (src/UFEMISM_main_model.f90:369,797) */
  if (c->dt_thermo > 0.0 && r->n_steps == 0 && r->dtc[ORA_T_THERMO] != c->dt_thermo) {
    r->dtc[ORA_T_THERMO] = c->dt_thermo;
    r->t1[ORA_T_THERMO] = r->t0[ORA_T_THERMO] + c->dt_thermo;
  }
  while (r->time < t_end && (max_steps <= 0 || steps < max_steps)) {
    /* run_ELRA_model (src/bedrock_ELRA_module.f90:22-66): benchmark -> t0_ELRA = time; realistic -> only when the deformation rate is due */
    if (c->benchmark != ORA_BM_NONE || r->do_[ORA_T_ELRA]) r->t0[ORA_T_ELRA] = r->time;
    ora_calculate_ice_thickness_change(m, ice, c, r->dt);
    ora_update_general_ice_model_data(m, ice, c, r->time);
    if (r->do_[ORA_T_SIA]) { ora_solve_SIA(m, ice, c); r->t0[ORA_T_SIA] = r->time; r->n_sia++; }
    if (r->do_[ORA_T_SSA]) {
      ora_ssa_stats st;
      int rc = ora_solve_SSA(m, ice, c, &st);
      if (rc < 0) return rc;
      r->t0[ORA_T_SSA] = r->time; r->n_ssa++; r->n_sor_total += st.n_inner_total; r->n_outer_total += st.n_outer;
    }
    if (r->do_[ORA_T_CLIMATE]) r->t0[ORA_T_CLIMATE] = r->time;
    if (r->do_[ORA_T_SMB]) { ora_run_SMB_benchmark(m, ice, c, r->time, r->H0, r->R0, r->lambda); r->t0[ORA_T_SMB] = r->time; }
    if (r->do_[ORA_T_BMB]) r->t0[ORA_T_BMB] = r->time;
    if (r->do_[ORA_T_THERMO]) {
      /* update_ice_temperature (src/thermodynamics_module.f90:23-202): the EISMINT experiments and realistic runs; with
       * c->thermo == 0 only its solve_SIA_3D U/V half runs (what the critical time step reads) */
      if ((c->benchmark >= ORA_BM_EISMINT_1 && c->benchmark <= ORA_BM_EISMINT_6) || c->benchmark == ORA_BM_NONE) {
        if (c->thermo) { int nu; int rc = ora_update_ice_temperature(m, ice, c, &nu); if (rc < 0) return rc - 10; }
        else ora_solve_SIA_3D_UV(m, ice, c);
      }
      r->t0[ORA_T_THERMO] = r->time;
    }
    if (r->do_[ORA_T_OUTPUT]) r->t0[ORA_T_OUTPUT] = r->time;
    /* determine_timesteps_and_actions */
    double d3[3];
    ora_determine_timesteps(m, ice, c, d3);
    double dt_D_2D_min = d3[0], dt_V_2D_SSA_min = d3[1], dt_V_3D_SIA_min = d3[2];
    r->dt_crit_last[0] = d3[0]; r->dt_crit_last[1] = d3[1]; r->dt_crit_last[2] = d3[2];
    r->dt = fmin(fmin(fmin(dt_D_2D_min, dt_V_2D_SSA_min), dt_V_3D_SIA_min), c->dt_max);
    if (fabs(1.0 - r->dt / r->dt_prev) > 0.1) r->dt_prev = r->dt;
    r->dtc[ORA_T_SIA] = fmin(c->dt_max, fmin(dt_D_2D_min, dt_V_3D_SIA_min));
    r->dtc[ORA_T_SSA] = fmin(c->dt_max, dt_V_2D_SSA_min);
    double t_next_action = 0.0;
    for (int k = 0; k < ORA_NT; k++) { r->t1[k] = r->t0[k] + r->dtc[k]; if (k == 0 || r->t1[k] < t_next_action) t_next_action = r->t1[k]; }
    r->dt = t_next_action - r->time;
    for (int k = 0; k < ORA_NT; k++) r->do_[k] = (t_next_action == r->t1[k]);
    if (t_next_action >= t_end) {
      r->dt = t_end - r->time;
      r->do_[ORA_T_SIA] = r->do_[ORA_T_SSA] = r->do_[ORA_T_THERMO] = r->do_[ORA_T_CLIMATE] = r->do_[ORA_T_SMB] = r->do_[ORA_T_BMB] = 1;
    }
    r->time = r->time + r->dt;
    steps++; r->n_steps++;
  }
  return 0;
}
