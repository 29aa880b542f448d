/*
 * ufm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the UFEMISM v1.1.1 ice-dynamics hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY PIN -- what it is and what it is not.  The reference is Fortran 90 + MPI + NetCDF; this image has no Fortran compiler,
 * MPI or NetCDF, and the reference ships no tests, fixtures or golden vectors: there is no
 * reference BUILD to compare with and, by the letter of the task statement, the oracle is "parity unpinned".  It is pinned as
 * far as this environment allows:
 *  (a) to the REFERENCE'S OWN SOURCE TEXT: oracle/f90py.py translates 78 routines of /root/reference/src (the SOR sweep, the whole
 *      solve_SSA with the grounding-line flux, masks, gradients, SIA, thickness update, critical time steps, neighbour functions,
 *      Ac / AaAc mesh construction, five-colouring, Voronoi areas, connection widths, update_ice_temperature with its DGTSV going to
 *      a real LAPACK) statement by statement into Python and runs
 *      them on the golden mesh; the oracle and the mesh substrate agree with them BIT FOR BIT (tests/test_reference_source.py), and
 *      their outputs are committed as golden vectors (tests/golden/reference_source_600.npz) so that the pin also holds on machines
 *      without the reference, incl. the B200 box, where the CUDA path is compared with them.  This check found -- and this file
 *      now has fixed -- one discrepancy: gfortran's NORM2 is libgfortran's scaled sum of squares, not hypot();
 *  (b) by following the Fortran statement by statement (file:line cited at every function);
 *  (c) by model runs against the two analytic solutions hard-coded in the reference (Halfar, Bueler:
 *      src/reference_fields_module.f90:707-795, src/SMB_module.f90:240-283) and against two closed-form steady states of the heat
 *      equation (pure conduction; the Robin profile the reference codes in replace_Ti_with_robin_solution);
 *  (d) for the one third-party routine on the path, LAPACK DGTSV, by a bit-for-bit comparison with scipy's bundled LAPACK
 *      (tests/test_oracle.py).
 * Not covered by (a): what a translation cannot show -- code generation choices of gfortran itself (assumed: -O3 without -ffast-math, no FMA contraction, libm pow / tan / exp).
 *
 * Layout = the reference's: column-major, 1-based indices stored in the integer arrays, padded
 * ELL rows (width nC_mem for connectivity, nC_mem+1 for neighbour functions).
 * Arithmetic = gfortran -O3 without -ffast-math on x86-64: no reassociation, no FMA,
 * SUM() left to right, x**2 -> x*x, x**real -> libm pow.
 *
 * The field lists below are X-macros so that oracle/oracle.py can build matching ctypes
 * structures by parsing this header.  X(type, name, rows, cols): rows/cols in
 * {NV, NAC, NAA (=nV+nAc), NZ, NCM (=nC_mem), NCM1 (=nC_mem+1), N1, N2, N4, N5}.
 */
#ifndef UFM_ORACLE_H
#define UFM_ORACLE_H

#define ORA_MESH_FIELDS(X) \
  X(double, V, NV, N2) X(double, A, NV, N1) X(int, nC, NV, N1) X(int, C, NV, NCM) X(double, Cw, NV, NCM) \
  X(int, edge_index, NV, N1) X(double, Nx, NV, NCM1) X(double, Ny, NV, NCM1) \
  X(int, Aci, NAC, N4) X(int, iAci, NV, NCM) X(int, edge_index_Ac, NAC, N1) \
  X(double, Nx_Ac, NAC, N4) X(double, Ny_Ac, NAC, N4) X(double, No_Ac, NAC, N4) X(double, Np_Ac, NAC, N1) \
  X(int, nCAaAc, NAA, N1) X(int, CAaAc, NAA, NCM) \
  X(double, Nx_AaAc, NAA, NCM1) X(double, Ny_AaAc, NAA, NCM1) X(double, Nxx_AaAc, NAA, NCM1) \
  X(double, Nxy_AaAc, NAA, NCM1) X(double, Nyy_AaAc, NAA, NCM1) \
  X(int, colour_vi, NAA, N5) X(int, colour_nV, N5, N1) \
  /* read only by thermodynamics (upwind derivative, src/mesh_derivatives_module.f90:435-483) */ \
  X(double, R, NV, N1) X(int, Tri, NTRI, N3) X(int, niTri, NV, N1) X(int, iTri, NV, NCM) X(double, NxTri, NTRI, N3) X(double, NyTri, NTRI, N3)

#define ORA_ICE_FIELDS(X) \
  /* Aa, src/data_types_module.f90:15-214 */ \
  X(double, Hi, NV, N1) X(double, Hb, NV, N1) X(double, Hs, NV, N1) X(double, SL, NV, N1) \
  X(double, dHb_dt, NV, N1) X(double, dHi_dt, NV, N1) X(double, dHs_dt, NV, N1) X(double, Hi_prev, NV, N1) \
  X(double, dHi_dx, NV, N1) X(double, dHi_dy, NV, N1) X(double, dHs_dx, NV, N1) X(double, dHs_dy, NV, N1) \
  X(double, dHs_dx_shelf, NV, N1) X(double, dHs_dy_shelf, NV, N1) X(double, A_flow_mean, NV, N1) \
  X(double, U_SIA, NV, N1) X(double, V_SIA, NV, N1) X(double, D_SIA, NV, N1) X(double, U_SSA, NV, N1) X(double, V_SSA, NV, N1) \
  X(double, SMB_year, NV, N1) X(double, BMB, NV, N1) X(int, mask_noice, NV, N1) \
  X(int, mask_land, NV, N1) X(int, mask_ocean, NV, N1) X(int, mask_lake, NV, N1) X(int, mask_ice, NV, N1) \
  X(int, mask_sheet, NV, N1) X(int, mask_shelf, NV, N1) X(int, mask_coast, NV, N1) X(int, mask_margin, NV, N1) \
  X(int, mask_gl, NV, N1) X(int, mask_cf, NV, N1) X(int, mask, NV, N1) \
  X(double, Ti, NV, NZ) X(double, A_flow, NV, NZ) X(double, U_3D, NV, NZ) X(double, V_3D, NV, NZ) X(double, dVi_in, NV, NCM) \
  /* thermodynamics, src/data_types_module.f90:160-183 */ \
  X(double, W_3D, NV, NZ) X(double, Ti_pmp, NV, NZ) X(double, Cpi, NV, NZ) X(double, Ki, NV, NZ) \
  X(double, dzeta_dt, NV, NZ) X(double, dzeta_dx, NV, NZ) X(double, dzeta_dy, NV, NZ) X(double, dzeta_dz, NV, N1) \
  X(double, frictional_heating, NV, N1) X(double, GHF, NV, N1) X(double, T2m, NV, N12) X(double, Ti_new, NV, NZ) \
  /* Ac */ \
  X(double, Hi_Ac, NAC, N1) X(double, Hb_Ac, NAC, N1) X(double, Hs_Ac, NAC, N1) X(double, SL_Ac, NAC, N1) \
  X(double, dHi_dx_Ac, NAC, N1) X(double, dHi_dy_Ac, NAC, N1) X(double, dHi_dp_Ac, NAC, N1) X(double, dHi_do_Ac, NAC, N1) \
  X(double, dHb_dx_Ac, NAC, N1) X(double, dHb_dy_Ac, NAC, N1) X(double, dHb_dp_Ac, NAC, N1) X(double, dHb_do_Ac, NAC, N1) \
  X(double, dHs_dx_Ac, NAC, N1) X(double, dHs_dy_Ac, NAC, N1) X(double, dHs_dp_Ac, NAC, N1) X(double, dHs_do_Ac, NAC, N1) \
  X(double, dSL_dx_Ac, NAC, N1) X(double, dSL_dy_Ac, NAC, N1) X(double, dSL_dp_Ac, NAC, N1) X(double, dSL_do_Ac, NAC, N1) \
  X(double, dHs_dx_shelf_Ac, NAC, N1) X(double, dHs_dy_shelf_Ac, NAC, N1) X(double, A_flow_mean_Ac, NAC, N1) \
  X(double, Ux_SIA_Ac, NAC, N1) X(double, Uy_SIA_Ac, NAC, N1) X(double, Up_SIA_Ac, NAC, N1) X(double, Uo_SIA_Ac, NAC, N1) \
  X(double, D_SIA_Ac, NAC, N1) \
  X(double, Ux_SSA_Ac, NAC, N1) X(double, Uy_SSA_Ac, NAC, N1) X(double, Up_SSA_Ac, NAC, N1) X(double, Uo_SSA_Ac, NAC, N1) \
  X(double, Qabs_GL_Ac, NAC, N1) X(double, Qp_GL_Ac, NAC, N1) \
  X(int, mask_land_Ac, NAC, N1) X(int, mask_ocean_Ac, NAC, N1) X(int, mask_lake_Ac, NAC, N1) X(int, mask_ice_Ac, NAC, N1) \
  X(int, mask_sheet_Ac, NAC, N1) X(int, mask_shelf_Ac, NAC, N1) X(int, mask_coast_Ac, NAC, N1) X(int, mask_margin_Ac, NAC, N1) \
  X(int, mask_gl_Ac, NAC, N1) X(int, mask_cf_Ac, NAC, N1) X(int, mask_Ac, NAC, N1) \
  X(double, Ti_Ac, NAC, NZ) X(double, A_flow_Ac, NAC, NZ) X(double, D_SIA_3D_Ac, NAC, NZ) \
  /* AaAc, allocation list src/ice_dynamics_module.f90:1151-1176 */ \
  X(double, Hi_AaAc, NAA, N1) X(double, Hb_AaAc, NAA, N1) X(double, SL_AaAc, NAA, N1) \
  X(double, dHs_dx_shelf_AaAc, NAA, N1) X(double, dHs_dy_shelf_AaAc, NAA, N1) X(double, A_flow_mean_AaAc, NAA, N1) \
  X(double, U_SSA_AaAc, NAA, N1) X(double, V_SSA_AaAc, NAA, N1) X(double, N_AaAc, NAA, N1) X(double, N_AaAc_prev, NAA, N1) \
  X(double, eta_AaAc, NAA, N1) X(double, dU_SSA_dx_AaAc, NAA, N1) X(double, dU_SSA_dy_AaAc, NAA, N1) \
  X(double, dV_SSA_dx_AaAc, NAA, N1) X(double, dV_SSA_dy_AaAc, NAA, N1) X(double, S_AaAc, NAA, N1) \
  X(double, tau_c_AaAc, NAA, N1) X(double, phi_fric_AaAc, NAA, N1) X(double, RHSx_AaAc, NAA, N1) X(double, RHSy_AaAc, NAA, N1) \
  X(double, eu_i_AaAc, NAA, N1) X(double, ev_i_AaAc, NAA, N1) X(double, LHSx_AaAc, NAA, N1) X(double, LHSy_AaAc, NAA, N1) \
  X(double, resU_AaAc, NAA, N1) X(double, resV_AaAc, NAA, N1)

typedef struct {
  int nV, nAc, nVAaAc, nC_mem, nTri;
#define X(t, n, r, c) t *n;
  ORA_MESH_FIELDS(X)
#undef X
} ora_mesh;

typedef struct {
#define X(t, n, r, c) t *n;
  ORA_ICE_FIELDS(X)
#undef X
} ora_ice;

/* choice_benchmark_experiment; 0 = do_benchmark_experiment .FALSE. */
enum { ORA_BM_NONE = 0, ORA_BM_EISMINT_1 = 1, ORA_BM_EISMINT_2, ORA_BM_EISMINT_3, ORA_BM_EISMINT_4, ORA_BM_EISMINT_5,
       ORA_BM_EISMINT_6, ORA_BM_HALFAR = 7, ORA_BM_BUELER = 8, ORA_BM_MISMIP_MOD = 9, ORA_BM_MESH_GENERATION_TEST = 10,
       ORA_BM_SSA_ICESTREAM = 11 };

/* the C%... scalars the hot path reads; defaults: src/configuration_module.f90:37,78,124-126,169-184 */
typedef struct {
  int nZ;
  double zeta[32];
  double m_enh_sia, m_enh_ssa;
  int use_analytical_GL_flux;
  double SSA_RN_tol;
  int SSA_max_outer_loops;
  double SSA_max_residual_UV, SSA_SOR_omega;
  int SSA_max_inner_loops;
  double dt_max;
  int benchmark;
  int nthreads; /* how many MPI ranks the run is split into (OpenMP threads here) */
  double dt_thermo; /* C%dt_thermo, src/configuration_module.f90:38 */
  int thermo;       /* ora_run_model: 1 = the whole update_ice_temperature on the thermodynamics timer; 0 = its U_3D / V_3D half only */
} ora_config;

typedef struct {
  int n_outer, n_inner_total, n_inner_last, did_reset, rc; /* rc: 0 ok, 1 SOR hit max_inner_loops (WARNING), -1 unstable twice */
  double last_max_residual, last_RN;
} ora_ssa_stats;

void ora_config_defaults(ora_config *c);
void ora_partition_list(int ntot, int i, int n, int *i1, int *i2);

void ora_calculate_ice_thickness_change(const ora_mesh *m, ora_ice *ice, const ora_config *c, double dt);
void ora_update_general_ice_model_data(const ora_mesh *m, ora_ice *ice, const ora_config *c, double time);
void ora_solve_SIA(const ora_mesh *m, ora_ice *ice, const ora_config *c);
void ora_solve_SIA_3D_UV(const ora_mesh *m, ora_ice *ice, const ora_config *c);
/* thermodynamics: the whole solve_SIA_3D (U, V and W), and update_ice_temperature.
 * ora_update_ice_temperature returns 0, -1 (more than 1 % of the vertices unstable: the reference STOPs), -2 (DGTSV info /= 0:
 * STOP) or -3 (no upwind triangle found: the reference prints an ERROR and then indexes NxTri(0,:)); *n_unstable = number of
 * columns replaced by the Robin solution. */
void ora_solve_SIA_3D(const ora_mesh *m, ora_ice *ice, const ora_config *c);
int  ora_update_ice_temperature(const ora_mesh *m, ora_ice *ice, const ora_config *c, int *n_unstable);
void ora_replace_Ti_with_robin_solution(const ora_mesh *m, ora_ice *ice, const ora_config *c, int vi);
/* LAPACK DGTSV (netlib reference algorithm, NRHS = 1), called by tridiagonal_solve (src/thermodynamics_module.f90:313-358);
 * dl, d, du, b are overwritten as LAPACK does.  Pinned against scipy's bundled LAPACK in tests/test_oracle.py. */
void ora_dgtsv(int n, double *dl, double *d, double *du, double *b, int *info);
int  ora_solve_SSA(const ora_mesh *m, ora_ice *ice, const ora_config *c, ora_ssa_stats *st);
void ora_determine_timesteps(const ora_mesh *m, const ora_ice *ice, const ora_config *c, double out3[3]);

/* pieces of solve_SSA, exposed for kernel-level parity tests */
void ora_basal_yield_stress(const ora_mesh *m, ora_ice *ice, const ora_config *c);
void ora_calculate_GL_flux(const ora_mesh *m, ora_ice *ice, const ora_config *c);
void ora_SSA_gather_AaAc(const ora_mesh *m, ora_ice *ice, const ora_config *c);
void ora_SSA_effective_viscosity(const ora_mesh *m, ora_ice *ice, const ora_config *c);
void ora_SSA_sliding_term(const ora_mesh *m, ora_ice *ice, const ora_config *c);
/* max_inner_override > 0 replaces C%SSA_max_inner_loops; force_iters != 0 disables the stop tests */
int  ora_solve_SSA_linearised(const ora_mesh *m, ora_ice *ice, const ora_config *c, int max_inner_override, int force_iters,
                              int *n_inner, double *max_residual, int *did_reset);
void ora_apply_Neumann_boundary_AaAc(const ora_mesh *m, const ora_config *c, double *d_AaAc);
void ora_get_mesh_derivatives(const ora_mesh *m, const ora_config *c, const double *d, double *ddx, double *ddy);
void ora_remap_cons_2D(int order, int nV_dst, const int *vli1, const int *vli2, const int *vi, const double *w0, const double *w1x, const double *w1y,
                       const double *d_src, const double *ddx_src, const double *ddy_src, double *d_dst);
void ora_map_Ac_to_Aa(const ora_mesh *m, const ora_config *c, const double *d_Ac, double *d_Aa);

/* benchmark forcing + analytic solutions (host side in the reference too) */
double ora_Halfar_solution(double H0, double R0, double x, double y, double t);
double ora_Bueler_solution(double H0, double R0, double lambda, double x, double y, double t);
double ora_Bueler_solution_MB(double H0, double R0, double lambda, double x, double y, double t);
void ora_run_SMB_benchmark(const ora_mesh *m, ora_ice *ice, const ora_config *c, double time, double H0, double R0, double lambda);

/* region time loop for benchmark physics: run_model + determine_timesteps_and_actions.
 * Eight action timers as in type_model_region (src/data_types_module.f90:927-960). */
enum { ORA_T_SIA = 0, ORA_T_SSA, ORA_T_THERMO, ORA_T_CLIMATE, ORA_T_SMB, ORA_T_BMB, ORA_T_ELRA, ORA_T_OUTPUT, ORA_NT };
typedef struct {
  double time, dt, dt_prev;
  double t0[ORA_NT], t1[ORA_NT], dtc[ORA_NT]; /* dtc[SIA], dtc[SSA] = region%dt_SIA, dt_SSA; others C%dt_* */
  int do_[ORA_NT];
  double H0, R0, lambda;                 /* Halfar / Bueler parameters */
  long n_steps, n_sia, n_ssa, n_sor_total, n_outer_total;
  double dt_crit_last[3];
} ora_region;
void ora_region_init(ora_region *r, double start_time);
int  ora_run_model(const ora_mesh *m, ora_ice *ice, const ora_config *c, ora_region *r, double t_end, long max_steps);

#endif
