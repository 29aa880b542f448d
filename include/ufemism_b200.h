/*
 * ufemism_b200.h -- C ABI of libufemism_b200.so: the B200-native ice-dynamics hot path of UFEMISM.
 *
 * The reference (IMAU-paleo/UFEMISM v1.1.1, Fortran 90 + MPI shared memory) has no plugin or FFI
 * interface for this path; it sits behind four module procedures called from run_model
 * (src/UFEMISM_main_model.f90:90,115,124,132).  This header IS the drop-in boundary: every entry
 * point names the reference routine whose body it replaces.  The ISO_C_BINDING interfaces that bind
 * it are in ufemism_b200/fortran/ufemism_b200_shim.f90; see INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; all reals are IEEE fp64 (dp = KIND(1.0D0), src/configuration_module.f90:24),
 *    all integers 32-bit, LOGICAL passed as int.
 *  - host arrays are in the reference's own layout: column-major, 1-based indices stored in the
 *    arrays, leading dimension = number of rows given in the descriptor.  The library copies; the
 *    host keeps ownership of every pointer it passes.
 *  - return code 0 = ok; > 0 = warning the reference only prints (WRITE(0,*)) and carries on;
 *    < 0 = fatal (the reference calls MPI_ABORT); <= -100 = CUDA error (-100 - cudaError_t).
 *    ufm_last_error() returns the message.  No exceptions cross the boundary.
 *  - one handle = one model region on one GPU, one CUDA stream, single host thread, non-reentrant.
 *    In the SPMD host only par%master (or one rank per GPU) enters the library, followed by CALL sync.
 *  - there is no CPU fallback: without a usable CUDA device ufm_create fails.
 */
#ifndef UFEMISM_B200_H
#define UFEMISM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define UFM_ABI_VERSION 2
#define UFM_MAX_NZ 32

typedef struct ufm_handle ufm_handle;

/* choice_benchmark_experiment (src/configuration_module.f90:69-70); 0 = do_benchmark_experiment .FALSE. */
enum ufm_benchmark {
  UFM_BM_NONE = 0, UFM_BM_EISMINT_1 = 1, UFM_BM_EISMINT_2, UFM_BM_EISMINT_3, UFM_BM_EISMINT_4, UFM_BM_EISMINT_5,
  UFM_BM_EISMINT_6, UFM_BM_HALFAR = 7, UFM_BM_BUELER = 8, UFM_BM_MISMIP_MOD = 9, UFM_BM_MESH_GENERATION_TEST = 10,
  UFM_BM_SSA_ICESTREAM = 11
};

/* The C%... scalars this path reads (src/configuration_module.f90:37,124-126,169-184).  Physical
 * constants of src/parameters_module.f90:9-21 are compile-time constants in the kernels, as they
 * are PARAMETERs in the reference. */
typedef struct ufm_params {
  int    nZ;                      /* C%nZ, <= UFM_MAX_NZ */
  double zeta[UFM_MAX_NZ];        /* C%zeta(1:nZ) */
  double m_enh_sia, m_enh_ssa;    /* C%m_enh_sia, C%m_enh_ssa */
  int    use_analytical_GL_flux;  /* C%use_analytical_GL_flux */
  double SSA_RN_tol;              /* 1e-5 */
  int    SSA_max_outer_loops;     /* 50 */
  double SSA_max_residual_UV;     /* 2.5 */
  double SSA_SOR_omega;           /* 1.2 */
  int    SSA_max_inner_loops;     /* 10000 */
  double dt_max;                  /* 10 */
  int    benchmark;               /* enum ufm_benchmark */
  int    exact_xy;                /* 1 (default): evaluate the 3*V_xy / 3*U_xy cross terms of the SOR
                                     sweep coefficient by coefficient exactly as
                                     get_mesh_curvatures_vertex_AaAc does (bit-identical results);
                                     0: use the pre-summed row sum (saves 8(n+1) B per vertex and sweep,
                                     differs from the reference by O(1 ulp) per update). */
  double dt_thermo;               /* C%dt_thermo (10 yr, src/configuration_module.f90:38): time step of the heat equation */
} ufm_params;

/* type_mesh fields used by the path (src/data_types_module.f90:216-345).  ld* = allocated rows. */
typedef struct ufm_mesh_desc {
  int nV, nAc, nC_mem;            /* mesh%nV, mesh%nAc, mesh%nC_mem (= C%nconmax, 16) */
  int ldV, ldAc, ldAaAc;          /* leading dimensions of the (nV,..), (nAc,..), (nV+nAc,..) arrays */
  const double *V;                /* (ldV,2) */
  const double *A;                /* (nV) Voronoi cell areas */
  const int    *nC, *C;           /* (nV), (ldV,nC_mem) */
  const double *Cw;               /* (ldV,nC_mem) */
  const int    *edge_index;       /* (nV) 0..8 */
  const double *Nx, *Ny;          /* (ldV,nC_mem+1) */
  const int    *Aci;              /* (ldAc,4) [vi,vj,vl,vr] */
  const int    *iAci;             /* (ldV,nC_mem) */
  const int    *edge_index_Ac;    /* (nAc) */
  const double *Nx_Ac, *Ny_Ac, *No_Ac; /* (ldAc,4) */
  const double *Np_Ac;            /* (nAc) */
  const int    *nCAaAc, *CAaAc;   /* (nV+nAc), (ldAaAc,nC_mem) */
  const double *Nx_AaAc, *Ny_AaAc, *Nxx_AaAc, *Nxy_AaAc, *Nyy_AaAc; /* (ldAaAc,nC_mem+1) */
  const int    *colour_vi;        /* (ldAaAc,5) */
  const int    *colour_nV;        /* (5) */
  /* optional, read only by ufm_update_ice_temperature (upwind derivative of the temperature,
   * src/mesh_derivatives_module.f90:435-483).  Tri == NULL: thermodynamics is not available on this mesh. */
  int nTri, ldTri;                /* mesh%nTri and the leading dimension of the (nTri,3) arrays */
  const int    *Tri;              /* (ldTri,3) */
  const int    *niTri, *iTri;     /* (nV), (ldV,nC_mem) */
  const double *R;                /* (nV) resolution */
  const double *NxTri, *NyTri;    /* (ldTri,3) */
} ufm_mesh_desc;

/* Fields of type_ice_model (src/data_types_module.f90:15-214) that can cross the boundary.
 * AA = on vertices (nV), AC = staggered (nAc), AAAC = combined (nV+nAc); I = INTEGER array. */
enum ufm_field {
  /* inputs written by CPU components (ELRA, SMB, BMB, remapping) */
  UFM_F_HI = 0, UFM_F_HB, UFM_F_SL, UFM_F_DHB_DT, UFM_F_SMB_YEAR, UFM_F_BMB, UFM_F_MASK_NOICE /*I*/,
  /* Aa outputs */
  UFM_F_HS, UFM_F_DHI_DT, UFM_F_DHS_DT, UFM_F_HI_PREV, UFM_F_DHI_DX, UFM_F_DHI_DY, UFM_F_DHS_DX, UFM_F_DHS_DY,
  UFM_F_DHS_DX_SHELF, UFM_F_DHS_DY_SHELF, UFM_F_A_FLOW_MEAN, UFM_F_U_SIA, UFM_F_V_SIA, UFM_F_D_SIA, UFM_F_U_SSA, UFM_F_V_SSA,
  UFM_F_MASK_LAND /*I*/, UFM_F_MASK_OCEAN, UFM_F_MASK_LAKE, UFM_F_MASK_ICE, UFM_F_MASK_SHEET, UFM_F_MASK_SHELF,
  UFM_F_MASK_COAST, UFM_F_MASK_MARGIN, UFM_F_MASK_GL, UFM_F_MASK_CF, UFM_F_MASK,
  /* Ac outputs */
  UFM_F_HI_AC, UFM_F_HB_AC, UFM_F_HS_AC, UFM_F_SL_AC,
  UFM_F_DHI_DX_AC, UFM_F_DHI_DY_AC, UFM_F_DHI_DP_AC, UFM_F_DHI_DO_AC,
  UFM_F_DHB_DX_AC, UFM_F_DHB_DY_AC, UFM_F_DHB_DP_AC, UFM_F_DHB_DO_AC,
  UFM_F_DHS_DX_AC, UFM_F_DHS_DY_AC, UFM_F_DHS_DP_AC, UFM_F_DHS_DO_AC,
  UFM_F_DSL_DX_AC, UFM_F_DSL_DY_AC, UFM_F_DSL_DP_AC, UFM_F_DSL_DO_AC,
  UFM_F_DHS_DX_SHELF_AC, UFM_F_DHS_DY_SHELF_AC, UFM_F_A_FLOW_MEAN_AC,
  UFM_F_UX_SIA_AC, UFM_F_UY_SIA_AC, UFM_F_UP_SIA_AC, UFM_F_UO_SIA_AC, UFM_F_D_SIA_AC,
  UFM_F_UX_SSA_AC, UFM_F_UY_SSA_AC, UFM_F_UP_SSA_AC, UFM_F_UO_SSA_AC, UFM_F_QABS_GL_AC, UFM_F_QP_GL_AC,
  UFM_F_MASK_LAND_AC /*I*/, UFM_F_MASK_OCEAN_AC, UFM_F_MASK_LAKE_AC, UFM_F_MASK_ICE_AC, UFM_F_MASK_SHEET_AC, UFM_F_MASK_SHELF_AC,
  UFM_F_MASK_COAST_AC, UFM_F_MASK_MARGIN_AC, UFM_F_MASK_GL_AC, UFM_F_MASK_CF_AC, UFM_F_MASK_AC,
  /* AaAc (SSA work arrays, src/ice_dynamics_module.f90:1151-1176) */
  UFM_F_U_SSA_AAAC, UFM_F_V_SSA_AAAC, UFM_F_ETA_AAAC, UFM_F_N_AAAC, UFM_F_S_AAAC, UFM_F_TAU_C_AAAC, UFM_F_PHI_FRIC_AAAC,
  UFM_F_RHSX_AAAC, UFM_F_RHSY_AAAC, UFM_F_EU_I_AAAC, UFM_F_EV_I_AAAC,
  UFM_F_DU_DX_AAAC, UFM_F_DU_DY_AAAC, UFM_F_DV_DX_AAAC, UFM_F_DV_DY_AAAC,
  /* (nV,nZ) */
  UFM_F_U_3D, UFM_F_V_3D,
  UFM_F_TI, /* englacial temperature, input of the Arrhenius flow factor when do_benchmark_experiment is .FALSE. */
  /* thermodynamics (meshes uploaded with Tri): W_3D (nV,nZ) out; GHF (nV) in; T2m (nV,12) in = climate%applied%T2m;
   * frictional_heating (nV) out */
  UFM_F_W_3D, UFM_F_GHF, UFM_F_T2M, UFM_F_FRICTIONAL_HEATING,
  UFM_F_COUNT
};

/* what solve_SSA would have printed had its WRITE statements not been commented out
 * (src/ice_dynamics_module.f90:519,666,675); parity is defined "after the same SOR iteration count". */
typedef struct ufm_ssa_stats {
  int    n_outer;            /* viscosity iterations started */
  int    n_inner_total;      /* SOR iterations summed over all linear solves */
  int    n_inner_last;       /* SOR iterations of the last linear solve */
  int    did_reset;          /* velocities were reset to zero once (:679-684) */
  int    rc;                 /* 0 ok, 1 = "WARNING - SSA SOR solver doesnt converge!" (:686), -1 = unstable twice (:537) */
  double last_max_residual;  /* max(|resU|,|resV|) of the last SOR iteration */
  double last_RN;            /* sqrt(sum (N-Nprev)^2 / sum N^2) of the last viscosity iteration */
} ufm_ssa_stats;

/* work counters / timers since ufm_create or the last ufm_counters_reset */
typedef struct ufm_counters {
  long long kernel_launches;       /* kernels launched by this library */
  long long sor_iterations;        /* SOR iterations executed */
  double    sor_ms;                /* device time inside the SOR kernel (CUDA events on the handle's stream) */
  long long sor_launches;
  double    sor_bytes_per_iteration; /* algorithmic bytes of one full 5-colour iteration: sum_i (80 + 20 n_i) */
  double    h2d_bytes, d2h_bytes;  /* bytes moved by ufm_state_upload / ufm_state_download */
} ufm_counters;

/* ---- life cycle ---------------------------------------------------------------------------- */
/* after initialize_main_constants (src/UFEMISM_program.f90:103): copies the scalars, selects the device */
int ufm_create(int device, const ufm_params *params, ufm_handle **out);
int ufm_destroy(ufm_handle *h);
int ufm_set_params(ufm_handle *h, const ufm_params *params);
/* use an existing CUDA stream (cudaStream_t / CUstream) instead of the handle's own; NULL restores it */
int ufm_set_stream(ufm_handle *h, void *cuda_stream);
int ufm_synchronize(ufm_handle *h);
const char *ufm_last_error(void);
int ufm_abi_version(void);

/* ---- mesh: end of create_final_mesh_from_merged_submesh (src/mesh_creation_module.f90:1737),
 *      read_mesh_from_restart_file (src/restart_module.f90:102), mesh swap (src/UFEMISM_main_model.f90:294).
 *      Copies, renumbers (colour-major, degree-sliced, space-filling order) and compacts the mesh;
 *      allocates and zero-fills all state (U_SSA is not remapped by the reference: :1212-1213).
 *      A second call replaces the mesh (device re-upload after a CPU mesh update). ---- */
int ufm_mesh_upload(ufm_handle *h, const ufm_mesh_desc *mesh);
int ufm_mesh_free(ufm_handle *h);

/* ---- device re-upload from PRIMARY mesh data: what a mesh update (src/mesh_update_module.f90) or
 *      read_mesh_from_restart_file (src/restart_module.f90:31-116) leaves before the reference rebuilds the secondary data on the CPU
 *      (:88-103).  The library derives Voronoi areas, connection widths, the Ac and AaAc meshes and the five-colouring on the host
 *      (same results as find_Voronoi_cell_areas / find_connection_widths / make_Ac_mesh / make_combined_AaAc_mesh /
 *      calculate_five_colouring_AaAc) and all neighbour functions on the device, then proceeds as ufm_mesh_upload.
 *      ufm_mesh_secondary_get lends the derived host arrays (valid until the next upload / ufm_mesh_free / ufm_destroy; the
 *      neighbour-function members of *out are NULL, ldAc > nAc) so that the host need not recompute the ones it reads. ---- */
typedef struct ufm_mesh_primary {
  int nV, nTri, nC_mem;
  int ldV, ldTri;                   /* leading dimensions of the (nV,..) / (nTri,..) arrays; 0 = nV / nTri */
  double xmin, xmax, ymin, ymax;    /* mesh%xmin .. mesh%ymax: the model domain */
  const double *V;                  /* (ldV,2) */
  const int    *nC, *C;             /* (nV), (ldV,nC_mem) */
  const int    *niTri, *iTri;       /* (nV), (ldV,nC_mem) */
  const int    *edge_index;         /* (nV) */
  const int    *Tri;                /* (ldTri,3), counter-clockwise */
  int thermo;                       /* != 0: also derive R, NxTri, NyTri so that ufm_update_ice_temperature is available */
} ufm_mesh_primary;
int ufm_mesh_upload_primary(ufm_handle *h, const ufm_mesh_primary *mesh);
/* the host-only half on its own (no device needed): derive, look, free; *derived is an opaque object */
int ufm_mesh_derive_secondary(const ufm_mesh_primary *mesh, void **derived);
/* the same with *derived_inout holding the object of an earlier call (or NULL): its buffers are reused and its contents replaced; on an
 * error the old object is gone and *derived_inout is NULL */
int ufm_mesh_derive_secondary_reuse(const ufm_mesh_primary *mesh, void **derived_inout);
int ufm_mesh_derived_get(const void *derived, ufm_mesh_desc *out, const double **Tricc, const int **Tri_edge_index, const double **VAc,
                         const double **VAaAc, const int **colour);
void ufm_mesh_derived_free(void *derived);
int ufm_mesh_secondary_get(ufm_handle *h, ufm_mesh_desc *out, const double **Tricc, const int **Tri_edge_index, const double **VAc,
                           const double **VAaAc, const int **colour);

/* ---- vertex-partitioned runs over the GPUs of one NVSwitch domain (replaces the index-range split of
 *      partition_list, src/mesh_help_functions_module.f90:1475-1496, and the MPI_BARRIER / MPI_ALLREDUCE of the SOR loop,
 *      src/ice_dynamics_module.f90:662,673).  One process per GPU.  Every rank uploads the same mesh and holds the whole
 *      state; the SSA solve (viscosity, linear-system setup, SOR sweeps) is partitioned into x-strips balanced by row
 *      count.  After each colour sweep the rows a neighbour strip reads are pushed straight into that GPU's (U,V) array
 *      by the sweep kernel itself (NVLink peer stores through CUDA-IPC mappings) and epochs are exchanged through
 *      per-GPU mailboxes; results are bit-identical to the single-GPU run.
 *      Protocol:  ufm_partition_set -> ufm_mesh_upload -> ufm_comm_export -> (all-gather the blobs, e.g. with
 *      torch.distributed / MPI_ALLGATHER) -> ufm_comm_connect -> (host barrier) -> compute calls, collectively. ---- */
#define UFM_MAX_RANKS 8
#define UFM_COMM_BLOB_BYTES 256
int ufm_partition_set(ufm_handle *h, int rank, int nranks);
/* host-only: owner rank (0..nranks-1) of each of the nV+nAc combined-mesh vertices, as ufm_mesh_upload will assign them */
int ufm_partition_owners(const ufm_mesh_desc *mesh, int nranks, unsigned char *owner_out);
/* host-only (planning / tests): the device row order ufm_mesh_upload derives from these per-row keys -- block (1..5 = colour of
 * calculate_five_colouring_AaAc, src/mesh_five_colour_module.f90:18-137; 6 = domain-edge row), owner rank, partition-boundary flag,
 * "late" flag, degree, Morton code, x coordinate.  n_bands = 0: (degree, Morton) inside a group, the default; n_bands > 0: the
 * experimental x-band order selected by the environment variable UFM_ROW_ORDER=bands:<n>[:<window>].  Results of every kernel are
 * independent of this order (rows of one colour are never neighbours). */
int ufm_plan_row_order(int M, const unsigned char *block, const unsigned char *owner, const unsigned char *boundary, const unsigned char *late,
                       const unsigned char *degree, const unsigned *morton, const double *X, int n_bands, int deg_window, int *order_out);
/* Owner rank of every combined-mesh vertex of the resident mesh, reference order: entries 1..nV the Aa vertices, nV+1..nV+nAc the
 * staggered ones (x-strips holding equally many vertices, cf. partition_domain_x_balanced, src/mesh_help_functions_module.f90:1337-1404).
 * Returns 1 when the per-step kernels are partitioned too: thickness update, geometry, SIA, yield stress, scatter and critical time
 * steps then run for the rank's own elements only (halo values exchanged over NVLink), and ufm_state_download is valid for OWNED
 * elements only -- what each MPI rank of the reference writes into the shared window; 0 when they are replicated (single GPU,
 * experiments with column thermodynamics, UFM_PARTITION_STEP=0) and every rank holds every field completely. */
int ufm_partition_owner_of(ufm_handle *h, unsigned char *owner_out /* nV + nAc */);
/* host-only planning: the number of Aa / Ac values rank s sends to rank q in one halo exchange of the partitioned per-step kernels
 * (cnt[s * nranks + q]) -- what replaces the reference's MPI shared-memory window for the strip boundaries */
int ufm_partition_halo_counts(const ufm_mesh_desc *mesh, int nranks, int *cnt_aa /* nranks^2 */, int *cnt_ac /* nranks^2 */);
int ufm_comm_export(ufm_handle *h, void *blob /* UFM_COMM_BLOB_BYTES */);
int ufm_comm_connect(ufm_handle *h, const void *blobs /* nranks * UFM_COMM_BLOB_BYTES, in rank order */);

/* ---- state: explicit, field-granular, reference vertex order ---- */
int ufm_state_upload(ufm_handle *h, int field, const void *host);
int ufm_state_download(ufm_handle *h, int field, void *host);
/* 1 when `field` can be copied with the resident mesh and parameters (Ti needs realistic flow factors or a thermodynamics mesh,
 * W_3D / GHF / T2m a thermodynamics mesh), 0 when not, < 0 on a bad handle / field id */
int ufm_field_resident(ufm_handle *h, int field);
/* sizes of what is resident: dims[0] = 1 when a mesh is resident (else 0 and the rest 0), dims[1] = nV, dims[2] = nAc,
 * dims[3] = nVAaAc = nV + nAc (mesh%nV, mesh%nAc, mesh%nVAaAc: src/data_types_module.f90:216-345), dims[4] = C%nZ */
int ufm_resident_dims(ufm_handle *h, int dims[5]);
/* bit 0 (1): the device evaluates x**y with the bits of this host's libm pow; bit 1 (2): likewise tan on 0.07 <= |x| <= 0.78 (the friction
 * angles of the yield stress).  In detail -- 1: the device evaluates x**y with the bits of this host's libm pow (glibc's algorithm re-stated on the device with the tables of the
 * running libm, validated at ufm_create), so that viscosity, sliding law, SIA diffusivity and grounding-line flux are bit-identical with a
 * CPU run on the same machine; 0: CUDA's pow (within 2 ulp) because the tables were not found / did not validate / UFM_POW_EXACT=0 */
int ufm_pow_mode(ufm_handle *h);
/* host: the same evaluation (libm's pow off its main path); for tests */
double ufm_pow_host(double x, double y);
/* host-only: what ufm_pow_mode would report on this host (the tables are looked for in the running libm.so.6 and validated on first use) */
int ufm_powtab_status(void);
double ufm_tan_host(double x);
/* host twin of the device's x / n for a vertex degree n (k_sia_aa: map_Ac_to_Aa, src/mesh_ArakawaC_module.f90:770-791, divides every term
 * by nC(vi)): one multiply and two fused multiply-adds with 1.0 / n, the bits of the IEEE division; for tests */
double ufm_div_small_host(double x, int n);

/* Page-lock a host array (e.g. one of the Fortran host's MPI shared-memory windows, src/parallel_module.f90:144-160) so that
 * ufm_state_upload / ufm_state_download DMA it directly instead of bouncing through a staging buffer.  Optional. */
int ufm_host_register(ufm_handle *h, void *host, unsigned long long bytes);
int ufm_host_unregister(ufm_handle *h, void *host);

/* ---- mesh update data path (row N3): APPLICATION of the conservative remapping on the device, so that Hi (and the other
 *      fields the reference remaps, src/ice_dynamics_module.f90:1208-1218) need not travel to the host when the CPU
 *      replaces the mesh.  The remapping weights (type_remapping_conservative, src/data_types_module.f90:544-556) are still
 *      BUILT on the CPU (src/mesh_mapping_module.f90:673-3843, out of scope) and handed over as they are.
 *      Protocol: ufm_remap_stash(field) on the old mesh -> ufm_mesh_upload(new mesh) -> ufm_remap_apply(field, map, order).
 *      Replaces remap_cons_1st_order_2D / remap_cons_2nd_order_2D (src/mesh_mapping_module.f90:3964-3983, 4010-4043). ---- */
typedef struct ufm_remap_cons {
  int nV_dst;               /* vertices of the new mesh */
  int n_tot;                /* entries of vi / w0 / w1x / w1y */
  const int *vli1, *vli2;   /* (nV_dst) 1-based inclusive entry range of each destination vertex */
  const int *vi;            /* (n_tot) 1-based source vertex */
  const double *w0, *w1x, *w1y; /* (n_tot); w1x, w1y may be NULL for order 1 */
} ufm_remap_cons;
int ufm_remap_stash(ufm_handle *h, int field);
int ufm_remap_apply(ufm_handle *h, int field, const ufm_remap_cons *map, int order);

/* ---- the four drop-in entry points ---- */
/* body of calculate_ice_thickness_change (src/ice_dynamics_module.f90:31-237) */
int ufm_thickness_update(ufm_handle *h, double dt);
/* body of update_general_ice_model_data (src/general_ice_model_data_module.f90:23-96) */
int ufm_update_general(ufm_handle *h, double time);
/* body of solve_SIA (src/ice_dynamics_module.f90:240-314) */
int ufm_solve_SIA(ufm_handle *h);
/* U_3D / V_3D half of solve_SIA_3D (src/ice_dynamics_module.f90:317-367, called from update_ice_temperature,
 * src/thermodynamics_module.f90:71) incl. apply_Neumann_boundary_3D; the vertical velocity W_3D feeds thermodynamics only
 * and stays on the host.  U_3D / V_3D set the third critical time step. */
int ufm_solve_SIA_3D(ufm_handle *h);
/* body of solve_SSA (src/ice_dynamics_module.f90:408-557) */
int ufm_solve_SSA(ufm_handle *h, ufm_ssa_stats *stats);
/* critical time steps of determine_timesteps_and_actions (src/UFEMISM_main_model.f90:738-778):
 * out = {dt_D_2D_min, dt_V_2D_SSA_min, dt_V_3D_SIA_min}, each already multiplied by 0.9 */
int ufm_cfl(ufm_handle *h, double out3[3]);

/* ---- pieces of solve_SSA, for kernel-level parity tests and profiling ---- */
/* basal_yield_stress [+ calculate_GL_flux] + gather into the AaAc arrays (:468-496) */
int ufm_ssa_prepare(ufm_handle *h);
/* SSA_effective_viscosity (:695-726) + the two sums of :512-513; sums2 = {sum_DN_sq, sum_N_sq} */
int ufm_ssa_viscosity(ufm_handle *h, double sums2[2]);
/* SSA_sliding_term (:727-779) + RHS and centre coefficients of solve_SSA_linearised (:581-596) */
int ufm_ssa_sliding_and_setup(ufm_handle *h);
/* SOR loop of solve_SSA_linearised (:598-692).  max_inner_override > 0 replaces C%SSA_max_inner_loops;
 * force_iters != 0 disables the stop tests so exactly that many iterations run. */
int ufm_ssa_sor(ufm_handle *h, int max_inner_override, int force_iters, ufm_ssa_stats *stats);
/* scatter AaAc -> U_SSA, V_SSA, Ux/Uy_SSA_Ac and rotate_xy_to_po (:548-555) */
int ufm_ssa_finish(ufm_handle *h);

/* ---- region time loop for benchmark physics (row N1: run_model, src/UFEMISM_main_model.f90:78-214,
 *      with determine_timesteps_and_actions :708-843 and the closed-form benchmark SMB,
 *      src/SMB_module.f90:55-97,172-238): no host round trip of fields per step ---- */
enum { UFM_T_SIA = 0, UFM_T_SSA, UFM_T_THERMO, UFM_T_CLIMATE, UFM_T_SMB, UFM_T_BMB, UFM_T_ELRA, UFM_T_OUTPUT, UFM_NT };
typedef struct ufm_region {
  double time, dt, dt_prev;
  double t0[UFM_NT], t1[UFM_NT], dtc[UFM_NT];
  int    do_[UFM_NT];
  double H0, R0, lambda;
  long   n_steps, n_sia, n_ssa, n_sor_total, n_outer_total;
  double dt_crit_last[3];
} ufm_region;
int ufm_region_init(ufm_region *r, double start_time);
int ufm_run_model(ufm_handle *h, ufm_region *r, double t_end, long max_steps);

/* The same loop in drop-in mode: the Fortran host owns the fields (its MPI shared-memory windows), so every
 * step uploads what CPU components may have changed (ELRA: Hb, dHb_dt, SL; SMB/BMB; remapped Hi; mask_noice)
 * and downloads what they read afterwards (Hi, Hi_prev, dHi_dt, Hs, U/V_SSA, U/V_SIA, D_SIA, mask).
 * All pointers are host arrays of length nV in reference vertex order; NULL members are skipped. */
typedef struct ufm_host_ice {
  const double *Hi, *Hb, *SL, *dHb_dt, *SMB_year, *BMB; const int *mask_noice;                 /* in  */
  double *Hi_out, *Hi_prev, *dHi_dt, *Hs, *U_SSA, *V_SSA, *U_SIA, *V_SIA, *D_SIA; int *mask;    /* out */
} ufm_host_ice;
int ufm_run_model_host(ufm_handle *h, ufm_region *r, double t_end, long max_steps, const ufm_host_ice *host);

/* ---- thermodynamics ----
 * ufm_update_ice_temperature replaces the body of update_ice_temperature (src/thermodynamics_module.f90:23-202): the
 * whole solve_SIA_3D (U, V and W, src/ice_dynamics_module.f90:317-405), bottom_frictional_heating (:281-311), the zeta
 * Jacobians (src/zeta_module.f90:86-112), one implicit step of the heat equation per column (DGTSV), the Neumann pass, and
 * the Robin-solution safety net (:204-279).  Inputs on the device: GHF, T2m, Ti (+ what update_general / solve_SSA left).
 * rc 0 ok (also for the benchmarks that skip thermodynamics, :44-64); -8: more than 1 % of the columns unstable (the
 * reference STOPs, :195-199); -9: singular column system (DGTSV info /= 0, STOP at :353); -10: no upwind triangle. */
typedef struct ufm_thermo_stats { int n_unstable; int rc; } ufm_thermo_stats;
int ufm_update_ice_temperature(ufm_handle *h, ufm_thermo_stats *st);
/* pieces, for kernel-level parity tests: W_3D from the resident U_3D / V_3D; the heat-equation step from the resident 3-D velocities */
int ufm_thermo_w3d(ufm_handle *h);
int ufm_thermo_heat(ufm_handle *h, ufm_thermo_stats *st);

/* ---- restart and help_fields files in the reference's own on-disk format.  NetCDF classic (CDF-1, or the
 *      64-bit-offset variant when an offset exceeds 2 GiB), written and read without a NetCDF library; dimension / variable names,
 *      order, types and attributes are those of create_restart_file_mesh / create_help_fields_file_mesh
 *      (src/netcdf_module.f90:489-820), so the MATLAB tooling (MATLAB/ReadMeshFromFile.m) and a restart of the Fortran model
 *      keep working.  All arrays are column-major with leading dimension = the mesh size given; NULL arrays are written as
 *      the NetCDF fill value.  rc: -11 I/O error, -12 not a restart file / variable missing or mismatching (the reference STOPs
 *      in inquire_*_var), -13 file exists (the reference aborts rather than overwrite, :508-512), -14 nZ mismatch (:3070),
 *      -15 time_to_restart_from outside the file's range (:3154), -16 unknown help field (:1035). ---- */
typedef struct ufm_nc_mesh {
  int nV, nTri, nC_mem, nAc, nV_transect, nVAaAc, nTriAaAc;
  const double *V;              /* (nV,2) */
  const int    *Tri;            /* (nTri,3) */
  const int    *nC, *C;         /* (nV), (nV,nC_mem) */
  const int    *niTri, *iTri;   /* (nV), (nV,nC_mem) */
  const int    *edge_index;     /* (nV) */
  const double *Tricc;          /* (nTri,2) */
  const int    *TriC;           /* (nTri,3) */
  const int    *Tri_edge_index; /* (nTri) */
  const double *VAc;            /* (nAc,2) */
  const int    *Aci, *iAci;     /* (nAc,4), (nV,nC_mem) */
  const double *VAaAc;          /* (nVAaAc,2) */
  const int    *TriAaAc;        /* (nTriAaAc,3) */
  const double *A, *R;          /* (nV) */
  const int    *vi_transect;    /* (nV_transect,2) */
  const double *w_transect;     /* (nV_transect,2) */
} ufm_nc_mesh;
typedef struct ufm_restart_frame {      /* what write_to_restart_file_mesh writes per time frame (src/netcdf_module.f90:195-205) */
  const double *Hi, *Hb, *Hs, *U_SIA, *V_SIA, *U_SSA, *V_SSA;   /* (nV) */
  const double *Ti;                      /* (nV,nZ) */
  const double *FirnDepth;               /* (nV,12) */
  const double *MeltPreviousYear;        /* (nV) */
} ufm_restart_frame;
typedef struct ufm_restart_frame_out {  /* what read_restart_file_init reads (src/netcdf_module.f90:3173-3180) */
  double *Hi, *Hb, *Hs, *Ti, *U_SSA, *V_SSA, *MeltPreviousYear, *FirnDepth;
} ufm_restart_frame_out;
/* create_restart_file_mesh (src/netcdf_module.f90:489-633) */
int ufm_restart_create(const char *filename, const ufm_nc_mesh *mesh, int nZ, const double *zeta);
/* write_to_restart_file_mesh (:180-214): append one time frame from host arrays / from the device; returns the frame index (>= 1) */
int ufm_restart_append(const char *filename, double time, const ufm_restart_frame *frame);
int ufm_restart_write(ufm_handle *h, const char *filename, double time, const double *FirnDepth, const double *MeltPreviousYear);
/* inquire_restart_file_mesh (:3012-3049), read_restart_file_mesh (:3104-3131): the primary mesh data */
int ufm_restart_inquire_mesh(const char *filename, int *nV, int *nTri, int *nC_mem);
int ufm_restart_read_mesh(const char *filename, double *V, int *nC, int *C, int *niTri, int *iTri, int *edge_index, int *Tri, double *Tricc,
                          int *TriC, int *Tri_edge_index);
/* inquire_restart_file_init (:3050-3103): rc 1 = zeta differs from the configuration (the reference only warns) */
int ufm_restart_inquire_init(const char *filename, int nZ, const double *zeta, int *nt);
/* read_restart_file_init (:3132-3185): the frame closest to time_to_restart_from, into host arrays / onto the device
 * (ufm_restart_load uploads Hi, Hb, Ti, U_SSA, V_SSA and returns the frame index) */
int ufm_restart_read_init(const char *filename, double time_to_restart_from, ufm_restart_frame_out *out, int *ti_out);
int ufm_restart_load(ufm_handle *h, const char *filename, double time_to_restart_from, double *FirnDepth, double *MeltPreviousYear);
/* create_help_fields_file_mesh (:634-820) with the fields of C%help_field_01..50; write_to_help_fields_file_mesh (:216-487):
 * host_data[k] NULL = take the field from the device (h may be NULL when every field comes from the host) */
int ufm_help_fields_create(const char *filename, const ufm_nc_mesh *mesh, int nZ, const double *zeta, int n_fields, const char *const *names);
int ufm_help_fields_write(ufm_handle *h, const char *filename, double time, int n_fields, const char *const *names, const void *const *host_data);
/* get_output_filenames (:66-174): first free <dir>restart_<NAM>_0000n.nc (kind 0) / <dir>help_fields_<NAM>_0000n.nc (kind 1); returns n */
int ufm_output_filename(const char *output_dir, const char *region_name, int kind, char *out, int out_len);

/* ---- instrumentation ---- */
int ufm_counters_get(ufm_handle *h, ufm_counters *out);
int ufm_counters_reset(ufm_handle *h);
/* tuning aid: with UFM_SOR_TRACE set in the environment the SOR kernel records, for its fourth iteration, per CTA and
 * phase (5 colours) the %globaltimer values {phase start, first warp done, last warp done, barrier left}; this copies up
 * to n_words 64-bit words of that record (layout [cta][6][4]) to `out`.  Returns the number of CTAs, <0 on error. */
int ufm_sor_trace_get(ufm_handle *h, unsigned long long *out, int n_words);

#ifdef __cplusplus
}
#endif
#endif
