// ufm_internal.cuh -- device-side data layout of libufemism_b200.so (not part of the ABI).
//
// HBM layout:
//  * three index spaces, each renumbered at upload for locality:
//      Aa   (nV vertices)      : Morton order of V
//      Ac   (nAc staggered)    : Morton order of the edge midpoint
//      AaAc (nV+nAc combined)  : colour-major; inside a colour by (degree, Morton); every colour
//                                block starts on a multiple of 32; domain-edge vertices (never swept)
//                                form a sixth block at the end.  Padding rows have deg = DEG_PAD.
//  * neighbour data are "sliced ELL": rows in groups of 32 (one warp), slice s has width w_s =
//    max degree in the slice and its entries live at off_s + c*32 + lane, so a warp reading column c
//    of its slice touches one contiguous 128 B (int) / 256 B (double) segment.
//  * everything is SoA except (U,V), (RHSx,RHSy), (e_u,e_v) which are double2 pairs because they
//    are always consumed together (one 16 B gather per neighbour instead of two 8 B gathers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#ifdef __cplusplus
#include <vector>
#endif

#include "../../include/ufemism_b200.h"

#define UFM_DEG_PAD 0xFF
#define UFM_SLICE 32
#define UFM_CHUNK 256       // rows per reduction chunk; every (block, owner) range of the AaAc layout is aligned to it
// mailbox words (one mailbox per GPU, written by its peers over NVLink)
#define MAIL_FLAG 0         // [q] epoch counter signalled by rank q
#define MAIL_RESID 8        // [3][8] max-residual bits of rank q for SOR iteration it%3
#define MAIL_EPOCH 40       // own epoch counter
#define MAIL_ABORT 41       // set when a peer wait timed out
#define MAIL_RED 64         // [2][8][4] small all-reduce slots: parity, rank, value (critical time steps, mask_sheet count)
#define MAIL_WORDS 128
// dataflow SOR sweep (k_ssa_sor_df): stage counters
#define DF_NONE 0xFFFFu
#define DF_CNT_STRIDE 8      // one 32 B sector per counter
#define DF_MAX_STAGES 256

// src/parameters_module.f90:9-21
#define UFM_PI 3.141592653589793
#define UFM_SEC_PER_YEAR 31556943.36
#define UFM_GRAV 9.81
#define UFM_N_FLOW 3.0
#define UFM_ICE_DENSITY 910.0
#define UFM_SEAWATER_DENSITY 1028.0
#define UFM_SMT 271.15

// mask bits (determine_masks, src/general_ice_model_data_module.f90:97-297); bits 12..15 = ice%mask code
enum { MB_LAND = 1, MB_OCEAN = 2, MB_LAKE = 4, MB_ICE = 8, MB_SHEET = 16, MB_SHELF = 32, MB_COAST = 64, MB_MARGIN = 128,
       MB_GL = 256, MB_CF = 512, MB_CODE_SHIFT = 12 };

// NORM2([x, y]) as gfortran evaluates it: libgfortran's _gfortran_norm2_r8 (m4/norm2.m4), a scaled sum of squares, NOT the
// hypotenuse function of libm / CUDA.  Every operation is an IEEE + * / sqrt, so device and host give the reference's bits
// (the library is built with -fmad=false).
__host__ __device__ __forceinline__ double ufm_norm2_2(const double x, const double y)
{
  double result = 0.0, scale = 1.0;
  if (x != 0.0) {
    const double a = fabs(x);
    if (scale < a) { const double val = scale / a; result = 1.0 + result * val * val; scale = a; }
    else { const double val = a / scale; result += val * val; }
  }
  if (y != 0.0) {
    const double a = fabs(y);
    if (scale < a) { const double val = scale / a; result = 1.0 + result * val * val; scale = a; }
    else { const double val = a / scale; result += val * val; }
  }
  return scale * sqrt(result);
}

// order-preserving integer image of a double (atomicMin on it = minimum of the doubles); critical time steps
__host__ __device__ __forceinline__ unsigned long long ord_key(double x)
{
  unsigned long long b;
#ifdef __CUDA_ARCH__
  b = (unsigned long long)__double_as_longlong(x);
#else
  memcpy(&b, &x, sizeof(b));
#endif
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double ord_unkey(unsigned long long k)
{
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double x;
#ifdef __CUDA_ARCH__
  x = __longlong_as_double((long long)b);
#else
  memcpy(&x, &b, sizeof(x));
#endif
  return x;
}
#ifdef __CUDACC__
// minimum over a 256-thread CTA, then one atomicMin on the key (every thread of the CTA must call this)
__device__ __forceinline__ void block_min_to_key(double v, unsigned long long *key)
{
  __shared__ double sh_min[8];
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh_min[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sh_min[threadIdx.x] : 1000.0;
    for (int o = 4; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0 && v < 1000.0) atomicMin(key, ord_key(v));   // the keys start at 1000 (the reference's initial value, UFEMISM_main_model.f90:736-738)
  }
}
#endif

// partitioned per-step kernels: does the calling rank own element idx?  own == NULL: single GPU / replicated kernels, everything is "owned"
#define UFM_OWNED(own, idx, rank) (!(own) || (own)[idx] == (unsigned char)(rank))

struct SlicedEll {
  int n_rows = 0;            // padded to a multiple of 32
  int n_slices = 0;
  long long n_entries = 0;
  long long *off = nullptr;  // [n_slices+1] element offsets (device)
  unsigned char *deg = nullptr;  // [n_rows] (device) row degree, UFM_DEG_PAD for padding rows
};

// one triangle as the upwind search reads it: corner coordinates (is_in_triangle), first-order neighbour functions, corners
struct TriRec { double ax, ay, bx, by, cx, cy, nx[3], ny[3]; int v[3]; int pad; };

struct DevMesh {
  int nV = 0, nAc = 0, M = 0;     // reference sizes
  int nVp = 0, nAcp = 0, Mp = 0;  // padded device sizes
  // permutations (device + host copies): ref (0-based) -> device position, and back (-1 = padding)
  int *aa_ref2dev = nullptr, *aa_dev2ref = nullptr;
  int *ac_ref2dev = nullptr, *ac_dev2ref = nullptr;
  int *m_ref2dev = nullptr, *m_dev2ref = nullptr;
  // ---- Aa ----
  SlicedEll aa;               // rows = Aa vertices
  int *aa_C = nullptr;        // neighbour Aa (device idx)
  int *aa_iAci = nullptr;     // Ac of each connection (device idx); bit 31 set when this vertex is Aci(aci,1)
  double *aa_Nx = nullptr, *aa_Ny = nullptr;  // neighbour functions per connection
  double *aa_Nx0 = nullptr, *aa_Ny0 = nullptr;  // home coefficients Nx(vi,nC+1)
  double *aa_A = nullptr;     // Voronoi area
  double *aa_sqrtApi = nullptr;  // SQRT(A/pi) (CFL)
  double *aa_rmin = nullptr;     // MIN( SQRT(A/pi), shortest connection ): numerator of the vertex's SSA critical time step
  unsigned char *aa_edge = nullptr;  // edge_index
  // ---- Ac ----
  int4 *ac_Aci = nullptr;     // vi, vj, vl, vr (device Aa idx)
  double *ac_Nx[4] = {}, *ac_Ny[4] = {}, *ac_No[4] = {};
  double *ac_Np = nullptr;
  double *ac_Cw = nullptr;    // Cw( Aci(aci,1), ci ) as used at ice_dynamics_module.f90:95
  double *ac_Dx = nullptr, *ac_Dy = nullptr;  // V(vj)-V(vi)
  double *ac_dist2 = nullptr;    // dist**2 with dist = SQRT(Dx**2 + Dy**2), as at UFEMISM_main_model.f90:751-752
  // ---- AaAc ----
  SlicedEll m;
  int *m_idx = nullptr;                   // neighbour positions
  double *m_cU = nullptr, *m_cV = nullptr;  // 4Nxx+Nyy, 4Nyy+Nxx per neighbour
  double *m_nxy = nullptr;                // Nxy per neighbour
  double *m_nx = nullptr, *m_ny = nullptr;
  double *m_nxy0 = nullptr, *m_nxysum = nullptr;  // home Nxy; pre-summed row (exact_xy = 0)
  double *m_nx0 = nullptr, *m_ny0 = nullptr;
  double *m_cU0 = nullptr, *m_cV0 = nullptr;      // 4Nxx+Nyy, 4Nyy+Nxx home
  int *m_src = nullptr;        // >= 0: Aa device idx; < 0: ~(Ac device idx); INT_MIN: padding
  int *aa2m = nullptr, *ac2m = nullptr;  // Aa / Ac device idx -> AaAc position
  // vertex partition: rows of every block (colours 1..5, edge block) are grouped by owner rank, so rank r's
  // share of block b is the contiguous slice range rng[b][r] = [begin, end); x-strips balanced by row count
  int P = 1, rank = 0;
  int rng[6][UFM_MAX_RANKS][3] = {};   // [begin, boundary_begin, end): interior rows first, rows that read a peer-owned row last
  int *rng_dev = nullptr;            // [(b*P + r)*3 + {0,1,2}]
  int *rng_all_dev = nullptr;        // [b*3 + {0,1,2}]: whole blocks (all owners)
  unsigned char *m_xmask = nullptr;  // per row: bit q set -> rank q reads this row, push new (U,V) to it
  unsigned char *m_sowner = nullptr; // per slice: owner rank
  unsigned nbr_mask = 0;             // ranks this rank exchanges rows with (CommDev::nbr)
  // ---- per-step kernels partitioned by owner: storage stays replicated, every Aa / Ac element is computed by the rank whose
  //      x-strip holds it; what a neighbour strip reads (thickness, out-flux factors, edge velocities) is exchanged once per kernel ----
  bool part_step = false;
  unsigned char *own_aa = nullptr, *own_ac = nullptr;   // owner rank per Aa / Ac device index (255: padding)
  unsigned char *act_aa1 = nullptr;                     // 1: k_geom_aa1 runs here on this rank (owned, or read by an owned Aa / Ac element)
  // halo lists, CSR over the peers: this rank SENDS x?_s_idx[x?_s_ptr[q] .. x?_s_ptr[q+1]) to rank q and RECEIVES x?_r_idx[...] from it
  // (device indices; a receive list is the peer's send list in the same order).  xa: Aa vertices, xc: Ac vertices
  int *xa_s_idx = nullptr, *xa_r_idx = nullptr, *xc_s_idx = nullptr, *xc_r_idx = nullptr;
  int xa_s_ptr[UFM_MAX_RANKS + 1] = {}, xa_r_ptr[UFM_MAX_RANKS + 1] = {}, xc_s_ptr[UFM_MAX_RANKS + 1] = {}, xc_r_ptr[UFM_MAX_RANKS + 1] = {};
  int x_region = 0;                  // doubles per (parity, sender) region of the exchange buffer
  int bc_rng[UFM_MAX_RANKS + 1] = {};  // Neumann rows grouped by owner
  int corner_owner[4] = {};
  int n_chunks = 0;
  // Neumann boundary lists
  int adj5_end = 0;            // single-GPU layout: slices [rng[4][0][0], adj5_end) hold the colour-5 rows adjacent to a domain-edge row
  int n_bc = 0;                // edge vertices except corners 1..4
  int *bc_pos = nullptr, *bc_ptr = nullptr, *bc_nbr = nullptr;
  int corner_pos[4] = {}, corner_n[4] = {};
  int *corner_dev = nullptr;   // [8] corner_pos, corner_n
  int *corner_nbr = nullptr;   // [4*16] neighbour position
  int *corner_row = nullptr;   // [4*16] bc row of that neighbour, or -1 when it is not an edge vertex
  double sor_bytes = 0;        // sum_i (80 + 20 n_i) over swept vertices
  // dataflow SOR sweep (k_ssa_sor_df, single GPU): x-band row order, no "adjacent colour-5 rows first" grouping
  bool df_layout = false;
  unsigned short *df_need = nullptr, *df_need_bc = nullptr;   // [n_slices], [n_bc + 4]
  unsigned *df_stage_cnt = nullptr;                           // [DF_MAX_STAGES * DF_CNT_STRIDE]
  int df_ready_for = 0;        // warps in the grid for which df_need was computed (0: not yet)
  int df_n_stages = 0, df_base[5] = {}, df_K[5] = {}, df_act[5] = {};
  // ---- thermodynamics only (present when the mesh was uploaded with Tri) ----
  bool has_tri = false;
  int nTri = 0;
  int *aa_iTri = nullptr;      // triangles around each vertex in iTri order, same sliced-ELL shape as aa_C; -1 = none
  double2 *aa_xy = nullptr;    // vertex coordinates
  double *aa_R = nullptr;      // mesh%R
  TriRec *tri = nullptr;  // [nTri] reference triangle order
};


struct DevState {
  // Aa
  double *Hi = nullptr, *Hi_alt = nullptr, *Hb = nullptr, *SL = nullptr, *Hs = nullptr, *dHb_dt = nullptr, *dHi_dt = nullptr, *dHs_dt = nullptr;
  double *dHi_dx = nullptr, *dHi_dy = nullptr, *dHs_dx = nullptr, *dHs_dy = nullptr, *dHs_dx_shelf = nullptr, *dHs_dy_shelf = nullptr;
  double *U_SIA = nullptr, *V_SIA = nullptr, *D_SIA = nullptr, *U_SSA = nullptr, *V_SSA = nullptr, *SMB_year = nullptr, *BMB = nullptr;
  double *thk_factor = nullptr, *thk_smb = nullptr, *thk_flux = nullptr;
  double *U_3D = nullptr, *V_3D = nullptr;  // (nV,nZ) device layout k-major: [k*nVp + v]
  double *Ti = nullptr;                     // (nV,nZ) englacial temperature, k-major (realistic flow factor or thermodynamics)
  double *Ti_new = nullptr, *W_3D = nullptr;  // (nV,nZ) thermodynamics
  double *GHF = nullptr, *T2m = nullptr, *fric_heat = nullptr;  // (nV), (nV,12) month-major, (nV)
  double *A_mean = nullptr, *A_mean_Ac = nullptr;  // A_flow_mean on Aa / Ac (realistic flow factor only)
  double *Afac = nullptr;                   // (m_enh_ssa*0.5*A_flow_mean_AaAc)**(-1/n) per AaAc row (realistic only)
  bool realistic_A = false;                 // C%do_benchmark_experiment == .FALSE.
  int *mask_noice = nullptr;
  unsigned *mbits = nullptr;
  // Ac
  double *Hi_Ac = nullptr, *Hb_Ac = nullptr, *SL_Ac = nullptr, *Hs_Ac = nullptr;
  double *dHi_Ac[4] = {}, *dHb_Ac[4] = {}, *dHs_Ac[4] = {}, *dSL_Ac[4] = {};  // x, y, p, o
  double *dHs_dx_shelf_Ac = nullptr, *dHs_dy_shelf_Ac = nullptr;
  double *U_SIA_Ac[4] = {}, *U_SSA_Ac[4] = {};  // x, y, p, o
  double *D_SIA_Ac = nullptr, *Qabs_GL_Ac = nullptr, *Qp_GL_Ac = nullptr;
  unsigned *mbits_Ac = nullptr;
  // AaAc
  double2 *UV = nullptr, *RHS = nullptr, *E = nullptr, *rhsnum = nullptr, *dU = nullptr, *dV = nullptr;
  double *eta = nullptr, *N = nullptr, *S = nullptr, *tau_c = nullptr, *phi = nullptr, *Hm = nullptr;
  unsigned char *mflag = nullptr;  // bit0 grounded (not floating), bit1 held fixed (GL flux)
  double A_flow_const = 1.0e-16;   // benchmark flow factor (ice_physical_properties)
  // reductions / control
  double *partials = nullptr;      // [2*n_partial]
  double *red_scratch = nullptr;   // [2*64] block sums of the RN reduction tree
  unsigned long long *ctrl = nullptr;  // SOR control block
  unsigned long long *mail = nullptr;  // mailbox for peer GPUs
  double *xbuf = nullptr;              // halo exchange buffer the peers write into: [2 parities][P senders][x_region]
  double *scal = nullptr;          // small result scratch (device), mirrored in pinned host memory
  double *scal_h = nullptr;
};

// peer-mapped buffers of the other ranks of a partitioned run (CUDA IPC), index = rank; [own rank] = own buffers
struct CommDev {
  int P, rank;
  unsigned nbr;   // bit q: rank q reads rows of this rank or owns rows this rank reads (for x-strips: the two neighbours)
  double2 *uv[UFM_MAX_RANKS];
  double *partials[UFM_MAX_RANKS];
  unsigned long long *mail[UFM_MAX_RANKS];
  double *xbuf[UFM_MAX_RANKS];
};

// words of DevState::ctrl (unsigned long long[128]) -- one table so that no two users overlap:
//   [0..2] SOR max-residual slots   [8..10] SOR results   [12] SOR fused-Neumann counter   [16] mask_sheet sum
//   [24..26] CFL minima keys   [28..29] thermodynamics status   [30] RN-reduction ticket   [32..95] SOR grid barrier   [96..103] SCTL_*
#define CTRL_CFL_KEYS 24
#define CTRL_CFL_TMP 104          // [104..106] all-reduced copy of the CFL keys (partitioned per-step kernels)
#define CTRL_THERMO_STATUS 28
#define CTRL_RN_TICKET 30
#define UFM_XFER_SLOTS 20
#define UFM_SOR_CHUNK_DEFAULT 1
#define UFM_SOR_FUSE_BC_DEFAULT 1
#define UFM_SOR_BAR_DEFAULT 1
struct ufm_handle {
  int device = 0;
  ufm_params P;
  double zeta3[UFM_MAX_NZ];      // C%zeta**n_flow, evaluated with the host libm like the reference
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_pool[32] = {};
  int num_sms = 0;
  bool has_mesh = false;
  DevMesh mesh;
  DevState st;
  ufm_counters cnt;
  int sor_grid = 0, sor_block = 1024;
  size_t sor_smem = 0;
  int sor_chunk = 0, sor_fuse_bc = 0, sor_bar = 0;  // SOR scheduling switches (env UFM_SOR_CHUNK / _FUSE_BC / _BAR)
  // critical time steps (determine_timesteps_and_actions, UFEMISM_main_model.f90:747-773): each of the three minima only changes when
  // the field behind it does, so the kernels that write D_SIA_Ac / (U,V)_SSA / (U,V)_3D reduce it in their epilogue into
  // ctrl[CTRL_CFL_KEYS + k]; cfl_ok[k] == false (a field was uploaded, remapped, ...) makes ufm_cfl recompute it from the fields
  bool cfl_ok[3] = {false, false, false};
  // device-driven region loop (ufm_api.cu run_model_device): while it enqueues steps these point at the control block's gates
  // (run, SIA due, SMB due, thermodynamics timer due) and at the time step; NULL otherwise
  const int *gate[4] = {nullptr, nullptr, nullptr, nullptr};
  const double *dt_dev = nullptr;
  void *stepctl_dev = nullptr, *stepctl_host = nullptr;   // StepCtl in device / pinned host memory
  int pow_exact = 0;             // 1: device pow reproduces the host libm's bits (ufm_pow.cuh), 0: CUDA's pow (within 2 ulp)
  int pow_reason = 0;            // why not, when pow_exact == 0 (return code of ufm_powtab_build)
  int sor_df = 0;                // 1: dataflow sweep kernel (env UFM_SOR_DATAFLOW, needs the mesh's df_layout)
  const void *carveout_done_for = nullptr;   // kernel variant whose L1 carve-out preference has been set on this handle's device
  int visc_minb = -1;            // resident CTAs per SM of the viscosity kernel (env UFM_VISC_MINB)
  unsigned long long *sor_trace = nullptr;   // tuning aid, see SorArgs::trace
  int part_rank = 0, part_n = 1;   // set by ufm_partition_set before the mesh upload
  bool comm_connected = false;
  CommDev comm;
  void *ipc_opened[4 * UFM_MAX_RANKS] = {};
  int xparity = 0;                 // parity of the next halo exchange / small all-reduce (double-buffered regions)
  std::vector<unsigned char> *owner_ref = nullptr;   // owner rank of every AaAc vertex in REFERENCE order (partitioned runs; ufm_partition_owner_of)
  // remap stash: a field of the OLD mesh and its Aa gradients, reference order, survives ufm_mesh_upload
  struct Stash { int field = -1; int n = 0; double *d = nullptr, *ddx = nullptr, *ddy = nullptr; } stash[4];
  // host buffers page-locked with ufm_host_register (base, bytes): field copies from/to them are DMA'd directly
  void *pinned_base[64] = {};
  size_t pinned_bytes[64] = {};
  int n_pinned = 0;
  // device arena: every mesh / state array except the three CUDA-IPC-shared ones is carved out of a few large cudaMalloc
  // chunks that SURVIVE ufm_mesh_free, so a re-upload after a mesh update (src/UFEMISM_main_model.f90:294) pays no
  // allocation cost (cudaMalloc of ~100 arrays was 1-2 s of a 1 M-vertex upload)
  struct Chunk { char *base; size_t cap; };
  Chunk arena[64] = {};
  int arena_n = 0, arena_cur = 0;
  size_t arena_used = 0, arena_total = 0;
  // drop-in mode (ufm_run_model_host): per-field device / pinned-host slots and a copy stream, so that the H2D / D2H copies of a
  // step overlap each other and the SSA solve instead of being serialised with a host synchronisation each
  cudaStream_t xfer_stream = nullptr;      // host -> device copies of drop-in mode
  cudaStream_t xfer_stream_out = nullptr;  // device -> host copies (their own stream: the two directions use different copy engines)
  cudaEvent_t xfer_ev[2 * UFM_XFER_SLOTS] = {};
  char *xfer_dev = nullptr, *xfer_host = nullptr;   // UFM_XFER_SLOTS slots of xfer_slot_bytes each (host slots only if a buffer is not page-locked)
  size_t xfer_slot_bytes = 0;
  struct Pending { void *dst; const void *src; size_t bytes; } xfer_pending[UFM_XFER_SLOTS] = {};
  int xfer_n_pending = 0;
  void *secondary = nullptr;     // ufm_secondary: host arrays derived by ufm_mesh_upload_primary (ufm_mesh_primary.cpp)
  // the three buffers peers map through CUDA IPC keep allocations of their own (outside the arena); like the arena they survive
  // ufm_mesh_free and are reused by the next upload when large enough (cudaFree / cudaMalloc cost 0.2 s of a re-upload)
  struct OwnBuf { void *p = nullptr; size_t bytes = 0; } own_buf[4];
  double *scal_h_keep = nullptr;
  void *staging = nullptr;       // pinned host staging for upload/download permutation
  size_t staging_bytes = 0;
  void *dev_staging = nullptr;
  size_t dev_staging_bytes = 0;
};

int ufm_arena_alloc(ufm_handle *h, size_t bytes, void **out);
void ufm_arena_release(ufm_handle *h);
void ufm_secondary_free(ufm_handle *h);
void ufm_own_release(ufm_handle *h);
int ufm_set_error(int rc, const char *fmt, ...);
int ufm_cuda_check(cudaError_t e, const char *what);
#define UFM_CUDA(x) do { int rc__ = ufm_cuda_check((x), #x); if (rc__) return rc__; } while (0)

// launchers (each returns 0 or a negative rc)
int ufm_k_geom(ufm_handle *h, double time);
int ufm_k_sia(ufm_handle *h);
int ufm_k_sia3d(ufm_handle *h);
int ufm_k_neumann3d_pair(ufm_handle *h, double *A3, double *B3);
int ufm_k_thermo_w3d(ufm_handle *h);
int ufm_k_thermo_heat(ufm_handle *h, ufm_thermo_stats *st);
int ufm_k_remap_stash(ufm_handle *h, int slot, double *field_dev);
int ufm_k_remap_apply(ufm_handle *h, int slot, const ufm_remap_cons *map, int order, double *field_dev);
int ufm_k_thickness(ufm_handle *h, double dt);
int ufm_k_cfl(ufm_handle *h, double out3[3]);
int ufm_cfl_key_reset(ufm_handle *h, int which);
int ufm_k_cfl3d_enqueue(ufm_handle *h);
int ufm_k_ssa_prepare(ufm_handle *h);
int ufm_k_ssa_viscosity(ufm_handle *h, double sums2[2]);
int ufm_k_ssa_sliding_setup(ufm_handle *h);
int ufm_k_ssa_gradients(ufm_handle *h);
int ufm_comm_reset(ufm_handle *h);
int ufm_halo_exchange(ufm_handle *h, int kind, int narr, double *const *arrays);   // kind 0: Aa lists, 1: Ac lists; collective
int ufm_peer_allreduce(ufm_handle *h, unsigned long long *vals_dev, int n, int op);  // op 0: min (ordered keys), 1: sum; collective
int ufm_push_uv_halo(ufm_handle *h);
int ufm_k_ssa_sor(ufm_handle *h, int max_inner, int force_iters, ufm_ssa_stats *stats);
int ufm_k_ssa_finish(ufm_handle *h);
int ufm_k_ssa_outer_loop(ufm_handle *h, ufm_ssa_stats *stats);
int ufm_k_ssa_zero(ufm_handle *h);
int ufm_k_sum_mask_sheet(ufm_handle *h, long long *out);
int ufm_k_smb_benchmark(ufm_handle *h, double time, double H0, double R0, double lambda);
int ufm_k_permute(ufm_handle *h, int kind, int is_int, int to_device, void *dev_field, void *dev_staging, int n_ref);
int ufm_sor_configure(ufm_handle *h);
struct UfmPowTab;
int ufm_ssa_powtab_init(const UfmPowTab *t);
int ufm_geom_powtab_init(const UfmPowTab *t);
int ufm_thermo_powtab_init(const UfmPowTab *t);
extern "C" int ufm_powtab_build(UfmPowTab *out);
