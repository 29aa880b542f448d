// ufm_api.cu -- extern "C" entry points of libufemism_b200.so (include/ufemism_b200.h).
#include <math.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ufm_internal.cuh"
#include "ufm_pow.cuh"

int ufm_mesh_upload_impl(ufm_handle *h, const ufm_mesh_desc *d);
int ufm_mesh_free_impl(ufm_handle *h);
int ufm_perm_double(ufm_handle *h, int n, const int *r2d, double *dev, int stride, int comp, double *ref_dev, int to_device);
int ufm_perm_int(ufm_handle *h, int n, const int *r2d, int *dev, int *ref_dev, int to_device);
int ufm_perm_mask(ufm_handle *h, int n, const int *r2d, const unsigned *bits, unsigned arg, int mode, int *ref_dev);
int ufm_perm_3d(ufm_handle *h, int n, int nZ, int nVp, const int *r2d, double *dev, double *ref_dev, int to_device);

static thread_local char g_err[512] = "";

int ufm_set_error(int rc, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return rc;
}
int ufm_cuda_check(cudaError_t e, const char *what)
{
  if (e == cudaSuccess) return 0;
  return ufm_set_error(-100 - (int)e, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

static int check_params(const ufm_params *p)
{
  if (!p) return ufm_set_error(-2, "NULL params");
  if (p->nZ < 2 || p->nZ > UFM_MAX_NZ) return ufm_set_error(-2, "nZ = %d out of range [2,%d]", p->nZ, UFM_MAX_NZ);
  if (p->benchmark < 0 || p->benchmark > UFM_BM_SSA_ICESTREAM) return ufm_set_error(-2, "unknown benchmark id %d", p->benchmark);
  if (p->SSA_max_inner_loops < 1 || p->SSA_max_outer_loops < 1) return ufm_set_error(-2, "SSA loop limits must be >= 1");
  return 0;
}
static void derive_params(ufm_handle *h)
{
  // C%zeta**n_flow with the host libm, as the reference evaluates it (ice_dynamics_module.f90:277)
  for (int k = 0; k < h->P.nZ; k++) h->zeta3[k] = pow(h->P.zeta[k], UFM_N_FLOW);
}

extern "C" {

int ufm_abi_version(void) { return UFM_ABI_VERSION; }
const char *ufm_last_error(void) { return g_err; }

int ufm_create(int device, const ufm_params *params, ufm_handle **out)
{
  if (!out) return ufm_set_error(-2, "ufm_create: NULL out");
  *out = nullptr;
  int rc = check_params(params);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return ufm_set_error(e != cudaSuccess ? -100 - (int)e : -5, "ufm_create: no usable CUDA device (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return ufm_set_error(-2, "ufm_create: device %d out of range (0..%d)", device, ndev - 1);
  UFM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  UFM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return ufm_set_error(-5, "ufm_create: device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
  if (!prop.cooperativeLaunch) return ufm_set_error(-5, "ufm_create: device lacks cooperative launch");
  ufm_handle *h = new ufm_handle();
  h->device = device;
  h->P = *params;
  h->num_sms = prop.multiProcessorCount;
  memset(&h->cnt, 0, sizeof(h->cnt));
  derive_params(h);
  UFM_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  {
    // x**y with the host libm's bits on the device (ufm_pow.cuh): tables found in this process's libm and validated once per process,
    // then copied into every kernel translation unit on this handle's device
    static UfmPowTab tab;
    static int built = 0, reason = 0;
    if (!built) { reason = ufm_powtab_build(&tab); built = 1; }
    h->pow_reason = reason;
    if (tab.enabled) {
      if (ufm_ssa_powtab_init(&tab) || ufm_geom_powtab_init(&tab) || ufm_thermo_powtab_init(&tab)) { delete h; return ufm_set_error(-3, "ufm_create: could not upload the pow tables"); }
      h->pow_exact = tab.tan_enabled ? 3 : 1;
    }
  }
  UFM_CUDA(cudaEventCreate(&h->ev0));
  UFM_CUDA(cudaEventCreate(&h->ev1));
  *out = h;
  return 0;
}

int ufm_destroy(ufm_handle *h)
{
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  ufm_mesh_free_impl(h);
  ufm_secondary_free(h);
  ufm_own_release(h);
  ufm_arena_release(h);
  for (int k = 0; k < h->n_pinned; k++) cudaHostUnregister(h->pinned_base[k]);
  for (auto &q : h->stash) if (q.d) { cudaFree(q.d); cudaFree(q.ddx); cudaFree(q.ddy); }
  if (h->staging) cudaFreeHost(h->staging);
  if (h->dev_staging) cudaFree(h->dev_staging);
  if (h->xfer_dev) cudaFree(h->xfer_dev);
  if (h->xfer_host) cudaFreeHost(h->xfer_host);
  for (auto e : h->xfer_ev) if (e) cudaEventDestroy(e);
  if (h->xfer_stream) cudaStreamDestroy(h->xfer_stream);
  if (h->xfer_stream_out) cudaStreamDestroy(h->xfer_stream_out);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
  for (auto e : h->ev_pool) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(h->own_stream);
  delete h;
  return 0;
}

int ufm_set_params(ufm_handle *h, const ufm_params *params)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  int rc = check_params(params);
  if (rc) return rc;
  if (h->has_mesh && params->nZ != h->P.nZ) return ufm_set_error(-2, "ufm_set_params: nZ cannot change while a mesh is resident");
  if (h->has_mesh && (params->benchmark == UFM_BM_NONE) != (h->P.benchmark == UFM_BM_NONE))
    return ufm_set_error(-2, "ufm_set_params: do_benchmark_experiment cannot change while a mesh is resident");
  h->P = *params;
  derive_params(h);
  return h->has_mesh ? ufm_sor_configure(h) : 0;
}

int ufm_set_stream(ufm_handle *h, void *cuda_stream)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}
int ufm_synchronize(ufm_handle *h)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int ufm_mesh_upload(ufm_handle *h, const ufm_mesh_desc *mesh)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  UFM_CUDA(cudaSetDevice(h->device));
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  int rc = ufm_mesh_upload_impl(h, mesh);
  if (rc) ufm_mesh_free_impl(h);
  ufm_secondary_free(h);   // host arrays derived by an earlier ufm_mesh_upload_primary describe the mesh that was just replaced
  return rc;
}
int ufm_mesh_free(ufm_handle *h)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  ufm_secondary_free(h);
  return ufm_mesh_free_impl(h);
}

/* ---- partitioned runs: CUDA-IPC plumbing ---- */
int ufm_partition_set(ufm_handle *h, int rank, int nranks)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  if (nranks < 1 || nranks > UFM_MAX_RANKS || rank < 0 || rank >= nranks) return ufm_set_error(-2, "ufm_partition_set: rank %d of %d out of range (max %d ranks)", rank, nranks, UFM_MAX_RANKS);
  if (h->has_mesh) return ufm_set_error(-2, "ufm_partition_set must precede ufm_mesh_upload");
  h->part_rank = rank; h->part_n = nranks;
  return 0;
}
int ufm_comm_export(ufm_handle *h, void *blob)
{
  if (!h || !h->has_mesh || !blob) return ufm_set_error(-2, "ufm_comm_export: no mesh resident");
  UFM_CUDA(cudaSetDevice(h->device));
  static_assert(4 * sizeof(cudaIpcMemHandle_t) <= UFM_COMM_BLOB_BYTES, "blob too small");
  memset(blob, 0, UFM_COMM_BLOB_BYTES);
  cudaIpcMemHandle_t hd[4];
  memset(hd, 0, sizeof(hd));
  UFM_CUDA(cudaIpcGetMemHandle(&hd[0], h->st.UV));
  UFM_CUDA(cudaIpcGetMemHandle(&hd[1], h->st.partials));
  UFM_CUDA(cudaIpcGetMemHandle(&hd[2], h->st.mail));
  if (h->st.xbuf) UFM_CUDA(cudaIpcGetMemHandle(&hd[3], h->st.xbuf));
  memcpy(blob, hd, sizeof(hd));
  return 0;
}
int ufm_comm_connect(ufm_handle *h, const void *blobs)
{
  if (!h || !h->has_mesh || !blobs) return ufm_set_error(-2, "ufm_comm_connect: no mesh resident");
  UFM_CUDA(cudaSetDevice(h->device));
  ufm_comm_reset(h);
  const int P = h->mesh.P, me = h->mesh.rank;
  for (int q = 0; q < P; q++) {
    if (q == me) continue;
    cudaIpcMemHandle_t hd[4];
    memcpy(hd, (const char *)blobs + (size_t)q * UFM_COMM_BLOB_BYTES, sizeof(hd));
    void *ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    const int nh = h->st.xbuf ? 4 : 3;
    for (int k = 0; k < nh; k++) {
      UFM_CUDA(cudaIpcOpenMemHandle(&ptr[k], hd[k], cudaIpcMemLazyEnablePeerAccess));
      h->ipc_opened[4 * q + k] = ptr[k];
    }
    h->comm.uv[q] = (double2 *)ptr[0]; h->comm.partials[q] = (double *)ptr[1]; h->comm.mail[q] = (unsigned long long *)ptr[2];
    h->comm.xbuf[q] = (double *)ptr[3];
  }
  h->comm.xbuf[me] = h->st.xbuf;
  h->comm.nbr = h->mesh.nbr_mask & ~(1u << me);
  { const char *e = getenv("UFM_PEER_ALL"); if (e && atoi(e) != 0) h->comm.nbr = ((1u << P) - 1u) & ~(1u << me); }   // A/B: every phase signals every peer (round 1)
  h->comm_connected = true;
  return 0;
}
}  // extern "C"
int ufm_comm_reset(ufm_handle *h)
{
  for (int k = 0; k < 4 * UFM_MAX_RANKS; k++) if (h->ipc_opened[k]) { cudaIpcCloseMemHandle(h->ipc_opened[k]); h->ipc_opened[k] = nullptr; }
  h->comm_connected = false;
  return 0;
}
extern "C" {

/* ---- field table ---- */
enum { K_AA = 0, K_AC = 1, K_M = 2 };
struct FieldRef { int kind; int is_int; double *d; int stride, comp; int *i; const unsigned *bits; unsigned barg; int bmode; int is3d; };

static int field_ref(ufm_handle *h, int f, FieldRef *r)
{
  DevState &s = h->st;
  memset(r, 0, sizeof(*r));
  r->stride = 1;
#define D_(K, P) { r->kind = K; r->d = (P); return 0; }
#define D2_(P, C) { r->kind = K_M; r->d = (double *)(P); r->stride = 2; r->comp = C; return 0; }
#define B_(K, BITS, ARG) { r->kind = K; r->is_int = 1; r->bits = (BITS); r->barg = ARG; r->bmode = 0; return 0; }
  switch (f) {
    case UFM_F_HI: D_(K_AA, s.Hi) case UFM_F_HB: D_(K_AA, s.Hb) case UFM_F_SL: D_(K_AA, s.SL) case UFM_F_DHB_DT: D_(K_AA, s.dHb_dt)
    case UFM_F_SMB_YEAR: D_(K_AA, s.SMB_year) case UFM_F_BMB: D_(K_AA, s.BMB)
    case UFM_F_MASK_NOICE: { r->kind = K_AA; r->is_int = 1; r->i = s.mask_noice; return 0; }
    case UFM_F_HS: D_(K_AA, s.Hs) case UFM_F_DHI_DT: D_(K_AA, s.dHi_dt) case UFM_F_DHS_DT: D_(K_AA, s.dHs_dt) case UFM_F_HI_PREV: D_(K_AA, s.Hi_alt)
    case UFM_F_DHI_DX: D_(K_AA, s.dHi_dx) case UFM_F_DHI_DY: D_(K_AA, s.dHi_dy) case UFM_F_DHS_DX: D_(K_AA, s.dHs_dx) case UFM_F_DHS_DY: D_(K_AA, s.dHs_dy)
    case UFM_F_DHS_DX_SHELF: D_(K_AA, s.dHs_dx_shelf) case UFM_F_DHS_DY_SHELF: D_(K_AA, s.dHs_dy_shelf)
    case UFM_F_U_SIA: D_(K_AA, s.U_SIA) case UFM_F_V_SIA: D_(K_AA, s.V_SIA) case UFM_F_D_SIA: D_(K_AA, s.D_SIA)
    case UFM_F_U_SSA: D_(K_AA, s.U_SSA) case UFM_F_V_SSA: D_(K_AA, s.V_SSA)
    case UFM_F_MASK_LAND: B_(K_AA, s.mbits, MB_LAND) case UFM_F_MASK_OCEAN: B_(K_AA, s.mbits, MB_OCEAN) case UFM_F_MASK_LAKE: B_(K_AA, s.mbits, MB_LAKE)
    case UFM_F_MASK_ICE: B_(K_AA, s.mbits, MB_ICE) case UFM_F_MASK_SHEET: B_(K_AA, s.mbits, MB_SHEET) case UFM_F_MASK_SHELF: B_(K_AA, s.mbits, MB_SHELF)
    case UFM_F_MASK_COAST: B_(K_AA, s.mbits, MB_COAST) case UFM_F_MASK_MARGIN: B_(K_AA, s.mbits, MB_MARGIN) case UFM_F_MASK_GL: B_(K_AA, s.mbits, MB_GL)
    case UFM_F_MASK_CF: B_(K_AA, s.mbits, MB_CF)
    case UFM_F_MASK: { r->kind = K_AA; r->is_int = 1; r->bits = s.mbits; r->bmode = 1; return 0; }
    case UFM_F_HI_AC: D_(K_AC, s.Hi_Ac) case UFM_F_HB_AC: D_(K_AC, s.Hb_Ac) case UFM_F_HS_AC: D_(K_AC, s.Hs_Ac) case UFM_F_SL_AC: D_(K_AC, s.SL_Ac)
    case UFM_F_DHI_DX_AC: D_(K_AC, s.dHi_Ac[0]) case UFM_F_DHI_DY_AC: D_(K_AC, s.dHi_Ac[1]) case UFM_F_DHI_DP_AC: D_(K_AC, s.dHi_Ac[2]) case UFM_F_DHI_DO_AC: D_(K_AC, s.dHi_Ac[3])
    case UFM_F_DHB_DX_AC: D_(K_AC, s.dHb_Ac[0]) case UFM_F_DHB_DY_AC: D_(K_AC, s.dHb_Ac[1]) case UFM_F_DHB_DP_AC: D_(K_AC, s.dHb_Ac[2]) case UFM_F_DHB_DO_AC: D_(K_AC, s.dHb_Ac[3])
    case UFM_F_DHS_DX_AC: D_(K_AC, s.dHs_Ac[0]) case UFM_F_DHS_DY_AC: D_(K_AC, s.dHs_Ac[1]) case UFM_F_DHS_DP_AC: D_(K_AC, s.dHs_Ac[2]) case UFM_F_DHS_DO_AC: D_(K_AC, s.dHs_Ac[3])
    case UFM_F_DSL_DX_AC: D_(K_AC, s.dSL_Ac[0]) case UFM_F_DSL_DY_AC: D_(K_AC, s.dSL_Ac[1]) case UFM_F_DSL_DP_AC: D_(K_AC, s.dSL_Ac[2]) case UFM_F_DSL_DO_AC: D_(K_AC, s.dSL_Ac[3])
    case UFM_F_DHS_DX_SHELF_AC: D_(K_AC, s.dHs_dx_shelf_Ac) case UFM_F_DHS_DY_SHELF_AC: D_(K_AC, s.dHs_dy_shelf_Ac)
    case UFM_F_UX_SIA_AC: D_(K_AC, s.U_SIA_Ac[0]) case UFM_F_UY_SIA_AC: D_(K_AC, s.U_SIA_Ac[1]) case UFM_F_UP_SIA_AC: D_(K_AC, s.U_SIA_Ac[2]) case UFM_F_UO_SIA_AC: D_(K_AC, s.U_SIA_Ac[3])
    case UFM_F_D_SIA_AC: D_(K_AC, s.D_SIA_Ac)
    case UFM_F_UX_SSA_AC: D_(K_AC, s.U_SSA_Ac[0]) case UFM_F_UY_SSA_AC: D_(K_AC, s.U_SSA_Ac[1]) case UFM_F_UP_SSA_AC: D_(K_AC, s.U_SSA_Ac[2]) case UFM_F_UO_SSA_AC: D_(K_AC, s.U_SSA_Ac[3])
    case UFM_F_QABS_GL_AC: D_(K_AC, s.Qabs_GL_Ac) case UFM_F_QP_GL_AC: D_(K_AC, s.Qp_GL_Ac)
    case UFM_F_MASK_LAND_AC: B_(K_AC, s.mbits_Ac, MB_LAND) case UFM_F_MASK_OCEAN_AC: B_(K_AC, s.mbits_Ac, MB_OCEAN) case UFM_F_MASK_LAKE_AC: B_(K_AC, s.mbits_Ac, MB_LAKE)
    case UFM_F_MASK_ICE_AC: B_(K_AC, s.mbits_Ac, MB_ICE) case UFM_F_MASK_SHEET_AC: B_(K_AC, s.mbits_Ac, MB_SHEET) case UFM_F_MASK_SHELF_AC: B_(K_AC, s.mbits_Ac, MB_SHELF)
    case UFM_F_MASK_COAST_AC: B_(K_AC, s.mbits_Ac, MB_COAST) case UFM_F_MASK_MARGIN_AC: B_(K_AC, s.mbits_Ac, MB_MARGIN) case UFM_F_MASK_GL_AC: B_(K_AC, s.mbits_Ac, MB_GL)
    case UFM_F_MASK_CF_AC: B_(K_AC, s.mbits_Ac, MB_CF)
    case UFM_F_MASK_AC: { r->kind = K_AC; r->is_int = 1; r->bits = s.mbits_Ac; r->bmode = 1; return 0; }
    case UFM_F_U_SSA_AAAC: D2_(s.UV, 0) case UFM_F_V_SSA_AAAC: D2_(s.UV, 1)
    case UFM_F_ETA_AAAC: D_(K_M, s.eta) case UFM_F_N_AAAC: D_(K_M, s.N) case UFM_F_S_AAAC: D_(K_M, s.S) case UFM_F_TAU_C_AAAC: D_(K_M, s.tau_c)
    case UFM_F_PHI_FRIC_AAAC: D_(K_M, s.phi)
    case UFM_F_RHSX_AAAC: D2_(s.RHS, 0) case UFM_F_RHSY_AAAC: D2_(s.RHS, 1) case UFM_F_EU_I_AAAC: D2_(s.E, 0) case UFM_F_EV_I_AAAC: D2_(s.E, 1)
    case UFM_F_DU_DX_AAAC: D2_(s.dU, 0) case UFM_F_DU_DY_AAAC: D2_(s.dU, 1) case UFM_F_DV_DX_AAAC: D2_(s.dV, 0) case UFM_F_DV_DY_AAAC: D2_(s.dV, 1)
    case UFM_F_U_3D: { r->kind = K_AA; r->d = s.U_3D; r->is3d = h->P.nZ; return 0; }
    case UFM_F_V_3D: { r->kind = K_AA; r->d = s.V_3D; r->is3d = h->P.nZ; return 0; }
    case UFM_F_TI: {
      if (!s.Ti) return ufm_set_error(-2, "Ti is only resident when do_benchmark_experiment is .FALSE. or the mesh carries Tri (thermodynamics)");
      r->kind = K_AA; r->d = s.Ti; r->is3d = h->P.nZ; return 0;
    }
    case UFM_F_W_3D: case UFM_F_GHF: case UFM_F_T2M: case UFM_F_FRICTIONAL_HEATING: {
      if (!h->mesh.has_tri) return ufm_set_error(-2, "thermodynamics fields need a mesh uploaded with Tri / iTri / R / NxTri / NyTri");
      r->kind = K_AA;
      if (f == UFM_F_W_3D) { r->d = s.W_3D; r->is3d = h->P.nZ; }
      else if (f == UFM_F_T2M) { r->d = s.T2m; r->is3d = 12; }
      else r->d = f == UFM_F_GHF ? s.GHF : s.fric_heat;
      return 0;
    }
    case UFM_F_A_FLOW_MEAN: D_(K_AA, s.A_mean) case UFM_F_A_FLOW_MEAN_AC: D_(K_AC, s.A_mean_Ac)
    default: break;
  }
#undef D_
#undef D2_
#undef B_
  return ufm_set_error(-2, "unknown field id %d", f);
}

static bool is_pinned(const ufm_handle *h, const void *p, size_t bytes)
{
  for (int k = 0; k < h->n_pinned; k++)
    if ((const char *)p >= (const char *)h->pinned_base[k] && (const char *)p + bytes <= (const char *)h->pinned_base[k] + h->pinned_bytes[k]) return true;
  return false;
}

// a field behind one of the cached critical time steps is about to change from outside the kernels that reduce them (ufm_k_cfl)
static void cfl_invalidate(ufm_handle *h, int field)
{
  if (field == UFM_F_D_SIA_AC) h->cfl_ok[0] = false;
  if (field == UFM_F_U_SSA || field == UFM_F_V_SSA) h->cfl_ok[1] = false;
  if (field == UFM_F_U_3D || field == UFM_F_V_3D) h->cfl_ok[2] = false;
}

static int field_copy(ufm_handle *h, int field, void *host, int to_device)
{
  if (!h || !h->has_mesh) return ufm_set_error(-2, "no mesh resident");
  if (!host) return ufm_set_error(-2, "NULL host pointer");
  UFM_CUDA(cudaSetDevice(h->device));
  DevMesh &m = h->mesh;
  if ((field == UFM_F_A_FLOW_MEAN || field == UFM_F_A_FLOW_MEAN_AC) && !h->st.realistic_A) {
    // benchmark flow factor: a scalar on the device (ice_physical_properties, general_ice_model_data_module.f90:321-368)
    if (to_device) return ufm_set_error(-2, "A_flow_mean is an output");
    const int nn = field == UFM_F_A_FLOW_MEAN ? m.nV : m.nAc;
    for (int i = 0; i < nn; i++) ((double *)host)[i] = h->st.A_flow_const;
    return 0;
  }
  FieldRef r;
  int rc = field_ref(h, field, &r);
  if (rc) return rc;
  const int n = r.kind == K_AA ? m.nV : (r.kind == K_AC ? m.nAc : m.M);
  const int *r2d = r.kind == K_AA ? m.aa_ref2dev : (r.kind == K_AC ? m.ac_ref2dev : m.m_ref2dev);
  const size_t bytes = r.is3d ? (size_t)n * r.is3d * sizeof(double) : (size_t)n * (r.is_int ? sizeof(int) : sizeof(double));
  if (!to_device && field >= UFM_F_DU_DX_AAAC && field <= UFM_F_DV_DY_AAAC) { if ((rc = ufm_k_ssa_gradients(h))) return rc; }
  if (to_device) {
    if (r.bits) return ufm_set_error(-2, "mask fields are outputs");
    cfl_invalidate(h, field);
    const void *src = host;
    if (!is_pinned(h, host, bytes)) { memcpy(h->staging, host, bytes); src = h->staging; }
    UFM_CUDA(cudaMemcpyAsync(h->dev_staging, src, bytes, cudaMemcpyHostToDevice, h->stream));
    h->cnt.h2d_bytes += (double)bytes;
  }
  if (r.is3d) rc = ufm_perm_3d(h, n, r.is3d, m.nVp, r2d, r.d, (double *)h->dev_staging, to_device);
  else if (r.bits) rc = ufm_perm_mask(h, n, r2d, r.bits, r.barg, r.bmode, (int *)h->dev_staging);
  else if (r.is_int) rc = ufm_perm_int(h, n, r2d, r.i, (int *)h->dev_staging, to_device);
  else rc = ufm_perm_double(h, n, r2d, r.d, r.stride, r.comp, (double *)h->dev_staging, to_device);
  if (rc) return rc;
  if (!to_device) {
    const bool direct = is_pinned(h, host, bytes);
    UFM_CUDA(cudaMemcpyAsync(direct ? host : h->staging, h->dev_staging, bytes, cudaMemcpyDeviceToHost, h->stream));
    UFM_CUDA(cudaStreamSynchronize(h->stream));
    if (!direct) memcpy(host, h->staging, bytes);
    h->cnt.d2h_bytes += (double)bytes;
  } else {
    UFM_CUDA(cudaStreamSynchronize(h->stream));  // staging buffer is reused by the next call
  }
  return 0;
}

int ufm_host_register(ufm_handle *h, void *host, unsigned long long bytes)
{
  if (!h || !host || !bytes) return ufm_set_error(-2, "ufm_host_register: bad argument");
  if (is_pinned(h, host, (size_t)bytes)) return 0;
  if (h->n_pinned >= 64) return ufm_set_error(-2, "ufm_host_register: too many registered buffers");
  UFM_CUDA(cudaSetDevice(h->device));
  UFM_CUDA(cudaHostRegister(host, (size_t)bytes, cudaHostRegisterPortable));
  h->pinned_base[h->n_pinned] = host; h->pinned_bytes[h->n_pinned] = (size_t)bytes; h->n_pinned++;
  return 0;
}
int ufm_host_unregister(ufm_handle *h, void *host)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  for (int k = 0; k < h->n_pinned; k++)
    if (h->pinned_base[k] == host) {
      cudaHostUnregister(host);
      h->pinned_base[k] = h->pinned_base[h->n_pinned - 1]; h->pinned_bytes[k] = h->pinned_bytes[h->n_pinned - 1]; h->n_pinned--;
      return 0;
    }
  return ufm_set_error(-2, "ufm_host_unregister: pointer was not registered");
}
static int stash_slot(ufm_handle *h, int field, bool create)
{
  for (int k = 0; k < 4; k++) if (h->stash[k].field == field) return k;
  if (!create) return -1;
  for (int k = 0; k < 4; k++) if (h->stash[k].field < 0) { h->stash[k].field = field; return k; }
  return -1;
}
int ufm_remap_stash(ufm_handle *h, int field)
{
  if (!h || !h->has_mesh) return ufm_set_error(-2, "no mesh resident");
  UFM_CUDA(cudaSetDevice(h->device));
  FieldRef r;
  int rc = field_ref(h, field, &r);
  if (rc) return rc;
  if (r.kind != K_AA || r.is_int || r.is3d || r.stride != 1 || !r.d) return ufm_set_error(-2, "ufm_remap_stash: only fp64 fields on the Aa vertices can be remapped");
  const int slot = stash_slot(h, field, true);
  if (slot < 0) return ufm_set_error(-2, "ufm_remap_stash: at most 4 fields can be stashed");
  return ufm_k_remap_stash(h, slot, r.d);
}
int ufm_remap_apply(ufm_handle *h, int field, const ufm_remap_cons *map, int order)
{
  if (!h || !h->has_mesh) return ufm_set_error(-2, "no mesh resident");
  UFM_CUDA(cudaSetDevice(h->device));
  if (!map || !map->vli1 || !map->vli2 || !map->vi || !map->w0) return ufm_set_error(-2, "ufm_remap_apply: NULL pointer in the remapping arrays");
  if (order != 1 && order != 2) return ufm_set_error(-2, "ufm_remap_apply: order must be 1 or 2");
  if (order == 2 && (!map->w1x || !map->w1y)) return ufm_set_error(-2, "ufm_remap_apply: 2nd order needs w1x and w1y");
  if (map->nV_dst != h->mesh.nV) return ufm_set_error(-2, "ufm_remap_apply: map is for %d destination vertices, the resident mesh has %d", map->nV_dst, h->mesh.nV);
  const int slot = stash_slot(h, field, false);
  if (slot < 0 || !h->stash[slot].d) return ufm_set_error(-2, "ufm_remap_apply: field %d was not stashed on the old mesh", field);
  cfl_invalidate(h, field);
  for (int i = 0; i < map->n_tot; i++) if (map->vi[i] < 1 || map->vi[i] > h->stash[slot].n) return ufm_set_error(-2, "ufm_remap_apply: source vertex index out of range");
  FieldRef r;
  int rc = field_ref(h, field, &r);
  if (rc) return rc;
  rc = ufm_k_remap_apply(h, slot, map, order, r.d);
  if (rc) return rc;
  cudaFree(h->stash[slot].d); cudaFree(h->stash[slot].ddx); cudaFree(h->stash[slot].ddy);
  h->stash[slot] = ufm_handle::Stash();
  return 0;
}
int ufm_state_upload(ufm_handle *h, int field, const void *host) { return field_copy(h, field, (void *)host, 1); }
int ufm_state_download(ufm_handle *h, int field, void *host) { return field_copy(h, field, host, 0); }
int ufm_pow_mode(ufm_handle *h) { return h ? h->pow_exact : ufm_set_error(-2, "NULL handle"); }
int ufm_partition_owner_of(ufm_handle *h, unsigned char *owner_out)
{
  if (!h || !h->has_mesh || !owner_out) return ufm_set_error(-2, "ufm_partition_owner_of: no mesh resident");
  const int M = h->mesh.M;
  if (h->mesh.P <= 1 || !h->owner_ref) { memset(owner_out, 0, (size_t)M); return 0; }
  memcpy(owner_out, h->owner_ref->data(), (size_t)M);
  return h->mesh.part_step ? 1 : 0;
}
int ufm_resident_dims(ufm_handle *h, int dims[5])
{
  if (!h || !dims) return ufm_set_error(-2, "NULL argument");
  dims[0] = h->has_mesh ? 1 : 0;
  dims[1] = h->has_mesh ? h->mesh.nV : 0; dims[2] = h->has_mesh ? h->mesh.nAc : 0; dims[3] = h->has_mesh ? h->mesh.M : 0;
  dims[4] = h->P.nZ;
  return 0;
}
int ufm_field_resident(ufm_handle *h, int field)
{
  if (!h || !h->has_mesh) return ufm_set_error(-2, "no mesh resident");
  if (field < 0 || field >= UFM_F_COUNT) return ufm_set_error(-2, "unknown field id %d", field);
  if (field == UFM_F_A_FLOW_MEAN || field == UFM_F_A_FLOW_MEAN_AC) return 1;   // a scalar for the benchmarks, expanded on download
  FieldRef r;
  return field_ref(h, field, &r) == 0 ? 1 : 0;
}

#define NEED_MESH(h) do { if (!(h) || !(h)->has_mesh) return ufm_set_error(-2, "no mesh resident"); UFM_CUDA(cudaSetDevice((h)->device)); } while (0)

int ufm_thickness_update(ufm_handle *h, double dt)
{
  NEED_MESH(h);
  const int b = h->P.benchmark;
  // the reference aborts for benchmark names it does not know here (ice_dynamics_module.f90:192-218)
  if (b == UFM_BM_MESH_GENERATION_TEST)
    return ufm_set_error(-1, "benchmark experiment \"mesh_generation_test\" not implemented in calculate_ice_thickness_change!");
  return ufm_k_thickness(h, dt);
}
int ufm_update_general(ufm_handle *h, double time) { NEED_MESH(h); return ufm_k_geom(h, time); }
int ufm_solve_SIA(ufm_handle *h) { NEED_MESH(h); return ufm_k_sia(h); }
int ufm_solve_SIA_3D(ufm_handle *h) { NEED_MESH(h); return ufm_k_sia3d(h); }
int ufm_cfl(ufm_handle *h, double out3[3]) { NEED_MESH(h); return ufm_k_cfl(h, out3); }
int ufm_ssa_prepare(ufm_handle *h) { NEED_MESH(h); return ufm_k_ssa_prepare(h); }
int ufm_ssa_viscosity(ufm_handle *h, double sums2[2]) { NEED_MESH(h); return ufm_k_ssa_viscosity(h, sums2); }
int ufm_ssa_sliding_and_setup(ufm_handle *h) { NEED_MESH(h); return ufm_k_ssa_sliding_setup(h); }
int ufm_ssa_finish(ufm_handle *h) { NEED_MESH(h); return ufm_k_ssa_finish(h); }
int ufm_ssa_sor(ufm_handle *h, int max_inner_override, int force_iters, ufm_ssa_stats *stats)
{
  NEED_MESH(h);
  ufm_ssa_stats tmp;
  if (!stats) stats = &tmp;
  memset(stats, 0, sizeof(*stats));
  int rc = ufm_k_ssa_sor(h, max_inner_override > 0 ? max_inner_override : h->P.SSA_max_inner_loops, force_iters, stats);
  stats->n_inner_total = stats->n_inner_last;
  return rc;
}

/* solve_SSA, src/ice_dynamics_module.f90:408-557: the outer (viscosity) loop stays on the host,
 * one scalar round trip (RN) per outer iteration; each linear solve is one persistent kernel. */
int ufm_solve_SSA(ufm_handle *h, ufm_ssa_stats *stats)
{
  NEED_MESH(h);
  ufm_ssa_stats tmp;
  if (!stats) stats = &tmp;
  memset(stats, 0, sizeof(*stats));
  const int b = h->P.benchmark;
  bool set_zero = false;
  if (b != UFM_BM_NONE) {
    if ((b >= UFM_BM_EISMINT_1 && b <= UFM_BM_EISMINT_6) || b == UFM_BM_HALFAR || b == UFM_BM_BUELER) set_zero = true;
    else if (b == UFM_BM_MISMIP_MOD || b == UFM_BM_MESH_GENERATION_TEST || b == UFM_BM_SSA_ICESTREAM) { /* SSA is solved */ }
    else { stats->rc = -2; return ufm_set_error(-2, "benchmark experiment %d not implemented in solve_SSA!", b); }
  }
  if (!set_zero) {
    long long n_sheet = 0;
    int rc = ufm_k_sum_mask_sheet(h, &n_sheet);
    if (rc) return rc;
    if (n_sheet == 0) set_zero = true;
  }
  if (set_zero) return ufm_k_ssa_zero(h);
  int rc = ufm_k_ssa_prepare(h);
  if (rc) return rc;
  if ((rc = ufm_k_ssa_outer_loop(h, stats))) { ufm_k_ssa_finish(h); return rc; }
  if (stats->rc == -1) {
    ufm_k_ssa_finish(h);
    return ufm_set_error(-1, "solve_SSA - ERROR: SSA remains unstable after resetting velocities to zero!");
  }
  if ((rc = ufm_k_ssa_finish(h))) return rc;
  if (stats->rc == 1) { ufm_set_error(1, " WARNING - SSA SOR solver doesnt converge!"); return 1; }
  return 0;
}

/* ---- region loop: run_model + determine_timesteps_and_actions for benchmark physics ---- */
int ufm_region_init(ufm_region *r, double start_time)
{
  if (!r) return ufm_set_error(-2, "NULL region");
  memset(r, 0, sizeof(*r));
  r->time = start_time;
  for (int k = 0; k < UFM_NT; k++) { r->t0[k] = start_time; r->t1[k] = start_time; r->do_[k] = 1; }
  r->dtc[UFM_T_THERMO] = 10.0; r->dtc[UFM_T_CLIMATE] = 10.0; r->dtc[UFM_T_SMB] = 10.0; r->dtc[UFM_T_BMB] = 10.0;
  r->dtc[UFM_T_ELRA] = 100.0; r->dtc[UFM_T_OUTPUT] = 5000.0;
  r->t1[UFM_T_THERMO] = start_time + r->dtc[UFM_T_THERMO];
  r->do_[UFM_T_THERMO] = 0;
  r->dt = 0.0; r->dt_prev = 1000.0;
  r->H0 = 5000.0; r->R0 = 300000.0; r->lambda = 5.0;
  return 0;
}

static int run_model_impl(ufm_handle *h, ufm_region *r, double t_end, long max_steps, const ufm_host_ice *host);
int ufm_run_model(ufm_handle *h, ufm_region *r, double t_end, long max_steps) { return run_model_impl(h, r, t_end, max_steps, nullptr); }
int ufm_run_model_host(ufm_handle *h, ufm_region *r, double t_end, long max_steps, const ufm_host_ice *host)
{
  if (!host) return ufm_set_error(-2, "NULL host state");
  return run_model_impl(h, r, t_end, max_steps, host);
}
}  // extern "C"

// ---- overlapped field transfers of the drop-in loop ------------------------------------------------------------------
// One slot per field and direction.  Upload: H2D on the copy stream -> event -> permutation kernel on the compute stream.
// Download: permutation kernel on the compute stream (right after the field's producer) -> event -> D2H on the copy stream,
// which therefore runs underneath whatever the compute stream does next (the SSA solve).  The host synchronises once per
// step (xfer_finish).  Only Aa fields (nV doubles / ints) travel this way.
static int xfer_prepare(ufm_handle *h, bool need_host_slots)
{
  const size_t slot = (((size_t)h->mesh.nV * sizeof(double)) + 255) & ~(size_t)255;
  if (!h->xfer_stream) {
    UFM_CUDA(cudaStreamCreateWithFlags(&h->xfer_stream, cudaStreamNonBlocking));
    UFM_CUDA(cudaStreamCreateWithFlags(&h->xfer_stream_out, cudaStreamNonBlocking));
    for (auto &e : h->xfer_ev) UFM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (h->xfer_slot_bytes < slot) {
    UFM_CUDA(cudaStreamSynchronize(h->xfer_stream));
    UFM_CUDA(cudaStreamSynchronize(h->xfer_stream_out));
    if (h->xfer_dev) cudaFree(h->xfer_dev);
    if (h->xfer_host) cudaFreeHost(h->xfer_host);
    h->xfer_dev = h->xfer_host = nullptr; h->xfer_slot_bytes = 0;
    UFM_CUDA(cudaMalloc((void **)&h->xfer_dev, slot * UFM_XFER_SLOTS));
    h->xfer_slot_bytes = slot;
  }
  if (need_host_slots && !h->xfer_host) UFM_CUDA(cudaMallocHost((void **)&h->xfer_host, h->xfer_slot_bytes * UFM_XFER_SLOTS));
  h->xfer_n_pending = 0;
  return 0;
}
// phase: 3 = copy and permutation (default); 1 = only enqueue the host -> device copy; 2 = only the permutation kernel that waits for it
// (so that a step can put all its copies on the wire first and let the compute stream pick the fields up in the order it needs them)
static int xfer_begin(ufm_handle *h, int field, void *host, int to_device, int slot, int phase = 3)
{
  FieldRef r;
  int rc = field_ref(h, field, &r);
  if (rc) return rc;
  DevMesh &m = h->mesh;
  if (r.kind != K_AA || r.is3d || slot < 0 || slot >= UFM_XFER_SLOTS) return ufm_set_error(-2, "xfer_begin: field %d does not travel through the overlapped path", field);
  const int n = m.nV;
  const size_t bytes = (size_t)n * (r.is_int ? sizeof(int) : sizeof(double));
  char *dslot = h->xfer_dev + (size_t)slot * h->xfer_slot_bytes;
  const bool direct = is_pinned(h, host, bytes);
  char *hslot = direct ? (char *)host : h->xfer_host + (size_t)slot * h->xfer_slot_bytes;
  cudaEvent_t ev = h->xfer_ev[slot];
  if (to_device) {
    if (r.bits) return ufm_set_error(-2, "mask fields are outputs");
    if (phase & 1) {
      if (!direct) memcpy(hslot, host, bytes);
      UFM_CUDA(cudaMemcpyAsync(dslot, hslot, bytes, cudaMemcpyHostToDevice, h->xfer_stream));
      UFM_CUDA(cudaEventRecord(ev, h->xfer_stream));
      h->cnt.h2d_bytes += (double)bytes;
    }
    if (phase & 2) {
      UFM_CUDA(cudaStreamWaitEvent(h->stream, ev, 0));
      rc = r.is_int ? ufm_perm_int(h, n, m.aa_ref2dev, r.i, (int *)dslot, 1) : ufm_perm_double(h, n, m.aa_ref2dev, r.d, r.stride, r.comp, (double *)dslot, 1);
    }
    return rc;
  }
  if (r.bits) rc = ufm_perm_mask(h, n, m.aa_ref2dev, r.bits, r.barg, r.bmode, (int *)dslot);
  else if (r.is_int) rc = ufm_perm_int(h, n, m.aa_ref2dev, r.i, (int *)dslot, 0);
  else rc = ufm_perm_double(h, n, m.aa_ref2dev, r.d, r.stride, r.comp, (double *)dslot, 0);
  if (rc) return rc;
  UFM_CUDA(cudaEventRecord(ev, h->stream));
  UFM_CUDA(cudaStreamWaitEvent(h->xfer_stream_out, ev, 0));
  UFM_CUDA(cudaMemcpyAsync(hslot, dslot, bytes, cudaMemcpyDeviceToHost, h->xfer_stream_out));
  if (!direct) h->xfer_pending[h->xfer_n_pending++] = {host, hslot, bytes};
  h->cnt.d2h_bytes += (double)bytes;
  return 0;
}
// end of a step: every copy has landed; results that went through a pinned slot are handed to the caller's pageable arrays
static int xfer_finish(ufm_handle *h)
{
  UFM_CUDA(cudaStreamSynchronize(h->xfer_stream));
  UFM_CUDA(cudaStreamSynchronize(h->xfer_stream_out));
  for (int k = 0; k < h->xfer_n_pending; k++) memcpy(h->xfer_pending[k].dst, h->xfer_pending[k].src, h->xfer_pending[k].bytes);
  h->xfer_n_pending = 0;
  return 0;
}

// ---- device-driven region loop ---------------------------------------------------------------------------------------------
// For the benchmark experiments whose step needs nothing from the host (no SSA solve: solve_SSA sets the velocities to zero; a
// time-independent closed-form mass balance; no thermodynamics on the mesh): determine_timesteps_and_actions
// (src/UFEMISM_main_model.f90:708-843) runs in a one-thread kernel at the end of every step, the step's kernels read the time step
// and their "due" flags from device memory, and the host enqueues UFM_DEVICE_BATCH steps between synchronisations instead of one.
// Same arithmetic, same order: identical trajectories (tests/test_gpu_parity.py::test_device_loop_matches_host_loop).
struct StepCtl {
  ufm_region r;
  double t_end, dt_max;
  long long steps_left;
  int g_run, g_sia, g_smb, g_thermo;   // gates of the next step's kernels
  int sia3d_on_thermo_timer;           // EISMINT experiments (update_ice_temperature's velocity half feeds the critical time step)
};
__global__ void k_step_control(StepCtl *c, unsigned long long *keys)
{
  if (!c->g_run) return;
  ufm_region &r = c->r;
  // what the step that has just run did (run_model, :78-214)
  r.t0[UFM_T_ELRA] = r.time;
  if (r.do_[UFM_T_SIA]) { r.t0[UFM_T_SIA] = r.time; r.n_sia++; }
  if (r.do_[UFM_T_SSA]) { r.t0[UFM_T_SSA] = r.time; r.n_ssa++; }
  if (r.do_[UFM_T_CLIMATE]) r.t0[UFM_T_CLIMATE] = r.time;
  if (r.do_[UFM_T_SMB]) r.t0[UFM_T_SMB] = r.time;
  if (r.do_[UFM_T_BMB]) r.t0[UFM_T_BMB] = r.time;
  if (r.do_[UFM_T_THERMO]) r.t0[UFM_T_THERMO] = r.time;
  if (r.do_[UFM_T_OUTPUT]) r.t0[UFM_T_OUTPUT] = r.time;
  // determine_timesteps_and_actions
  const double dt_correction_factor = 0.9;
  const double dt_D_2D_min = ord_unkey(keys[0]) * dt_correction_factor, dt_V_2D_SSA_min = ord_unkey(keys[1]) * dt_correction_factor,
               dt_V_3D_SIA_min = ord_unkey(keys[2]) * dt_correction_factor;
  r.dt_crit_last[0] = dt_D_2D_min; r.dt_crit_last[1] = dt_V_2D_SSA_min; r.dt_crit_last[2] = dt_V_3D_SIA_min;
  r.dt = fmin(fmin(fmin(dt_D_2D_min, dt_V_2D_SSA_min), dt_V_3D_SIA_min), c->dt_max);
  if (fabs(1.0 - r.dt / r.dt_prev) > 0.1) r.dt_prev = r.dt;
  r.dtc[UFM_T_SIA] = fmin(c->dt_max, fmin(dt_D_2D_min, dt_V_3D_SIA_min));
  r.dtc[UFM_T_SSA] = fmin(c->dt_max, dt_V_2D_SSA_min);
  double t_next_action = 0.0;
  for (int k = 0; k < UFM_NT; k++) { r.t1[k] = r.t0[k] + r.dtc[k]; if (k == 0 || r.t1[k] < t_next_action) t_next_action = r.t1[k]; }
  r.dt = t_next_action - r.time;
  for (int k = 0; k < UFM_NT; k++) r.do_[k] = (t_next_action == r.t1[k]);
  if (t_next_action >= c->t_end) {
    r.dt = c->t_end - r.time;
    r.do_[UFM_T_SIA] = r.do_[UFM_T_SSA] = r.do_[UFM_T_THERMO] = r.do_[UFM_T_CLIMATE] = r.do_[UFM_T_SMB] = r.do_[UFM_T_BMB] = 1;
  }
  r.time = r.time + r.dt;
  r.n_steps++;
  c->steps_left--;
  // gates of the next step and the minima its producers will reduce again
  c->g_run = (r.time < c->t_end && c->steps_left != 0) ? 1 : 0;
  c->g_sia = (c->g_run && r.do_[UFM_T_SIA]) ? 1 : 0;
  c->g_smb = (c->g_run && r.do_[UFM_T_SMB]) ? 1 : 0;
  c->g_thermo = (c->g_run && r.do_[UFM_T_THERMO] && c->sia3d_on_thermo_timer) ? 1 : 0;
  if (c->g_sia) keys[0] = ord_key(1000.0);
  if (c->g_thermo) keys[2] = ord_key(1000.0);
}

static bool device_loop_applies(const ufm_handle *h, const ufm_host_ice *host)
{
  const int b = h->P.benchmark;
  const char *e = getenv("UFM_DEVICE_LOOP");
  if (e && atoi(e) == 0) return false;
  // no SSA solve, a mass balance that does not depend on time (the time-dependent ones go through the host's libm: sin, pow), no
  // column thermodynamics on the device, one GPU
  return !host && (b == UFM_BM_EISMINT_1 || b == UFM_BM_EISMINT_4 || b == UFM_BM_HALFAR) && !h->mesh.has_tri && !h->st.realistic_A && h->mesh.P == 1;
}

#define UFM_DEVICE_BATCH 64
static int run_model_device(ufm_handle *h, ufm_region *r, double t_end, long max_steps)
{
  int rc;
  const int b = h->P.benchmark;
  if (!h->stepctl_dev) {
    UFM_CUDA(cudaMalloc(&h->stepctl_dev, sizeof(StepCtl)));
    UFM_CUDA(cudaMallocHost(&h->stepctl_host, sizeof(StepCtl)));
  }
  StepCtl *cd = (StepCtl *)h->stepctl_dev, *ch = (StepCtl *)h->stepctl_host;
  if (!(r->time < t_end) || max_steps < 0) return 0;
  // a clean start: zero SSA velocities (what solve_SSA does for these experiments whenever it is due) and every cached critical
  // time step valid for the fields as they are now
  if ((rc = ufm_k_ssa_zero(h))) return rc;
  { double d3[3]; if ((rc = ufm_k_cfl(h, d3))) return rc; }
  const bool eismint = b >= UFM_BM_EISMINT_1 && b <= UFM_BM_EISMINT_6;
  memset(ch, 0, sizeof(*ch));
  ch->r = *r; ch->t_end = t_end; ch->dt_max = h->P.dt_max; ch->steps_left = max_steps > 0 ? max_steps : -1;
  ch->sia3d_on_thermo_timer = eismint ? 1 : 0;
  ch->g_run = 1; ch->g_sia = r->do_[UFM_T_SIA] ? 1 : 0; ch->g_smb = r->do_[UFM_T_SMB] ? 1 : 0;
  ch->g_thermo = (r->do_[UFM_T_THERMO] && eismint) ? 1 : 0;
  UFM_CUDA(cudaMemcpyAsync(cd, ch, sizeof(StepCtl), cudaMemcpyHostToDevice, h->stream));
  if ((rc = ufm_cfl_key_reset(h, (ch->g_sia ? 1 : 0) | (ch->g_thermo ? 4 : 0)))) return rc;
  h->gate[0] = &cd->g_run; h->gate[1] = &cd->g_sia; h->gate[2] = &cd->g_smb; h->gate[3] = &cd->g_thermo;
  h->dt_dev = &cd->r.dt;
  long enq = 0;
  const long steps0 = r->n_steps;
  bool done = false;
  rc = 0;
  auto enqueue_step = [&]() -> int {
    int rc_;
    if ((rc_ = ufm_k_thickness(h, 0.0))) return rc_;
    if ((rc_ = ufm_k_geom(h, 0.0))) return rc_;
    if ((rc_ = ufm_k_sia(h))) return rc_;
    if ((rc_ = ufm_k_smb_benchmark(h, 0.0, r->H0, r->R0, r->lambda))) return rc_;
    if (eismint) { if ((rc_ = ufm_k_sia3d(h))) return rc_; if ((rc_ = ufm_k_cfl3d_enqueue(h))) return rc_; }
    k_step_control<<<1, 1, 0, h->stream>>>(cd, h->st.ctrl + CTRL_CFL_KEYS);
    h->cnt.kernel_launches++;
    return ufm_cuda_check(cudaGetLastError(), "k_step_control");
  };
  // Two consecutive steps as ONE CUDA graph (two, because the thickness update swaps the Hi / Hi_prev buffers: after a pair the
  // pointers are back where they were and the same graph can be launched again): a step of a 10 k-vertex mesh is a dozen kernels of a
  // few microseconds each, so launching them one by one from the host costs more than running them.  UFM_DEVICE_GRAPH=0 or a stream
  // that cannot be captured: plain launches.
  cudaGraphExec_t exec = nullptr;
  {
    const char *e = getenv("UFM_DEVICE_GRAPH");
    const long long before = h->cnt.kernel_launches;
    if ((!e || atoi(e) != 0) && cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int rc1 = enqueue_step();
      if (!rc1) rc1 = enqueue_step();
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      if (rc1 || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
      if (graph) cudaGraphDestroy(graph);
    }
    cudaGetLastError();
    const long long per_pair = h->cnt.kernel_launches - before;   // kernels in one graph launch
    h->cnt.kernel_launches = before;
    while (!done && !rc) {
      const long nb = (max_steps > 0 && max_steps - enq < UFM_DEVICE_BATCH) ? max_steps - enq : UFM_DEVICE_BATCH;
      if (exec) {
        for (long k = 0; k < nb && !rc; k += 2) { rc = ufm_cuda_check(cudaGraphLaunch(exec, h->stream), "device loop: graph launch"); enq += 2; h->cnt.kernel_launches += per_pair; }
      } else {
        for (long k = 0; k < nb && !rc; k++) { rc = enqueue_step(); enq++; }
      }
      if (rc) break;
      if ((rc = ufm_cuda_check(cudaMemcpyAsync(ch, cd, sizeof(StepCtl), cudaMemcpyDeviceToHost, h->stream), "device loop: read back"))) break;
      if ((rc = ufm_cuda_check(cudaStreamSynchronize(h->stream), "device loop: synchronise"))) break;
      done = !ch->g_run || (max_steps > 0 && enq >= max_steps);
    }
    if (exec) { cudaGraphExecDestroy(exec); enq = 0; }   // the host's Hi / Hi_prev pointers saw an even number of swaps (the captured pair)
  }
  h->gate[0] = h->gate[1] = h->gate[2] = h->gate[3] = nullptr;
  h->dt_dev = nullptr;
  if (rc) return rc;
  // ufm_k_thickness swaps Hi / Hi_prev on the host at every enqueued step; steps enqueued after the end did not run
  if ((enq - (ch->r.n_steps - steps0)) & 1) { double *t = h->st.Hi; h->st.Hi = h->st.Hi_alt; h->st.Hi_alt = t; }
  *r = ch->r;
  h->cfl_ok[0] = h->cfl_ok[1] = h->cfl_ok[2] = true;   // reduced by the last step that changed the fields behind them
  return 0;
}

static int run_model_impl(ufm_handle *h, ufm_region *r, double t_end, long max_steps, const ufm_host_ice *host)
{
  NEED_MESH(h);
  if (!r) return ufm_set_error(-2, "NULL region");
  const int b = h->P.benchmark;
  if (!host && b == UFM_BM_NONE)
    return ufm_set_error(-4, "ufm_run_model: without a benchmark experiment SMB and BMB come from the host's climate / SMB / BMB models each dt_SMB; use ufm_run_model_host or the step-wise entry points");
  long steps = 0;
  int rc;
  // run_ELRA_model knows every benchmark experiment but this one (bedrock_ELRA_module.f90:35-52)
  if (b == UFM_BM_MESH_GENERATION_TEST)
    return ufm_set_error(-1, "benchmark experiment \"mesh_generation_test\" not implemented in run_ELRA_model!");
  // drop-in mode moves the host's fields at the top of every step and the host's own components (ELRA bedrock update, climate / SMB /
  // BMB on their timers) run between steps: one step per call
  if (host && max_steps != 1) return ufm_set_error(-2, "ufm_run_model_host: max_steps must be 1 (the host's components run between the steps)");
  // the thermodynamics timer runs on C%dt_thermo (UFEMISM_main_model.f90:369,797), the step the heat equation integrates over
  if (h->P.dt_thermo > 0.0 && r->dtc[UFM_T_THERMO] != h->P.dt_thermo) {
    if (r->n_steps != 0) return ufm_set_error(-2, "ufm_run_model: region timer dt_thermo = %g but the handle's parameters say %g", r->dtc[UFM_T_THERMO], h->P.dt_thermo);
    r->dtc[UFM_T_THERMO] = h->P.dt_thermo;
    r->t1[UFM_T_THERMO] = r->t0[UFM_T_THERMO] + h->P.dt_thermo;
  }
  if (device_loop_applies(h, host)) return run_model_device(h, r, t_end, max_steps);
  // drop-in mode: UFM_XFER_OVERLAP=0 falls back to one synchronous copy per field (A/B measurements)
  bool overlap = false;
  if (host) {
    const char *e = getenv("UFM_XFER_OVERLAP");
    overlap = !e || atoi(e) != 0;
    if (overlap) {
      const size_t nb = (size_t)h->mesh.nV * sizeof(double), ni = (size_t)h->mesh.nV * sizeof(int);
      const struct { const void *p; size_t b; } all[] = {{host->Hi, nb}, {host->Hb, nb}, {host->SL, nb}, {host->dHb_dt, nb}, {host->SMB_year, nb}, {host->BMB, nb},
          {host->mask_noice, ni}, {host->Hi_out, nb}, {host->Hi_prev, nb}, {host->dHi_dt, nb}, {host->Hs, nb}, {host->U_SSA, nb}, {host->V_SSA, nb},
          {host->U_SIA, nb}, {host->V_SIA, nb}, {host->D_SIA, nb}, {host->mask, ni}};
      bool need_host_slots = false;
      for (auto &q : all) if (q.p && !is_pinned(h, q.p, q.b)) need_host_slots = true;
      if ((rc = xfer_prepare(h, need_host_slots))) return rc;
    }
  }
  // what leaves the device, and after which stage of the step it is final
  enum { AFTER_THK = 0, AFTER_GENERAL, AFTER_SIA, AFTER_SSA };
  struct Out { int f; void *p; int stage; };
  const Out outs[] = {{UFM_F_HI, host ? host->Hi_out : nullptr, AFTER_THK}, {UFM_F_HI_PREV, host ? host->Hi_prev : nullptr, AFTER_THK},
                      {UFM_F_DHI_DT, host ? host->dHi_dt : nullptr, AFTER_THK}, {UFM_F_HS, host ? host->Hs : nullptr, AFTER_GENERAL},
                      {UFM_F_U_SSA, host ? host->U_SSA : nullptr, AFTER_SSA}, {UFM_F_V_SSA, host ? host->V_SSA : nullptr, AFTER_SSA},
                      {UFM_F_U_SIA, host ? host->U_SIA : nullptr, AFTER_SIA}, {UFM_F_V_SIA, host ? host->V_SIA : nullptr, AFTER_SIA},
                      {UFM_F_D_SIA, host ? host->D_SIA : nullptr, AFTER_SIA}, {UFM_F_MASK, host ? host->mask : nullptr, AFTER_GENERAL}};
  const int n_outs = (int)(sizeof(outs) / sizeof(outs[0]));
  auto start_downloads = [&](int stage) -> int {
    if (!overlap) return 0;
    for (int k = 0; k < n_outs; k++)
      if (outs[k].p && outs[k].stage == stage) { int rc_ = xfer_begin(h, outs[k].f, outs[k].p, 0, 8 + k); if (rc_) return rc_; }
    return 0;
  };
  while (r->time < t_end && (max_steps <= 0 || steps < max_steps)) {
    // run_ELRA_model (bedrock_ELRA_module.f90:22-66): the benchmark branch only moves the timer on; a realistic run does so when the
    // deformation rate is due (the host computes it and updates Hb between the steps, then uploads Hb / dHb_dt with the step's inputs)
    if (b != UFM_BM_NONE || r->do_[UFM_T_ELRA]) r->t0[UFM_T_ELRA] = r->time;
    // Drop-in mode, order of the step's transfers: every host -> device copy goes on the wire at once, the inputs of the thickness update
    // first (Hi, SMB, BMB, mask_noice); the compute stream picks those up, updates the thickness and hands its results to the
    // device -> host stream while the geometry's inputs (Hb, SL, dHb_dt) are still arriving in the other direction.  Outputs whose
    // producer does not run in this step (velocities between two solves) are unchanged since the last step and leave first.
    const struct { int f; const void *p; int late; } in[] = {{UFM_F_HI, host ? host->Hi : nullptr, 0}, {UFM_F_SMB_YEAR, host ? host->SMB_year : nullptr, 0},
                                                             {UFM_F_BMB, host ? host->BMB : nullptr, 0}, {UFM_F_MASK_NOICE, host ? host->mask_noice : nullptr, 0},
                                                             {UFM_F_HB, host ? host->Hb : nullptr, 1}, {UFM_F_SL, host ? host->SL : nullptr, 1},
                                                             {UFM_F_DHB_DT, host ? host->dHb_dt : nullptr, 1}};
    const int n_in = (int)(sizeof(in) / sizeof(in[0]));
    auto pick_up = [&](int late) -> int {
      for (int k = 0; k < n_in; k++)
        if (in[k].p && in[k].late == late) { int rc_ = xfer_begin(h, in[k].f, (void *)in[k].p, 1, k, 2); if (rc_) return rc_; }
      return 0;
    };
    if (host && overlap) {
      if (!r->do_[UFM_T_SSA] && (rc = start_downloads(AFTER_SSA))) return rc;
      if (!r->do_[UFM_T_SIA] && (rc = start_downloads(AFTER_SIA))) return rc;
      for (int k = 0; k < n_in; k++) if (in[k].p && (rc = xfer_begin(h, in[k].f, (void *)in[k].p, 1, k, 1))) return rc;
      if ((rc = pick_up(0))) return rc;
    } else if (host) {
      for (int k = 0; k < n_in; k++) if (in[k].p && (rc = ufm_state_upload(h, in[k].f, in[k].p))) return rc;
    }
    if ((rc = ufm_thickness_update(h, r->dt))) return rc;
    if ((rc = start_downloads(AFTER_THK))) return rc;
    if (host && overlap && (rc = pick_up(1))) return rc;
    if ((rc = ufm_update_general(h, r->time))) return rc;
    if ((rc = start_downloads(AFTER_GENERAL))) return rc;
    if (r->do_[UFM_T_SIA]) {
      if ((rc = ufm_solve_SIA(h))) return rc;
      r->t0[UFM_T_SIA] = r->time; r->n_sia++;
      if ((rc = start_downloads(AFTER_SIA))) return rc;
    }
    if (r->do_[UFM_T_SSA]) {
      ufm_ssa_stats st;
      rc = ufm_solve_SSA(h, &st);
      if (rc < 0) return rc;
      r->t0[UFM_T_SSA] = r->time; r->n_ssa++; r->n_sor_total += st.n_inner_total; r->n_outer_total += st.n_outer;
      if ((rc = start_downloads(AFTER_SSA))) return rc;
    }
    // climate / BMB: no-ops for the dynamics in the benchmark experiments (BMB = 0, src/BMB_module.f90:51-69)
    if (r->do_[UFM_T_CLIMATE]) r->t0[UFM_T_CLIMATE] = r->time;
    if (r->do_[UFM_T_SMB]) {   // run_SMB_model, benchmark branches: closed forms evaluated on the device (no host round trip)
      if (!host && (rc = ufm_k_smb_benchmark(h, r->time, r->H0, r->R0, r->lambda))) return rc;
      r->t0[UFM_T_SMB] = r->time;
    }
    if (r->do_[UFM_T_BMB]) r->t0[UFM_T_BMB] = r->time;
    if (r->do_[UFM_T_THERMO]) {
      // update_ice_temperature (thermodynamics_module.f90:23-202), EISMINT and realistic runs: the whole routine when the
      // mesh carries the triangle data its upwind advection reads, otherwise only the U_3D / V_3D refresh the time step needs
      if ((b >= UFM_BM_EISMINT_1 && b <= UFM_BM_EISMINT_6) || b == UFM_BM_NONE) {
        if (h->mesh.has_tri) { ufm_thermo_stats ts; if ((rc = ufm_update_ice_temperature(h, &ts))) return rc; }
        else if ((rc = ufm_solve_SIA_3D(h))) return rc;
      }
      r->t0[UFM_T_THERMO] = r->time;
    }
    if (r->do_[UFM_T_OUTPUT]) r->t0[UFM_T_OUTPUT] = r->time;
    double d3[3];
    if ((rc = ufm_cfl(h, d3))) return rc;
    const double dt_D_2D_min = d3[0], dt_V_2D_SSA_min = d3[1], dt_V_3D_SIA_min = d3[2];
    r->dt_crit_last[0] = d3[0]; r->dt_crit_last[1] = d3[1]; r->dt_crit_last[2] = d3[2];
    r->dt = fmin(fmin(fmin(dt_D_2D_min, dt_V_2D_SSA_min), dt_V_3D_SIA_min), h->P.dt_max);
    if (fabs(1.0 - r->dt / r->dt_prev) > 0.1) r->dt_prev = r->dt;
    r->dtc[UFM_T_SIA] = fmin(h->P.dt_max, fmin(dt_D_2D_min, dt_V_3D_SIA_min));
    r->dtc[UFM_T_SSA] = fmin(h->P.dt_max, dt_V_2D_SSA_min);
    double t_next_action = 0.0;
    for (int k = 0; k < UFM_NT; k++) { r->t1[k] = r->t0[k] + r->dtc[k]; if (k == 0 || r->t1[k] < t_next_action) t_next_action = r->t1[k]; }
    r->dt = t_next_action - r->time;
    for (int k = 0; k < UFM_NT; k++) r->do_[k] = (t_next_action == r->t1[k]);
    if (t_next_action >= t_end) {
      r->dt = t_end - r->time;
      r->do_[UFM_T_SIA] = r->do_[UFM_T_SSA] = r->do_[UFM_T_THERMO] = r->do_[UFM_T_CLIMATE] = r->do_[UFM_T_SMB] = r->do_[UFM_T_BMB] = 1;
    }
    r->time = r->time + r->dt;
    steps++; r->n_steps++;
    if (host && overlap) {
      if ((rc = xfer_finish(h))) return rc;
    } else if (host) {
      for (int k = 0; k < n_outs; k++) if (outs[k].p && (rc = ufm_state_download(h, outs[k].f, outs[k].p))) return rc;
    }
  }
  return 0;
}

extern "C" {
int ufm_counters_get(ufm_handle *h, ufm_counters *out)
{
  if (!h || !out) return ufm_set_error(-2, "NULL argument");
  *out = h->cnt;
  return 0;
}
int ufm_sor_trace_get(ufm_handle *h, unsigned long long *out, int n_words)
{
  if (!h || !out) return ufm_set_error(-2, "NULL argument");
  if (!h->sor_trace) return ufm_set_error(-2, "ufm_sor_trace_get: UFM_SOR_TRACE was not set when the SOR kernel was configured");
  if (n_words > 4096 * 24) n_words = 4096 * 24;
  UFM_CUDA(cudaStreamSynchronize(h->stream));
  UFM_CUDA(cudaMemcpy(out, h->sor_trace, (size_t)n_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return h->sor_grid;
}
int ufm_counters_reset(ufm_handle *h)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  double keep = h->cnt.sor_bytes_per_iteration;
  memset(&h->cnt, 0, sizeof(h->cnt));
  h->cnt.sor_bytes_per_iteration = keep;
  return 0;
}

}  // extern "C"
