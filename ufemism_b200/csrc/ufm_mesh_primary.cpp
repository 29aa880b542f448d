// ufm_mesh_primary.cpp -- device re-upload from PRIMARY mesh data (second half).
//
// After a mesh update (src/mesh_update_module.f90) or a restart (read_mesh_from_restart_file, src/restart_module.f90:31-116) the
// reference holds only the primary mesh data -- V, nC, C, niTri, iTri, edge_index, Tri -- and rebuilds everything else on the
// CPU (src/restart_module.f90:88-103, src/mesh_creation_module.f90:1724-1737): find_Voronoi_cell_areas, find_connection_widths,
// make_Ac_mesh, get_neighbour_functions, determine_mesh_resolution, make_combined_AaAc_mesh, calculate_five_colouring_AaAc.
// ufm_mesh_upload_primary takes the primary data as they are and does that work inside the library:
//   * Voronoi areas / connection widths, the staggered Ac mesh, the combined AaAc connectivity and the five-colouring on the
//     host (csrc/mesh_host.c; the colouring is sequential by construction and its order is part of the parity contract:
//     linked-list queues instead of the reference's array shifts, same result, linear time);
//   * every neighbour function (Aa, Ac, AaAc: 5 x 17 doubles per AaAc vertex, the bulk of the data) on the device
//     (k_derive_nf_Aa / _Ac / _AaAc in ufm_upload.cu), never materialised on the host;
// and keeps the derived host arrays so that the Fortran host can copy the few it still reads itself (A for ice volumes, Aci /
// iAci for output, the colour lists) with ufm_mesh_secondary_get instead of recomputing them.
#include <algorithm>
#include <chrono>
#include <parallel/algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "ufm_internal.cuh"

extern "C" {
int ufm_mesh_geometry(int nV, int nTri, int nC_mem, const double *V, const int *Tri, const int *nC, const int *C, const int *niTri, const int *iTri,
                      const int *edge_index, double xmin, double xmax, double ymin, double ymax, double *Tricc, int *Tri_edge_index, double *A, double *Cw);
int ufm_mesh_make_Ac(int nV, int nTri, int nC_mem, int nAc_max, const double *V, const int *Tri, const int *nC, const int *C, const int *niTri,
                     const int *iTri, const int *edge_index, int *iAci, int *Aci, double *VAc, double *Nx_Ac, double *Ny_Ac, double *Np_Ac,
                     double *No_Ac, int *edge_index_Ac);
int ufm_mesh_make_AaAc(int nV, int nAc, int ldAc, int nC_mem, const double *V, const double *VAc, const int *nC, const int *C, const int *iAci,
                       const int *Aci, const int *edge_index_Ac, double *VAaAc, int *nCAaAc, int *CAaAc);
int ufm_mesh_five_colouring_labelled(int M, int nC_mem, const int *nCAaAc, const int *CAaAc, const int *label, int *colour, int *colour_vi,
                                     int *colour_nV);
size_t ufm_mesh_five_colouring_ws_bytes(int M, int nC_mem);
int ufm_mesh_five_colouring_ws(int M, int nC_mem, const int *nCAaAc, const int *CAaAc, const int *label, int *colour, int *colour_vi,
                               int *colour_nV, void *ws);
}

struct ufm_secondary {
  int nV = 0, nTri = 0, nAc = 0, ldAc = 0, W = 0;
  std::vector<double> V, A, Cw, Tricc, VAc, VAaAc, R, NxTri, NyTri;
  std::vector<int> nC, C, niTri, iTri, edge_index, Tri, Tri_edge_index, iAci, Aci, edge_index_Ac, nCAaAc, CAaAc, colour, colour_vi, colour_nV;
  ufm_mesh_desc desc;
  // the five-colouring may still be running on its own thread (ufm_mesh_upload_primary overlaps it with the colour-independent half of the
  // upload): colour, colour_vi and colour_nV are final once colour_join() has returned
  std::thread colour_thread;
  std::vector<int> label;
  std::vector<unsigned long long> sort_keys;   // scratch of the Morton relabelling, kept for the next mesh
  void *colour_ws = nullptr;                   // scratch of the five-colouring (uninitialised memory, kept for the next mesh)
  size_t colour_ws_bytes = 0;
  int colour_rc = 0;
  int colour_join() { if (colour_thread.joinable()) colour_thread.join(); return colour_rc; }
  ~ufm_secondary() { colour_join(); free(colour_ws); }
};
extern thread_local std::function<int()> *g_ufm_colour_wait;   // ufm_upload.cu

void ufm_secondary_free(ufm_handle *h)
{
  delete (ufm_secondary *)h->secondary;
  h->secondary = nullptr;
}

namespace {
// Size a host array without touching more memory than needed.  The arrays of one derivation are hundreds of megabytes; a std::vector
// that is created (or assigned) anew value-initialises every element on one thread and page-faults its way through fresh memory --
// measured as more than half of the time of the phases below.  ufm_mesh_upload_primary therefore hands the previous mesh's object back
// (`recycle`): with some head room in the capacity a resize to a similar size touches nothing, and the routines that fill an array
// completely (or fill_par below) run on all cores.
template <class T>
void grow(std::vector<T> &v, size_t n)
{
  if (v.capacity() < n) { std::vector<T>().swap(v); v.reserve(n + n / 16 + 64); }
  v.resize(n);
}
template <class T>
void fill_par(std::vector<T> &v, size_t n, T value)
{
  grow(v, n);
  T *q = v.data();
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; i++) q[i] = value;
}
// copy an (n, cols) column-major array with leading dimension ld into a dense one
template <class T>
void compact(std::vector<T> &dst, const T *src, int n, int cols, int ld)
{
  grow(dst, (size_t)n * cols);
  T *q = dst.data();
#pragma omp parallel for schedule(static)
  for (int c = 0; c < cols; c++) memcpy(q + (size_t)c * n, src + (size_t)c * ld, sizeof(T) * (size_t)n);
}
template <class T>
void copy_par(std::vector<T> &dst, const T *src, size_t n)
{
  grow(dst, n);
  T *q = dst.data();
  const int parts = 16;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < parts; k++) { const size_t lo = n * k / parts, hi = n * (k + 1) / parts; memcpy(q + lo, src + lo, sizeof(T) * (hi - lo)); }
}
}  // namespace

// host-only half: everything ufm_mesh_upload_primary derives before it touches the device.  async_colour: return while the five-colouring
// (sequential by construction, a third of the whole re-upload) is still running on its own thread
static int derive_secondary_impl(const ufm_mesh_primary *p, void **derived, bool async_colour, ufm_secondary *recycle);
extern "C" int ufm_mesh_derive_secondary(const ufm_mesh_primary *p, void **derived)
{
  int rc = derive_secondary_impl(p, derived, false, nullptr);
  return rc;
}
// host-only, like ufm_mesh_derive_secondary, but *derived may hold the object of an earlier call, whose buffers are then reused (what
// ufm_mesh_upload_primary does with the previous mesh's object)
extern "C" int ufm_mesh_derive_secondary_reuse(const ufm_mesh_primary *p, void **derived)
{
  if (!derived) return ufm_set_error(-2, "ufm_mesh_derive_secondary_reuse: NULL argument");
  ufm_secondary *old = (ufm_secondary *)*derived;
  return derive_secondary_impl(p, derived, false, old);
}
// recycle: the object of the previous mesh (or NULL); its buffers are reused, its contents are gone when this returns, whatever it returns
static int derive_secondary_impl(const ufm_mesh_primary *p, void **derived, bool async_colour, ufm_secondary *recycle)
{
  struct Recycled { ufm_secondary *s; ~Recycled() { delete s; } } recycled{recycle};   // deleted on every early return below
  if (!derived) return ufm_set_error(-2, "ufm_mesh_derive_secondary: NULL output");
  *derived = nullptr;
  if (!p || !p->V || !p->nC || !p->C || !p->niTri || !p->iTri || !p->edge_index || !p->Tri)
    return ufm_set_error(-2, "ufm_mesh_upload_primary: NULL pointer in the primary mesh data");
  const int N = p->nV, T = p->nTri, W = p->nC_mem;
  if (N < 5 || T < 4 || W < 3 || W > 32) return ufm_set_error(-2, "ufm_mesh_upload_primary: implausible sizes nV=%d nTri=%d nC_mem=%d", N, T, W);
  if (!(p->xmax > p->xmin) || !(p->ymax > p->ymin)) return ufm_set_error(-2, "ufm_mesh_upload_primary: empty domain [%g,%g] x [%g,%g]", p->xmin, p->xmax, p->ymin, p->ymax);
  const int ldV = p->ldV ? p->ldV : N, ldT = p->ldTri ? p->ldTri : T;
  if (ldV < N || ldT < T) return ufm_set_error(-2, "ufm_mesh_upload_primary: leading dimension smaller than the mesh");
  const bool timing = getenv("UFM_UPLOAD_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[ufm_mesh_upload_primary] %-24s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  if (recycle) recycle->colour_join();
  ufm_secondary *s = recycle ? recycle : new ufm_secondary();
  recycled.s = nullptr;
  struct Guard { ufm_secondary *s; ~Guard() { delete s; } } guard{s};
  s->colour_rc = 0;
  s->R.clear(); s->NxTri.clear(); s->NyTri.clear();
  s->nV = N; s->nTri = T; s->W = W;
  compact(s->V, p->V, N, 2, ldV);
  compact(s->C, p->C, N, W, ldV);
  compact(s->iTri, p->iTri, N, W, ldV);
  compact(s->Tri, p->Tri, T, 3, ldT);
  copy_par(s->nC, p->nC, (size_t)N); copy_par(s->niTri, p->niTri, (size_t)N); copy_par(s->edge_index, p->edge_index, (size_t)N);
  // the host routines below index with what the arrays hold: validate first
  int bad_v = 0, bad_t = 0;
#pragma omp parallel for schedule(static)
  for (int v = 0; v < N; v++) {
    const int n = s->nC[v], nt = s->niTri[v], ei = s->edge_index[v];
    bool ok = !(n < 2 || n > W || nt < 1 || nt > W || ei < 0 || ei > 8);
    for (int c = 0; ok && c < n; c++) { const int q = s->C[(size_t)c * N + v]; ok = q >= 1 && q <= N && q != v + 1; }
    for (int c = 0; ok && c < nt; c++) { const int q = s->iTri[(size_t)c * N + v]; ok = q >= 1 && q <= T; }
    if (!ok) bad_v = v + 1;
  }
  if (bad_v) return ufm_set_error(-2, "ufm_mesh_upload_primary: vertex %d: nC=%d niTri=%d edge_index=%d, or an entry of its C / iTri row, out of range",
                                  bad_v, s->nC[bad_v - 1], s->niTri[bad_v - 1], s->edge_index[bad_v - 1]);
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; t++)
    for (int k = 0; k < 3; k++) { const int q = s->Tri[(size_t)k * T + t]; if (q < 1 || q > N) bad_t = t + 1; }
  if (bad_t) return ufm_set_error(-2, "ufm_mesh_upload_primary: Tri(%d,:) out of range", bad_t);
  lap("copy + validate");

  int rc = 0;
  // a planar triangulation of a simply connected domain has exactly nV + nTri - 1 edges (Euler)
  const int nAc_max = N + T;
  s->ldAc = nAc_max;
  // make_Ac_mesh writes every entry of iAci and rows 1..nAc of the others; their rows nAc+1..nAc_max are never read
  grow(s->iAci, (size_t)N * W); grow(s->Aci, (size_t)nAc_max * 4); grow(s->VAc, (size_t)nAc_max * 2); grow(s->edge_index_Ac, (size_t)nAc_max);
  const int nAc = ufm_mesh_make_Ac(N, T, W, nAc_max, s->V.data(), s->Tri.data(), s->nC.data(), s->C.data(), s->niTri.data(), s->iTri.data(),
                                   s->edge_index.data(), s->iAci.data(), s->Aci.data(), s->VAc.data(), nullptr, nullptr, nullptr, nullptr, s->edge_index_Ac.data());
  if (nAc <= 0) return ufm_set_error(-2, "ufm_mesh_upload_primary: make_Ac_mesh failed (%d): %s", nAc, nAc == -1 ? "more edges than nV + nTri" : "an edge without its triangle(s)");
  s->nAc = nAc;
  lap("make_Ac_mesh");

  const int M = N + nAc;
  grow(s->VAaAc, (size_t)M * 2); grow(s->nCAaAc, (size_t)M); grow(s->CAaAc, (size_t)M * W);   // written completely by make_combined_AaAc_mesh
  rc = ufm_mesh_make_AaAc(N, nAc, nAc_max, W, s->V.data(), s->VAc.data(), s->nC.data(), s->C.data(), s->iAci.data(), s->Aci.data(), s->edge_index_Ac.data(),
                          s->VAaAc.data(), s->nCAaAc.data(), s->CAaAc.data());
  if (rc) return ufm_set_error(-2, "ufm_mesh_upload_primary: make_combined_AaAc_mesh failed (%d)", rc);
  lap("make_combined_AaAc_mesh");

  // The colouring's delete loop hops from a vertex to its neighbours; meshes come numbered in refinement order (random in space), so
  // run it on a Morton relabelling of the graph: same decisions, same colours (see ufm_mesh_five_colouring_labelled), rows in cache.
  std::vector<int> &label = s->label;
  grow(label, (size_t)M);
  {
    std::vector<unsigned long long> &key = s->sort_keys;
    grow(key, (size_t)M);
    const double *X = s->VAaAc.data(), *Y = X + M;
    const double sx = 65535.0 / (p->xmax - p->xmin), sy = 65535.0 / (p->ymax - p->ymin);
    auto spread = [](unsigned long long v) {
      v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555;
      return v;
    };
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; i++) {
      const double fx = std::min(65535.0, std::max(0.0, (X[i] - p->xmin) * sx)), fy = std::min(65535.0, std::max(0.0, (Y[i] - p->ymin) * sy));
      key[i] = ((spread((unsigned long long)fx) | (spread((unsigned long long)fy) << 1)) << 32) | (unsigned long long)i;
    }
    __gnu_parallel::sort(key.begin(), key.end());
#pragma omp parallel for schedule(static)
    for (int k = 0; k < M; k++) label[(int)(key[k] & 0xffffffffull)] = k + 1;
  }
  lap("Morton labels");
  grow(s->colour, (size_t)M); grow(s->colour_vi, (size_t)M * 5); s->colour_nV.assign(5, 0);   // the colouring zeroes and fills colour_vi itself
  {
    const size_t need = ufm_mesh_five_colouring_ws_bytes(M, W);
    if (s->colour_ws_bytes < need) {
      free(s->colour_ws);
      s->colour_ws_bytes = need + need / 16;
      s->colour_ws = malloc(s->colour_ws_bytes);
      if (!s->colour_ws) { s->colour_ws_bytes = 0; return ufm_set_error(-3, "ufm_mesh_upload_primary: out of host memory (five-colouring scratch)"); }
    }
  }
  s->colour_thread = std::thread([s, M, W]() {
    s->colour_rc = ufm_mesh_five_colouring_ws(M, W, s->nCAaAc.data(), s->CAaAc.data(), s->label.data(), s->colour.data(), s->colour_vi.data(), s->colour_nV.data(), s->colour_ws);
  });
  // ... while this thread goes on with what does not need the colours
  grow(s->A, (size_t)N); fill_par(s->Cw, (size_t)N * W, 0.0); grow(s->Tricc, (size_t)T * 2); grow(s->Tri_edge_index, (size_t)T);
  rc = ufm_mesh_geometry(N, T, W, s->V.data(), s->Tri.data(), s->nC.data(), s->C.data(), s->niTri.data(), s->iTri.data(), s->edge_index.data(),
                             p->xmin, p->xmax, p->ymin, p->ymax, s->Tricc.data(), s->Tri_edge_index.data(), s->A.data(), s->Cw.data());
  if (rc) return ufm_set_error(-2, "ufm_mesh_upload_primary: Voronoi areas / connection widths failed (%d): C / iTri / Tri are inconsistent", rc);
  lap("Voronoi areas, Cw");

  if (!async_colour) {
    if ((rc = s->colour_join())) return ufm_set_error(-2, "ufm_mesh_upload_primary: five-colouring failed (%d)%s", rc, rc == -2 ? " (the reference aborts in IDENTIFY)" : "");
    lap("five-colouring (rest)");
  }

  ufm_mesh_desc &d = s->desc;
  memset(&d, 0, sizeof(d));
  d.nV = N; d.nAc = nAc; d.nC_mem = W; d.ldV = N; d.ldAc = nAc_max; d.ldAaAc = M;
  d.V = s->V.data(); d.A = s->A.data(); d.nC = s->nC.data(); d.C = s->C.data(); d.Cw = s->Cw.data(); d.edge_index = s->edge_index.data();
  d.Aci = s->Aci.data(); d.iAci = s->iAci.data(); d.edge_index_Ac = s->edge_index_Ac.data();
  d.nCAaAc = s->nCAaAc.data(); d.CAaAc = s->CAaAc.data(); d.colour_vi = s->colour_vi.data(); d.colour_nV = s->colour_nV.data();
  // Nx .. Nyy_AaAc stay NULL: derived on the device
  if (p->thermo) {
    // determine_mesh_resolution (src/mesh_help_functions_module.f90:193-216), triangle neighbour functions (src/mesh_derivatives_module.f90:30-47)
    s->R.assign(N, p->xmax - p->xmin);
    const double *V = s->V.data();
#pragma omp parallel for schedule(static)
    for (int v = 0; v < N; v++)
      for (int c = 0; c < s->nC[v]; c++) {
        const int q = s->C[(size_t)c * N + v] - 1;
        const double dx = V[q] - V[v], dy = V[N + q] - V[N + v];
        s->R[v] = fmin(s->R[v], sqrt(dx * dx + dy * dy));
      }
    s->NxTri.resize((size_t)T * 3); s->NyTri.resize((size_t)T * 3);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; t++) {
      const int a = s->Tri[t] - 1, b = s->Tri[(size_t)T + t] - 1, c = s->Tri[(size_t)2 * T + t] - 1;
      const double ax = V[a], ay = V[N + a], bx = V[b], by = V[N + b], cx = V[c], cy = V[N + c];
      const double D = ax * (by - cy) + bx * (cy - ay) + cx * (ay - by);
      s->NxTri[t] = (by - cy) / D; s->NxTri[(size_t)T + t] = (cy - ay) / D; s->NxTri[(size_t)2 * T + t] = (ay - by) / D;
      s->NyTri[t] = (cx - bx) / D; s->NyTri[(size_t)T + t] = (ax - cx) / D; s->NyTri[(size_t)2 * T + t] = (bx - ax) / D;
    }
    d.nTri = T; d.ldTri = T; d.Tri = s->Tri.data(); d.niTri = s->niTri.data(); d.iTri = s->iTri.data(); d.R = s->R.data();
    d.NxTri = s->NxTri.data(); d.NyTri = s->NyTri.data();
    lap("resolution, NxTri/NyTri");
  }
  *derived = s;
  guard.s = nullptr;
  return 0;
}
extern "C" void ufm_mesh_derived_free(void *derived) { delete (ufm_secondary *)derived; }

extern "C" int ufm_mesh_upload_primary(ufm_handle *h, const ufm_mesh_primary *p)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  void *derived = nullptr;
  // the previous mesh's derived arrays are valid "until the next upload": this is it -- their buffers are reused
  ufm_secondary *old = (ufm_secondary *)h->secondary;
  h->secondary = nullptr;
  int rc = derive_secondary_impl(p, &derived, true, old);
  if (rc) return rc;
  ufm_secondary *s = (ufm_secondary *)derived;
  const auto t0 = std::chrono::steady_clock::now();
  // ufm_mesh_upload sorts, fills and uploads the Aa / Ac arrays first and asks for the colours only then
  std::function<int()> wait = [s]() {
    const int rc_c = s->colour_join();
    return rc_c ? ufm_set_error(-2, "ufm_mesh_upload_primary: five-colouring failed (%d)%s", rc_c, rc_c == -2 ? " (the reference aborts in IDENTIFY)" : "") : 0;
  };
  g_ufm_colour_wait = &wait;
  rc = ufm_mesh_upload(h, &s->desc);   // drops the arrays derived for the previous mesh
  g_ufm_colour_wait = nullptr;
  if (getenv("UFM_UPLOAD_TIMING"))
    fprintf(stderr, "[ufm_mesh_upload_primary] %-24s %8.1f ms\n", "ufm_mesh_upload", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  if (rc) { delete s; return rc; }
  h->secondary = s;
  if (getenv("UFM_UPLOAD_TIMING"))
    fprintf(stderr, "[ufm_mesh_upload_primary] %-24s %8.1f ms\n", "total since derive", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return 0;
}

// the derived arrays of the last ufm_mesh_upload_primary, valid until the next upload or ufm_destroy.  Neighbour-function members
// are NULL (they exist only on the device); ldAc > nAc.
extern "C" int ufm_mesh_derived_get(const void *derived, ufm_mesh_desc *out, const double **Tricc, const int **Tri_edge_index, const double **VAc,
                                    const double **VAaAc, const int **colour)
{
  if (!derived || !out) return ufm_set_error(-2, "ufm_mesh_derived_get: NULL argument");
  const ufm_secondary *s = (const ufm_secondary *)derived;
  *out = s->desc;
  if (Tricc) *Tricc = s->Tricc.data();
  if (Tri_edge_index) *Tri_edge_index = s->Tri_edge_index.data();
  if (VAc) *VAc = s->VAc.data();
  if (VAaAc) *VAaAc = s->VAaAc.data();
  if (colour) *colour = s->colour.data();
  return 0;
}
extern "C" int ufm_mesh_secondary_get(ufm_handle *h, ufm_mesh_desc *out, const double **Tricc, const int **Tri_edge_index, const double **VAc,
                                      const double **VAaAc, const int **colour)
{
  if (!h || !out) return ufm_set_error(-2, "ufm_mesh_secondary_get: NULL argument");
  if (!h->secondary) return ufm_set_error(-2, "ufm_mesh_secondary_get: the resident mesh was not uploaded with ufm_mesh_upload_primary");
  return ufm_mesh_derived_get(h->secondary, out, Tricc, Tri_edge_index, VAc, VAaAc, colour);
}
