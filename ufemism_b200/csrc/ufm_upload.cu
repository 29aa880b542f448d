// ufm_upload.cu -- ufm_mesh_upload: copy, renumber and compact a reference mesh into the device layout.
//
// Host-side work only (runs once per mesh; mesh generation stays on the CPU per north_star and
// triggers this re-upload, src/UFEMISM_main_model.f90:294).  Coefficients that the reference
// combines inside the SOR sweep with a fixed expression -- 4*Nxx+Nyy, 4*Nyy+Nxx
// (src/ice_dynamics_module.f90:642-643) -- are combined here with the same two fp64 operations
// (this file is compiled without FMA contraction), so the products the sweep forms are bit-identical.
#include <algorithm>
#include <chrono>
#include <parallel/algorithm>
#include <omp.h>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>

#include "ufm_internal.cuh"

namespace {

template <class T>
int dev_upload(ufm_handle *h, const std::vector<T> &v, T **out)
{
  *out = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  int rc = ufm_arena_alloc(h, bytes, (void **)out);
  if (rc) return rc;
  if (!v.empty()) UFM_CUDA(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
template <class T>
int dev_zeros(ufm_handle *h, size_t n, T **out, int own_slot = -1)
{
  *out = nullptr;
  size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  if (own_slot >= 0) {   // buffers exported through CUDA IPC keep an allocation of their own, reused across mesh updates
    ufm_handle::OwnBuf &b = h->own_buf[own_slot];
    if (b.bytes < bytes) {
      if (b.p) cudaFree(b.p);
      b.p = nullptr; b.bytes = 0;
      UFM_CUDA(cudaMalloc(&b.p, bytes));
      b.bytes = bytes;
    }
    *out = (T *)b.p;
  } else { int rc = ufm_arena_alloc(h, bytes, (void **)out); if (rc) return rc; }
  UFM_CUDA(cudaMemset(*out, 0, bytes));
  return 0;
}
#define UP(vec, ptr) do { int rc_ = dev_upload(h, vec, &(ptr)); if (rc_) return rc_; } while (0)
#define ZE(n, ptr) do { int rc_ = dev_zeros(h, (size_t)(n), &(ptr)); if (rc_) return rc_; } while (0)
#define ZE_OWN(n, ptr, slot) do { int rc_ = dev_zeros(h, (size_t)(n), &(ptr), slot); if (rc_) return rc_; } while (0)

inline uint32_t part1by1(uint32_t x)
{
  x &= 0x0000ffff;
  x = (x ^ (x << 8)) & 0x00ff00ff;
  x = (x ^ (x << 4)) & 0x0f0f0f0f;
  x = (x ^ (x << 2)) & 0x33333333;
  x = (x ^ (x << 1)) & 0x55555555;
  return x;
}
inline uint32_t morton2(double x, double y, double x0, double y0, double sx, double sy)
{
  double fx = (x - x0) * sx, fy = (y - y0) * sy;
  uint32_t ix = (uint32_t)std::min(65535.0, std::max(0.0, fx)), iy = (uint32_t)std::min(65535.0, std::max(0.0, fy));
  return part1by1(ix) | (part1by1(iy) << 1);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// slice offsets for rows with the given degrees (UFM_DEG_PAD rows count as degree 0)
void build_slices(const std::vector<unsigned char> &deg, std::vector<long long> &off)
{
  int n_slices = (int)deg.size() / UFM_SLICE;
  off.assign(n_slices + 1, 0);
  for (int s = 0; s < n_slices; s++) {
    int w = 0;
    for (int l = 0; l < UFM_SLICE; l++) {
      unsigned char d = deg[(size_t)s * UFM_SLICE + l];
      if (d != UFM_DEG_PAD) w = std::max(w, (int)d);
    }
    off[s + 1] = off[s] + (long long)w * UFM_SLICE;
  }
}

}  // namespace

// bump allocator over a few large device chunks (see ufm_handle::arena); 256 B alignment keeps every array 128 B-line aligned
int ufm_arena_alloc(ufm_handle *h, size_t bytes, void **out)
{
  bytes = (bytes + 255) & ~(size_t)255;
  while (h->arena_cur < h->arena_n) {
    ufm_handle::Chunk &c = h->arena[h->arena_cur];
    if (h->arena_used + bytes <= c.cap) { *out = c.base + h->arena_used; h->arena_used += bytes; return 0; }
    h->arena_cur++; h->arena_used = 0;
  }
  if (h->arena_n >= 64) return ufm_set_error(-3, "device arena: too many chunks");
  size_t cap = std::max(bytes, std::min<size_t>(std::max<size_t>(h->arena_total, (size_t)32 << 20), (size_t)1 << 30));
  char *p = nullptr;
  UFM_CUDA(cudaMalloc((void **)&p, cap));
  h->arena[h->arena_n] = {p, cap};
  h->arena_cur = h->arena_n++;
  h->arena_total += cap;
  *out = p; h->arena_used = bytes;
  return 0;
}
void ufm_own_release(ufm_handle *h)
{
  for (auto &b : h->own_buf) { if (b.p) cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
  if (h->scal_h_keep) cudaFreeHost(h->scal_h_keep);
  h->scal_h_keep = nullptr;
}
void ufm_arena_release(ufm_handle *h)
{
  for (int k = 0; k < h->arena_n; k++) cudaFree(h->arena[k].base);
  h->arena_n = h->arena_cur = 0; h->arena_used = h->arena_total = 0;
}

#define F2(a, i, j, ld) (a)[((size_t)((j) - 1)) * (size_t)(ld) + (size_t)((i) - 1)]


// ---------------------------------------------------------------------------------------------
// Neighbour functions of the combined AaAc mesh derived ON THE DEVICE (second half): when the host
// passes no Nx_AaAc ... Nyy_AaAc, the five (nV+nAc, nC_mem+1) arrays -- 2.7 GB at 1 M vertices, the bulk of the upload
// and 1.6 s of host arithmetic -- are never built, renumbered or copied; one thread per AaAc row evaluates
// get_neighbour_functions_vertex_gr (src/mesh_derivatives_module.f90:76-313, averaged-gradient approach; free and
// boundary vertices) from the vertex coordinates and writes the pre-combined sweep coefficients straight into the sliced
// ELL arrays.  Expression order follows the reference statement by statement (this file is compiled with -fmad=false),
// so the coefficients are bit-identical to host-built ones (tests/test_gpu_parity.py::test_device_derived_neighbour_functions).
// ---------------------------------------------------------------------------------------------
#define NF_MAX 17
struct NfArgs {
  int n_slices;
  const long long *off;
  const unsigned char *deg, *is_edge;
  const int *idx;
  const double2 *xy;
  double *cU, *cV, *nxy, *nx, *ny, *cU0, *cV0, *nxy0, *nx0, *ny0, *nxysum;
};
__device__ void d_neighbour_functions_vertex_gr(const double xi, const double yi, const int n, const double2 *V_vc, const bool is_edge,
                                                double *Nx, double *Ny, double *Nxx, double *Nxy, double *Nyy)
{
  double NxTri[NF_MAX][3], NyTri[NF_MAX][3], NxSub[NF_MAX][3], NySub[NF_MAX][3];
  for (int c = 0; c <= n; c++) { Nx[c] = 0; Ny[c] = 0; Nxx[c] = 0; Nxy[c] = 0; Nyy[c] = 0; }
#define NX(c) Nx[(c) - 1]
#define NY(c) Ny[(c) - 1]
#define NXX(c) Nxx[(c) - 1]
#define NXY(c) Nxy[(c) - 1]
#define NYY(c) Nyy[(c) - 1]
  const int nTri = is_edge ? n - 1 : n, nSub = is_edge ? n - 2 : n;
  for (int ti = 1; ti <= nTri; ti++) {
    int tip1s = ti + 1; if (tip1s > n) tip1s -= n;
    const double xt = V_vc[ti - 1].x, yt = V_vc[ti - 1].y, xtp1s = V_vc[tip1s - 1].x, ytp1s = V_vc[tip1s - 1].y;
    const double nzt = ((xt - xi) * (ytp1s - yi)) - ((yt - yi) * (xtp1s - xi));
    NxTri[ti - 1][0] = (yt - ytp1s) / nzt; NxTri[ti - 1][1] = (ytp1s - yi) / nzt; NxTri[ti - 1][2] = (yi - yt) / nzt;
    NyTri[ti - 1][0] = (xtp1s - xt) / nzt; NyTri[ti - 1][1] = (xi - xtp1s) / nzt; NyTri[ti - 1][2] = (xt - xi) / nzt;
  }
  for (int si = 1; si <= nSub; si++) {
    int sip1s = si + 1; if (sip1s > n) sip1s -= n;
    int sip2s = sip1s + 1; if (sip2s > n) sip2s -= n;
    const double xs = V_vc[si - 1].x, ys = V_vc[si - 1].y;
    const double xsp1s = V_vc[sip1s - 1].x, ysp1s = V_vc[sip1s - 1].y;
    const double xsp2s = V_vc[sip2s - 1].x, ysp2s = V_vc[sip2s - 1].y;
    const double third = 1.0 / 3.0;
    const double nzs = (third * (xs + xsp1s - 2 * xi) * (ysp1s + ysp2s - 2 * yi)) -
                       (third * (ys + ysp1s - 2 * yi) * (xsp1s + xsp2s - 2 * xi));
    NxSub[si - 1][0] = (ys - ysp2s) / nzs; NxSub[si - 1][1] = (ysp1s + ysp2s - 2.0 * yi) / nzs; NxSub[si - 1][2] = (2.0 * yi - ys - ysp1s) / nzs;
    NySub[si - 1][0] = (xsp2s - xs) / nzs; NySub[si - 1][1] = (2.0 * xi - xsp1s - xsp2s) / nzs; NySub[si - 1][2] = (xs + xsp1s - 2.0 * xi) / nzs;
  }
  if (!is_edge) {
    const double rn = 1.0 / (double)n;
    double sx = 0, sy = 0;
    for (int t = 0; t < n; t++) { sx += NxTri[t][0]; sy += NyTri[t][0]; }
    NX(n + 1) = rn * sx; NY(n + 1) = rn * sy;
    for (int ci = 1; ci <= n; ci++) {
      int cim1s = ci - 1; if (cim1s == 0) cim1s += n;
      NX(ci) = rn * (NxTri[ci - 1][1] + NxTri[cim1s - 1][2]);
      NY(ci) = rn * (NyTri[ci - 1][1] + NyTri[cim1s - 1][2]);
    }
    for (int si = 1; si <= n; si++) {
      int sip1s = si + 1; if (sip1s > n) sip1s -= n;
      NXX(n + 1) = NXX(n + 1) + (rn * ((NxSub[si - 1][0] * NX(n + 1)) + (NxSub[si - 1][1] * NxTri[si - 1][0]) + (NxSub[si - 1][2] * NxTri[sip1s - 1][0])));
      NXY(n + 1) = NXY(n + 1) + (rn * ((NySub[si - 1][0] * NX(n + 1)) + (NySub[si - 1][1] * NxTri[si - 1][0]) + (NySub[si - 1][2] * NxTri[sip1s - 1][0])));
      NYY(n + 1) = NYY(n + 1) + (rn * ((NySub[si - 1][0] * NY(n + 1)) + (NySub[si - 1][1] * NyTri[si - 1][0]) + (NySub[si - 1][2] * NyTri[sip1s - 1][0])));
    }
    double sNxSub1 = 0, sNySub1 = 0;
    for (int q = 0; q < n; q++) { sNxSub1 += NxSub[q][0]; sNySub1 += NySub[q][0]; }
    for (int si = 1; si <= n; si++) {
      int sim1s = si - 1; if (sim1s < 1) sim1s += n;
      int sim2s = sim1s - 1; if (sim2s < 1) sim2s += n;
      NXX(si) = rn * ((NxTri[si - 1][1] * (NxSub[si - 1][1] + NxSub[sim1s - 1][2]) +
                      (NxTri[sim1s - 1][2] * (NxSub[sim1s - 1][1] + NxSub[sim2s - 1][2]) + (NX(si) * sNxSub1))));
      NXY(si) = rn * ((NxTri[si - 1][1] * (NySub[si - 1][1] + NySub[sim1s - 1][2]) +
                      (NxTri[sim1s - 1][2] * (NySub[sim1s - 1][1] + NySub[sim2s - 1][2]) + (NX(si) * sNySub1))));
      NYY(si) = rn * ((NyTri[si - 1][1] * (NySub[si - 1][1] + NySub[sim1s - 1][2]) +
                      (NyTri[sim1s - 1][2] * (NySub[sim1s - 1][1] + NySub[sim2s - 1][2]) + (NY(si) * sNySub1))));
    }
  } else {
    const double rn1 = 1.0 / (double)(n - 1), rn2 = 1.0 / (double)(n - 2);
    double sx = 0, sy = 0;
    for (int t = 0; t < n - 1; t++) { sx += NxTri[t][0]; sy += NyTri[t][0]; }
    NX(n + 1) = rn1 * sx; NY(n + 1) = rn1 * sy;
    for (int ci = 1; ci <= n; ci++) {
      if (ci == 1) { NX(ci) = rn1 * NxTri[0][1]; NY(ci) = rn1 * NyTri[0][1]; }
      else if (ci == n) { NX(ci) = rn1 * NxTri[n - 2][2]; NY(ci) = rn1 * NyTri[n - 2][2]; }
      else { NX(ci) = rn1 * (NxTri[ci - 1][1] + NxTri[ci - 2][2]); NY(ci) = rn1 * (NyTri[ci - 1][1] + NyTri[ci - 2][2]); }
    }
    for (int si = 1; si <= n - 2; si++) {
      NXX(n + 1) = NXX(n + 1) + (rn2 * ((NxSub[si - 1][0] * NX(n + 1)) + (NxSub[si - 1][1] * NxTri[si - 1][0]) + (NxSub[si - 1][2] * NxTri[si][0])));
      NXY(n + 1) = NXY(n + 1) + (rn2 * ((NySub[si - 1][0] * NX(n + 1)) + (NySub[si - 1][1] * NxTri[si - 1][0]) + (NySub[si - 1][2] * NxTri[si][0])));
      NYY(n + 1) = NYY(n + 1) + (rn2 * ((NySub[si - 1][0] * NY(n + 1)) + (NySub[si - 1][1] * NyTri[si - 1][0]) + (NySub[si - 1][2] * NyTri[si][0])));
    }
    double sNxSub1 = 0, sNySub1 = 0;
    for (int q = 0; q < n - 2; q++) { sNxSub1 += NxSub[q][0]; sNySub1 += NySub[q][0]; }
    for (int si = 1; si <= n; si++) {
      double Axx = 0, Axy = 0, Ayy = 0, Bxx = 0, Bxy = 0, Byy = 0, Cxx = 0, Cxy = 0, Cyy = 0;
      if (si < n - 1) {
        Axx = NxSub[si - 1][1] * NxTri[si - 1][1]; Axy = NySub[si - 1][1] * NxTri[si - 1][1]; Ayy = NySub[si - 1][1] * NyTri[si - 1][1];
      }
      if (si > 1 && si < n) {
        Bxx = (NxSub[si - 2][1] * NxTri[si - 2][2]) + (NxSub[si - 2][2] * NxTri[si - 1][1]);
        Bxy = (NySub[si - 2][1] * NxTri[si - 2][2]) + (NySub[si - 2][2] * NxTri[si - 1][1]);
        Byy = (NySub[si - 2][1] * NyTri[si - 2][2]) + (NySub[si - 2][2] * NyTri[si - 1][1]);
      }
      if (si > 2) {
        Cxx = NxSub[si - 3][2] * NxTri[si - 2][2]; Cxy = NySub[si - 3][2] * NxTri[si - 2][2]; Cyy = NySub[si - 3][2] * NyTri[si - 2][2];
      }
      NXX(si) = rn2 * (Axx + Bxx + Cxx + (NX(si) * sNxSub1));
      NXY(si) = rn2 * (Axy + Bxy + Cxy + (NX(si) * sNySub1));
      NYY(si) = rn2 * (Ayy + Byy + Cyy + (NY(si) * sNySub1));
    }
  }
#undef NX
#undef NY
#undef NXX
#undef NXY
#undef NYY
}

__global__ void __launch_bounds__(128) k_derive_nf_AaAc(NfArgs a)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < a.n_slices; s += nw) {
    const long long o = a.off[s];
    const int w = (int)((a.off[s + 1] - o) >> 5);
    const int p = s * 32 + lane;
    const int n = a.deg[p];
    double Nx[NF_MAX], Ny[NF_MAX], Nxx[NF_MAX], Nxy[NF_MAX], Nyy[NF_MAX];
    const bool live = n != UFM_DEG_PAD && n < NF_MAX;
    if (live) {
      double2 V_vc[NF_MAX];
      for (int c = 0; c < n; c++) V_vc[c] = a.xy[a.idx[o + (long long)c * 32 + lane]];
      const double2 vi = a.xy[p];
      d_neighbour_functions_vertex_gr(vi.x, vi.y, n, V_vc, a.is_edge[p] != 0, Nx, Ny, Nxx, Nxy, Nyy);
    }
    for (int c = 0; c < w; c++) {
      const long long e = o + (long long)c * 32 + lane;
      const bool in = live && c < n;
      a.cU[e] = in ? 4.0 * Nxx[c] + Nyy[c] : 0.0;
      a.cV[e] = in ? 4.0 * Nyy[c] + Nxx[c] : 0.0;
      a.nxy[e] = in ? Nxy[c] : 0.0;
      a.nx[e] = in ? Nx[c] : 0.0;
      a.ny[e] = in ? Ny[c] : 0.0;
    }
    double cu0 = 0.0, cv0 = 0.0, xy0 = 0.0, x0 = 0.0, y0 = 0.0, ssum = 0.0;
    if (live) {
      cu0 = 4.0 * Nxx[n] + Nyy[n]; cv0 = 4.0 * Nyy[n] + Nxx[n]; xy0 = Nxy[n]; x0 = Nx[n]; y0 = Ny[n];
      ssum = xy0;
      for (int c = 0; c < n; c++) ssum = ssum + Nxy[c];
    }
    a.cU0[p] = cu0; a.cV0[p] = cv0; a.nxy0[p] = xy0; a.nx0[p] = x0; a.ny0[p] = y0; a.nxysum[p] = ssum;
  }
}

// Aa neighbour functions Nx, Ny (the only two the path reads on the Aa mesh) from the vertex coordinates
__global__ void __launch_bounds__(128) k_derive_nf_Aa(int n_slices, const long long *off, const unsigned char *deg, const unsigned char *edge,
                                                      const int *C, const double2 *xy, double *Nx_o, double *Ny_o, double *Nx0, double *Ny0)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = wg; s < n_slices; s += nw) {
    const long long o = off[s];
    const int w = (int)((off[s + 1] - o) >> 5);
    const int p = s * 32 + lane;
    const int n = deg[p];
    double Nx[NF_MAX], Ny[NF_MAX], Nxx[NF_MAX], Nxy[NF_MAX], Nyy[NF_MAX];
    const bool live = n != UFM_DEG_PAD && n < NF_MAX;
    if (live) {
      double2 V_vc[NF_MAX];
      for (int c = 0; c < n; c++) V_vc[c] = xy[C[o + (long long)c * 32 + lane]];
      const double2 vi = xy[p];
      d_neighbour_functions_vertex_gr(vi.x, vi.y, n, V_vc, edge[p] != 0, Nx, Ny, Nxx, Nxy, Nyy);
    }
    for (int c = 0; c < w; c++) {
      const long long e = o + (long long)c * 32 + lane;
      Nx_o[e] = (live && c < n) ? Nx[c] : 0.0;
      Ny_o[e] = (live && c < n) ? Ny[c] : 0.0;
    }
    Nx0[p] = live ? Nx[n] : 0.0; Ny0[p] = live ? Ny[n] : 0.0;
  }
}

// Ac neighbour functions (make_Ac_mesh, src/mesh_ArakawaC_module.f90:24-234): gradients of the two triangles adjacent to the
// edge, averaged; boundary segments have one triangle only.  One thread per staggered vertex.
__device__ __forceinline__ bool d_is_boundary_segment(const int a, const int b)
{
  if (a == 0 || b == 0) return false;
  if ((a == 1 || a == 2 || a == 8) && (b == 1 || b == 2 || b == 8)) return true;
  if ((a == 2 || a == 3 || a == 4) && (b == 2 || b == 3 || b == 4)) return true;
  if ((a == 4 || a == 5 || a == 6) && (b == 4 || b == 5 || b == 6)) return true;
  if ((a == 6 || a == 7 || a == 8) && (b == 6 || b == 7 || b == 8)) return true;
  return false;
}
struct NfAcArgs { int nAc; const int4 *Aci; const double2 *xy; const unsigned char *edge; double *Nx[4], *Ny[4], *No[4], *Np; };
__global__ void __launch_bounds__(256) k_derive_nf_Ac(NfAcArgs a)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.nAc) return;
  const int4 v = a.Aci[p];
  const double2 Vi = a.xy[v.x], Vj = a.xy[v.y], Vl = a.xy[v.z];
  double Nxl[4], Nyl[4], Nx[4], Ny[4];
  Nxl[0] = Vl.y - Vj.y; Nxl[1] = Vi.y - Vl.y; Nxl[2] = Vj.y - Vi.y; Nxl[3] = 0.0;
  Nyl[0] = Vj.x - Vl.x; Nyl[1] = Vl.x - Vi.x; Nyl[2] = Vi.x - Vj.x; Nyl[3] = 0.0;
  const double Nzl = ((Vj.x - Vi.x) * (Vl.y - Vi.y)) - ((Vj.y - Vi.y) * (Vl.x - Vi.x));
  if (!d_is_boundary_segment(a.edge[v.x], a.edge[v.y])) {
    const double2 Vr = a.xy[v.w];
    double Nxr[4], Nyr[4];
    Nxr[0] = Vj.y - Vr.y; Nxr[1] = Vr.y - Vi.y; Nxr[2] = 0.0; Nxr[3] = Vi.y - Vj.y;
    Nyr[0] = Vr.x - Vj.x; Nyr[1] = Vi.x - Vr.x; Nyr[2] = 0.0; Nyr[3] = Vj.x - Vi.x;
    const double Nzr = ((Vr.x - Vi.x) * (Vj.y - Vi.y)) - ((Vr.y - Vi.y) * (Vj.x - Vi.x));
    for (int k = 0; k < 4; k++) {
      Nx[k] = -((Nxl[k] / Nzl) + (Nxr[k] / Nzr)) / 2.0;
      Ny[k] = -((Nyl[k] / Nzl) + (Nyr[k] / Nzr)) / 2.0;
    }
  } else {
    for (int k = 0; k < 4; k++) { Nx[k] = -Nxl[k] / Nzl; Ny[k] = -Nyl[k] / Nzl; }
  }
  const double Ux = Vj.x - Vi.x, Uy = Vj.y - Vi.y, U = sqrt(Ux * Ux + Uy * Uy);
  a.Np[p] = 1.0 / U;
  for (int k = 0; k < 4; k++) {
    a.Nx[k][p] = Nx[k]; a.Ny[k][p] = Ny[k];
    a.No[k][p] = (Ny[k] * Ux - Nx[k] * Uy) / U;
  }
}

// owner rank of every AaAc row from its x coordinate: P strips holding equally many rows
static void ufm_partition_owners_impl(const std::vector<double> &X, int P, std::vector<unsigned char> &owner)
{
  const int M = (int)X.size();
  owner.assign(M, 0);
  if (P <= 1) return;
  std::vector<int> byx(M);
  std::iota(byx.begin(), byx.end(), 0);
  __gnu_parallel::stable_sort(byx.begin(), byx.end(), [&](int a, int b) { return X[a] < X[b]; });
  for (int k = 0; k < M; k++) owner[byx[k]] = (unsigned char)(((long long)k * P) / M);
}

// Who reads what across the partition (per-step kernels): rd_aa[v*P + q] = 1 when rank q reads Aa vertex v -- v is a neighbour of one
// of q's vertices (gradients, masks, thickness fluxes) or one of the four vertices of one of q's staggered vertices (map_Aa_to_Ac,
// get_mesh_derivatives_Ac); rd_ac[a*P + q] = 1 when staggered vertex a lies on a connection of one of q's vertices (thickness fluxes,
// map_Ac_to_Aa).  Reference (0-based) indices.
static void ufm_partition_reads_impl(const ufm_mesh_desc *d, const std::vector<unsigned char> &owner, int P, std::vector<unsigned char> &rd_aa,
                                     std::vector<unsigned char> &rd_ac)
{
  const int N = d->nV, E = d->nAc, ldV = d->ldV ? d->ldV : N, ldAc = d->ldAc ? d->ldAc : E;
  rd_aa.assign((size_t)N * P, 0); rd_ac.assign((size_t)E * P, 0);
#pragma omp parallel for schedule(static)
  for (int v = 0; v < N; v++) {
    const int q = owner[v];
    for (int c = 1; c <= d->nC[v]; c++) {
      const int u = F2(d->C, v + 1, c, ldV) - 1, a = F2(d->iAci, v + 1, c, ldV) - 1;
      if (u >= 0 && u < N) rd_aa[(size_t)u * P + q] = 1;      // benign race: every writer stores 1
      if (a >= 0 && a < E) rd_ac[(size_t)a * P + q] = 1;
    }
  }
#pragma omp parallel for schedule(static)
  for (int a = 0; a < E; a++) {
    const int q = owner[N + a];
    for (int k = 1; k <= 4; k++) { const int u = F2(d->Aci, a + 1, k, ldAc) - 1; if (u >= 0 && u < N) rd_aa[(size_t)u * P + q] = 1; }
  }
}

// host-only planning entry point (no device work): the partition ufm_mesh_upload will use
extern "C" int ufm_partition_owners(const ufm_mesh_desc *d, int nranks, unsigned char *owner_out)
{
  if (!d || !d->V || !d->Aci || !owner_out) return ufm_set_error(-2, "ufm_partition_owners: NULL argument");
  if (nranks < 1 || nranks > UFM_MAX_RANKS) return ufm_set_error(-2, "ufm_partition_owners: nranks out of range");
  const int N = d->nV, E = d->nAc, ldV = d->ldV ? d->ldV : N, ldAc = d->ldAc ? d->ldAc : E;
  std::vector<double> X((size_t)N + E);
  for (int v = 1; v <= N; v++) X[v - 1] = F2(d->V, v, 1, ldV);
  for (int a = 1; a <= E; a++) X[N + a - 1] = 0.5 * (X[F2(d->Aci, a, 1, ldAc) - 1] + X[F2(d->Aci, a, 2, ldAc) - 1]);
  std::vector<unsigned char> owner;
  ufm_partition_owners_impl(X, nranks, owner);
  memcpy(owner_out, owner.data(), owner.size());
  return 0;
}

// host-only planning entry point: how many Aa / Ac values rank s sends to rank q in one halo exchange of the partitioned per-step
// kernels (cnt[s*P + q]; the diagonal is 0)
extern "C" int ufm_partition_halo_counts(const ufm_mesh_desc *d, int nranks, int *cnt_aa, int *cnt_ac)
{
  if (!d || !d->V || !d->Aci || !d->C || !d->nC || !d->iAci || !cnt_aa || !cnt_ac) return ufm_set_error(-2, "ufm_partition_halo_counts: NULL argument");
  if (nranks < 1 || nranks > UFM_MAX_RANKS) return ufm_set_error(-2, "ufm_partition_halo_counts: nranks out of range");
  const int N = d->nV, E = d->nAc, ldV = d->ldV ? d->ldV : N, ldAc = d->ldAc ? d->ldAc : E, P = nranks;
  std::vector<double> X((size_t)N + E);
  for (int v = 1; v <= N; v++) X[v - 1] = F2(d->V, v, 1, ldV);
  for (int a = 1; a <= E; a++) X[N + a - 1] = 0.5 * (X[F2(d->Aci, a, 1, ldAc) - 1] + X[F2(d->Aci, a, 2, ldAc) - 1]);
  std::vector<unsigned char> owner, rd_aa, rd_ac;
  ufm_partition_owners_impl(X, P, owner);
  ufm_partition_reads_impl(d, owner, P, rd_aa, rd_ac);
  for (int k = 0; k < P * P; k++) cnt_aa[k] = cnt_ac[k] = 0;
  for (int v = 0; v < N; v++) for (int q = 0; q < P; q++) if (q != owner[v] && rd_aa[(size_t)v * P + q]) cnt_aa[owner[v] * P + q]++;
  for (int a = 0; a < E; a++) for (int q = 0; q < P; q++) if (q != owner[N + a] && rd_ac[(size_t)a * P + q]) cnt_ac[owner[N + a] * P + q]++;
  return 0;
}

// Device row order of the AaAc mesh: block-major (blocks 1..5 = colours, 6 = domain-edge rows), then owner rank, partition-boundary rows
// last, the colour-5 rows the Neumann pass reads first, and inside such a group
//   n_bands == 0: (degree, Morton)                                            -- the order measured in DESIGN.md section 4
//   n_bands  > 0: n_bands x-bands, Morton inside a band, the degree sorted only inside windows of deg_window rows (experimental; it
//                 puts the rows a row depends on at nearly the same relative position of the previous colour block, which a sweep
//                 without grid barriers between the colours needs: tools/sor_dataflow_analysis.py).
// Results never depend on the order inside a colour block.  Host only; checked against numpy by tests/test_abi.py.
// Stage 1 of the default row order, independent of the colouring: rows by (owner, boundary, degree, Morton, index).  One packed key per
// row instead of four indirections per comparison.
static void ufm_row_presort(int M, const unsigned char *owner, const unsigned char *isb, const unsigned char *degv, const uint32_t *mort, std::vector<int> &pre)
{
  pre.resize(M);
  std::iota(pre.begin(), pre.end(), 0);
  std::vector<uint64_t> key(M);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; i++) key[i] = ((uint64_t)owner[i] << 42) | ((uint64_t)(isb[i] != 0) << 41) | ((uint64_t)degv[i] << 32) | (uint64_t)mort[i];
  __gnu_parallel::sort(pre.begin(), pre.end(), [&](int a, int b) { return key[a] != key[b] ? key[a] < key[b] : a < b; });
}
// Stage 2: a stable distribution of the pre-sorted rows over the buckets (block, owner, boundary, late).  Inside a bucket owner and
// boundary are constant, so the rows keep the (degree, Morton, index) order of stage 1: the result is the lexicographic order by
// (block, owner, boundary, late, degree, Morton, index).
static void ufm_row_order_from_presorted(int M, const std::vector<int> &pre, const unsigned char *blkv, const unsigned char *owner, const unsigned char *isb,
                                         const unsigned char *late, std::vector<int> &m_order)
{
  int P = 1;
  for (int i = 0; i < M; i++) P = std::max(P, (int)owner[i] + 1);
  auto bucket = [&](int i) { return (((int)blkv[i] * P + (int)owner[i]) * 2 + (isb[i] != 0)) * 2 + (late[i] != 0); };
  const int nb = 256 * P * 4;
  const int T = std::max(1, omp_get_max_threads());
  std::vector<size_t> cnt((size_t)T * nb, 0);
  // per-thread counts over contiguous pieces of the pre-sorted list, then a bucket-major / thread-minor prefix sum: stable
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < T; t++) {   // T pieces, however many threads actually run
    const size_t lo = (size_t)M * t / T, hi = (size_t)M * (t + 1) / T;
    size_t *c = cnt.data() + (size_t)t * nb;
    for (size_t k = lo; k < hi; k++) c[bucket(pre[k])]++;
  }
  size_t run = 0;
  for (int b = 0; b < nb; b++)
    for (int t = 0; t < T; t++) { const size_t c = cnt[(size_t)t * nb + b]; cnt[(size_t)t * nb + b] = run; run += c; }
  m_order.resize(M);
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < T; t++) {
    const size_t lo = (size_t)M * t / T, hi = (size_t)M * (t + 1) / T;
    size_t *c = cnt.data() + (size_t)t * nb;
    for (size_t k = lo; k < hi; k++) { const int i = pre[k]; m_order[c[bucket(i)]++] = i; }
  }
}

static void ufm_row_order_impl(int M, const unsigned char *blkv, const unsigned char *owner, const unsigned char *isb, const unsigned char *late,
                               const unsigned char *degv, const uint32_t *mort, const double *X, int n_bands, int deg_window, std::vector<int> &m_order)
{
  m_order.resize(M);
  std::iota(m_order.begin(), m_order.end(), 0);
  n_bands = std::max(0, std::min(n_bands, 65535));
  deg_window = std::max(UFM_SLICE, deg_window);
  auto group_less = [&](int a, int b, bool &equal) {
    equal = false;
    if (blkv[a] != blkv[b]) return blkv[a] < blkv[b];
    if (owner[a] != owner[b]) return owner[a] < owner[b];
    if (isb[a] != isb[b]) return isb[a] < isb[b];
    if (late[a] != late[b]) return late[a] < late[b];
    equal = true;
    return false;
  };
  if (n_bands == 0) {
    // the default order (block, owner, boundary, late, degree, Morton, index) in two stages, so that ufm_mesh_upload can run the expensive
    // one while the five-colouring -- which decides block and late -- is still being computed on another thread
    std::vector<int> pre;
    ufm_row_presort(M, owner, isb, degv, mort, pre);
    ufm_row_order_from_presorted(M, pre, blkv, owner, isb, late, m_order);
    return;
  }
  double x0 = 1e300, x1 = -1e300;
  for (int i = 0; i < M; i++) { x0 = std::min(x0, X[i]); x1 = std::max(x1, X[i]); }
  std::vector<unsigned short> band(M);
  const double bs = (double)n_bands / std::max(x1 - x0, 1e-300);
  for (int i = 0; i < M; i++) band[i] = (unsigned short)std::min((double)(n_bands - 1), std::max(0.0, (X[i] - x0) * bs));
  __gnu_parallel::stable_sort(m_order.begin(), m_order.end(), [&](int a, int b) {
    bool eq;
    const bool lt = group_less(a, b, eq);
    if (!eq) return lt;
    if (band[a] != band[b]) return band[a] < band[b];
    return mort[a] < mort[b];
  });
  // window index of every row inside its group, then the degree sorted inside the windows (stable: band / Morton order survives)
  std::vector<int> window(M);
  for (int k = 0, start = 0; k < M; k++) {
    bool eq = true;
    if (k > 0) group_less(m_order[k - 1], m_order[k], eq);
    if (!eq) start = k;
    window[m_order[k]] = (k - start) / deg_window;
  }
  __gnu_parallel::stable_sort(m_order.begin(), m_order.end(), [&](int a, int b) {
    bool eq;
    const bool lt = group_less(a, b, eq);
    if (!eq) return lt;
    if (window[a] != window[b]) return window[a] < window[b];
    return degv[a] < degv[b];
  });
}

// host-only planning / test entry point (no device work): the row order ufm_mesh_upload would use for these keys
extern "C" int ufm_plan_row_order(int M, const unsigned char *block, const unsigned char *owner, const unsigned char *boundary, const unsigned char *late,
                                  const unsigned char *degree, const unsigned *morton, const double *X, int n_bands, int deg_window, int *order_out)
{
  if (M < 0 || !block || !owner || !boundary || !late || !degree || !morton || !X || !order_out) return ufm_set_error(-2, "ufm_plan_row_order: bad argument");
  std::vector<int> o;
  ufm_row_order_impl(M, block, owner, boundary, late, degree, morton, X, n_bands, deg_window, o);
  memcpy(order_out, o.data(), sizeof(int) * (size_t)M);
  return 0;
}

// set by ufm_mesh_upload_primary for the duration of its ufm_mesh_upload call: blocks until colour_vi / colour_nV of the descriptor are
// final (they are computed concurrently with the colour-independent half of the upload) and returns the colouring's return code
thread_local std::function<int()> *g_ufm_colour_wait = nullptr;

int ufm_mesh_free_impl(ufm_handle *h)
{
  cudaDeviceSynchronize();   // nothing may still be reading the arrays that are handed back to the arena
  ufm_comm_reset(h);
  DevState &s = h->st;
  // s.UV, s.partials, s.mail (CUDA IPC) and the pinned scalars stay allocated in h->own_buf / h->scal_h_keep for the next mesh
  h->arena_cur = 0; h->arena_used = 0;   // every other array lives in the arena, which is kept for the next mesh
  h->mesh = DevMesh();
  h->st = DevState();
  h->has_mesh = false;
  h->cfl_ok[0] = h->cfl_ok[1] = h->cfl_ok[2] = false;   // cached critical time steps belong to the state that goes away
  return 0;
}

int ufm_mesh_upload_impl(ufm_handle *h, const ufm_mesh_desc *d)
{
  if (!d || !d->V || !d->nC || !d->C || !d->Aci || !d->iAci || !d->nCAaAc || !d->CAaAc || !d->colour_vi || !d->colour_nV)
    return ufm_set_error(-2, "ufm_mesh_upload: NULL pointer in mesh descriptor");
  const int N = d->nV, E = d->nAc, M = N + E, W = d->nC_mem;
  const int ldV = d->ldV ? d->ldV : N, ldAc = d->ldAc ? d->ldAc : E, ldM = d->ldAaAc ? d->ldAaAc : M;
  if (N < 5 || E < 4 || W < 3 || W > 64) return ufm_set_error(-2, "ufm_mesh_upload: implausible sizes nV=%d nAc=%d nC_mem=%d", N, E, W);
  const auto t_free0 = std::chrono::steady_clock::now();
  ufm_mesh_free_impl(h);   // also after a failed upload: hands every array back to the arena
  const double ms_free = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_free0).count();
  // The host-side renumbering below is OpenMP-parallel.  Launchers such as torchrun export OMP_NUM_THREADS=1 for every
  // rank; one rank per GPU still has (cores / ranks) cores to itself, so use them for the duration of the upload.
  struct OmpGuard {
    int old;
    explicit OmpGuard(int ranks) : old(omp_get_max_threads()) {
      const char *e = getenv("UFM_UPLOAD_THREADS");
      const int t = e ? atoi(e) : std::max(old, omp_get_num_procs() / std::max(1, ranks));
      omp_set_num_threads(std::max(1, t));
    }
    ~OmpGuard() { omp_set_num_threads(old); }
  } omp_guard(h->part_n);
  DevMesh &m = h->mesh;
  m.nV = N; m.nAc = E; m.M = M;
  const bool timing = getenv("UFM_UPLOAD_TIMING") != nullptr;
  if (timing) fprintf(stderr, "[ufm_mesh_upload] %-28s %8.1f ms\n", "free previous mesh", ms_free);
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[ufm_mesh_upload] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };

  // ---- coordinates of all AaAc vertices, bounding box ----
  std::vector<double> X(M), Y(M);
  double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
#pragma omp parallel for schedule(static) reduction(min : x0, y0) reduction(max : x1, y1)
  for (int v = 1; v <= N; v++) {
    X[v - 1] = F2(d->V, v, 1, ldV); Y[v - 1] = F2(d->V, v, 2, ldV);
    x0 = std::min(x0, X[v - 1]); x1 = std::max(x1, X[v - 1]); y0 = std::min(y0, Y[v - 1]); y1 = std::max(y1, Y[v - 1]);
  }
  int bad_aci = 0;
#pragma omp parallel for schedule(static) reduction(max : bad_aci)
  for (int a = 1; a <= E; a++) {
    int vi = F2(d->Aci, a, 1, ldAc), vj = F2(d->Aci, a, 2, ldAc);
    if (vi < 1 || vi > N || vj < 1 || vj > N) { bad_aci = std::max(bad_aci, a); continue; }
    X[N + a - 1] = 0.5 * (X[vi - 1] + X[vj - 1]); Y[N + a - 1] = 0.5 * (Y[vi - 1] + Y[vj - 1]);
  }
  if (bad_aci) return ufm_set_error(-2, "ufm_mesh_upload: Aci out of range at aci=%d", bad_aci);
  const double sx = 65535.0 / std::max(x1 - x0, 1e-300), sy = 65535.0 / std::max(y1 - y0, 1e-300);
  std::vector<uint32_t> mort(M);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; i++) mort[i] = morton2(X[i], Y[i], x0, y0, sx, sy);

  lap("coordinates + morton");
  // ---- Aa and Ac orders: Morton ----
  std::vector<int> aa_order(N), ac_order(E);
  std::iota(aa_order.begin(), aa_order.end(), 0);
  std::iota(ac_order.begin(), ac_order.end(), 0);
  __gnu_parallel::stable_sort(aa_order.begin(), aa_order.end(), [&](int a, int b) { return mort[a] < mort[b]; });
  __gnu_parallel::stable_sort(ac_order.begin(), ac_order.end(), [&](int a, int b) { return mort[N + a] < mort[N + b]; });
  m.nVp = round_up(N, UFM_SLICE); m.nAcp = round_up(E, UFM_SLICE);
  std::vector<int> aa_r2d(N), aa_d2r(m.nVp, -1), ac_r2d(E), ac_d2r(m.nAcp, -1);
  for (int p = 0; p < N; p++) { aa_r2d[aa_order[p]] = p; aa_d2r[p] = aa_order[p]; }
  for (int p = 0; p < E; p++) { ac_r2d[ac_order[p]] = p; ac_d2r[p] = ac_order[p]; }

  lap("Aa/Ac sort");
  // The Aa and Ac arrays below do not depend on the colouring: ufm_mesh_upload_primary computes it on another thread meanwhile.
  // ---- Aa sliced ELL ----
  {
    std::vector<unsigned char> deg(m.nVp, UFM_DEG_PAD), edge(m.nVp, 0);
    for (int p = 0; p < N; p++) {
      int n = d->nC[aa_d2r[p]];
      if (n < 2 || n > W) return ufm_set_error(-2, "ufm_mesh_upload: nC(%d) = %d out of range", aa_d2r[p] + 1, n);
      deg[p] = (unsigned char)n;
      edge[p] = (unsigned char)d->edge_index[aa_d2r[p]];
    }
    std::vector<long long> off;
    build_slices(deg, off);
    m.aa.n_rows = m.nVp; m.aa.n_slices = m.nVp / UFM_SLICE; m.aa.n_entries = off.back();
    size_t ne = (size_t)off.back();
    std::vector<int> Cn(ne), iA(ne, 0);
    const bool derive_aa = !d->Nx || !d->Ny;   // no Aa neighbour functions from the host: derived on the device below
    std::vector<double> nx(derive_aa ? 0 : ne, 0.0), ny(derive_aa ? 0 : ne, 0.0), nx0(derive_aa ? 0 : m.nVp, 0.0), ny0(derive_aa ? 0 : m.nVp, 0.0), A(m.nVp, 1.0), sA(m.nVp, 1.0);
    int bad_vertex = 0;
    // thermodynamics: triangles around every vertex (iTri order is the search order of get_upwind_derivative_vertex_3D)
    const bool has_tri = d->Tri && d->niTri && d->iTri && d->R && d->NxTri && d->NyTri && d->nTri > 0;
    m.has_tri = has_tri; m.nTri = has_tri ? d->nTri : 0;
    std::vector<int> iT(has_tri ? ne : 0, -1);
    std::vector<double2> xy(m.nVp, make_double2(0.0, 0.0));   // vertex coordinates: benchmark SMB closed forms, upwind search
    std::vector<double> rmin(m.nVp, 1.0);
    std::vector<double> Rr(has_tri ? m.nVp : 0, 1.0);
#pragma omp parallel for schedule(static)
    for (int s = 0; s < m.aa.n_slices; s++) {
      int w = (int)((off[s + 1] - off[s]) / UFM_SLICE);
      for (int l = 0; l < UFM_SLICE; l++) {
        int p = s * UFM_SLICE + l, vi = aa_d2r[p];
        for (int c = 0; c < w; c++) Cn[(size_t)off[s] + (size_t)c * UFM_SLICE + l] = p < N ? p : 0;
        if (vi < 0) continue;
        int n = deg[p];
        for (int c = 1; c <= n; c++) {
          size_t e = (size_t)off[s] + (size_t)(c - 1) * UFM_SLICE + l;
          int vc = F2(d->C, vi + 1, c, ldV), aci = F2(d->iAci, vi + 1, c, ldV);
          if (vc < 1 || vc > N || aci < 1 || aci > E) { bad_vertex = vi + 1; continue; }
          Cn[e] = aa_r2d[vc - 1];
          int first = (F2(d->Aci, aci, 1, ldAc) == vi + 1);
          iA[e] = ac_r2d[aci - 1] | (first ? (int)0x80000000u : 0);
          if (!derive_aa) { nx[e] = F2(d->Nx, vi + 1, c, ldV); ny[e] = F2(d->Ny, vi + 1, c, ldV); }
        }
        if (!derive_aa) { nx0[p] = F2(d->Nx, vi + 1, n + 1, ldV); ny0[p] = F2(d->Ny, vi + 1, n + 1, ldV); }
        if (has_tri) {
          const int nt = d->niTri[vi];
          if (nt < 0 || nt > n) { bad_vertex = vi + 1; continue; }
          for (int c = 1; c <= nt; c++) {
            const int ti = F2(d->iTri, vi + 1, c, ldV);
            if (ti < 1 || ti > d->nTri) { bad_vertex = vi + 1; continue; }
            iT[(size_t)off[s] + (size_t)(c - 1) * UFM_SLICE + l] = ti - 1;
          }
          Rr[p] = d->R[vi];
        }
        xy[p] = make_double2(F2(d->V, vi + 1, 1, ldV), F2(d->V, vi + 1, 2, ldV));
        A[p] = d->A[vi];
        sA[p] = std::sqrt(d->A[vi] / UFM_PI);
        // numerator of the vertex's SSA critical time step: SQRT(A/pi) or its shortest connection (UFEMISM_main_model.f90:751-762)
        double rm = sA[p];
        for (int c = 1; c <= n; c++) {
          const int vc = F2(d->C, vi + 1, c, ldV);
          if (vc < 1 || vc > N) continue;
          const double ddx = F2(d->V, vc, 1, ldV) - xy[p].x, ddy = F2(d->V, vc, 2, ldV) - xy[p].y;
          rm = std::min(rm, std::sqrt(ddx * ddx + ddy * ddy));
        }
        rmin[p] = rm;
      }
    }
    if (bad_vertex) return ufm_set_error(-2, "ufm_mesh_upload: C/iAci out of range at vertex %d", bad_vertex);
    lap("Aa ELL fill");
    UP(off, m.aa.off); UP(deg, m.aa.deg); UP(Cn, m.aa_C); UP(iA, m.aa_iAci);
    UP(A, m.aa_A); UP(sA, m.aa_sqrtApi); UP(rmin, m.aa_rmin); UP(edge, m.aa_edge); UP(xy, m.aa_xy);
    if (!derive_aa) { UP(nx, m.aa_Nx); UP(ny, m.aa_Ny); UP(nx0, m.aa_Nx0); UP(ny0, m.aa_Ny0); }
    else {
      double **q4[] = {&m.aa_Nx, &m.aa_Ny};
      for (double **q : q4) { int rc_ = ufm_arena_alloc(h, std::max<size_t>(ne, 1) * sizeof(double), (void **)q); if (rc_) return rc_; }
      double **q2[] = {&m.aa_Nx0, &m.aa_Ny0};
      for (double **q : q2) { int rc_ = ufm_arena_alloc(h, (size_t)m.nVp * sizeof(double), (void **)q); if (rc_) return rc_; }
      k_derive_nf_Aa<<<(m.aa.n_slices * 32 + 127) / 128, 128, 0, h->stream>>>(m.aa.n_slices, m.aa.off, m.aa.deg, m.aa_edge, m.aa_C, m.aa_xy,
                                                                              m.aa_Nx, m.aa_Ny, m.aa_Nx0, m.aa_Ny0);
      UFM_CUDA(cudaGetLastError());
      h->cnt.kernel_launches++;
      UFM_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (has_tri) {
      const int nT = d->nTri, ldT = d->ldTri ? d->ldTri : nT;
      std::vector<TriRec> tr(nT);
      int bad_tri = 0;
#pragma omp parallel for schedule(static)
      for (int t = 0; t < nT; t++) {
        TriRec q;
        int v[3];
        bool ok = true;
        for (int k = 0; k < 3; k++) {
          v[k] = F2(d->Tri, t + 1, k + 1, ldT);
          if (v[k] < 1 || v[k] > N) { ok = false; v[k] = 1; }
          q.v[k] = aa_r2d[v[k] - 1];
          q.nx[k] = F2(d->NxTri, t + 1, k + 1, ldT);
          q.ny[k] = F2(d->NyTri, t + 1, k + 1, ldT);
        }
        if (!ok) bad_tri = t + 1;
        q.ax = F2(d->V, v[0], 1, ldV); q.ay = F2(d->V, v[0], 2, ldV);
        q.bx = F2(d->V, v[1], 1, ldV); q.by = F2(d->V, v[1], 2, ldV);
        q.cx = F2(d->V, v[2], 1, ldV); q.cy = F2(d->V, v[2], 2, ldV);
        q.pad = 0;
        tr[t] = q;
      }
      if (bad_tri) return ufm_set_error(-2, "ufm_mesh_upload: Tri(%d,:) out of range", bad_tri);
      UP(iT, m.aa_iTri); UP(Rr, m.aa_R); UP(tr, m.tri);
    }
  }

  lap("Aa (+ triangle) upload");
  // ---- Ac arrays ----
  {
    std::vector<int4> aci(m.nAcp, make_int4(0, 0, 0, 0));
    const bool derive_ac = !d->Nx_Ac || !d->Ny_Ac || !d->No_Ac || !d->Np_Ac;   // derived on the device below
    std::vector<double> c4[3][4], np(derive_ac ? 0 : m.nAcp, 0.0), cw(m.nAcp, 0.0), dx(m.nAcp, 1.0), dy(m.nAcp, 0.0), dist2(m.nAcp, 1.0);
    for (int q = 0; q < 3; q++) for (int k = 0; k < 4; k++) c4[q][k].assign(derive_ac ? 0 : m.nAcp, 0.0);
    int bad_ac = 0;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < E; p++) {
      int a = ac_d2r[p] + 1;
      int v[4];
      bool ok = true;
      for (int k = 0; k < 4; k++) {
        v[k] = F2(d->Aci, a, k + 1, ldAc);
        if (v[k] < 1 || v[k] > N) { ok = false; v[k] = 1; }
        if (!derive_ac) { c4[0][k][p] = F2(d->Nx_Ac, a, k + 1, ldAc); c4[1][k][p] = F2(d->Ny_Ac, a, k + 1, ldAc); c4[2][k][p] = F2(d->No_Ac, a, k + 1, ldAc); }
      }
      aci[p] = make_int4(aa_r2d[v[0] - 1], aa_r2d[v[1] - 1], aa_r2d[v[2] - 1], aa_r2d[v[3] - 1]);
      if (!derive_ac) np[p] = d->Np_Ac[a - 1];
      int ci = 0;
      for (int c = 1; c <= d->nC[v[0] - 1]; c++) if (F2(d->C, v[0], c, ldV) == v[1]) { ci = c; break; }
      if (!ci || !ok) { bad_ac = a; continue; }
      cw[p] = F2(d->Cw, v[0], ci, ldV);
      dx[p] = F2(d->V, v[1], 1, ldV) - F2(d->V, v[0], 1, ldV);
      dy[p] = F2(d->V, v[1], 2, ldV) - F2(d->V, v[0], 2, ldV);
      const double dist = std::sqrt(dx[p] * dx[p] + dy[p] * dy[p]);
      dist2[p] = dist * dist;   // dist**2 of UFEMISM_main_model.f90:752
    }
    if (bad_ac) return ufm_set_error(-2, "ufm_mesh_upload: Aci(%d,:) is out of range or not a connection in C", bad_ac);
    lap("Ac fill");
    UP(aci, m.ac_Aci); UP(cw, m.ac_Cw); UP(dx, m.ac_Dx); UP(dy, m.ac_Dy); UP(dist2, m.ac_dist2);
    if (!derive_ac) {
      UP(np, m.ac_Np);
      for (int k = 0; k < 4; k++) { UP(c4[0][k], m.ac_Nx[k]); UP(c4[1][k], m.ac_Ny[k]); UP(c4[2][k], m.ac_No[k]); }
    } else {
      NfAcArgs a;
      a.nAc = E; a.Aci = m.ac_Aci; a.xy = m.aa_xy; a.edge = m.aa_edge;
      const size_t bytes = (size_t)m.nAcp * sizeof(double);
      int rc_ = ufm_arena_alloc(h, bytes, (void **)&m.ac_Np);
      if (rc_) return rc_;
      UFM_CUDA(cudaMemset(m.ac_Np, 0, bytes));
      for (int k = 0; k < 4; k++) {
        double **q3[] = {&m.ac_Nx[k], &m.ac_Ny[k], &m.ac_No[k]};
        for (double **q : q3) { if ((rc_ = ufm_arena_alloc(h, bytes, (void **)q))) return rc_; UFM_CUDA(cudaMemset(*q, 0, bytes)); }
        a.Nx[k] = m.ac_Nx[k]; a.Ny[k] = m.ac_Ny[k]; a.No[k] = m.ac_No[k];
      }
      a.Np = m.ac_Np;
      UFM_CUDA(cudaDeviceSynchronize());
      k_derive_nf_Ac<<<(E + 255) / 256, 256, 0, h->stream>>>(a);
      UFM_CUDA(cudaGetLastError());
      h->cnt.kernel_launches++;
      UFM_CUDA(cudaStreamSynchronize(h->stream));
    }
  }
  // ---- colour-independent half of the AaAc row order: degree, domain edge, owner strip, partition boundary, and the rows sorted by
  //      (owner, boundary, degree, Morton) -- the colouring below only distributes them over the colour blocks ----
  std::vector<unsigned char> is_edge(M), degv(M);
  {
    int bad_row0 = 0;
#pragma omp parallel for schedule(static)
    for (int ai = 0; ai < M; ai++) {
      is_edge[ai] = ai < N ? (d->edge_index[ai] > 0) : (d->edge_index_Ac[ai - N] > 0);
      const int n = d->nCAaAc[ai];
      degv[ai] = (unsigned char)(n < 0 ? 0 : (n > 255 ? 255 : n));
      if (n < 1 || n > W) bad_row0 = ai + 1;
    }
    if (bad_row0) return ufm_set_error(-2, "ufm_mesh_upload: nCAaAc(%d) = %d out of range", bad_row0, d->nCAaAc[bad_row0 - 1]);
    int bad_c = 0;
#pragma omp parallel for schedule(static)
    for (int ai = 0; ai < M; ai++)
      for (int c = 1; c <= degv[ai]; c++) {
        const int ac = F2(d->CAaAc, ai + 1, c, ldM);
        if (ac < 1 || ac > M) bad_c = ai + 1;
      }
    if (bad_c) return ufm_set_error(-2, "ufm_mesh_upload: CAaAc out of range");
  }
  // owner rank of every AaAc row: x-strips balanced by row count (cf. partition_domain_x_balanced,
  // src/mesh_help_functions_module.f90:1337-1404, which the reference uses for mesh generation)
  const int P = h->part_n;
  m.P = P; m.rank = h->part_rank;
  std::vector<unsigned char> owner(M, 0);
  ufm_partition_owners_impl(X, P, owner);
  // rows that read a row owned by another rank ("boundary" rows of the partition) are swept last in every phase, so
  // that the wait for the peers' pushes of the previous phase hides behind the interior rows
  std::vector<unsigned char> isb(M, 0);
  if (P > 1) {
#pragma omp parallel for schedule(static)
    for (int ai = 0; ai < M; ai++)
      for (int c = 1; c <= degv[ai]; c++)
        if (owner[F2(d->CAaAc, ai + 1, c, ldM) - 1] != owner[ai]) { isb[ai] = 1; break; }
  }
  const char *env_order = getenv("UFM_ROW_ORDER");
  {
    const char *e = getenv("UFM_SOR_DATAFLOW");
    m.df_layout = P == 1 && e && atoi(e) != 0;
  }
  // row order inside the colour blocks: (degree, Morton), or the experimental x-band order (UFM_ROW_ORDER=bands:<n>[:<window>])
  int n_bands = m.df_layout ? 64 : 0, deg_window = 4096;
  if (env_order) {
    if (sscanf(env_order, "bands:%d:%d", &n_bands, &deg_window) < 1) n_bands = m.df_layout ? 64 : 0;
  }
  std::vector<int> m_presorted;
  if (n_bands == 0) ufm_row_presort(M, owner.data(), isb.data(), degv.data(), mort.data(), m_presorted);
  lap("AaAc rows pre-sorted");
  // ---- the five-colouring is needed from here on (ufm_mesh_upload_primary: wait for the thread that computes it) ----
  if (g_ufm_colour_wait) { const int rc_w = (*g_ufm_colour_wait)(); if (rc_w) return rc_w; }
  lap("wait for the five-colouring");
  // ---- AaAc order: colour-major, (degree, Morton) inside a colour, edge block last ----
  std::vector<int> colour(M, 0);
  for (int c = 1; c <= 5; c++) {
    int nc = d->colour_nV[c - 1];
    for (int k = 1; k <= nc; k++) {
      int ai = F2(d->colour_vi, k, c, ldM);
      if (ai < 1 || ai > M) return ufm_set_error(-2, "ufm_mesh_upload: colour_vi out of range");
      colour[ai - 1] = c;
    }
  }
  int bad_row = 0;
#pragma omp parallel for schedule(static)
  for (int ai = 0; ai < M; ai++) if (colour[ai] == 0) bad_row = ai + 1;
  if (bad_row) return ufm_set_error(-2, "ufm_mesh_upload: AaAc vertex %d has no colour", bad_row);
  // the sweep relies on same-coloured vertices being non-adjacent (check_solution, mesh_five_colour_module.f90:318-343)
  int bad_a = 0, bad_b = 0;
#pragma omp parallel for schedule(static)
  for (int ai = 0; ai < M; ai++)
    for (int c = 1; c <= degv[ai]; c++) {
      const int ac = F2(d->CAaAc, ai + 1, c, ldM);
      if (colour[ac - 1] == colour[ai]) { bad_a = ai + 1; bad_b = ac; }
    }
  if (bad_a) return ufm_set_error(-2, "ufm_mesh_upload: invalid five-colouring (vertices %d, %d)", bad_a, bad_b);
  lap("colour + validity checks");
  std::vector<int> m_order;
  auto blk = [&](int ai) { return is_edge[ai] ? 6 : colour[ai]; };
  // single-GPU layout: the colour-5 rows that the Neumann pass reads (non-edge rows adjacent to a domain-edge row) lead
  // their colour block, so that the pass can run inside the fifth colour phase as soon as those few slices are done
  // ... unless the dataflow sweep (k_ssa_sor_df, opt-in: UFM_SOR_DATAFLOW=1; measured slower than the barrier kernel)
  // will run: it orders the rows of every colour block by x-band / Morton alone, so that a row's lower-coloured neighbours sit at
  // about the same relative position of their blocks
  std::vector<unsigned char> late(M, 1);
  int n_adj5 = 0;
  if (P == 1 && !m.df_layout) {
#pragma omp parallel for schedule(static) reduction(+ : n_adj5)
    for (int ai = 0; ai < M; ai++) {
      if (colour[ai] != 5 || is_edge[ai]) continue;
      for (int c = 1; c <= degv[ai]; c++)
        if (is_edge[F2(d->CAaAc, ai + 1, c, ldM) - 1]) { late[ai] = 0; n_adj5++; break; }
    }
  }
  {
    std::vector<unsigned char> blkv(M);
#pragma omp parallel for schedule(static)
    for (int ai = 0; ai < M; ai++) blkv[ai] = (unsigned char)blk(ai);
    if (n_bands == 0) ufm_row_order_from_presorted(M, m_presorted, blkv.data(), owner.data(), isb.data(), late.data(), m_order);
    else ufm_row_order_impl(M, blkv.data(), owner.data(), isb.data(), late.data(), degv.data(), mort.data(), X.data(), n_bands, deg_window, m_order);
    std::vector<int>().swap(m_presorted);
  }
  // layout: blocks 1..5 = colours (swept rows), 6 = domain-edge rows; inside a block one group per (owner, interior / boundary), every
  // group padded to a whole chunk.  m_order is sorted by group, so a group's rows are one run of it: sizes by a parallel count, starts by a
  // prefix sum, then every row is placed independently.
  std::vector<int> m_r2d(M), m_d2r;
  {
    const int NG = 6 * P * 2;
    auto grp = [&](int ai) { return ((blk(ai) - 1) * P + owner[ai]) * 2 + (isb[ai] != 0); };
    std::vector<long long> gcnt(NG, 0);
    {
      const int T = std::max(1, omp_get_max_threads());
      std::vector<long long> part((size_t)T * NG, 0);
#pragma omp parallel for schedule(static, 1)
      for (int t = 0; t < T; t++) {
        const size_t lo = (size_t)M * t / T, hi = (size_t)M * (t + 1) / T;
        long long *c = part.data() + (size_t)t * NG;
        for (size_t k = lo; k < hi; k++) c[grp(m_order[k])]++;
      }
      for (int t = 0; t < T; t++) for (int q = 0; q < NG; q++) gcnt[q] += part[(size_t)t * NG + q];
    }
    std::vector<long long> first_k(NG + 1, 0), start(NG, 0);
    long long size = 0;
    for (int b = 1; b <= 6; b++)
      for (int r = 0; r < P; r++) {
        m.rng[b - 1][r][0] = (int)(size / UFM_SLICE);
        for (int bd = 0; bd < 2; bd++) {
          const int q = ((b - 1) * P + r) * 2 + bd;
          if (bd == 1) m.rng[b - 1][r][1] = (int)(size / UFM_SLICE);
          start[q] = size; first_k[q + 1] = first_k[q] + gcnt[q];
          size += gcnt[q];
          if (bd == 1) m.rng[b - 1][r][2] = (int)((size + UFM_SLICE - 1) / UFM_SLICE);
          size = (size + UFM_CHUNK - 1) / UFM_CHUNK * UFM_CHUNK;
        }
      }
    if (first_k[NG] != M) return ufm_set_error(-2, "ufm_mesh_upload: internal error in the AaAc row layout");
    m_d2r.assign((size_t)size, -1);
    bool sorted_ok = true;
#pragma omp parallel for schedule(static) reduction(&& : sorted_ok)
    for (int k = 0; k < M; k++) {
      const int ai = m_order[k], q = grp(ai);
      if (k < first_k[q] || k >= first_k[q + 1]) { sorted_ok = false; continue; }   // m_order not grouped: cannot happen
      const long long pos = start[q] + (k - first_k[q]);
      m_r2d[ai] = (int)pos; m_d2r[(size_t)pos] = ai;
    }
    if (!sorted_ok) return ufm_set_error(-2, "ufm_mesh_upload: internal error in the AaAc row order");
  }
  m.Mp = (int)m_d2r.size();
  m.n_chunks = m.Mp / UFM_CHUNK;
  m.adj5_end = m.rng[4][0][0] + (n_adj5 + UFM_SLICE - 1) / UFM_SLICE;

  lap("AaAc sort + layout");
  // ---- AaAc sliced ELL ----
  {
    std::vector<unsigned char> deg(m.Mp, UFM_DEG_PAD);
#pragma omp parallel for schedule(static)
    for (int p = 0; p < m.Mp; p++) if (m_d2r[p] >= 0) deg[p] = degv[m_d2r[p]];
    std::vector<long long> off;
    build_slices(deg, off);
    m.m.n_rows = m.Mp; m.m.n_slices = m.Mp / UFM_SLICE; m.m.n_entries = off.back();
    size_t ne = (size_t)off.back();
    std::vector<int> idx(ne);
    // no neighbour functions from the host: they are derived on the device below (k_derive_nf_AaAc)
    const bool derive_nf = !d->Nx_AaAc || !d->Ny_AaAc || !d->Nxx_AaAc || !d->Nxy_AaAc || !d->Nyy_AaAc;
    const size_t ne_h = derive_nf ? 0 : ne, np_h = derive_nf ? 0 : (size_t)m.Mp;
    std::vector<double> cU(ne_h, 0.0), cV(ne_h, 0.0), nxy(ne_h, 0.0), nx(ne_h, 0.0), ny(ne_h, 0.0);
    std::vector<double> nxy0(np_h, 0.0), nxysum(np_h, 0.0), nx0(np_h, 0.0), ny0(np_h, 0.0), cU0(np_h, 0.0), cV0(np_h, 0.0);
    std::vector<int> src(m.Mp, INT_MIN);
    double bytes = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bytes)
    for (int s = 0; s < m.m.n_slices; s++) {
      int w = (int)((off[s + 1] - off[s]) / UFM_SLICE);
      for (int l = 0; l < UFM_SLICE; l++) {
        int p = s * UFM_SLICE + l, ai = m_d2r[p];
        for (int c = 0; c < w; c++) idx[(size_t)off[s] + (size_t)c * UFM_SLICE + l] = p;  // padding entries point home
        if (ai < 0) continue;
        int n = degv[ai];
        src[p] = ai < N ? aa_r2d[ai] : ~ac_r2d[ai - N];
        if (!is_edge[ai]) bytes += 80.0 + 20.0 * n;
        if (derive_nf) {
          for (int c = 1; c <= n; c++) idx[(size_t)off[s] + (size_t)(c - 1) * UFM_SLICE + l] = m_r2d[F2(d->CAaAc, ai + 1, c, ldM) - 1];
          continue;
        }
        for (int c = 1; c <= n; c++) {
          size_t e = (size_t)off[s] + (size_t)(c - 1) * UFM_SLICE + l;
          idx[e] = m_r2d[F2(d->CAaAc, ai + 1, c, ldM) - 1];
          double Nxx = F2(d->Nxx_AaAc, ai + 1, c, ldM), Nyy = F2(d->Nyy_AaAc, ai + 1, c, ldM);
          cU[e] = 4.0 * Nxx + Nyy;
          cV[e] = 4.0 * Nyy + Nxx;
          nxy[e] = F2(d->Nxy_AaAc, ai + 1, c, ldM);
          nx[e] = F2(d->Nx_AaAc, ai + 1, c, ldM);
          ny[e] = F2(d->Ny_AaAc, ai + 1, c, ldM);
        }
        double Nxx0 = F2(d->Nxx_AaAc, ai + 1, n + 1, ldM), Nyy0 = F2(d->Nyy_AaAc, ai + 1, n + 1, ldM);
        cU0[p] = 4.0 * Nxx0 + Nyy0;
        cV0[p] = 4.0 * Nyy0 + Nxx0;
        nxy0[p] = F2(d->Nxy_AaAc, ai + 1, n + 1, ldM);
        nx0[p] = F2(d->Nx_AaAc, ai + 1, n + 1, ldM);
        ny0[p] = F2(d->Ny_AaAc, ai + 1, n + 1, ldM);
        double ssum = nxy0[p];
        for (int c = 1; c <= n; c++) ssum = ssum + F2(d->Nxy_AaAc, ai + 1, c, ldM);
        nxysum[p] = ssum;
      }
    }
    m.sor_bytes = bytes;
    lap("AaAc ELL fill");
    // who reads whom across the partition: every row reads its neighbours (sweep, viscosity, Neumann pass); a corner row
    // additionally reads the non-edge neighbours of its edge neighbours (it recomputes their boundary value)
    std::vector<unsigned char> xmask(m.Mp, 0), sowner(m.m.n_slices, 0);
    for (int ai = 0; ai < M && P > 1; ai++) {
      const int r = owner[ai];
      for (int c = 1; c <= degv[ai]; c++) {
        const int ac = F2(d->CAaAc, ai + 1, c, ldM) - 1;
        if (owner[ac] != r) xmask[m_r2d[ac]] |= (unsigned char)(1u << r);
        if (ai < 4 && is_edge[ac])
          for (int c2 = 1; c2 <= degv[ac]; c2++) {
            const int q = F2(d->CAaAc, ac + 1, c2, ldM) - 1;
            if (owner[q] != r) xmask[m_r2d[q]] |= (unsigned char)(1u << r);
          }
      }
    }
    // the ranks this rank exchanges rows with: the readers of its rows and the owners of the rows it reads
    m.nbr_mask = 0;
    for (int ai = 0; ai < M && P > 1; ai++) {
      const unsigned xm = xmask[m_r2d[ai]];
      if (!xm) continue;
      if (owner[ai] == m.rank) m.nbr_mask |= xm;
      else if ((xm >> m.rank) & 1u) m.nbr_mask |= 1u << owner[ai];
    }
#pragma omp parallel for schedule(static)
    for (int sl = 0; sl < m.m.n_slices; sl++)
      for (int l = 0; l < UFM_SLICE; l++) { int ai = m_d2r[sl * UFM_SLICE + l]; if (ai >= 0) { sowner[sl] = owner[ai]; break; } }
    UP(xmask, m.m_xmask); UP(sowner, m.m_sowner);
    UP(off, m.m.off); UP(deg, m.m.deg); UP(idx, m.m_idx); UP(src, m.m_src);
    if (!derive_nf) {
      UP(cU, m.m_cU); UP(cV, m.m_cV); UP(nxy, m.m_nxy); UP(nx, m.m_nx); UP(ny, m.m_ny);
      UP(nxy0, m.m_nxy0); UP(nxysum, m.m_nxysum); UP(nx0, m.m_nx0); UP(ny0, m.m_ny0); UP(cU0, m.m_cU0); UP(cV0, m.m_cV0);
    } else {
      double **big[] = {&m.m_cU, &m.m_cV, &m.m_nxy, &m.m_nx, &m.m_ny};
      double **row[] = {&m.m_nxy0, &m.m_nxysum, &m.m_nx0, &m.m_ny0, &m.m_cU0, &m.m_cV0};
      for (double **q : big) { int rc_ = ufm_arena_alloc(h, std::max<size_t>(ne, 1) * sizeof(double), (void **)q); if (rc_) return rc_; }
      for (double **q : row) { int rc_ = ufm_arena_alloc(h, (size_t)m.Mp * sizeof(double), (void **)q); if (rc_) return rc_; }
      std::vector<double2> xy(m.Mp, make_double2(0.0, 0.0));
      std::vector<unsigned char> edge_row(m.Mp, 0);
#pragma omp parallel for schedule(static)
      for (int p = 0; p < m.Mp; p++) {
        const int ai = m_d2r[p];
        if (ai >= 0) { xy[p] = make_double2(X[ai], Y[ai]); edge_row[p] = is_edge[ai]; }
      }
      double2 *xy_dev = nullptr;
      unsigned char *edge_dev = nullptr;
      UP(xy, xy_dev); UP(edge_row, edge_dev);
      NfArgs a;
      a.n_slices = m.m.n_slices; a.off = m.m.off; a.deg = m.m.deg; a.is_edge = edge_dev; a.idx = m.m_idx; a.xy = xy_dev;
      a.cU = m.m_cU; a.cV = m.m_cV; a.nxy = m.m_nxy; a.nx = m.m_nx; a.ny = m.m_ny;
      a.cU0 = m.m_cU0; a.cV0 = m.m_cV0; a.nxy0 = m.m_nxy0; a.nx0 = m.m_nx0; a.ny0 = m.m_ny0; a.nxysum = m.m_nxysum;
      k_derive_nf_AaAc<<<(m.m.n_slices * 32 + 127) / 128, 128, 0, h->stream>>>(a);
      UFM_CUDA(cudaGetLastError());
      h->cnt.kernel_launches++;
      UFM_CUDA(cudaStreamSynchronize(h->stream));
    }
    std::vector<int> aa2m(m.nVp, 0), ac2m(m.nAcp, 0);
#pragma omp parallel for schedule(static)
    for (int v = 0; v < N; v++) aa2m[aa_r2d[v]] = m_r2d[v];
#pragma omp parallel for schedule(static)
    for (int a = 0; a < E; a++) ac2m[ac_r2d[a]] = m_r2d[N + a];
    UP(aa2m, m.aa2m); UP(ac2m, m.ac2m);
  }

  lap("AaAc upload");
  // ---- Neumann boundary lists (apply_Neumann_boundary_AaAc, mesh_ArakawaC_module.f90:660-724) ----
  {
    std::vector<int> bc_pos, bc_ptr(1, 0), bc_nbr, row_of(M, -1);
    for (int r = 0; r < P; r++) {
      m.bc_rng[r] = (int)bc_pos.size();
      for (int ai = 4; ai < M; ai++) {  // ai = MAX(5,..) .. in 1-based terms
        if (!is_edge[ai] || owner[ai] != r) continue;
        row_of[ai] = (int)bc_pos.size();
        bc_pos.push_back(m_r2d[ai]);
        for (int c = 1; c <= degv[ai]; c++) {
          int ac = F2(d->CAaAc, ai + 1, c, ldM) - 1;
          if (is_edge[ac]) continue;
          bc_nbr.push_back(m_r2d[ac]);
        }
        bc_ptr.push_back((int)bc_nbr.size());
      }
    }
    m.bc_rng[P] = (int)bc_pos.size();
    for (int k = 0; k < 4; k++) m.corner_owner[k] = owner[k];
    m.n_bc = (int)bc_pos.size();
    std::vector<int> cn(4 * 16, 0), cr(4 * 16, -1);
    for (int k = 0; k < 4; k++) {
      int ai = k;
      m.corner_pos[k] = m_r2d[ai];
      m.corner_n[k] = degv[ai];
      if (degv[ai] > 16) return ufm_set_error(-2, "ufm_mesh_upload: corner vertex with more than 16 connections");
      for (int c = 1; c <= degv[ai]; c++) {
        int ac = F2(d->CAaAc, ai + 1, c, ldM) - 1;
        cn[k * 16 + c - 1] = m_r2d[ac];
        cr[k * 16 + c - 1] = (ac >= 4 && is_edge[ac]) ? row_of[ac] : -1;
        if (ac < 4) return ufm_set_error(-2, "ufm_mesh_upload: corner vertices adjacent on the AaAc mesh");
      }
    }
    std::vector<int> rngv((size_t)6 * P * 3), cornv(8);
    for (int b = 0; b < 6; b++) for (int r = 0; r < P; r++) for (int q = 0; q < 3; q++) rngv[((size_t)b * P + r) * 3 + q] = m.rng[b][r][q];
    for (int k = 0; k < 4; k++) { cornv[k] = m.corner_pos[k]; cornv[4 + k] = m.corner_n[k]; }
    std::vector<int> rnga(18);
    for (int b = 0; b < 6; b++) { rnga[3 * b] = m.rng[b][0][0]; rnga[3 * b + 1] = rnga[3 * b + 2] = (b < 5 ? m.rng[b + 1][0][0] : m.Mp / UFM_SLICE); }
    UP(rngv, m.rng_dev); UP(rnga, m.rng_all_dev); UP(cornv, m.corner_dev);
    if (m.df_layout) {
      int rc_;
      if ((rc_ = ufm_arena_alloc(h, sizeof(unsigned short) * (size_t)m.m.n_slices, (void **)&m.df_need))) return rc_;
      if ((rc_ = ufm_arena_alloc(h, sizeof(unsigned short) * ((size_t)m.n_bc + 4), (void **)&m.df_need_bc))) return rc_;
      if ((rc_ = ufm_arena_alloc(h, sizeof(unsigned) * DF_CNT_STRIDE * DF_MAX_STAGES, (void **)&m.df_stage_cnt))) return rc_;
    }
    UP(bc_pos, m.bc_pos); UP(bc_ptr, m.bc_ptr); UP(bc_nbr, m.bc_nbr); UP(cn, m.corner_nbr); UP(cr, m.corner_row);
  }

  lap("boundary lists");
  UP(aa_r2d, m.aa_ref2dev); UP(aa_d2r, m.aa_dev2ref); UP(ac_r2d, m.ac_ref2dev); UP(ac_d2r, m.ac_dev2ref);
  UP(m_r2d, m.m_ref2dev); UP(m_d2r, m.m_dev2ref);

  lap("Ac + permutation upload");
  // ---- per-step kernels partitioned by owner: owner of every Aa / Ac element, where stage 1 of the geometry has to run
  //      redundantly (the vertices an owned element reads), and the halo lists.  Every rank derives every list from the same owner
  //      array, so a receive list is the peer's send list by construction. ----
  {
    const int b_ = h->P.benchmark;
    const char *e = getenv("UFM_PARTITION_STEP");
    // experiments with a column model on the device (update_ice_temperature / solve_SIA_3D on the thermodynamics timer) keep the
    // replicated per-step kernels: their 3-D fields are not exchanged
    m.part_step = P > 1 && !(e && atoi(e) == 0) && b_ != UFM_BM_NONE && !(b_ >= UFM_BM_EISMINT_1 && b_ <= UFM_BM_EISMINT_6) && !m.has_tri;
    if (h->owner_ref) { delete h->owner_ref; h->owner_ref = nullptr; }
    if (P > 1) h->owner_ref = new std::vector<unsigned char>(owner);
  }
  if (m.part_step) {
    const int me = m.rank;
    std::vector<unsigned char> own_aa(m.nVp, 255), own_ac(m.nAcp, 255), act1(m.nVp, 0);
    for (int v = 0; v < N; v++) own_aa[aa_r2d[v]] = owner[v];
    for (int a = 0; a < E; a++) own_ac[ac_r2d[a]] = owner[N + a];
    // reads[q] marks (reference index) the Aa vertices / Ac vertices that rank q reads: neighbours of its vertices, the four vertices of
    // its staggered vertices; the staggered vertices around its vertices
    std::vector<std::vector<int>> sa(P), ra(P), sc(P), rc(P);
    std::vector<unsigned char> rd_aa, rd_ac;
    ufm_partition_reads_impl(d, owner, P, rd_aa, rd_ac);
    // lists in device-index order (deterministic on every rank)
    for (int p_ = 0; p_ < m.nVp; p_++) {
      const int v = aa_d2r[p_];
      if (v < 0) continue;
      const int o = owner[v];
      if (o == me || rd_aa[(size_t)v * P + me]) act1[p_] = 1;
      for (int q = 0; q < P; q++) {
        if (q == o || !rd_aa[(size_t)v * P + q]) continue;
        if (o == me) sa[q].push_back(p_);
        if (q == me) ra[o].push_back(p_);
      }
    }
    for (int p_ = 0; p_ < m.nAcp; p_++) {
      const int a = ac_d2r[p_];
      if (a < 0) continue;
      const int o = owner[N + a];
      for (int q = 0; q < P; q++) {
        if (q == o || !rd_ac[(size_t)a * P + q]) continue;
        if (o == me) sc[q].push_back(p_);
        if (q == me) rc[o].push_back(p_);
      }
    }
    auto flat = [&](const std::vector<std::vector<int>> &L, int *ptr, std::vector<int> &out) {
      out.clear(); ptr[0] = 0;
      for (int q = 0; q < P; q++) { out.insert(out.end(), L[q].begin(), L[q].end()); ptr[q + 1] = (int)out.size(); }
      if (out.empty()) out.push_back(0);
    };
    std::vector<int> f1, f2, f3, f4;
    flat(sa, m.xa_s_ptr, f1); flat(ra, m.xa_r_ptr, f2); flat(sc, m.xc_s_ptr, f3); flat(rc, m.xc_r_ptr, f4);
    UP(own_aa, m.own_aa); UP(own_ac, m.own_ac); UP(act1, m.act_aa1);
    UP(f1, m.xa_s_idx); UP(f2, m.xa_r_idx); UP(f3, m.xc_s_idx); UP(f4, m.xc_r_idx);
    // one region of the exchange buffer holds the largest message of any ordered pair of ranks: up to 4 arrays per Ac entry
    // (solve_SIA: Up, Ux, Uy, D), 1 per Aa entry -- sized from the global maxima so that every rank uses the same stride
    size_t worst = 1;
    {
      std::vector<size_t> cnt_a((size_t)P * P, 0), cnt_c((size_t)P * P, 0);
      for (int v = 0; v < N; v++) for (int q = 0; q < P; q++) if (q != owner[v] && rd_aa[(size_t)v * P + q]) cnt_a[(size_t)owner[v] * P + q]++;
      for (int a = 0; a < E; a++) for (int q = 0; q < P; q++) if (q != owner[N + a] && rd_ac[(size_t)a * P + q]) cnt_c[(size_t)owner[N + a] * P + q]++;
      for (size_t k = 0; k < cnt_a.size(); k++) worst = std::max(worst, std::max(cnt_a[k], 4 * cnt_c[k]));
    }
    m.x_region = (int)((worst + 31) & ~(size_t)31);
  }
  lap("partition: owners + halo lists");
  // ---- state, zero-filled ----
  DevState &s = h->st;
  const size_t nv = m.nVp, na = m.nAcp, nm = m.Mp, nz = (size_t)h->P.nZ;
  double **aa_d[] = {&s.Hi, &s.Hi_alt, &s.Hb, &s.SL, &s.Hs, &s.dHb_dt, &s.dHi_dt, &s.dHs_dt, &s.dHi_dx, &s.dHi_dy, &s.dHs_dx, &s.dHs_dy,
                     &s.dHs_dx_shelf, &s.dHs_dy_shelf, &s.U_SIA, &s.V_SIA, &s.D_SIA, &s.U_SSA, &s.V_SSA, &s.SMB_year, &s.BMB, &s.thk_factor, &s.thk_smb};
  for (double **p : aa_d) ZE(nv, *p);
  ZE(nv * nz, s.U_3D); ZE(nv * nz, s.V_3D);
  s.realistic_A = (h->P.benchmark == UFM_BM_NONE);
  if (s.realistic_A) { ZE(nv * nz, s.Ti); ZE(nv, s.A_mean); ZE(na, s.A_mean_Ac); ZE(nm, s.Afac); }
  if (m.has_tri) {
    if (!s.Ti) ZE(nv * nz, s.Ti);
    ZE(nv * nz, s.Ti_new); ZE(nv * nz, s.W_3D); ZE(nv, s.GHF); ZE(nv * 12, s.T2m); ZE(nv, s.fric_heat);
  }
  ZE(nv, s.mask_noice); ZE(nv, s.mbits);
  double **ac_d[] = {&s.Hi_Ac, &s.Hb_Ac, &s.SL_Ac, &s.Hs_Ac, &s.dHs_dx_shelf_Ac, &s.dHs_dy_shelf_Ac, &s.D_SIA_Ac, &s.Qabs_GL_Ac, &s.Qp_GL_Ac, &s.thk_flux};
  for (double **p : ac_d) ZE(na, *p);
  for (int k = 0; k < 4; k++) { ZE(na, s.dHi_Ac[k]); ZE(na, s.dHb_Ac[k]); ZE(na, s.dHs_Ac[k]); ZE(na, s.dSL_Ac[k]); ZE(na, s.U_SIA_Ac[k]); ZE(na, s.U_SSA_Ac[k]); }
  ZE(na, s.mbits_Ac);
  lap("state: Aa/Ac arrays");
  ZE_OWN(nm, s.UV, 0); ZE(nm, s.RHS); ZE(nm, s.E); ZE(nm, s.rhsnum); ZE(nm, s.dU); ZE(nm, s.dV);
  ZE(nm, s.eta); ZE(nm, s.N); ZE(nm, s.S); ZE(nm, s.tau_c); ZE(nm, s.phi); ZE(nm, s.Hm); ZE(nm, s.mflag);
  ZE_OWN(2 * (size_t)m.m.n_slices + 2, s.partials, 1); ZE(128, s.ctrl); ZE(64, s.scal); ZE(128, s.red_scratch); ZE_OWN(MAIL_WORDS, s.mail, 2);
  if (m.part_step) ZE_OWN(2 * (size_t)P * (size_t)m.x_region, s.xbuf, 3);
  lap("state: AaAc arrays, IPC bufs");
  if (!h->scal_h_keep) UFM_CUDA(cudaMallocHost((void **)&h->scal_h_keep, 64 * sizeof(double)));
  s.scal_h = h->scal_h_keep;
  lap("state: pinned scalars");

  // staging for permuted upload/download of one field
  size_t need = std::max<size_t>((size_t)M, nv * std::max<size_t>(nz, 12)) * sizeof(double);
  if (h->staging_bytes < need) {
    if (h->staging) cudaFreeHost(h->staging);
    if (h->dev_staging) cudaFree(h->dev_staging);
    UFM_CUDA(cudaMallocHost(&h->staging, need));
    UFM_CUDA(cudaMalloc(&h->dev_staging, need));
    h->staging_bytes = h->dev_staging_bytes = need;
  }
  lap("state: staging");
  h->has_mesh = true;
  // single-GPU view of the peer tables: this rank only
  memset(&h->comm, 0, sizeof(h->comm));
  h->comm.P = m.P; h->comm.rank = m.rank;
  h->comm.uv[m.rank] = s.UV; h->comm.partials[m.rank] = s.partials; h->comm.mail[m.rank] = s.mail; h->comm.xbuf[m.rank] = s.xbuf;
  h->comm_connected = false;
  h->xparity = 0;
  h->cnt.sor_bytes_per_iteration = m.sor_bytes;
  const int rc_cfg = ufm_sor_configure(h);
  lap("SOR configure");
  return rc_cfg;
}
