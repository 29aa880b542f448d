/*
 * mesh_host.c -- CPU mesh substrate for the B200 ice-dynamics path.
 *
 * north_star keeps mesh creation/refinement on the CPU; in a real deployment the
 * Fortran host hands these arrays to ufm_mesh_upload().  For tests and benchmarks
 * (no Fortran compiler in this image) this file derives the *secondary* mesh data
 * of UFEMISM's type_mesh from a primary triangulation, following the routines that
 * create_final_mesh_from_merged_submesh calls
 * (reference: src/mesh_creation_module.f90:1724-1737):
 *
 *   connectivity (nC, C, niTri, iTri)        contract of documentation :1202-1214,
 *                                             example src/mesh_creation_module.f90:1898-1960
 *   Voronoi areas A, connection widths Cw    src/mesh_help_functions_module.f90:15-49,101-171
 *   staggered Ac mesh + operators            src/mesh_ArakawaC_module.f90:24-235
 *   Ac edge indices                          src/mesh_ArakawaC_module.f90:236-285
 *   combined AaAc mesh                       src/mesh_ArakawaC_module.f90:286-540
 *   neighbour functions (averaged gradient)  src/mesh_derivatives_module.f90:76-313
 *   five-colouring (Williams 1985)           src/mesh_five_colour_module.f90:18-735
 *
 * All arrays use the reference layout: column-major, 1-based indices stored in the
 * arrays, padded ELL rows of width nC_mem (connectivity) / nC_mem+1 (neighbour
 * functions), unused slots = 0.
 *
 * Plain C, no FMA contraction (x86-64 baseline), so coefficient bits follow the
 * operation order written here.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double mh_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

#define I2(a, i, j, ld) (a)[((size_t)((j) - 1)) * (size_t)(ld) + (size_t)((i) - 1)]

/* ------------------------------------------------------------------------------------------
 * 1. Connectivity from a ccw triangle list.
 *    Tri: nTri x 3 column-major, 1-based, counter-clockwise.
 *    For every vertex the neighbours are listed counter-clockwise; for a boundary vertex the
 *    list runs edge-to-edge (first and last entries are its boundary neighbours) and
 *    iTri(vi,k) is the triangle spanned by C(vi,k), C(vi,k+1).
 *    Returns 0, or -1 if a vertex exceeds nC_mem connections, -2 on a broken fan.
 * ------------------------------------------------------------------------------------------ */
int ufm_mesh_connectivity(int nV, int nTri, const int *Tri, int nC_mem,
                          int *nC, int *C, int *niTri, int *iTri)
{
  int *cnt = (int *)calloc((size_t)nV + 2, sizeof(int));
  int *ptr = (int *)calloc((size_t)nV + 2, sizeof(int));
  if (!cnt || !ptr) return -3;
  for (int t = 1; t <= nTri; t++)
    for (int k = 1; k <= 3; k++) cnt[I2(Tri, t, k, nTri)]++;
  ptr[1] = 0;
  for (int v = 1; v <= nV; v++) ptr[v + 1] = ptr[v] + cnt[v];
  int *inc = (int *)malloc(sizeof(int) * (size_t)ptr[nV + 1] * 3); /* (tri, b, c) with a->b->c ccw */
  int *fill = (int *)calloc((size_t)nV + 2, sizeof(int));
  for (int t = 1; t <= nTri; t++) {
    for (int k = 1; k <= 3; k++) {
      int a = I2(Tri, t, k, nTri);
      int b = I2(Tri, t, k % 3 + 1, nTri);
      int c = I2(Tri, t, (k + 1) % 3 + 1, nTri);
      int p = ptr[a] + fill[a]++;
      inc[3 * p + 0] = t; inc[3 * p + 1] = b; inc[3 * p + 2] = c;
    }
  }
  memset(nC, 0, sizeof(int) * (size_t)nV);
  memset(niTri, 0, sizeof(int) * (size_t)nV);
  memset(C, 0, sizeof(int) * (size_t)nV * nC_mem);
  memset(iTri, 0, sizeof(int) * (size_t)nV * nC_mem);
  int rc = 0;
  for (int a = 1; a <= nV && rc == 0; a++) {
    int n = cnt[a];
    if (n == 0) { rc = -2; break; }
    const int *q = inc + 3 * ptr[a];
    /* start: boundary -> the b that is nobody's c; interior -> smallest neighbour index */
    int start = -1;
    for (int i = 0; i < n; i++) {
      int isc = 0;
      for (int j = 0; j < n; j++) if (q[3 * j + 2] == q[3 * i + 1]) { isc = 1; break; }
      if (!isc) { start = i; break; }
    }
    int boundary = (start >= 0);
    if (!boundary) {
      start = 0;
      for (int i = 1; i < n; i++) if (q[3 * i + 1] < q[3 * start + 1]) start = i;
    }
    int nn = boundary ? n + 1 : n;
    if (nn > nC_mem) { rc = -1; break; }
    int cur = start;
    for (int k = 1; k <= n; k++) {
      I2(C, a, k, nV) = q[3 * cur + 1];
      I2(iTri, a, k, nV) = q[3 * cur + 0];
      int nextb = q[3 * cur + 2];
      if (k == n) {
        if (boundary) I2(C, a, n + 1, nV) = nextb;
        break;
      }
      int nxt = -1;
      for (int j = 0; j < n; j++) if (q[3 * j + 1] == nextb) { nxt = j; break; }
      if (nxt < 0) { rc = -2; break; }
      cur = nxt;
    }
    nC[a - 1] = nn;
    niTri[a - 1] = n;
  }
  free(cnt); free(ptr); free(inc); free(fill);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * 2. Triangle circumcentres, Voronoi cell areas, connection widths.
 *    Boundary Voronoi cells are closed along the domain edge; circumcentres that fall outside
 *    the domain are clamped (the generator keeps boundary triangles well shaped so the
 *    reference's crop_circumcenter branch, mesh_help_functions_module.f90:451-487, is inactive).
 * ------------------------------------------------------------------------------------------ */
static void circumcentre(const double *p, const double *q, const double *r, double *cc)
{
  double ax = p[0], ay = p[1], bx = q[0] - ax, by = q[1] - ay, cx = r[0] - ax, cy = r[1] - ay;
  double d = 2.0 * (bx * cy - by * cx);
  double b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
  cc[0] = ax + (cy * b2 - by * c2) / d;
  cc[1] = ay + (bx * c2 - cx * b2) / d;
}

/* NORM2 of a 2-vector as gfortran evaluates it: libgfortran's _gfortran_norm2_r8 (m4/norm2.m4), a scaled sum of squares; differs
 * from sqrt(x*x + y*y) in the last bit for about one pair in three (tests/test_reference_source.py) */
static double norm2_2(double x, double y)
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

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

int ufm_mesh_geometry(int nV, int nTri, int nC_mem, const double *V, const int *Tri,
                      const int *nC, const int *C, const int *niTri, const int *iTri,
                      const int *edge_index, double xmin, double xmax, double ymin, double ymax,
                      double *Tricc, int *Tri_edge_index, double *A, double *Cw)
{
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int t = 1; t <= nTri; t++) {
    double p[3][2];
    int side[3];
    for (int k = 1; k <= 3; k++) {
      int v = I2(Tri, t, k, nTri);
      p[k - 1][0] = I2(V, v, 1, nV); p[k - 1][1] = I2(V, v, 2, nV);
      side[k - 1] = edge_index[v - 1];
    }
    double cc[2];
    circumcentre(p[0], p[1], p[2], cc);
    I2(Tricc, t, 1, nTri) = cc[0]; I2(Tricc, t, 2, nTri) = cc[1];
    /* boundary triangle: two consecutive vertices on the same domain side */
    int tei = 0;
    for (int k = 0; k < 3 && !tei; k++) {
      int a = side[k], b = side[(k + 1) % 3];
      if (!a || !b) continue;
      if ((a == 8 || a == 1 || a == 2) && (b == 8 || b == 1 || b == 2)) tei = 1;
      else if ((a == 2 || a == 3 || a == 4) && (b == 2 || b == 3 || b == 4)) tei = 3;
      else if ((a == 4 || a == 5 || a == 6) && (b == 4 || b == 5 || b == 6)) tei = 5;
      else if ((a == 6 || a == 7 || a == 8) && (b == 6 || b == 7 || b == 8)) tei = 7;
    }
    Tri_edge_index[t - 1] = tei;
  }
  memset(Cw, 0, sizeof(double) * (size_t)nV * nC_mem);
#pragma omp parallel for schedule(static)
  for (int vi = 1; vi <= nV; vi++) {
    double x0 = I2(V, vi, 1, nV), y0 = I2(V, vi, 2, nV);
    int nt = niTri[vi - 1], ei = edge_index[vi - 1];
    double vor[40][2];
    int nv = 0;
    if (nt + 3 > 40) { err = -3; continue; }
    for (int k = 1; k <= nt; k++) {
      int t = I2(iTri, vi, k, nV);
      double cx = clampd(I2(Tricc, t, 1, nTri), xmin, xmax), cy = clampd(I2(Tricc, t, 2, nTri), ymin, ymax);
      if (k == 1 && ei > 0) {
        /* projection of the first circumcentre on the side shared with C(vi,1) */
        int vb = I2(C, vi, 1, nV);
        double xb = I2(V, vb, 1, nV), yb = I2(V, vb, 2, nV);
        if (yb == y0 && (y0 == ymax || y0 == ymin)) { vor[nv][0] = cx; vor[nv][1] = y0; nv++; }
        else if (xb == x0 && (x0 == xmax || x0 == xmin)) { vor[nv][0] = x0; vor[nv][1] = cy; nv++; }
      }
      vor[nv][0] = cx; vor[nv][1] = cy; nv++;
      if (k == nt && ei > 0) {
        int vb = I2(C, vi, nC[vi - 1], nV);
        double xb = I2(V, vb, 1, nV), yb = I2(V, vb, 2, nV);
        if (yb == y0 && (y0 == ymax || y0 == ymin)) { vor[nv][0] = cx; vor[nv][1] = y0; nv++; }
        else if (xb == x0 && (x0 == xmax || x0 == xmin)) { vor[nv][0] = x0; vor[nv][1] = cy; nv++; }
      }
    }
    if (ei == 0) { vor[nv][0] = vor[0][0]; vor[nv][1] = vor[0][1]; nv++; }
    /* fan of triangles (vi, Vor(n-1), Vor(n)); cf. mesh_help_functions_module.f90:33-37 */
    double area = 0.0;
    for (int n = 1; n < nv; n++) {
      double ax = vor[n][0] - x0, ay = vor[n][1] - y0, bx = vor[n - 1][0] - x0, by = vor[n - 1][1] - y0;
      area += fabs(ax * by - ay * bx) / 2.0;
    }
    A[vi - 1] = area;
    /* connection widths; cf. mesh_help_functions_module.f90:116-167 */
    for (int ci = 1; ci <= nC[vi - 1]; ci++) {
      int vj = I2(C, vi, ci, nV), t1 = 0, t2 = 0;
      for (int k = 1; k <= nt; k++) {
        int t = I2(iTri, vi, k, nV), has = 0;
        for (int n = 1; n <= 3; n++) if (I2(Tri, t, n, nTri) == vj) has = 1;
        if (has) { if (!t1) t1 = t; else t2 = t; }
      }
      if (!t1) { err = -1; break; }
      double w;
      if (t2) {
        double dx = I2(Tricc, t1, 1, nTri) - I2(Tricc, t2, 1, nTri), dy = I2(Tricc, t1, 2, nTri) - I2(Tricc, t2, 2, nTri);
        w = norm2_2(dx, dy);   /* norm2(mesh%Tricc(t1,:) - mesh%Tricc(t2,:)) */
      } else {
        int tei = Tri_edge_index[t1 - 1];
        if (tei == 1) w = fmax(0.0, ymax - I2(Tricc, t1, 2, nTri));
        else if (tei == 3) w = fmax(0.0, xmax - I2(Tricc, t1, 1, nTri));
        else if (tei == 5) w = fmax(0.0, I2(Tricc, t1, 2, nTri) - ymin);
        else if (tei == 7) w = fmax(0.0, I2(Tricc, t1, 1, nTri) - xmin);
        else { err = -2; break; }
      }
      I2(Cw, vi, ci, nV) = w;
    }
  }
  return err;
}

/* ------------------------------------------------------------------------------------------
 * 3. Neighbour functions of one vertex -- averaged-gradient approach.
 *    Restates get_neighbour_functions_vertex_gr, src/mesh_derivatives_module.f90:76-313,
 *    expression by expression (evaluation order kept).  V_vc is n x 2 row pairs (x,y).
 *    Outputs are rows of length nC_mem+1 with stride ld (column-major ELL): entry c at [ (c-1)*ld ].
 * ------------------------------------------------------------------------------------------ */
#define NMAX 64
static void neighbour_functions_vertex_gr(const double *V_vi, int n, const double (*V_vc)[2], int is_edge,
                                          int width, size_t ld, double *Nx, double *Ny, double *Nxx, double *Nxy, double *Nyy)
{
  double NxTri[NMAX][3], NyTri[NMAX][3], NxSub[NMAX][3], NySub[NMAX][3];
  for (int c = 0; c < width; c++) { Nx[c * ld] = 0; Ny[c * ld] = 0; Nxx[c * ld] = 0; Nxy[c * ld] = 0; Nyy[c * ld] = 0; }
#define NX(c) Nx[((c) - 1) * ld]
#define NY(c) Ny[((c) - 1) * ld]
#define NXX(c) Nxx[((c) - 1) * ld]
#define NXY(c) Nxy[((c) - 1) * ld]
#define NYY(c) Nyy[((c) - 1) * ld]
  int nTri = is_edge ? n - 1 : n, nSub = is_edge ? n - 2 : n;
  double xi = V_vi[0], yi = V_vi[1];
  for (int ti = 1; ti <= nTri; ti++) {
    int tip1s = ti + 1; if (tip1s > n) tip1s -= n;
    double xt = V_vc[ti - 1][0], yt = V_vc[ti - 1][1], xtp1s = V_vc[tip1s - 1][0], ytp1s = V_vc[tip1s - 1][1];
    double nzt = ((xt - xi) * (ytp1s - yi)) - ((yt - yi) * (xtp1s - xi));
    NxTri[ti - 1][0] = (yt - ytp1s) / nzt; NxTri[ti - 1][1] = (ytp1s - yi) / nzt; NxTri[ti - 1][2] = (yi - yt) / nzt;
    NyTri[ti - 1][0] = (xtp1s - xt) / nzt; NyTri[ti - 1][1] = (xi - xtp1s) / nzt; NyTri[ti - 1][2] = (xt - xi) / nzt;
  }
  for (int si = 1; si <= nSub; si++) {
    int sip1s = si + 1; if (sip1s > n) sip1s -= n;
    int sip2s = sip1s + 1; if (sip2s > n) sip2s -= n;
    double xs = V_vc[si - 1][0], ys = V_vc[si - 1][1];
    double xsp1s = V_vc[sip1s - 1][0], ysp1s = V_vc[sip1s - 1][1];
    double xsp2s = V_vc[sip2s - 1][0], ysp2s = V_vc[sip2s - 1][1];
    double third = 1.0 / 3.0;
    double nzs = (third * (xs + xsp1s - 2 * xi) * (ysp1s + ysp2s - 2 * yi)) -
                 (third * (ys + ysp1s - 2 * yi) * (xsp1s + xsp2s - 2 * xi));
    NxSub[si - 1][0] = (ys - ysp2s) / nzs; NxSub[si - 1][1] = (ysp1s + ysp2s - 2.0 * yi) / nzs; NxSub[si - 1][2] = (2.0 * yi - ys - ysp1s) / nzs;
    NySub[si - 1][0] = (xsp2s - xs) / nzs; NySub[si - 1][1] = (2.0 * xi - xsp1s - xsp2s) / nzs; NySub[si - 1][2] = (xs + xsp1s - 2.0 * xi) / nzs;
  }
  if (!is_edge) {
    double rn = 1.0 / (double)n, sx = 0, sy = 0;
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
    for (int s = 0; s < n; s++) { sNxSub1 += NxSub[s][0]; sNySub1 += NySub[s][0]; }
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
    double rn1 = 1.0 / (double)(n - 1), rn2 = 1.0 / (double)(n - 2), sx = 0, sy = 0;
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
    for (int s = 0; s < n - 2; s++) { sNxSub1 += NxSub[s][0]; sNySub1 += NySub[s][0]; }
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

/* Neighbour functions on a vertex set with connectivity (n, Cn); restates the driver loops of
 * get_neighbour_functions (mesh_derivatives_module.f90:50-72) and of make_combined_AaAc_mesh
 * (mesh_ArakawaC_module.f90:503-537). */
int ufm_mesh_neighbour_functions(int n_vert, int nC_mem, const double *V, const int *nCn, const int *Cn,
                                 const int *is_edge, double *Nx, double *Ny, double *Nxx, double *Nxy, double *Nyy)
{
  int width = nC_mem + 1;
#pragma omp parallel for schedule(static)
  for (int vi = 1; vi <= n_vert; vi++) {
    double V_vi[2] = {I2(V, vi, 1, n_vert), I2(V, vi, 2, n_vert)};
    double V_vc[NMAX][2];
    int n = nCn[vi - 1];
    for (int ci = 1; ci <= n; ci++) {
      int vc = I2(Cn, vi, ci, n_vert);
      V_vc[ci - 1][0] = I2(V, vc, 1, n_vert); V_vc[ci - 1][1] = I2(V, vc, 2, n_vert);
    }
    neighbour_functions_vertex_gr(V_vi, n, (const double (*)[2])V_vc, is_edge[vi - 1] > 0, width, (size_t)n_vert,
                                  Nx + (vi - 1), Ny + (vi - 1), Nxx + (vi - 1), Nxy + (vi - 1), Nyy + (vi - 1));
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * 4. Staggered (Ac) mesh.  Restates make_Ac_mesh + find_Ac_edge_indices,
 *    src/mesh_ArakawaC_module.f90:24-285: numbering walk (vi ascending, ci ascending, skip
 *    connections met from the other end), vl/vr from the ccw triangles, neighbour functions.
 *    Caller allocates for the upper bound nAc_max = 3*nTri; returns nAc (>0) or <0.
 * ------------------------------------------------------------------------------------------ */
static int is_boundary_segment(const int *edge_index, int v1, int v2)
{
  /* src/mesh_help_functions_module.f90 is_boundary_segment: both on the same domain side */
  int a = edge_index[v1 - 1], b = edge_index[v2 - 1];
  if (a == 0 || b == 0) return 0;
  if ((a == 1 || a == 2 || a == 8) && (b == 1 || b == 2 || b == 8)) return 1;
  if ((a == 2 || a == 3 || a == 4) && (b == 2 || b == 3 || b == 4)) return 1;
  if ((a == 4 || a == 5 || a == 6) && (b == 4 || b == 5 || b == 6)) return 1;
  if ((a == 6 || a == 7 || a == 8) && (b == 6 || b == 7 || b == 8)) return 1;
  return 0;
}

int ufm_mesh_make_Ac(int nV, int nTri, int nC_mem, int nAc_max, const double *V, const int *Tri,
                     const int *nC, const int *C, const int *niTri, const int *iTri, const int *edge_index,
                     int *iAci, int *Aci, double *VAc, double *Nx_Ac, double *Ny_Ac, double *Np_Ac, double *No_Ac,
                     int *edge_index_Ac)
{
  /* The reference numbers an edge the first time its walk (vi ascending, ci ascending) meets it and skips edges already
   * numbered from the other end, i.e. every edge is numbered at its lower-index end.  The number of an edge is therefore
   * (edges owned by lower vertices) + (its rank among vi's connections to higher vertices): a prefix sum, after which the
   * vertices can be processed independently with the reference's numbering reproduced exactly. */
  int err = 0;
  const int W = nC_mem;
  int *first = (int *)malloc(sizeof(int) * ((size_t)nV + 2));
  /* Row-major scratch copies of what the loop below reaches through a NEIGHBOUR's index (C, iTri, Tri, V): the reference's arrays are
   * column-major, so one vertex's connection list is spread over nC cache lines nV*4 bytes apart; here it is one line.  The edge
   * numbers are written row-major too (a vertex also writes into its higher neighbours' rows) and transposed into iAci at the end. */
  int *Cr = (int *)malloc(sizeof(int) * (size_t)nV * W), *iTr = (int *)malloc(sizeof(int) * (size_t)nV * W);
  int *Tr = (int *)malloc(sizeof(int) * (size_t)nTri * 3), *iAr = (int *)malloc(sizeof(int) * (size_t)nV * W);
  double *XY = (double *)malloc(sizeof(double) * (size_t)nV * 2);
  if (!first || !Cr || !iTr || !Tr || !iAr || !XY) { free(first); free(Cr); free(iTr); free(Tr); free(iAr); free(XY); return -3; }
#pragma omp parallel for schedule(static)
  for (int vi = 1; vi <= nV; vi++) {
    int k = 0;
    int *cr = Cr + (size_t)(vi - 1) * W, *tr = iTr + (size_t)(vi - 1) * W, *ar = iAr + (size_t)(vi - 1) * W;
    for (int ci = 1; ci <= W; ci++) { cr[ci - 1] = ci <= nC[vi - 1] ? I2(C, vi, ci, nV) : 0; ar[ci - 1] = 0; }
    for (int ci = 1; ci <= W; ci++) tr[ci - 1] = ci <= niTri[vi - 1] ? I2(iTri, vi, ci, nV) : 0;
    for (int ci = 0; ci < nC[vi - 1]; ci++) if (cr[ci] > vi) k++;
    first[vi] = k;
    XY[2 * (size_t)(vi - 1)] = I2(V, vi, 1, nV); XY[2 * (size_t)(vi - 1) + 1] = I2(V, vi, 2, nV);
  }
#pragma omp parallel for schedule(static)
  for (int ti = 1; ti <= nTri; ti++)
    for (int n = 1; n <= 3; n++) Tr[3 * (size_t)(ti - 1) + n - 1] = I2(Tri, ti, n, nTri);
  long long tot = 0;
  for (int vi = 1; vi <= nV; vi++) { int k = first[vi]; first[vi] = (int)tot; tot += k; }
  if (tot > nAc_max) { free(first); free(Cr); free(iTr); free(Tr); free(iAr); free(XY); return -1; }
  const int nAc_total = (int)tot;
#define VX(v) XY[2 * (size_t)((v) - 1)]
#define VY(v) XY[2 * (size_t)((v) - 1) + 1]
#define TRI(t, n) Tr[3 * (size_t)((t) - 1) + (n) - 1]
#pragma omp parallel for schedule(dynamic, 1024)
  for (int vi = 1; vi <= nV; vi++) {
    int nAc = first[vi];
    const int *cr = Cr + (size_t)(vi - 1) * W, *tr = iTr + (size_t)(vi - 1) * W;
    for (int ci = 1; ci <= nC[vi - 1]; ci++) {
      int vj = cr[ci - 1];
      if (vj <= vi) continue;
      nAc++;
      iAr[(size_t)(vi - 1) * W + ci - 1] = nAc;
      I2(VAc, nAc, 1, nAc_max) = (VX(vi) + VX(vj)) / 2.0;
      I2(VAc, nAc, 2, nAc_max) = (VY(vi) + VY(vj)) / 2.0;
      {
        const int *cj_row = Cr + (size_t)(vj - 1) * W;
        for (int cj = 1; cj <= nC[vj - 1]; cj++)
          if (cj_row[cj - 1] == vi) { iAr[(size_t)(vj - 1) * W + cj - 1] = nAc; break; }
      }
      int vl = 0, vr = 0;
      double Nxl[4], Nyl[4], Nxr[4], Nyr[4], Nx[4], Ny[4], Nzl, Nzr;
      if (!is_boundary_segment(edge_index, vi, vj)) {
        for (int iti = 1; iti <= niTri[vi - 1]; iti++) {
          int ti = tr[iti - 1];
          for (int n1 = 1; n1 <= 3; n1++) {
            int n2 = n1 + 1; if (n2 == 4) n2 = 1;
            int n3 = n2 + 1; if (n3 == 4) n3 = 1;
            if (TRI(ti, n1) == vi && TRI(ti, n2) == vj) vl = TRI(ti, n3);
            else if (TRI(ti, n1) == vj && TRI(ti, n2) == vi) vr = TRI(ti, n3);
          }
        }
        if (!vl || !vr) { err = -2; continue; }
        I2(Aci, nAc, 1, nAc_max) = vi; I2(Aci, nAc, 2, nAc_max) = vj; I2(Aci, nAc, 3, nAc_max) = vl; I2(Aci, nAc, 4, nAc_max) = vr;
        if (Nx_Ac && Ny_Ac && No_Ac) {
          Nxl[0] = VY(vl) - VY(vj); Nxl[1] = VY(vi) - VY(vl); Nxl[2] = VY(vj) - VY(vi); Nxl[3] = 0.0;
          Nyl[0] = VX(vj) - VX(vl); Nyl[1] = VX(vl) - VX(vi); Nyl[2] = VX(vi) - VX(vj); Nyl[3] = 0.0;
          Nxr[0] = VY(vj) - VY(vr); Nxr[1] = VY(vr) - VY(vi); Nxr[2] = 0.0; Nxr[3] = VY(vi) - VY(vj);
          Nyr[0] = VX(vr) - VX(vj); Nyr[1] = VX(vi) - VX(vr); Nyr[2] = 0.0; Nyr[3] = VX(vj) - VX(vi);
          Nzl = ((VX(vj) - VX(vi)) * (VY(vl) - VY(vi))) - ((VY(vj) - VY(vi)) * (VX(vl) - VX(vi)));
          Nzr = ((VX(vr) - VX(vi)) * (VY(vj) - VY(vi))) - ((VY(vr) - VY(vi)) * (VX(vj) - VX(vi)));
          for (int k = 0; k < 4; k++) {
            Nx[k] = -((Nxl[k] / Nzl) + (Nxr[k] / Nzr)) / 2.0;
            Ny[k] = -((Nyl[k] / Nzl) + (Nyr[k] / Nzr)) / 2.0;
          }
        }
      } else {
        for (int iti = 1; iti <= niTri[vi - 1]; iti++) {
          int ti = tr[iti - 1];
          for (int n1 = 1; n1 <= 3; n1++) {
            int n2 = n1 + 1; if (n2 == 4) n2 = 1;
            int n3 = n2 + 1; if (n3 == 4) n3 = 1;
            if ((TRI(ti, n1) == vi && TRI(ti, n2) == vj) || (TRI(ti, n1) == vj && TRI(ti, n2) == vi)) vl = TRI(ti, n3);
          }
        }
        if (!vl) { err = -2; continue; }
        I2(Aci, nAc, 1, nAc_max) = vi; I2(Aci, nAc, 2, nAc_max) = vj; I2(Aci, nAc, 3, nAc_max) = vl; I2(Aci, nAc, 4, nAc_max) = 1;
        if (Nx_Ac && Ny_Ac && No_Ac) {
          Nxl[0] = VY(vl) - VY(vj); Nxl[1] = VY(vi) - VY(vl); Nxl[2] = VY(vj) - VY(vi); Nxl[3] = 0.0;
          Nyl[0] = VX(vj) - VX(vl); Nyl[1] = VX(vl) - VX(vi); Nyl[2] = VX(vi) - VX(vj); Nyl[3] = 0.0;
          Nzl = ((VX(vj) - VX(vi)) * (VY(vl) - VY(vi))) - ((VY(vj) - VY(vi)) * (VX(vl) - VX(vi)));
          for (int k = 0; k < 4; k++) { Nx[k] = -Nxl[k] / Nzl; Ny[k] = -Nyl[k] / Nzl; }
        }
      }
      /* the four operator arrays are optional: ufm_mesh_upload_primary lets the device derive them (k_derive_nf_Ac) */
      if (Np_Ac || (Nx_Ac && Ny_Ac && No_Ac)) {
        double Ux = VX(vj) - VX(vi), Uy = VY(vj) - VY(vi), U = sqrt(Ux * Ux + Uy * Uy);
        if (Np_Ac) Np_Ac[nAc - 1] = 1.0 / U;
        for (int k = 0; k < 4 && Nx_Ac && Ny_Ac && No_Ac; k++) {
          I2(Nx_Ac, nAc, k + 1, nAc_max) = Nx[k];
          I2(Ny_Ac, nAc, k + 1, nAc_max) = Ny[k];
          I2(No_Ac, nAc, k + 1, nAc_max) = (Ny[k] * Ux - Nx[k] * Uy) / U;
        }
      }
    }
  }
#undef TRI
  /* edge numbers back into the reference's column-major iAci (columns past nC stay 0) */
  int unnumbered = 0;   /* a connection that its other end does not list (C not symmetric) was never numbered */
#pragma omp parallel for schedule(static) reduction(| : unnumbered)
  for (int vi = 1; vi <= nV; vi++) {
    const int *ar = iAr + (size_t)(vi - 1) * W;
    for (int ci = 1; ci <= W; ci++) { I2(iAci, vi, ci, nV) = ar[ci - 1]; if (ci <= nC[vi - 1] && ar[ci - 1] < 1) unnumbered = 1; }
  }
  free(Cr); free(iTr); free(Tr); free(iAr); free(XY);
  if (unnumbered && !err) err = -2;
  free(first);
  if (err) return err;
  const int nAc = nAc_total;
  /* find_Ac_edge_indices, mesh_ArakawaC_module.f90:236-285 */
#pragma omp parallel for schedule(static)
  for (int aci = 1; aci <= nAc; aci++) {
    int a = edge_index[I2(Aci, aci, 1, nAc_max) - 1], b = edge_index[I2(Aci, aci, 2, nAc_max) - 1], e = 0;
    if ((a == 8 || a == 1 || a == 2) && (b == 8 || b == 1 || b == 2)) e = 1;
    else if ((a == 2 || a == 3 || a == 4) && (b == 2 || b == 3 || b == 4)) e = 3;
    else if ((a == 4 || a == 5 || a == 6) && (b == 4 || b == 5 || b == 6)) e = 5;
    else if ((a == 6 || a == 7 || a == 8) && (b == 6 || b == 7 || b == 8)) e = 7;
    edge_index_Ac[aci - 1] = e;
  }
  return nAc;
#undef VX
#undef VY
}

/* ------------------------------------------------------------------------------------------
 * 5. Combined AaAc mesh connectivity.  Restates make_combined_AaAc_mesh,
 *    src/mesh_ArakawaC_module.f90:286-429 (connectivity part; ldA = leading dim of Aci etc.).
 * ------------------------------------------------------------------------------------------ */
int ufm_mesh_make_AaAc(int nV, int nAc, int ldAc, int nC_mem, const double *V, const double *VAc,
                       const int *nC, const int *C, const int *iAci, const int *Aci, const int *edge_index_Ac,
                       double *VAaAc, int *nCAaAc, int *CAaAc)
{
  int M = nV + nAc, err = 0;
  const int W = nC_mem;
  /* (neighbour vertex, edge number) pairs of every vertex in one row-major scratch row: the Ac loop below looks both up through the
   * index of a THIRD vertex (the far corner of a triangle), which in the reference's column-major C / iAci costs one cache line per
   * connection and array; here it is two lines per vertex */
  double t0_ = mh_now();
  int *CA = (int *)malloc(sizeof(int) * (size_t)nV * W * 2);
  if (!CA) return -3;
#pragma omp parallel for schedule(static)
  for (int vi = 1; vi <= nV; vi++) {
    I2(VAaAc, vi, 1, M) = I2(V, vi, 1, nV); I2(VAaAc, vi, 2, M) = I2(V, vi, 2, nV);
    nCAaAc[vi - 1] = nC[vi - 1];
    int *row = CA + (size_t)(vi - 1) * W * 2;
    for (int ci = 1; ci <= W; ci++) {
      const int in = ci <= nC[vi - 1];
      const int c = in ? I2(C, vi, ci, nV) : 0, ia = in ? I2(iAci, vi, ci, nV) : 0;
      row[2 * (ci - 1)] = c; row[2 * (ci - 1) + 1] = ia;
      I2(CAaAc, vi, ci, M) = in ? ia + nV : 0;
    }
  }
  double t1_ = mh_now();
#define NB(v, ci) CA[((size_t)((v) - 1) * W + (ci) - 1) * 2]
#define EN(v, ci) CA[((size_t)((v) - 1) * W + (ci) - 1) * 2 + 1]
#pragma omp parallel for schedule(static)
  for (int aci = 1; aci <= nAc; aci++) {
    int ai = aci + nV;
    I2(VAaAc, ai, 1, M) = I2(VAc, aci, 1, ldAc); I2(VAaAc, ai, 2, M) = I2(VAc, aci, 2, ldAc);
    int e = edge_index_Ac[aci - 1];
    int out[6] = {0, 0, 0, 0, 0, 0}, n_out = 0;
    if (e > 0) {
      int vi = I2(Aci, aci, 1, ldAc), vj = I2(Aci, aci, 2, ldAc), sw = 0;
      if (e == 1) { if (I2(V, vi, 1, nV) > I2(V, vj, 1, nV)) sw = 1; }
      else if (e == 3) { if (I2(V, vi, 2, nV) < I2(V, vj, 2, nV)) sw = 1; }
      else if (e == 5) { if (I2(V, vi, 1, nV) < I2(V, vj, 1, nV)) sw = 1; }
      else if (e == 7) { if (I2(V, vi, 2, nV) > I2(V, vj, 2, nV)) sw = 1; }
      if (sw) { int t = vi; vi = vj; vj = t; }
      int vk = I2(Aci, aci, 3, ldAc), aci1 = 0, aci2 = 0;
      for (int ci = 1; ci <= nC[vk - 1]; ci++) {
        if (NB(vk, ci) == vi) aci1 = EN(vk, ci);
        else if (NB(vk, ci) == vj) aci2 = EN(vk, ci);
      }
      if (!aci1 || !aci2) { err = -1; nCAaAc[ai - 1] = 0; }
      else { n_out = 4; out[0] = vi; out[1] = aci1 + nV; out[2] = aci2 + nV; out[3] = vj; }
    } else {
      int vi = I2(Aci, aci, 1, ldAc), vj = I2(Aci, aci, 2, ldAc), vl = I2(Aci, aci, 3, ldAc), vr = I2(Aci, aci, 4, ldAc);
      int aci1 = 0, aci2 = 0, aci3 = 0, aci4 = 0;
      for (int ci = 1; ci <= nC[vr - 1]; ci++) {
        if (NB(vr, ci) == vi) aci1 = EN(vr, ci);
        else if (NB(vr, ci) == vj) aci2 = EN(vr, ci);
      }
      for (int ci = 1; ci <= nC[vl - 1]; ci++) {
        if (NB(vl, ci) == vj) aci3 = EN(vl, ci);
        else if (NB(vl, ci) == vi) aci4 = EN(vl, ci);
      }
      if (!aci1 || !aci2 || !aci3 || !aci4) { err = -1; nCAaAc[ai - 1] = 0; }
      else { n_out = 6; out[0] = vi; out[1] = aci1 + nV; out[2] = aci2 + nV; out[3] = vj; out[4] = aci3 + nV; out[5] = aci4 + nV; }
    }
    if (n_out) nCAaAc[ai - 1] = n_out;
    for (int k = 1; k <= W; k++) I2(CAaAc, ai, k, M) = k <= n_out ? out[k - 1] : 0;   /* every column written here: no memset pass over CAaAc */
  }
#undef NB
#undef EN
  double t2_ = mh_now();
  free(CA);
  if (getenv("UFM_UPLOAD_TIMING")) fprintf(stderr, "[make_AaAc] Aa loop %.1f ms, Ac loop %.1f ms, free %.1f ms\n", (t1_ - t0_) * 1e3, (t2_ - t1_) * 1e3, (mh_now() - t2_) * 1e3);
  return err;
}

/* ------------------------------------------------------------------------------------------
 * 6. Five-colouring.  Restates calculate_five_colouring_AaAc (delete-only path; the IDENTIFY
 *    branch of the reference aborts unconditionally, mesh_five_colour_module.f90:430,452, so
 *    reaching it here returns -2).  The reference's array queues (top = last entry, removal
 *    shifts the tail down, :631-639) are order-equivalent to doubly linked lists, used here so
 *    that 10^7-vertex graphs colour in linear time.  Vertex removal from an adjacency row keeps
 *    the remaining order (:377-379), as the stack snapshot S_L depends on it.
 *    Outputs: colour(M), colour_vi(M,5) ascending per colour, colour_nV(5).
 * ------------------------------------------------------------------------------------------ */
typedef struct { int *prev, *next, head, tail, n; } dlist;
static void dl_push(dlist *q, int v) { q->prev[v] = q->tail; q->next[v] = 0; if (q->tail) q->next[q->tail] = v; else q->head = v; q->tail = v; q->n++; }
static void dl_remove(dlist *q, int v)
{
  int p = q->prev[v], n = q->next[v];
  if (p) q->next[p] = n; else q->head = n;
  if (n) q->prev[n] = p; else q->tail = p;
  q->prev[v] = q->next[v] = 0; q->n--;
}

/* `label` (optional): a permutation of 1..M.  The algorithm then runs on the relabelled graph -- vertex v is stored, queued and
 * stacked as label[v-1] -- with the initial queues filled in ORIGINAL index order, so that every decision it takes is the one it
 * would take without relabelling (adjacency rows keep their order; queue and stack discipline do not look at the labels).  The
 * colouring is identical; what changes is where the rows live in memory: with a space-filling-curve labelling the delete loop,
 * which walks from a deleted vertex to its neighbours, stays in cache (reference-ordered meshes are numbered in refinement
 * order, i.e. randomly in space). */
/* bytes of scratch ufm_mesh_five_colouring_ws needs (a caller that colours mesh after mesh keeps the block: a fresh one costs a page
 * fault per 4 KB, most of them on the one thread the algorithm runs on) */
size_t ufm_mesh_five_colouring_ws_bytes(int M, int nC_mem)
{
  const size_t m1 = ((size_t)M + 1 + 15) & ~(size_t)15;
  return sizeof(int) * (m1 * 6 + (size_t)M * nC_mem + (size_t)M * 5) + m1 + 256;
}
int ufm_mesh_five_colouring_ws(int M, int nC_mem, const int *nCAaAc, const int *CAaAc, const int *label,
                               int *colour, int *colour_vi, int *colour_nV, void *ws)
{
  const double tc0_ = mh_now();
  const size_t m1 = ((size_t)M + 1 + 15) & ~(size_t)15;
  int *ip = (int *)(((uintptr_t)ws + 63) & ~(uintptr_t)63);
  dlist Q4 = {0}, Q5 = {0};
  Q4.prev = ip; ip += m1; Q4.next = ip; ip += m1; Q5.prev = ip; ip += m1; Q5.next = ip; ip += m1;   /* these four and inq start at zero */
  int *deg = ip; ip += m1;
  int *S_vi = ip; ip += m1;
  int *L = ip; ip += (size_t)M * nC_mem;                               /* row-major copy: L[(v-1)*nC_mem + c] */
  int *S_L = ip; ip += (size_t)M * 5;                                  /* deleted vertices have deg <= 5 */
  char *inq = (char *)ip;                                              /* 0 none, 4 in Q4, 5 in Q5 */
  {
    int *z = Q4.prev;
    const long long nz = (long long)m1 * 4;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < nz; i++) z[i] = 0;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)m1; i++) inq[i] = 0;
  }
  int rc = 0, Sn = 0, noofvert = M;
#define LAB(v) (label ? label[(v) - 1] : (v))
#pragma omp parallel for schedule(static)
  for (int v = 1; v <= M; v++) {
    const int q = LAB(v), n = nCAaAc[v - 1];
    deg[q] = n;
    int *row = L + (size_t)(q - 1) * nC_mem;
    for (int c = 1; c <= nC_mem; c++) {   /* columns past the degree hold nothing (0 in the reference's array): not read */
      const int w = c <= n ? I2(CAaAc, v, c, M) : 0;
      row[c - 1] = w > 0 ? LAB(w) : 0;
    }
  }
#define CHECK(w) do { int d_ = deg[w]; \
    if (d_ <= 4) { if (inq[w] == 5) { dl_remove(&Q5, w); inq[w] = 0; } if (inq[w] != 4) { dl_push(&Q4, w); inq[w] = 4; } } \
    else if (d_ == 5) { if (inq[w] == 4) { dl_remove(&Q4, w); inq[w] = 0; } if (inq[w] != 5) { dl_push(&Q5, w); inq[w] = 5; } } \
    else { if (inq[w] == 4) { rc = -3; } if (inq[w] == 5) { dl_remove(&Q5, w); inq[w] = 0; } } } while (0)
  /* initial queues: every vertex of degree <= 4 (Q4) / == 5 (Q5) pushed in ORIGINAL index order.  The lists live at the labelled
   * positions, i.e. scattered; pieces of the index range are linked independently and then joined, which gives the lists the serial
   * loop of dl_push calls would give */
  {
    enum { NPIECE = 64 };
    int first4[NPIECE], last4[NPIECE], n4[NPIECE], first5[NPIECE], last5[NPIECE], n5[NPIECE];
#pragma omp parallel for schedule(static, 1)
    for (int pc = 0; pc < NPIECE; pc++) {
      const int lo = (int)((long long)M * pc / NPIECE) + 1, hi = (int)((long long)M * (pc + 1) / NPIECE);
      int f4 = 0, l4 = 0, c4 = 0, f5 = 0, l5 = 0, c5 = 0;
      for (int v0 = lo; v0 <= hi; v0++) {
        const int v = LAB(v0);
        if (deg[v] <= 4) { Q4.prev[v] = l4; if (l4) Q4.next[l4] = v; else f4 = v; l4 = v; c4++; inq[v] = 4; }
        else if (deg[v] == 5) { Q5.prev[v] = l5; if (l5) Q5.next[l5] = v; else f5 = v; l5 = v; c5++; inq[v] = 5; }
      }
      first4[pc] = f4; last4[pc] = l4; n4[pc] = c4; first5[pc] = f5; last5[pc] = l5; n5[pc] = c5;
    }
    for (int pc = 0; pc < NPIECE; pc++) {
      if (n4[pc]) { if (Q4.tail) { Q4.next[Q4.tail] = first4[pc]; Q4.prev[first4[pc]] = Q4.tail; } else Q4.head = first4[pc]; Q4.tail = last4[pc]; Q4.n += n4[pc]; }
      if (n5[pc]) { if (Q5.tail) { Q5.next[Q5.tail] = first5[pc]; Q5.prev[first5[pc]] = Q5.tail; } else Q5.head = first5[pc]; Q5.tail = last5[pc]; Q5.n += n5[pc]; }
    }
  }
  const double tc1_ = mh_now();
  while (noofvert > 5 && rc == 0) {
    if (Q4.n == 0) { rc = -2; break; }                 /* reference: IDENTIFY -> 'beep' + MPI_ABORT */
    int vi = Q4.tail;
    int *Lv = L + (size_t)(vi - 1) * nC_mem;
    for (int ci = 0; ci < deg[vi]; ci++) {   /* the neighbours' rows are scattered in memory: fetch them all before the first is used */
      int wi = Lv[ci];
      __builtin_prefetch(L + (size_t)(wi - 1) * nC_mem, 1);
      __builtin_prefetch(deg + wi, 1); __builtin_prefetch(inq + wi, 1);
      __builtin_prefetch(Q4.prev + wi, 1); __builtin_prefetch(Q4.next + wi, 1);
    }
    for (int ci = 0; ci < deg[vi]; ci++) {
      int wi = Lv[ci];
      int *Lw = L + (size_t)(wi - 1) * nC_mem;
      for (int c2 = 0; c2 < deg[wi]; c2++) {
        if (Lw[c2] == vi) {
          for (int k = c2; k < deg[wi] - 1; k++) Lw[k] = Lw[k + 1];
          Lw[deg[wi] - 1] = 0;
          deg[wi]--;
          CHECK(wi);
          break;
        }
      }
    }
    S_vi[Sn] = vi;
    for (int k = 0; k < 5; k++) S_L[(size_t)Sn * 5 + k] = (k < deg[vi]) ? Lv[k] : 0;
    Sn++;
    if (deg[vi] <= 4) { dl_remove(&Q4, vi); inq[vi] = 0; }
    else if (deg[vi] == 5) { dl_remove(&Q5, vi); inq[vi] = 0; }
    else { rc = -4; break; }
    noofvert--;
  }
  const double tc2_ = mh_now();
  if (rc == 0) {
    {
#pragma omp parallel for schedule(static)
      for (int i = 0; i < M; i++) colour[i] = 0;
    }
    int k = 0;
    for (int v = Q4.head; v && k < 5; v = Q4.next[v]) colour[v - 1] = ++k;     /* colour(Q4(1..5)) = 1..5 */
    if (k != 5 && M >= 5) rc = -5;
    while (Sn > 0 && rc == 0) {
      Sn--;
      int vi = S_vi[Sn], used[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < 5; c++) {
        int ui = S_L[(size_t)Sn * 5 + c];
        if (ui == 0) break;
        int col = colour[ui - 1];
        if (col) used[col] = 1;
      }
      int col = 0;
      for (int c = 1; c <= 5; c++) if (!used[c]) { col = c; break; }
      if (!col) { rc = -6; break; }
      colour[vi - 1] = col;
    }
  }
  const double tc3_ = mh_now();
  if (rc == 0 && label) {   /* back to the caller's numbering; S_vi doubles as scratch */
#pragma omp parallel for schedule(static)
    for (int v = 1; v <= M; v++) S_vi[v] = colour[label[v - 1] - 1];
#pragma omp parallel for schedule(static)
    for (int v = 1; v <= M; v++) colour[v - 1] = S_vi[v];
  }
#undef LAB
  if (rc == 0) {
    /* check_solution, :318-343 */
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int v = 1; v <= M; v++) {
      if (colour[v - 1] == 0) bad |= 1;
      for (int c = 1; c <= nCAaAc[v - 1]; c++) if (colour[I2(CAaAc, v, c, M) - 1] == colour[v - 1]) { bad |= 1; break; }
    }
    if (bad) rc = -7;
  }
  const double tc4_ = mh_now();
  if (rc == 0) {
    {
      const long long nz = (long long)M * 5;
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < nz; i++) colour_vi[i] = 0;
    }
    for (int c = 0; c < 5; c++) colour_nV[c] = 0;
    for (int v = 1; v <= M; v++) {
      int c = colour[v - 1];
      colour_nV[c - 1]++;
      I2(colour_vi, colour_nV[c - 1], c, M) = v;
    }
  }
  const double tc5_ = mh_now();
  if (getenv("UFM_UPLOAD_TIMING")) fprintf(stderr, "[five-colouring] set-up %.1f, delete loop %.1f, colour loop %.1f, relabel + check %.1f, colour_vi %.1f, free %.1f ms\n", (tc1_ - tc0_) * 1e3, (tc2_ - tc1_) * 1e3, (tc3_ - tc2_) * 1e3, (tc4_ - tc3_) * 1e3, (tc5_ - tc4_) * 1e3, (mh_now() - tc5_) * 1e3);
  return rc;
}

int ufm_mesh_five_colouring_labelled(int M, int nC_mem, const int *nCAaAc, const int *CAaAc, const int *label,
                                     int *colour, int *colour_vi, int *colour_nV)
{
  void *ws = malloc(ufm_mesh_five_colouring_ws_bytes(M, nC_mem));
  if (!ws) return -8;
  const int rc = ufm_mesh_five_colouring_ws(M, nC_mem, nCAaAc, CAaAc, label, colour, colour_vi, colour_nV, ws);
  free(ws);
  return rc;
}

int ufm_mesh_five_colouring(int M, int nC_mem, const int *nCAaAc, const int *CAaAc, int *colour, int *colour_vi, int *colour_nV)
{
  return ufm_mesh_five_colouring_labelled(M, nC_mem, nCAaAc, CAaAc, NULL, colour, colour_vi, colour_nV);
}
