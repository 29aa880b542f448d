// ufm_netcdf.cpp -- restart and help_fields files in the reference's own on-disk format.
//
// The reference writes its mesh output through netcdf-fortran with nf90_create(.., IOR(nf90_clobber, nf90_share), ..)
// (src/netcdf_module.f90:517,662), i.e. the NetCDF *classic* format (CDF-1: magic "CDF\x01", big-endian, 32-bit offsets).
// There is no NetCDF library in this image and the format is small, so this file implements it directly: a header
// (dimensions, attributes, variables) followed by the fixed-size variables in definition order and then the records, each
// record holding one slab of every record variable in definition order (every slab padded to 4 bytes -- never needed here,
// all data are 4- or 8-byte types).  When an offset does not fit 31 bits (meshes of several million vertices) the writer
// switches to the 64-bit-offset variant (CDF-2, magic "CDF\x02"), which every NetCDF library reads transparently.
//
// File layouts reproduced (dimension / variable names, order, types, long_name / units attributes):
//   restart_<region>_0000N.nc      create_restart_file_mesh      src/netcdf_module.f90:489-633
//                                  write_to_restart_file_mesh    src/netcdf_module.f90:180-214
//   help_fields_<region>_0000N.nc  create_help_fields_file_mesh  src/netcdf_module.f90:634-820 (+ create_help_field_mesh :821-1040)
//                                  write_to_help_fields_file_mesh / write_help_field_mesh  :216-487
// and read back by
//   inquire_restart_file_mesh :3012-3049, read_restart_file_mesh :3104-3131,
//   inquire_restart_file_init :3050-3103, read_restart_file_init :3132-3185 (nearest time frame to C%time_to_restart_from).
//
// Fortran dimension order is fastest-first, NetCDF file order slowest-first: a Fortran (vi, zeta, time) variable is the
// file variable (time, zeta, vi), and a column-major (nV, nZ) array is exactly one record slab of it -- arrays cross this
// boundary without transposition.
//
// Host-only code (no kernels); the two entry points that take a handle move fields with ufm_state_download / ufm_state_upload.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <initializer_list>
#include <string>
#include <vector>

#include "../../include/ufemism_b200.h"

int ufm_set_error(int rc, const char *fmt, ...);

namespace {

enum { T_BYTE = 1, T_CHAR = 2, T_SHORT = 3, T_INT = 4, T_FLOAT = 5, T_DOUBLE = 6 };
enum { TAG_DIM = 0x0A, TAG_VAR = 0x0B, TAG_ATT = 0x0C };
const double FILL_DOUBLE = 9.9692099683868690e+36;   // NC_FILL_DOUBLE
const int FILL_INT = -2147483647;                    // NC_FILL_INT

int type_size(int t) { return t == T_DOUBLE ? 8 : (t == T_INT || t == T_FLOAT) ? 4 : t == T_SHORT ? 2 : 1; }
long long pad4(long long n) { return (n + 3) & ~3LL; }

struct Dim { std::string name; long long len; };   // len 0 = the record dimension
struct Att { std::string name; int type; long long nelems; std::vector<unsigned char> raw; };   // raw = big-endian values, unpadded
struct Var {
  std::string name;
  std::vector<int> dimids;     // file order (slowest first)
  std::vector<Att> atts;
  int type = T_DOUBLE;
  bool is_rec = false;
  long long slab_elems = 0;    // elements of the whole variable (fixed) or of one record (record variable)
  long long vsize = 0, begin = 0;
};
struct File {
  int version = 1;
  long long numrecs = 0, recsize = 0, header_bytes = 0;
  int rec_dim = -1;
  std::vector<Dim> dims;
  std::vector<Att> gatts;
  std::vector<Var> vars;
  int dim_id(const char *name) const { for (size_t k = 0; k < dims.size(); k++) if (dims[k].name == name) return (int)k; return -1; }
  int var_id(const char *name) const { for (size_t k = 0; k < vars.size(); k++) if (vars[k].name == name) return (int)k; return -1; }
};

// ---- big-endian serialisation ----
void put32(std::vector<unsigned char> &b, uint32_t v) { for (int s = 24; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
void put64(std::vector<unsigned char> &b, uint64_t v) { for (int s = 56; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
void put_name(std::vector<unsigned char> &b, const std::string &s)
{
  put32(b, (uint32_t)s.size());
  b.insert(b.end(), s.begin(), s.end());
  while (b.size() & 3) b.push_back(0);
}
void put_atts(std::vector<unsigned char> &b, const std::vector<Att> &atts)
{
  if (atts.empty()) { put32(b, 0); put32(b, 0); return; }
  put32(b, TAG_ATT); put32(b, (uint32_t)atts.size());
  for (const Att &a : atts) {
    put_name(b, a.name);
    put32(b, (uint32_t)a.type); put32(b, (uint32_t)a.nelems);
    b.insert(b.end(), a.raw.begin(), a.raw.end());
    while (b.size() & 3) b.push_back(0);
  }
}
void serialise_header(const File &f, std::vector<unsigned char> &b)
{
  b.clear();
  b.push_back('C'); b.push_back('D'); b.push_back('F'); b.push_back((unsigned char)f.version);
  put32(b, (uint32_t)f.numrecs);
  if (f.dims.empty()) { put32(b, 0); put32(b, 0); }
  else {
    put32(b, TAG_DIM); put32(b, (uint32_t)f.dims.size());
    for (const Dim &d : f.dims) { put_name(b, d.name); put32(b, (uint32_t)d.len); }
  }
  put_atts(b, f.gatts);
  if (f.vars.empty()) { put32(b, 0); put32(b, 0); }
  else {
    put32(b, TAG_VAR); put32(b, (uint32_t)f.vars.size());
    for (const Var &v : f.vars) {
      put_name(b, v.name);
      put32(b, (uint32_t)v.dimids.size());
      for (int d : v.dimids) put32(b, (uint32_t)d);
      put_atts(b, v.atts);
      put32(b, (uint32_t)v.type);
      put32(b, v.vsize > 0xFFFFFFFCLL ? 0xFFFFFFFFu : (uint32_t)v.vsize);
      if (f.version == 1) put32(b, (uint32_t)v.begin); else put64(b, (uint64_t)v.begin);
    }
  }
}

// sizes, offsets and the format variant; returns 0 or an error
int layout(File &f)
{
  f.rec_dim = -1;
  for (size_t k = 0; k < f.dims.size(); k++)
    if (f.dims[k].len == 0) {
      if (f.rec_dim >= 0) return ufm_set_error(-2, "netcdf: more than one unlimited dimension (\"%s\" and \"%s\"); a dimension of length 0 is an unlimited one",
                                               f.dims[f.rec_dim].name.c_str(), f.dims[k].name.c_str());
      f.rec_dim = (int)k;
    }
  for (Var &v : f.vars) {
    v.is_rec = !v.dimids.empty() && v.dimids[0] == f.rec_dim;
    long long n = 1;
    for (size_t k = v.is_rec ? 1 : 0; k < v.dimids.size(); k++) {
      if (v.dimids[k] == f.rec_dim) return ufm_set_error(-2, "netcdf: variable \"%s\": the unlimited dimension must come first", v.name.c_str());
      n *= f.dims[v.dimids[k]].len;
    }
    v.slab_elems = n;
    v.vsize = pad4(n * type_size(v.type));
  }
  // UFM_NC_FORCE_64BIT_OFFSET: testing aid, writes the CDF-2 variant regardless of size
  for (f.version = getenv("UFM_NC_FORCE_64BIT_OFFSET") ? 2 : 1; f.version <= 2; f.version++) {
    std::vector<unsigned char> hdr;
    serialise_header(f, hdr);
    f.header_bytes = (long long)hdr.size();
    long long off = pad4(f.header_bytes);
    bool fits = true;
    for (Var &v : f.vars) if (!v.is_rec) { v.begin = off; off += v.vsize; if (v.begin > 0x7FFFFFFFLL) fits = false; }
    f.recsize = 0;
    for (Var &v : f.vars) if (v.is_rec) { v.begin = off; off += v.vsize; f.recsize += v.vsize; if (v.begin > 0x7FFFFFFFLL) fits = false; }
    if (fits || f.version == 2) break;
  }
  return 0;
}

// ---- raw file access ----
struct Fd {
  int fd = -1;
  ~Fd() { if (fd >= 0) close(fd); }
};
int write_at(int fd, const void *p, size_t n, long long off)
{
  const char *c = (const char *)p;
  while (n) {
    ssize_t w = pwrite(fd, c, n, (off_t)off);
    if (w < 0) { if (errno == EINTR) continue; return ufm_set_error(-11, "netcdf: write failed: %s", strerror(errno)); }
    c += w; n -= (size_t)w; off += w;
  }
  return 0;
}
int read_at(int fd, void *p, size_t n, long long off)
{
  char *c = (char *)p;
  while (n) {
    ssize_t r = pread(fd, c, n, (off_t)off);
    if (r < 0) { if (errno == EINTR) continue; return ufm_set_error(-11, "netcdf: read failed: %s", strerror(errno)); }
    if (r == 0) return ufm_set_error(-11, "netcdf: file is shorter than its header says");
    c += r; n -= (size_t)r; off += r;
  }
  return 0;
}
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

// `n` elements of `type` (4 or 8 bytes) between host memory and the file at `off`, byte-swapped in 1 MiB pieces.
// src == NULL on write: the NetCDF default fill value.
int put_elems(int fd, int type, long long off, const void *src, long long n)
{
  const int es = type_size(type);
  const long long chunk = (1 << 20) / es;
  std::vector<unsigned char> buf((size_t)std::min(chunk, std::max(n, 1LL)) * es);
  for (long long k = 0; k < n; k += chunk) {
    const long long m = std::min(chunk, n - k);
    if (es == 8) {
      uint64_t *o = (uint64_t *)buf.data();
      uint64_t fill; memcpy(&fill, &FILL_DOUBLE, 8); fill = bswap64(fill);
      const uint64_t *s = src ? (const uint64_t *)src + k : nullptr;
      for (long long i = 0; i < m; i++) o[i] = s ? bswap64(s[i]) : fill;
    } else {
      uint32_t *o = (uint32_t *)buf.data();
      const uint32_t fill = bswap32((uint32_t)FILL_INT);
      const uint32_t *s = src ? (const uint32_t *)src + k : nullptr;
      for (long long i = 0; i < m; i++) o[i] = s ? bswap32(s[i]) : fill;
    }
    int rc = write_at(fd, buf.data(), (size_t)m * es, off + k * es);
    if (rc) return rc;
  }
  return 0;
}
int get_elems(int fd, int type, long long off, void *dst, long long n)
{
  const int es = type_size(type);
  int rc = read_at(fd, dst, (size_t)n * es, off);
  if (rc) return rc;
  if (es == 8) { uint64_t *p = (uint64_t *)dst; for (long long i = 0; i < n; i++) p[i] = bswap64(p[i]); }
  else { uint32_t *p = (uint32_t *)dst; for (long long i = 0; i < n; i++) p[i] = bswap32(p[i]); }
  return 0;
}

// ---- header parser ----
struct Cursor {
  const unsigned char *p; size_t n, at = 0; bool bad = false;
  uint32_t u32() { if (at + 4 > n) { bad = true; return 0; } uint32_t v = ((uint32_t)p[at] << 24) | ((uint32_t)p[at + 1] << 16) | ((uint32_t)p[at + 2] << 8) | p[at + 3]; at += 4; return v; }
  uint64_t u64() { uint64_t hi = u32(); return (hi << 32) | u32(); }
  std::string name() { uint32_t l = u32(); if (bad || at + pad4(l) > n) { bad = true; return ""; } std::string s((const char *)p + at, l); at += (size_t)pad4(l); return s; }
  void skip(long long k) { if (at + (size_t)k > n) bad = true; else at += (size_t)k; }
};
bool parse_atts(Cursor &c, std::vector<Att> &atts)
{
  uint32_t tag = c.u32(), cnt = c.u32();
  if (c.bad) return false;
  if (tag == 0 && cnt == 0) return true;
  if (tag != TAG_ATT) return false;
  for (uint32_t k = 0; k < cnt && !c.bad; k++) {
    Att a;
    a.name = c.name(); a.type = (int)c.u32(); a.nelems = c.u32();
    if (a.type < T_BYTE || a.type > T_DOUBLE) return false;
    long long bytes = a.nelems * type_size(a.type);
    if (c.bad || c.at + (size_t)pad4(bytes) > c.n) return false;
    a.raw.assign(c.p + c.at, c.p + c.at + bytes);
    c.skip(pad4(bytes));
    atts.push_back(a);
  }
  return !c.bad;
}
// 1 = ok, 0 = need more bytes, -1 = not a classic NetCDF file
int parse_header(const unsigned char *p, size_t n, File &f)
{
  Cursor c{p, n};
  if (n < 4) return 0;
  if (p[0] != 'C' || p[1] != 'D' || p[2] != 'F' || (p[3] != 1 && p[3] != 2)) return -1;
  f = File();
  f.version = p[3]; c.at = 4;
  uint32_t nr = c.u32();
  f.numrecs = nr == 0xFFFFFFFFu ? -1 : (long long)nr;
  uint32_t tag = c.u32(), cnt = c.u32();
  if (c.bad) return 0;
  if (!(tag == 0 && cnt == 0)) {
    if (tag != TAG_DIM) return -1;
    for (uint32_t k = 0; k < cnt && !c.bad; k++) { Dim d; d.name = c.name(); d.len = c.u32(); f.dims.push_back(d); }
  }
  if (c.bad) return 0;
  if (!parse_atts(c, f.gatts)) return c.bad ? 0 : -1;
  tag = c.u32(); cnt = c.u32();
  if (c.bad) return 0;
  if (!(tag == 0 && cnt == 0)) {
    if (tag != TAG_VAR) return -1;
    for (uint32_t k = 0; k < cnt && !c.bad; k++) {
      Var v;
      v.name = c.name();
      uint32_t nd = c.u32();
      if (c.bad || nd > 1024) return c.bad ? 0 : -1;
      for (uint32_t q = 0; q < nd; q++) { int id = (int)c.u32(); if (!c.bad && (id < 0 || id >= (int)f.dims.size())) return -1; v.dimids.push_back(id); }
      if (!parse_atts(c, v.atts)) return c.bad ? 0 : -1;
      v.type = (int)c.u32();
      v.vsize = c.u32();
      v.begin = f.version == 1 ? (long long)c.u32() : (long long)c.u64();
      if (!c.bad && (v.type < T_BYTE || v.type > T_DOUBLE)) return -1;
      f.vars.push_back(v);
    }
  }
  if (c.bad) return 0;
  f.header_bytes = (long long)c.at;
  f.rec_dim = -1;
  for (size_t k = 0; k < f.dims.size(); k++) if (f.dims[k].len == 0) f.rec_dim = (int)k;
  f.recsize = 0;
  for (Var &v : f.vars) {
    v.is_rec = !v.dimids.empty() && v.dimids[0] == f.rec_dim;
    long long ne = 1;
    for (size_t k = v.is_rec ? 1 : 0; k < v.dimids.size(); k++) ne *= f.dims[v.dimids[k]].len;
    v.slab_elems = ne;
    v.vsize = pad4(ne * type_size(v.type));   // recomputed: the stored field saturates at 2^32-1
    if (v.is_rec) f.recsize += v.vsize;
  }
  return 1;
}

int open_file(const char *filename, bool writable, Fd &fd, File &f)
{
  if (!filename) return ufm_set_error(-2, "netcdf: NULL filename");
  fd.fd = open(filename, writable ? O_RDWR : O_RDONLY);
  if (fd.fd < 0) return ufm_set_error(-11, "netcdf: cannot open \"%s\": %s", filename, strerror(errno));
  struct stat st;
  if (fstat(fd.fd, &st)) return ufm_set_error(-11, "netcdf: fstat(\"%s\"): %s", filename, strerror(errno));
  std::vector<unsigned char> buf;
  size_t want = 1 << 16;
  for (;;) {
    want = std::min<size_t>(want, (size_t)st.st_size);
    buf.resize(want);
    int rc = want ? read_at(fd.fd, buf.data(), want, 0) : 0;
    if (rc) return rc;
    int ok = parse_header(buf.data(), buf.size(), f);
    if (ok == 1) break;
    if (ok < 0) return ufm_set_error(-12, "netcdf: \"%s\" is not a NetCDF classic / 64-bit-offset file (the reference writes classic files, src/netcdf_module.f90:517)", filename);
    if (want >= (size_t)st.st_size) return ufm_set_error(-12, "netcdf: \"%s\": truncated header", filename);
    want *= 4;
  }
  if (f.numrecs < 0) {   // "streaming" record count: derive it from the file size
    long long first = -1;
    for (const Var &v : f.vars) if (v.is_rec) { first = v.begin; break; }
    f.numrecs = (first >= 0 && f.recsize > 0 && st.st_size > first) ? (st.st_size - first) / f.recsize : 0;
  }
  return 0;
}

Att text_att(const char *name, const char *value)
{
  Att a; a.name = name; a.type = T_CHAR; a.nelems = (long long)strlen(value);
  a.raw.assign((const unsigned char *)value, (const unsigned char *)value + a.nelems);
  return a;
}
// create_double_var / create_int_var (src/netcdf_module.f90:3638-3677); dims in FORTRAN order (fastest first), as the reference lists them
void def_var(File &f, const char *name, int type, std::initializer_list<int> fortran_dims, const char *long_name, const char *units)
{
  Var v; v.name = name; v.type = type;
  v.dimids.assign(fortran_dims.begin(), fortran_dims.end());
  std::reverse(v.dimids.begin(), v.dimids.end());
  if (long_name) v.atts.push_back(text_att("long_name", long_name));
  if (units) v.atts.push_back(text_att("units", units));
  f.vars.push_back(v);
}
int def_dim(File &f, const char *name, long long len) { f.dims.push_back({name, len}); return (int)f.dims.size() - 1; }

int put_fixed(int fd, const File &f, const char *name, const void *data)
{
  const int id = f.var_id(name);
  if (id < 0) return ufm_set_error(-12, "netcdf: no variable \"%s\"", name);
  const Var &v = f.vars[id];
  return put_elems(fd, v.type, v.begin, data, v.slab_elems);
}
int put_record(int fd, const File &f, const char *name, long long rec, const void *data)
{
  const int id = f.var_id(name);
  if (id < 0) return ufm_set_error(-12, "netcdf: no variable \"%s\"", name);
  const Var &v = f.vars[id];
  if (!v.is_rec) return ufm_set_error(-12, "netcdf: variable \"%s\" has no time dimension", name);
  return put_elems(fd, v.type, v.begin + rec * f.recsize, data, v.slab_elems);
}
int set_numrecs(int fd, long long n)
{
  unsigned char b[4] = {(unsigned char)(n >> 24), (unsigned char)(n >> 16), (unsigned char)(n >> 8), (unsigned char)n};
  return write_at(fd, b, 4, 4);
}
// inquire_double_var / inquire_int_var (src/netcdf_module.f90:3694-3795): the variable must exist with this type and these dimensions
int inquire_var(const File &f, const char *name, int type, std::initializer_list<const char *> fortran_dims, const Var **out)
{
  const int id = f.var_id(name);
  if (id < 0) return ufm_set_error(-12, "ERROR: NetCDF: Variable not found concerning: %s", name);
  const Var &v = f.vars[id];
  if (v.type != type) return ufm_set_error(-12, "ERROR: Actual type of variable \"%s\" is not %s.", name, type == T_DOUBLE ? "nf90_DOUBLE" : "nf90_int");
  if (v.dimids.size() != fortran_dims.size())
    return ufm_set_error(-12, "ERROR: Actual number of dimensions(%d) of variable \"%s\": does not match required number of dimensions (%d).",
                         (int)v.dimids.size(), name, (int)fortran_dims.size());
  size_t k = fortran_dims.size();
  for (const char *dn : fortran_dims) {
    --k;
    if (f.dim_id(dn) < 0 || v.dimids[k] != f.dim_id(dn)) return ufm_set_error(-12, "ERROR: Actual dimensions of variable \"%s\" does not match required dimensions.", name);
  }
  if (out) *out = &v;
  return 0;
}
int inquire_dim(const File &f, const char *name, long long *len)
{
  const int id = f.dim_id(name);
  if (id < 0) return ufm_set_error(-12, "ERROR: NetCDF: Invalid dimension ID or name concerning: %s", name);
  *len = id == f.rec_dim ? f.numrecs : f.dims[id].len;
  return 0;
}

// ---- the part of both files that describes the mesh (src/netcdf_module.f90:521-575 and :666-720, identical) ----
struct MeshDims { int vi, ti, ci, aci, ciplusone, two, three, four, vii, ai, tai, zeta, month, time; };
int define_mesh_part(File &f, const ufm_nc_mesh *m, int nZ, MeshDims &d)
{
  if (!m) return ufm_set_error(-2, "netcdf: NULL mesh");
  if (m->nV < 1 || m->nTri < 1 || m->nC_mem < 1 || m->nAc < 1 || m->nVAaAc < 1 || m->nTriAaAc < 1)
    return ufm_set_error(-2, "netcdf: mesh sizes must be positive (nV=%d nTri=%d nC_mem=%d nAc=%d nVAaAc=%d nTriAaAc=%d)", m->nV, m->nTri, m->nC_mem, m->nAc, m->nVAaAc, m->nTriAaAc);
  if (m->nV_transect < 1)
    return ufm_set_error(-2, "netcdf: nV_transect = %d; a NetCDF dimension of length 0 is the unlimited one, and a classic file has only one (the reference aborts with NF90_EUNLIMIT)", m->nV_transect);
  if (nZ < 1) return ufm_set_error(-2, "netcdf: nZ = %d", nZ);
  d.vi = def_dim(f, "vi", m->nV);
  d.ti = def_dim(f, "ti", m->nTri);
  d.ci = def_dim(f, "ci", m->nC_mem);
  d.aci = def_dim(f, "aci", m->nAc);
  d.ciplusone = def_dim(f, "ciplusone", m->nC_mem + 1);
  d.two = def_dim(f, "two", 2);
  d.three = def_dim(f, "three", 3);
  d.four = def_dim(f, "four", 4);
  d.vii = def_dim(f, "vii", m->nV_transect);
  d.ai = def_dim(f, "ai", m->nVAaAc);
  d.tai = def_dim(f, "tai", m->nTriAaAc);
  def_var(f, "V", T_DOUBLE, {d.vi, d.two}, "Vertex coordinates", "m");
  def_var(f, "Tri", T_INT, {d.ti, d.three}, "Vertex indices", nullptr);
  def_var(f, "nC", T_INT, {d.vi}, "Number of connected vertices", nullptr);
  def_var(f, "C", T_INT, {d.vi, d.ci}, "Indices of connected vertices", nullptr);
  def_var(f, "niTri", T_INT, {d.vi}, "Number of inverse triangles", nullptr);
  def_var(f, "iTri", T_INT, {d.vi, d.ci}, "Indices of inverse triangles", nullptr);
  def_var(f, "edge_index", T_INT, {d.vi}, "Edge index", nullptr);
  def_var(f, "Tricc", T_DOUBLE, {d.ti, d.two}, "Triangle circumcenter", "m");
  def_var(f, "TriC", T_INT, {d.ti, d.three}, "Triangle neighbours", nullptr);
  def_var(f, "Tri_edge_index", T_INT, {d.ti}, "Triangle edge index", nullptr);
  def_var(f, "VAc", T_DOUBLE, {d.aci, d.two}, "Staggered vertex coordinates", "m");
  def_var(f, "Aci", T_INT, {d.aci, d.four}, "Staggered to regular vertex indices", nullptr);
  def_var(f, "iAci", T_INT, {d.vi, d.ci}, "Regular to staggered vertex indices", nullptr);
  def_var(f, "VAaAc", T_DOUBLE, {d.ai, d.two}, "Aa/Ac vertex coordinates", "m");
  def_var(f, "TriAaAc", T_INT, {d.tai, d.three}, "Aa/Ac vertex indices", nullptr);
  def_var(f, "A", T_DOUBLE, {d.vi}, "Vertex Voronoi cell area", "m^2");
  def_var(f, "R", T_DOUBLE, {d.vi}, "Vertex resolution", "m");
  def_var(f, "vi_transect", T_INT, {d.vii, d.two}, "Transect vertex pairs", nullptr);
  def_var(f, "w_transect", T_DOUBLE, {d.vii, d.two}, "Transect interpolation weights", nullptr);
  d.zeta = def_dim(f, "zeta", nZ);
  d.month = def_dim(f, "month", 12);
  d.time = def_dim(f, "time", 0);
  def_var(f, "time", T_DOUBLE, {d.time}, "Time", "years");
  def_var(f, "zeta", T_DOUBLE, {d.zeta}, "Vertical scaled coordinate", "unitless (0 = ice surface, 1 = bedrock)");
  def_var(f, "month", T_DOUBLE, {d.month}, "Month", "1-12");
  return 0;
}

// creates the file (refusing to overwrite, :508-512), writes header, mesh data, zeta and month
int create_and_write_mesh(const char *filename, File &f, const ufm_nc_mesh *m, int nZ, const double *zeta)
{
  if (!filename || !zeta) return ufm_set_error(-2, "netcdf: NULL argument");
  int rc = layout(f);
  if (rc) return rc;
  Fd fd;
  fd.fd = open(filename, O_RDWR | O_CREAT | O_EXCL, 0644);
  if (fd.fd < 0) {
    if (errno == EEXIST) return ufm_set_error(-13, "ERROR: %s already exists!", filename);
    return ufm_set_error(-11, "netcdf: cannot create \"%s\": %s", filename, strerror(errno));
  }
  std::vector<unsigned char> hdr;
  serialise_header(f, hdr);
  while (hdr.size() & 3) hdr.push_back(0);
  if ((rc = write_at(fd.fd, hdr.data(), hdr.size(), 0))) return rc;
  const struct { const char *name; const void *p; } fixed[] = {
      {"V", m->V}, {"Tri", m->Tri}, {"nC", m->nC}, {"C", m->C}, {"niTri", m->niTri}, {"iTri", m->iTri}, {"edge_index", m->edge_index},
      {"Tricc", m->Tricc}, {"TriC", m->TriC}, {"Tri_edge_index", m->Tri_edge_index}, {"VAc", m->VAc}, {"Aci", m->Aci}, {"iAci", m->iAci},
      {"VAaAc", m->VAaAc}, {"TriAaAc", m->TriAaAc}, {"A", m->A}, {"R", m->R}, {"vi_transect", m->vi_transect}, {"w_transect", m->w_transect},
      {"zeta", zeta}};
  for (const auto &q : fixed) if ((rc = put_fixed(fd.fd, f, q.name, q.p))) return rc;
  double month[12];
  for (int k = 0; k < 12; k++) month[k] = (double)(k + 1);
  if ((rc = put_fixed(fd.fd, f, "month", month))) return rc;
  // fixed-size variables defined after this point (help fields without a time dimension) start out filled
  bool after = false;
  for (const Var &v : f.vars) {
    if (v.name == "month") { after = true; continue; }
    if (after && !v.is_rec && (rc = put_elems(fd.fd, v.type, v.begin, nullptr, v.slab_elems))) return rc;
  }
  if (fsync(fd.fd)) return ufm_set_error(-11, "netcdf: fsync(\"%s\"): %s", filename, strerror(errno));   // nf90_sync, :629
  return 0;
}

// ---- help fields: name -> type, shape, attributes (create_help_field_mesh, src/netcdf_module.f90:821-1040) and where the
//      data of write_help_field_mesh (:285-487) live when they are resident on the device ----
enum Shape { S_V, S_VT, S_VZT, S_VMT };
enum Source { SRC_HOST = -1, SRC_TI_BASAL = -2, SRC_T2M_YEAR = -3, SRC_PHI_AA = -4, SRC_TAU_AA = -5 };
struct HelpField { const char *name; int type; Shape shape; const char *long_name, *units; int src; };
const HelpField HELP_FIELDS[] = {
    {"lat", T_DOUBLE, S_V, "Latitude", "degrees north", SRC_HOST},
    {"lon", T_DOUBLE, S_V, "Longitude", "degrees east", SRC_HOST},
    {"GHF", T_DOUBLE, S_V, "Geothermal heat flux", "J m^-2 yr^-1", UFM_F_GHF},
    {"Hi", T_DOUBLE, S_VT, "Ice thickness", "m", UFM_F_HI},
    {"Hb", T_DOUBLE, S_VT, "Bedrock elevation", "m w.r.t PD sealevel", UFM_F_HB},
    {"Hs", T_DOUBLE, S_VT, "Surface elevation", "m w.r.t PD sealevel", UFM_F_HS},
    {"SL", T_DOUBLE, S_VT, "Geoid elevation", "m w.r.t PD sealevel", UFM_F_SL},
    {"dHs_dx", T_DOUBLE, S_VT, "Surface slope in x-direction", "m/m", UFM_F_DHS_DX},
    {"dHs_dy", T_DOUBLE, S_VT, "Surface slope in y-direction", "m/m", UFM_F_DHS_DY},
    {"Ti", T_DOUBLE, S_VZT, "Englacial temperature", "K", UFM_F_TI},
    {"Cpi", T_DOUBLE, S_VZT, "Ice heat capacity", "J kg^-1 K^-1", SRC_HOST},
    {"Ki", T_DOUBLE, S_VZT, "Ice thermal conductivity", "J m^-1 K^-1 yr^-1", SRC_HOST},
    {"Ti_basal", T_DOUBLE, S_VT, "Ice basal temperature", "K", SRC_TI_BASAL},
    {"Ti_pmp", T_DOUBLE, S_VZT, "Ice pressure melting point temperature", "K", SRC_HOST},
    {"A_flow", T_DOUBLE, S_VZT, "Ice flow factor", "Pa^-3 y^-1", SRC_HOST},
    {"A_flow_mean", T_DOUBLE, S_VT, "Vertically averaged ice flow factor", "Pa^-3 y^-1", UFM_F_A_FLOW_MEAN},
    {"U_SIA", T_DOUBLE, S_VT, "Vertically averaged SIA ice x-velocity", "m/yr", UFM_F_U_SIA},
    {"V_SIA", T_DOUBLE, S_VT, "Vertically averaged SIA ice y-velocity", "m/yr", UFM_F_V_SIA},
    {"U_SSA", T_DOUBLE, S_VT, "Vertically averaged SSA ice x-velocity", "m/yr", UFM_F_U_SSA},
    {"V_SSA", T_DOUBLE, S_VT, "Vertically averaged SSA ice y-velocity", "m/yr", UFM_F_V_SSA},
    {"U_vav", T_DOUBLE, S_VT, "Vertically averaged ice x-velocity", "m/yr", SRC_HOST},
    {"V_vav", T_DOUBLE, S_VT, "Vertically averaged ice x-velocity", "m/yr", SRC_HOST},   // sic: the reference's long_name says "x"
    {"U_surf", T_DOUBLE, S_VT, "Surface ice x-velocity", "m/yr", SRC_HOST},
    {"V_surf", T_DOUBLE, S_VT, "Surface ice y-velocity", "m/yr", SRC_HOST},
    {"U_base", T_DOUBLE, S_VT, "Basal ice x-velocity", "m/yr", SRC_HOST},
    {"V_base", T_DOUBLE, S_VT, "Basal ice y-velocity", "m/yr", SRC_HOST},
    {"U_3D", T_DOUBLE, S_VZT, "3D ice x-velocity", "m/yr", UFM_F_U_3D},
    {"V_3D", T_DOUBLE, S_VZT, "3D ice y-velocity", "m/yr", UFM_F_V_3D},
    {"W_3D", T_DOUBLE, S_VZT, "3D ice z-velocity", "m/yr", UFM_F_W_3D},
    {"D_SIA", T_DOUBLE, S_VT, "SIA ice diffusivity", nullptr, UFM_F_D_SIA},
    {"D_SIA_3D", T_DOUBLE, S_VZT, "3D SIA ice diffusivity", nullptr, SRC_HOST},
    {"T2m", T_DOUBLE, S_VMT, "Monthly mean 2-m air temperature", "K", UFM_F_T2M},
    {"T2m_year", T_DOUBLE, S_VT, "Annual mean 2-m air temperature", "K", SRC_T2M_YEAR},
    {"Precip", T_DOUBLE, S_VMT, "Monthly total precipitation", "mm", SRC_HOST},
    {"Precip_year", T_DOUBLE, S_VT, "Annual total precipitation", "mm", SRC_HOST},
    {"Wind_WE", T_DOUBLE, S_VMT, "Monthly mean zonal wind", "m/s", SRC_HOST},
    {"Wind_WE_year", T_DOUBLE, S_VT, "Annual mean zonal wind", "m/s", SRC_HOST},
    {"Wind_SN", T_DOUBLE, S_VMT, "Monthly mean meridional wind", "m/s", SRC_HOST},
    {"Wind_SN_year", T_DOUBLE, S_VT, "Annual mean meridional wind", "m/s", SRC_HOST},
    {"SMB", T_DOUBLE, S_VMT, "Monthly surface mass balance", "m ice equivalent", SRC_HOST},
    {"SMB_year", T_DOUBLE, S_VT, "Annual surface mass balance", "m ice equivalent", UFM_F_SMB_YEAR},
    {"BMB_sheet", T_DOUBLE, S_VT, "Annual basal mass balance for grounded ice", "m ice equivalent", SRC_HOST},
    {"BMB_shelf", T_DOUBLE, S_VT, "Annual basal mass balance for floating ice", "m ice equivalent", SRC_HOST},
    {"BMB", T_DOUBLE, S_VT, "Annual basal mass balance", "m ice equivalent", UFM_F_BMB},
    {"Snowfall", T_DOUBLE, S_VMT, "Monthly total snowfall", "m water equivalent", SRC_HOST},
    {"Snowfall_year", T_DOUBLE, S_VT, "Annual total snowfall", "m water equivalent", SRC_HOST},
    {"Rainfall", T_DOUBLE, S_VMT, "Monthly total rainfall", "m water equivalent", SRC_HOST},
    {"Rainfall_year", T_DOUBLE, S_VT, "Annual total rainfall", "m water equivalent", SRC_HOST},
    {"AddedFirn", T_DOUBLE, S_VMT, "Monthly total added firn", "m water equivalent", SRC_HOST},
    {"AddedFirn_year", T_DOUBLE, S_VT, "Annual total added firn", "m water equivalent", SRC_HOST},
    {"Refreezing", T_DOUBLE, S_VMT, "Monthly total refreezing", "m water equivalent", SRC_HOST},
    {"Refreezing_year", T_DOUBLE, S_VT, "Annual total refreezing", "m water equivalent", SRC_HOST},
    {"Runoff", T_DOUBLE, S_VMT, "Monthly total runoff", "m water equivalent", SRC_HOST},
    {"Runoff_year", T_DOUBLE, S_VT, "Annual total runoff", "m water equivalent", SRC_HOST},
    {"Albedo", T_DOUBLE, S_VMT, "Monthly mean albedo", nullptr, SRC_HOST},
    {"Albedo_year", T_DOUBLE, S_VT, "Annual mean albedo", nullptr, SRC_HOST},
    {"FirnDepth", T_DOUBLE, S_VMT, "Monthly mean firn layer depth", "m water equivalent", SRC_HOST},
    {"FirnDepth_year", T_DOUBLE, S_VT, "Annual mean firn layer depth", "m water equivalent", SRC_HOST},
    {"mask", T_INT, S_VT, "mask", nullptr, UFM_F_MASK},
    {"mask_land", T_INT, S_VT, "land mask", nullptr, UFM_F_MASK_LAND},
    {"mask_ocean", T_INT, S_VT, "ocean mask", nullptr, UFM_F_MASK_OCEAN},
    {"mask_lake", T_INT, S_VT, "lake mask", nullptr, UFM_F_MASK_LAKE},
    {"mask_ice", T_INT, S_VT, "ice mask", nullptr, UFM_F_MASK_ICE},
    {"mask_sheet", T_INT, S_VT, "sheet mask", nullptr, UFM_F_MASK_SHEET},
    {"mask_shelf", T_INT, S_VT, "shelf mask", nullptr, UFM_F_MASK_SHELF},
    {"mask_coast", T_INT, S_VT, "coast mask", nullptr, UFM_F_MASK_COAST},
    {"mask_margin", T_INT, S_VT, "margin mask", nullptr, UFM_F_MASK_MARGIN},
    {"mask_gl", T_INT, S_VT, "grounding-line mask", nullptr, UFM_F_MASK_GL},
    {"mask_cf", T_INT, S_VT, "calving-front mask", nullptr, UFM_F_MASK_CF},
    {"phi_fric", T_DOUBLE, S_VT, "till friction angle", "degrees", SRC_PHI_AA},
    {"tau_yield", T_DOUBLE, S_VT, "basal yield stress", "Pa", SRC_TAU_AA},
    {"iso_ice", T_DOUBLE, S_VT, "Vertically averaged ice d18O", "per mille", SRC_HOST},
    {"iso_surf", T_DOUBLE, S_VT, "d18O of precipitation", "per mille", SRC_HOST},
    {"dHb", T_DOUBLE, S_VT, "Change in bedrock elevation w.r.t. PD", "m", SRC_HOST},
};
const HelpField *find_help_field(const char *name)
{
  for (const HelpField &h : HELP_FIELDS) if (!strcmp(h.name, name)) return &h;
  return nullptr;
}
bool skipped_help_field(const char *name) { return !name || !*name || !strcmp(name, "none") || !strcmp(name, "resolution"); }

}  // namespace

extern "C" {

/* create_restart_file_mesh (src/netcdf_module.f90:489-633) */
int ufm_restart_create(const char *filename, const ufm_nc_mesh *mesh, int nZ, const double *zeta)
{
  File f;
  MeshDims d;
  int rc = define_mesh_part(f, mesh, nZ, d);
  if (rc) return rc;
  def_var(f, "Hi", T_DOUBLE, {d.vi, d.time}, "Ice Thickness", "m");
  def_var(f, "Hb", T_DOUBLE, {d.vi, d.time}, "Bedrock Height", "m");
  def_var(f, "Hs", T_DOUBLE, {d.vi, d.time}, "Surface Height", "m");
  def_var(f, "U_SIA", T_DOUBLE, {d.vi, d.time}, "SIA ice x-velocity", "m/yr");
  def_var(f, "V_SIA", T_DOUBLE, {d.vi, d.time}, "SIA ice y-velocity", "m/yr");
  def_var(f, "U_SSA", T_DOUBLE, {d.vi, d.time}, "SSA ice x-velocity", "m/yr");
  def_var(f, "V_SSA", T_DOUBLE, {d.vi, d.time}, "SSA ice y-velocity", "m/yr");
  def_var(f, "Ti", T_DOUBLE, {d.vi, d.zeta, d.time}, "Ice temperature", "K");
  def_var(f, "FirnDepth", T_DOUBLE, {d.vi, d.month, d.time}, "Firn depth", "m");
  def_var(f, "MeltPreviousYear", T_DOUBLE, {d.vi, d.time}, "Melt during previous year", "mie");
  return create_and_write_mesh(filename, f, mesh, nZ, zeta);
}

/* write_to_restart_file_mesh (src/netcdf_module.f90:180-214) from host arrays: appends one time frame; returns its 1-based
 * index (netcdf%ti) or a negative rc.  NULL members are written as the NetCDF fill value. */
int ufm_restart_append(const char *filename, double time, const ufm_restart_frame *fr)
{
  if (!fr) return ufm_set_error(-2, "ufm_restart_append: NULL frame");
  Fd fd; File f;
  int rc = open_file(filename, true, fd, f);
  if (rc) return rc;
  const long long rec = f.numrecs;
  const struct { const char *name; const double *p; } vars[] = {
      {"time", &time}, {"Hi", fr->Hi}, {"Hb", fr->Hb}, {"Hs", fr->Hs}, {"U_SIA", fr->U_SIA}, {"V_SIA", fr->V_SIA}, {"U_SSA", fr->U_SSA},
      {"V_SSA", fr->V_SSA}, {"Ti", fr->Ti}, {"FirnDepth", fr->FirnDepth}, {"MeltPreviousYear", fr->MeltPreviousYear}};
  for (const auto &q : vars) if ((rc = put_record(fd.fd, f, q.name, rec, q.p))) return rc;
  if ((rc = set_numrecs(fd.fd, rec + 1))) return rc;
  return (int)(rec + 1);
}

/* the file must describe the mesh and the vertical grid that are resident on the handle: the frame buffers below are sized from the
 * file's dimensions while ufm_state_download / _upload move mesh.nV x P.nZ elements (a file of the previous mesh, after a mesh update,
 * or with another nZ would otherwise overflow them; the reference stops in NetCDF on such a shape mismatch) */
static int check_file_matches_handle(ufm_handle *h, const char *who, const char *filename, long long nV, long long nZ, long long nM = -1)
{
  int d[5];
  int rc = ufm_resident_dims(h, d);
  if (rc) return rc;
  if (!d[0]) return ufm_set_error(-2, "%s: no mesh resident", who);
  if (nV != d[1])
    return ufm_set_error(-12, "%s: \"%s\" holds %lld vertices, the resident mesh %d (a file of another mesh?)", who, filename, nV, d[1]);
  if (nM >= 0 && nM != d[3])
    return ufm_set_error(-12, "%s: \"%s\" holds %lld combined-mesh vertices, the resident mesh %d", who, filename, nM, d[3]);
  if (nZ != d[4]) return ufm_set_error(-14, "   ERROR: nZ in \"%s\" (%lld) doesnt match nZ in config (%d)!", filename, nZ, d[4]);
  return 0;
}

/* the same with the fields the device owns (Hi, Hb, Hs, U_SIA, V_SIA, U_SSA, V_SSA, Ti) downloaded from the handle;
 * FirnDepth (nV,12) and MeltPreviousYear (nV) belong to the host's SMB model (NULL: fill value) */
int ufm_restart_write(ufm_handle *h, const char *filename, double time, const double *FirnDepth, const double *MeltPreviousYear)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  Fd fd; File f;
  int rc = open_file(filename, true, fd, f);
  if (rc) return rc;
  long long nV = 0, nZ = 0;
  if ((rc = inquire_dim(f, "vi", &nV)) || (rc = inquire_dim(f, "zeta", &nZ))) return rc;
  if ((rc = check_file_matches_handle(h, "ufm_restart_write", filename, nV, nZ))) return rc;
  const long long rec = f.numrecs;
  std::vector<double> buf((size_t)(nV * nZ));
  if ((rc = put_record(fd.fd, f, "time", rec, &time))) return rc;
  const struct { const char *name; int field; } dev[] = {{"Hi", UFM_F_HI}, {"Hb", UFM_F_HB}, {"Hs", UFM_F_HS}, {"U_SIA", UFM_F_U_SIA},
                                                        {"V_SIA", UFM_F_V_SIA}, {"U_SSA", UFM_F_U_SSA}, {"V_SSA", UFM_F_V_SSA}, {"Ti", UFM_F_TI}};
  for (const auto &q : dev) {
    // Ti is resident only for realistic flow factors or thermodynamics meshes; the benchmarks that skip thermodynamics leave
    // the frame's Ti at the fill value
    const bool resident = ufm_field_resident(h, q.field) == 1;
    if (resident && (rc = ufm_state_download(h, q.field, buf.data()))) return rc;
    if ((rc = put_record(fd.fd, f, q.name, rec, resident ? buf.data() : nullptr))) return rc;
  }
  if ((rc = put_record(fd.fd, f, "FirnDepth", rec, FirnDepth))) return rc;
  if ((rc = put_record(fd.fd, f, "MeltPreviousYear", rec, MeltPreviousYear))) return rc;
  if ((rc = set_numrecs(fd.fd, rec + 1))) return rc;
  return (int)(rec + 1);
}

/* inquire_restart_file_mesh (src/netcdf_module.f90:3012-3049) */
int ufm_restart_inquire_mesh(const char *filename, int *nV, int *nTri, int *nC_mem)
{
  if (!nV || !nTri || !nC_mem) return ufm_set_error(-2, "ufm_restart_inquire_mesh: NULL output");
  Fd fd; File f;
  int rc = open_file(filename, false, fd, f);
  if (rc) return rc;
  long long a, b, c, dummy;
  if ((rc = inquire_dim(f, "vi", &a)) || (rc = inquire_dim(f, "ti", &b)) || (rc = inquire_dim(f, "ci", &c)) ||
      (rc = inquire_dim(f, "two", &dummy)) || (rc = inquire_dim(f, "three", &dummy))) return rc;
  if ((rc = inquire_var(f, "V", T_DOUBLE, {"vi", "two"}, nullptr)) || (rc = inquire_var(f, "nC", T_INT, {"vi"}, nullptr)) ||
      (rc = inquire_var(f, "C", T_INT, {"vi", "ci"}, nullptr)) || (rc = inquire_var(f, "niTri", T_INT, {"vi"}, nullptr)) ||
      (rc = inquire_var(f, "iTri", T_INT, {"vi", "ci"}, nullptr)) || (rc = inquire_var(f, "edge_index", T_INT, {"vi"}, nullptr)) ||
      (rc = inquire_var(f, "Tri", T_INT, {"ti", "three"}, nullptr)) || (rc = inquire_var(f, "Tricc", T_DOUBLE, {"ti", "two"}, nullptr)) ||
      (rc = inquire_var(f, "TriC", T_INT, {"ti", "three"}, nullptr)) || (rc = inquire_var(f, "Tri_edge_index", T_INT, {"ti"}, nullptr))) return rc;
  *nV = (int)a; *nTri = (int)b; *nC_mem = (int)c;
  return 0;
}

/* read_restart_file_mesh (src/netcdf_module.f90:3104-3131): the primary mesh data.  Arrays are (nV,..)/(nTri,..) column-major
 * with exactly the sizes ufm_restart_inquire_mesh returned; NULL outputs are skipped. */
int ufm_restart_read_mesh(const char *filename, double *V, int *nC, int *C, int *niTri, int *iTri, int *edge_index, int *Tri, double *Tricc,
                          int *TriC, int *Tri_edge_index)
{
  Fd fd; File f;
  int rc = open_file(filename, false, fd, f);
  if (rc) return rc;
  const struct { const char *name; int type; void *p; } vars[] = {
      {"V", T_DOUBLE, V}, {"nC", T_INT, nC}, {"C", T_INT, C}, {"niTri", T_INT, niTri}, {"iTri", T_INT, iTri}, {"edge_index", T_INT, edge_index},
      {"Tri", T_INT, Tri}, {"Tricc", T_DOUBLE, Tricc}, {"TriC", T_INT, TriC}, {"Tri_edge_index", T_INT, Tri_edge_index}};
  for (const auto &q : vars) {
    if (!q.p) continue;
    const int id = f.var_id(q.name);
    if (id < 0 || f.vars[id].type != q.type || f.vars[id].is_rec) return ufm_set_error(-12, "ufm_restart_read_mesh: variable \"%s\" missing or of the wrong type", q.name);
    if ((rc = get_elems(fd.fd, q.type, f.vars[id].begin, q.p, f.vars[id].slab_elems))) return rc;
  }
  return 0;
}

/* inquire_restart_file_init (src/netcdf_module.f90:3050-3103): nZ must match the configuration (fatal otherwise, :3070-3073);
 * a zeta that differs by more than 1e-4 only warns (rc 1).  *nt = number of time frames. */
int ufm_restart_inquire_init(const char *filename, int nZ, const double *zeta, int *nt)
{
  Fd fd; File f;
  int rc = open_file(filename, false, fd, f);
  if (rc) return rc;
  long long z, t, m;
  if ((rc = inquire_dim(f, "zeta", &z)) || (rc = inquire_dim(f, "time", &t)) || (rc = inquire_dim(f, "month", &m))) return rc;
  if (z != nZ) return ufm_set_error(-14, "   ERROR: nZ in restart file doesnt match nZ in config!");
  const Var *vz = nullptr;
  if ((rc = inquire_var(f, "zeta", T_DOUBLE, {"zeta"}, &vz)) || (rc = inquire_var(f, "time", T_DOUBLE, {"time"}, nullptr)) ||
      (rc = inquire_var(f, "month", T_DOUBLE, {"month"}, nullptr))) return rc;
  if ((rc = inquire_var(f, "Hi", T_DOUBLE, {"vi", "time"}, nullptr)) || (rc = inquire_var(f, "Hb", T_DOUBLE, {"vi", "time"}, nullptr)) ||
      (rc = inquire_var(f, "Hs", T_DOUBLE, {"vi", "time"}, nullptr)) || (rc = inquire_var(f, "Ti", T_DOUBLE, {"vi", "zeta", "time"}, nullptr)) ||
      (rc = inquire_var(f, "U_SSA", T_DOUBLE, {"vi", "time"}, nullptr)) || (rc = inquire_var(f, "V_SSA", T_DOUBLE, {"vi", "time"}, nullptr)) ||
      (rc = inquire_var(f, "MeltPreviousYear", T_DOUBLE, {"vi", "time"}, nullptr)) ||
      (rc = inquire_var(f, "FirnDepth", T_DOUBLE, {"vi", "month", "time"}, nullptr))) return rc;
  if (nt) *nt = (int)t;
  int warn = 0;
  if (zeta) {
    std::vector<double> zf((size_t)z);
    if ((rc = get_elems(fd.fd, T_DOUBLE, vz->begin, zf.data(), z))) return rc;
    for (long long k = 0; k < z; k++) if (fabs(zeta[k] - zf[k]) > 0.0001) warn = 1;
    if (warn) ufm_set_error(1, "  WARNING - vertical coordinate zeta in restart file doesnt match zeta in config!");
  }
  return warn;
}

/* read_restart_file_init (src/netcdf_module.f90:3132-3185): the time frame closest to time_to_restart_from, which must lie
 * inside the file's time range (fatal otherwise, :3154-3157).  NULL outputs are skipped; *ti_out = 1-based frame read. */
int ufm_restart_read_init(const char *filename, double time_to_restart_from, ufm_restart_frame_out *out, int *ti_out)
{
  if (!out) return ufm_set_error(-2, "ufm_restart_read_init: NULL output");
  Fd fd; File f;
  int rc = open_file(filename, false, fd, f);
  if (rc) return rc;
  const int idt = f.var_id("time");
  if (idt < 0 || !f.vars[idt].is_rec || f.vars[idt].type != T_DOUBLE) return ufm_set_error(-12, "ufm_restart_read_init: no time variable");
  const long long nt = f.numrecs;
  if (nt < 1) return ufm_set_error(-15, "ufm_restart_read_init: the restart file holds no time frame");
  std::vector<double> time((size_t)nt);
  for (long long k = 0; k < nt; k++) if ((rc = get_elems(fd.fd, T_DOUBLE, f.vars[idt].begin + k * f.recsize, &time[k], 1))) return rc;
  double tmin = time[0], tmax = time[0];
  for (double t : time) { tmin = std::min(tmin, t); tmax = std::max(tmax, t); }
  if (time_to_restart_from < tmin || time_to_restart_from > tmax)
    return ufm_set_error(-15, "  ERROR - time_to_restart_from %g outside range of restart file! (range = [%g - %g])", time_to_restart_from, tmin, tmax);
  long long ti_min = 0;
  double dt_min = 1e8;
  for (long long ti = 1; ti <= nt; ti++) {
    const double dt = fabs(time[ti - 1] - time_to_restart_from);
    if (dt < dt_min) { ti_min = ti; dt_min = dt; }
  }
  if (ti_min == 0) return ufm_set_error(-15, "ufm_restart_read_init: no time frame within 1e8 yr of time_to_restart_from");   // the reference would read frame 0 and stop in NetCDF
  const struct { const char *name; double *p; } vars[] = {{"Hi", out->Hi}, {"Hb", out->Hb}, {"Hs", out->Hs}, {"Ti", out->Ti}, {"U_SSA", out->U_SSA},
                                                          {"V_SSA", out->V_SSA}, {"MeltPreviousYear", out->MeltPreviousYear}, {"FirnDepth", out->FirnDepth}};
  for (const auto &q : vars) {
    if (!q.p) continue;
    const int id = f.var_id(q.name);
    if (id < 0 || !f.vars[id].is_rec || f.vars[id].type != T_DOUBLE) return ufm_set_error(-12, "ufm_restart_read_init: variable \"%s\" missing or of the wrong type", q.name);
    if ((rc = get_elems(fd.fd, T_DOUBLE, f.vars[id].begin + (ti_min - 1) * f.recsize, q.p, f.vars[id].slab_elems))) return rc;
  }
  if (ti_out) *ti_out = (int)ti_min;
  return 0;
}

/* read_init_data_from_restart_file (src/restart_module.f90:118-142) + what initialise_ice_model does with it: the frame's Hi, Hb,
 * Ti, U_SSA and V_SSA go straight to the device (Hs is recomputed by update_general_ice_model_data); FirnDepth (nV,12) and
 * MeltPreviousYear (nV) are handed to the host's SMB model (NULL: not read). */
int ufm_restart_load(ufm_handle *h, const char *filename, double time_to_restart_from, double *FirnDepth, double *MeltPreviousYear)
{
  if (!h) return ufm_set_error(-2, "NULL handle");
  Fd fd; File f;
  int rc = open_file(filename, false, fd, f);
  if (rc) return rc;
  long long nV = 0, nZ = 0;
  if ((rc = inquire_dim(f, "vi", &nV)) || (rc = inquire_dim(f, "zeta", &nZ))) return rc;
  if ((rc = check_file_matches_handle(h, "ufm_restart_load", filename, nV, nZ))) return rc;
  close(fd.fd); fd.fd = -1;
  std::vector<double> Hi((size_t)nV), Hb((size_t)nV), U((size_t)nV), V((size_t)nV), Ti((size_t)(nV * nZ));
  ufm_restart_frame_out o;
  memset(&o, 0, sizeof(o));
  o.Hi = Hi.data(); o.Hb = Hb.data(); o.U_SSA = U.data(); o.V_SSA = V.data(); o.Ti = Ti.data();
  o.FirnDepth = FirnDepth; o.MeltPreviousYear = MeltPreviousYear;
  int ti = 0;
  if ((rc = ufm_restart_read_init(filename, time_to_restart_from, &o, &ti))) return rc;
  if ((rc = ufm_state_upload(h, UFM_F_HI, Hi.data())) || (rc = ufm_state_upload(h, UFM_F_HB, Hb.data())) ||
      (rc = ufm_state_upload(h, UFM_F_U_SSA, U.data())) || (rc = ufm_state_upload(h, UFM_F_V_SSA, V.data()))) return rc;
  if (ufm_field_resident(h, UFM_F_TI) == 1 && (rc = ufm_state_upload(h, UFM_F_TI, Ti.data()))) return rc;
  return ti;
}

/* create_help_fields_file_mesh (src/netcdf_module.f90:634-820): the mesh part, then one variable per requested field
 * (C%help_field_01 .. _50; "none" and "resolution" define nothing, an unknown name is fatal as in create_help_field_mesh) */
int ufm_help_fields_create(const char *filename, const ufm_nc_mesh *mesh, int nZ, const double *zeta, int n_fields, const char *const *names)
{
  File f;
  MeshDims d;
  int rc = define_mesh_part(f, mesh, nZ, d);
  if (rc) return rc;
  if (n_fields < 0 || (n_fields > 0 && !names)) return ufm_set_error(-2, "ufm_help_fields_create: bad field list");
  for (int k = 0; k < n_fields; k++) {
    if (skipped_help_field(names[k])) continue;
    const HelpField *hf = find_help_field(names[k]);
    if (!hf) return ufm_set_error(-16, " ERROR: help field \"%s\" not implemented in create_help_field_mesh!", names[k]);
    if (f.var_id(hf->name) >= 0) return ufm_set_error(-16, "ERROR: NetCDF: String match to name in use concerning: %s", hf->name);
    switch (hf->shape) {
      case S_V: def_var(f, hf->name, hf->type, {d.vi}, hf->long_name, hf->units); break;
      case S_VT: def_var(f, hf->name, hf->type, {d.vi, d.time}, hf->long_name, hf->units); break;
      case S_VZT: def_var(f, hf->name, hf->type, {d.vi, d.zeta, d.time}, hf->long_name, hf->units); break;
      case S_VMT: def_var(f, hf->name, hf->type, {d.vi, d.month, d.time}, hf->long_name, hf->units); break;
    }
  }
  return create_and_write_mesh(filename, f, mesh, nZ, zeta);
}

/* write_to_help_fields_file_mesh (src/netcdf_module.f90:216-284): appends one time frame.  host_data[k] != NULL: that host array
 * (nV [, nZ | 12], column-major; INTEGER for the masks) is written; NULL: the field is taken from the device when it lives
 * there (Hi, Hb, Hs, SL, dHs_dx/dy, Ti, Ti_basal, A_flow_mean, U/V_SIA, U/V_SSA, U/V/W_3D, D_SIA, SMB_year, BMB, GHF, T2m,
 * T2m_year, the 11 masks, phi_fric, tau_yield), else rc -16.  Returns the 1-based frame index. */
int ufm_help_fields_write(ufm_handle *h, const char *filename, double time, int n_fields, const char *const *names, const void *const *host_data)
{
  Fd fd; File f;
  int rc = open_file(filename, true, fd, f);
  if (rc) return rc;
  if (n_fields < 0 || (n_fields > 0 && !names)) return ufm_set_error(-2, "ufm_help_fields_write: bad field list");
  long long nV = 0, nZ = 0;
  if ((rc = inquire_dim(f, "vi", &nV)) || (rc = inquire_dim(f, "zeta", &nZ))) return rc;
  const long long rec = f.numrecs;
  for (int k = 0; k < n_fields && h; k++) {   // any field that comes from the device: the file must describe the resident mesh (checked before the frame is started)
    if (skipped_help_field(names[k]) || (host_data && host_data[k])) continue;
    const HelpField *hf = find_help_field(names[k]);
    if (hf && hf->src != SRC_HOST) { if ((rc = check_file_matches_handle(h, "ufm_help_fields_write", filename, nV, nZ))) return rc; break; }
  }
  if ((rc = put_record(fd.fd, f, "time", rec, &time))) return rc;
  std::vector<double> buf, buf2;
  for (int k = 0; k < n_fields; k++) {
    if (skipped_help_field(names[k])) continue;
    const HelpField *hf = find_help_field(names[k]);
    if (!hf) return ufm_set_error(-16, " ERROR: help field \"%s\" not implemented in write_help_field_mesh!", names[k]);
    const void *src = host_data ? host_data[k] : nullptr;
    if (!src) {
      if (hf->src == SRC_HOST) return ufm_set_error(-16, "ufm_help_fields_write: field \"%s\" is not resident on the device; pass the host array", hf->name);
      if (!h) return ufm_set_error(-2, "ufm_help_fields_write: field \"%s\" needs a handle or a host array", hf->name);
      buf.resize((size_t)(nV * std::max(nZ, 12LL)));
      if (hf->src >= 0) {
        if ((rc = ufm_state_download(h, hf->src, buf.data()))) return rc;
      } else if (hf->src == SRC_TI_BASAL) {            // region%ice%Ti(:,C%nZ), :348
        buf2.resize((size_t)(nV * nZ));
        if ((rc = ufm_state_download(h, UFM_F_TI, buf2.data()))) return rc;
        memcpy(buf.data(), buf2.data() + (size_t)((nZ - 1) * nV), (size_t)nV * 8);
      } else if (hf->src == SRC_T2M_YEAR) {            // SUM(region%climate%applied%T2m,2)/12._dp, :388
        buf2.resize((size_t)(nV * 12));
        if ((rc = ufm_state_download(h, UFM_F_T2M, buf2.data()))) return rc;
        for (long long v = 0; v < nV; v++) {
          double s = 0.0;
          for (int m = 0; m < 12; m++) s += buf2[(size_t)(m * nV + v)];
          buf[(size_t)v] = s / 12.0;
        }
      } else {                                        // phi_fric_AaAc(1:nV), tau_c_AaAc(1:nV), :470-473
        long long nM = 0;
        if ((rc = inquire_dim(f, "ai", &nM))) return rc;
        if ((rc = check_file_matches_handle(h, "ufm_help_fields_write", filename, nV, nZ, nM))) return rc;
        buf2.resize((size_t)nM);
        if ((rc = ufm_state_download(h, hf->src == SRC_PHI_AA ? UFM_F_PHI_FRIC_AAAC : UFM_F_TAU_C_AAAC, buf2.data()))) return rc;
        memcpy(buf.data(), buf2.data(), (size_t)nV * 8);
      }
      src = buf.data();
    }
    const int id = f.var_id(hf->name);
    if (id < 0) return ufm_set_error(-12, "ufm_help_fields_write: the file has no variable \"%s\" (not in the list given to ufm_help_fields_create)", hf->name);
    const Var &v = f.vars[id];
    // fields without a time dimension are (re)written on every call, like the reference does (:300-307)
    if ((rc = put_elems(fd.fd, v.type, v.is_rec ? v.begin + rec * f.recsize : v.begin, src, v.slab_elems))) return rc;
  }
  // record variables not named in this call keep the fill value in the new frame
  for (const Var &v : f.vars) {
    if (!v.is_rec || v.name == "time") continue;
    bool named = false;
    for (int k = 0; k < n_fields && !named; k++) named = !skipped_help_field(names[k]) && v.name == names[k];
    if (!named && (rc = put_elems(fd.fd, v.type, v.begin + rec * f.recsize, nullptr, v.slab_elems))) return rc;
  }
  if ((rc = set_numrecs(fd.fd, rec + 1))) return rc;
  return (int)(rec + 1);
}

/* get_output_filenames (src/netcdf_module.f90:66-174): first free "<dir>restart_<NAM>_0000n.nc" / "<dir>help_fields_<NAM>_0000n.nc";
 * kind 0 = restart, 1 = help_fields.  out must hold at least 256 + strlen(output_dir) characters. */
int ufm_output_filename(const char *output_dir, const char *region_name, int kind, char *out, int out_len)
{
  if (!output_dir || !region_name || !out || strlen(region_name) != 3) return ufm_set_error(-2, "ufm_output_filename: bad argument (region names have 3 letters)");
  for (int n = 1; n < 100000000; n++) {
    const int w = snprintf(out, (size_t)out_len, "%s%s_%s_%05d.nc", output_dir, kind == 0 ? "restart" : "help_fields", region_name, n);
    if (w < 0 || w >= out_len) return ufm_set_error(-2, "ufm_output_filename: buffer too small");
    struct stat st;
    if (stat(out, &st) != 0) return n;
  }
  return ufm_set_error(-2, "ufm_output_filename: no free file name");
}

}  // extern "C"
