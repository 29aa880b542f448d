"""CPU mesh substrate: synthetic meshes carrying UFEMISM's ``type_mesh`` contract.

north_star keeps Delaunay mesh creation and refinement on the CPU (reference:
``src/mesh_creation_module.f90``, ``src/mesh_update_module.f90``); the device path only
consumes the resulting arrays.  There is no Fortran compiler in
this image, so tests and benchmarks obtain those arrays here: a point cloud is triangulated
with ``scipy.spatial.Delaunay`` and the secondary mesh data are derived by
``csrc/mesh_host.c`` which restates the reference routines called from
``create_final_mesh_from_merged_submesh`` (``src/mesh_creation_module.f90:1724-1737``).

Every array is Fortran-ordered (column-major) with 1-based indices *stored in* the arrays,
exactly what gfortran would hand to the C ABI (``include/ufemism_b200.h``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libufm_mesh.so")
_SRC = os.path.join(_HERE, "csrc", "mesh_host.c")
_lib = None

NC_MEM = 16  # C%nconmax, src/configuration_module.f90:78


def build_mesh_lib(force: bool = False) -> str:
    """Compile ``csrc/mesh_host.c`` into ``libufm_mesh.so`` (gcc, no fast-math, no FMA)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", _LIB_PATH, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build_mesh_lib()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@dataclass
class Mesh:
    """Arrays of ``type_mesh`` used by the hot path (``src/data_types_module.f90:216-345``)."""

    xmin: float
    xmax: float
    ymin: float
    ymax: float
    nV: int
    nTri: int
    nAc: int
    nC_mem: int
    V: np.ndarray
    Tri: np.ndarray
    Tricc: np.ndarray
    Tri_edge_index: np.ndarray
    nC: np.ndarray
    C: np.ndarray
    niTri: np.ndarray
    iTri: np.ndarray
    edge_index: np.ndarray
    A: np.ndarray
    Cw: np.ndarray
    Nx: np.ndarray
    Ny: np.ndarray
    Nxx: np.ndarray
    Nxy: np.ndarray
    Nyy: np.ndarray
    iAci: np.ndarray
    Aci: np.ndarray
    VAc: np.ndarray
    Nx_Ac: np.ndarray
    Ny_Ac: np.ndarray
    Np_Ac: np.ndarray
    No_Ac: np.ndarray
    edge_index_Ac: np.ndarray
    nVAaAc: int
    VAaAc: np.ndarray
    nCAaAc: np.ndarray
    CAaAc: np.ndarray
    Nx_AaAc: np.ndarray
    Ny_AaAc: np.ndarray
    Nxx_AaAc: np.ndarray
    Nxy_AaAc: np.ndarray
    Nyy_AaAc: np.ndarray
    colour: np.ndarray
    colour_vi: np.ndarray
    colour_nV: np.ndarray
    extra: dict = field(default_factory=dict)

    # ---- mesh data read only by thermodynamics (upwind temperature advection); derived on first use ----
    @property
    def R(self) -> np.ndarray:
        """``determine_mesh_resolution`` (``src/mesh_help_functions_module.f90:189-216``): distance to the nearest neighbour."""
        if "R" not in self.extra:
            R = np.full(self.nV, self.xmax - self.xmin, np.float64)
            for ci in range(self.nC_mem):
                vj = self.C[:, ci]
                has = vj > 0
                dx = self.V[vj[has] - 1, 0] - self.V[has, 0]
                dy = self.V[vj[has] - 1, 1] - self.V[has, 1]
                R[has] = np.minimum(R[has], np.sqrt(dx * dx + dy * dy))
            self.extra["R"] = R
        return self.extra["R"]

    def _tri_functions(self):
        """First-order neighbour functions on the triangles (``src/mesh_derivatives_module.f90:30-47``)."""
        if "NxTri" not in self.extra:
            t = self.Tri.astype(np.int64) - 1
            ax, ay = self.V[t[:, 0], 0], self.V[t[:, 0], 1]
            bx, by = self.V[t[:, 1], 0], self.V[t[:, 1], 1]
            cx, cy = self.V[t[:, 2], 0], self.V[t[:, 2], 1]
            D = ax * (by - cy) + bx * (cy - ay) + cx * (ay - by)
            self.extra["NxTri"] = np.asfortranarray(np.stack([(by - cy) / D, (cy - ay) / D, (ay - by) / D], 1))
            self.extra["NyTri"] = np.asfortranarray(np.stack([(cx - bx) / D, (ax - cx) / D, (bx - ax) / D], 1))
        return self.extra["NxTri"], self.extra["NyTri"]

    @property
    def NxTri(self) -> np.ndarray:
        return self._tri_functions()[0]

    @property
    def NyTri(self) -> np.ndarray:
        return self._tri_functions()[1]

    @property
    def TriC(self) -> np.ndarray:
        """Triangle neighbours, ``TriC(ti,n)`` = the triangle across from the n-th vertex of ``ti`` (0 on the domain boundary),
        ``src/data_types_module.f90:265``; written to the restart / help_fields files and read by ``MATLAB/ReadMeshFromFile.m``."""
        if "TriC" not in self.extra:
            t = self.Tri.astype(np.int64)
            nT = len(t)
            a = np.concatenate([t[:, 1], t[:, 2], t[:, 0]])          # edge opposite vertex n runs from vertex n+1 to vertex n+2
            b = np.concatenate([t[:, 2], t[:, 0], t[:, 1]])
            key = np.minimum(a, b) * (self.nV + 1) + np.maximum(a, b)
            order = np.argsort(key, kind="stable")
            ks = key[order]
            same_next = np.zeros(len(ks), bool)
            same_next[:-1] = ks[1:] == ks[:-1]
            partner = np.full(len(ks), -1, np.int64)
            i = np.nonzero(same_next)[0]
            partner[order[i]] = order[i + 1]
            partner[order[i + 1]] = order[i]
            TriC = np.where(partner >= 0, partner % nT + 1, 0).astype(np.int32).reshape(3, nT).T
            self.extra["TriC"] = np.asfortranarray(TriC)
        return self.extra["TriC"]

    def save(self, path: str) -> None:
        d = {k: v for k, v in self.__dict__.items() if k != "extra"}
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path: str) -> "Mesh":
        z = np.load(path)
        kw = {}
        for k in z.files:
            a = z[k]
            kw[k] = a.item() if a.ndim == 0 else np.asfortranarray(a)
        return Mesh(**kw)


def make_points(xmin, xmax, ymin, ymax, h, seed=20211103, jitter=0.18, order="random", warp=None):
    """Corner vertices 1..4 = SW, SE, NE, NW (``src/mesh_creation_module.f90:1956-1960``), then
    boundary points exactly on the domain edge, then a jittered hexagonal lattice of spacing
    ``h`` kept 0.7 h away from the edge.  ``order='random'`` shuffles vertices 5.. to mimic the
    poor index locality of refinement-ordered meshes; ``warp(x, y) -> (x, y)`` optionally grades
    the lattice."""
    rng = np.random.default_rng(seed)
    Lx, Ly = xmax - xmin, ymax - ymin
    nbx, nby = max(2, int(round(Lx / h))), max(2, int(round(Ly / h)))
    bx = xmin + Lx * (np.arange(1, nbx) / nbx)
    by = ymin + Ly * (np.arange(1, nby) / nby)
    corners = np.array([[xmin, ymin], [xmax, ymin], [xmax, ymax], [xmin, ymax]], dtype=np.float64)
    border = np.concatenate(
        [
            np.stack([bx, np.full_like(bx, ymin)], 1),
            np.stack([np.full_like(by, xmax), by], 1),
            np.stack([bx, np.full_like(bx, ymax)], 1),
            np.stack([np.full_like(by, xmin), by], 1),
        ]
    )
    dy = h * np.sqrt(3.0) / 2.0
    ny = int(np.ceil(Ly / dy)) + 2
    nx = int(np.ceil(Lx / h)) + 2
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    x = xmin + (ii + 0.5 * (jj % 2)) * h - 0.25 * h
    y = ymin + jj * dy
    x = x + rng.uniform(-jitter, jitter, x.shape) * h
    y = y + rng.uniform(-jitter, jitter, y.shape) * h
    pts = np.stack([x.ravel(), y.ravel()], 1)
    if warp is not None:
        wx, wy = warp(pts[:, 0], pts[:, 1])
        pts = np.stack([wx, wy], 1)
    m = 0.7 * h
    keep = (pts[:, 0] > xmin + m) & (pts[:, 0] < xmax - m) & (pts[:, 1] > ymin + m) & (pts[:, 1] < ymax - m)
    pts = pts[keep]
    rest = np.concatenate([border, pts])
    if order == "random":
        rest = rest[rng.permutation(len(rest))]
    elif order == "morton":   # locality-preserving numbering: Z-curve of the vertex coordinates (the CPU arm's best case)
        def spread(v):
            v = v.astype(np.uint64) & np.uint64(0xFFFF)
            for sh, msk in ((8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
                v = (v | (v << np.uint64(sh))) & np.uint64(msk)
            return v
        fx = np.clip((rest[:, 0] - xmin) / Lx * 65535.0, 0, 65535)
        fy = np.clip((rest[:, 1] - ymin) / Ly * 65535.0, 0, 65535)
        rest = rest[np.argsort(spread(fx) | (spread(fy) << np.uint64(1)), kind="stable")]
    elif order != "lattice":
        raise ValueError(order)
    return np.concatenate([corners, rest])


def build_mesh(points, xmin, xmax, ymin, ymax, nC_mem=NC_MEM) -> Mesh:
    """Triangulate ``points`` (vertices 1..4 must be the SW, SE, NE, NW corners) and derive all
    secondary mesh data of ``type_mesh``."""
    from scipy.spatial import Delaunay

    lib = _load()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    nV = len(pts)
    tri = Delaunay(pts).simplices.astype(np.int64)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    cross = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    # drop degenerate slivers along the straight domain boundary, make the rest counter-clockwise
    good = np.abs(cross) > 1e-12 * (xmax - xmin) * (ymax - ymin)
    tri, cross = tri[good], cross[good]
    flip = cross < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    nTri = len(tri)
    V = np.asfortranarray(pts)
    Tri = np.asfortranarray((tri + 1).astype(np.int32))

    x, y = pts[:, 0], pts[:, 1]
    N, E, S, W = y == ymax, x == xmax, y == ymin, x == xmin
    edge_index = np.zeros(nV, np.int32)
    edge_index[N] = 1
    edge_index[E] = 3
    edge_index[S] = 5
    edge_index[W] = 7
    edge_index[N & E] = 2
    edge_index[S & E] = 4
    edge_index[S & W] = 6
    edge_index[N & W] = 8

    nC = np.zeros(nV, np.int32)
    C = np.zeros((nV, nC_mem), np.int32, order="F")
    niTri = np.zeros(nV, np.int32)
    iTri = np.zeros((nV, nC_mem), np.int32, order="F")
    rc = lib.ufm_mesh_connectivity(nV, nTri, _p(Tri), nC_mem, _p(nC), _p(C), _p(niTri), _p(iTri))
    if rc:
        raise RuntimeError(f"ufm_mesh_connectivity failed rc={rc}")

    Tricc = np.zeros((nTri, 2), np.float64, order="F")
    Tei = np.zeros(nTri, np.int32)
    A = np.zeros(nV, np.float64)
    Cw = np.zeros((nV, nC_mem), np.float64, order="F")
    d = ctypes.c_double
    lib.ufm_mesh_geometry.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 7 + [d] * 4 + [ctypes.c_void_p] * 4
    rc = lib.ufm_mesh_geometry(nV, nTri, nC_mem, _p(V), _p(Tri), _p(nC), _p(C), _p(niTri), _p(iTri), _p(edge_index),
                               xmin, xmax, ymin, ymax, _p(Tricc), _p(Tei), _p(A), _p(Cw))
    if rc:
        raise RuntimeError(f"ufm_mesh_geometry failed rc={rc}")

    W1 = nC_mem + 1
    Nx, Ny, Nxx, Nxy, Nyy = (np.zeros((nV, W1), np.float64, order="F") for _ in range(5))
    lib.ufm_mesh_neighbour_functions(nV, nC_mem, _p(V), _p(nC), _p(C), _p(edge_index), _p(Nx), _p(Ny), _p(Nxx), _p(Nxy), _p(Nyy))

    nAc_max = 3 * nTri
    iAci = np.zeros((nV, nC_mem), np.int32, order="F")
    Aci = np.zeros((nAc_max, 4), np.int32, order="F")
    VAc = np.zeros((nAc_max, 2), np.float64, order="F")
    Nx_Ac, Ny_Ac, No_Ac = (np.zeros((nAc_max, 4), np.float64, order="F") for _ in range(3))
    Np_Ac = np.zeros(nAc_max, np.float64)
    eiAc = np.zeros(nAc_max, np.int32)
    nAc = lib.ufm_mesh_make_Ac(nV, nTri, nC_mem, nAc_max, _p(V), _p(Tri), _p(nC), _p(C), _p(niTri), _p(iTri), _p(edge_index),
                               _p(iAci), _p(Aci), _p(VAc), _p(Nx_Ac), _p(Ny_Ac), _p(Np_Ac), _p(No_Ac), _p(eiAc))
    if nAc <= 0:
        raise RuntimeError(f"ufm_mesh_make_Ac failed rc={nAc}")
    Aci = np.asfortranarray(Aci[:nAc])
    VAc = np.asfortranarray(VAc[:nAc])
    Nx_Ac, Ny_Ac, No_Ac = (np.asfortranarray(q[:nAc]) for q in (Nx_Ac, Ny_Ac, No_Ac))
    Np_Ac = Np_Ac[:nAc].copy()
    eiAc = eiAc[:nAc].copy()

    M = nV + nAc
    VAaAc = np.zeros((M, 2), np.float64, order="F")
    nCAaAc = np.zeros(M, np.int32)
    CAaAc = np.zeros((M, nC_mem), np.int32, order="F")
    rc = lib.ufm_mesh_make_AaAc(nV, nAc, nAc, nC_mem, _p(V), _p(VAc), _p(nC), _p(C), _p(iAci), _p(Aci), _p(eiAc),
                                _p(VAaAc), _p(nCAaAc), _p(CAaAc))
    if rc:
        raise RuntimeError(f"ufm_mesh_make_AaAc failed rc={rc}")
    is_edge = np.concatenate([edge_index, eiAc]).astype(np.int32)
    NxA, NyA, NxxA, NxyA, NyyA = (np.zeros((M, W1), np.float64, order="F") for _ in range(5))
    with np.errstate(all="ignore"):
        lib.ufm_mesh_neighbour_functions(M, nC_mem, _p(VAaAc), _p(nCAaAc), _p(CAaAc), _p(is_edge),
                                         _p(NxA), _p(NyA), _p(NxxA), _p(NxyA), _p(NyyA))

    colour = np.zeros(M, np.int32)
    colour_vi = np.zeros((M, 5), np.int32, order="F")
    colour_nV = np.zeros(5, np.int32)
    rc = lib.ufm_mesh_five_colouring(M, nC_mem, _p(nCAaAc), _p(CAaAc), _p(colour), _p(colour_vi), _p(colour_nV))
    if rc:
        raise RuntimeError(f"five-colouring failed rc={rc} (rc=-2: reference would abort in IDENTIFY)")

    return Mesh(xmin=float(xmin), xmax=float(xmax), ymin=float(ymin), ymax=float(ymax), nV=nV, nTri=nTri, nAc=int(nAc),
                nC_mem=nC_mem, V=V, Tri=Tri, Tricc=Tricc, Tri_edge_index=Tei, nC=nC, C=C, niTri=niTri, iTri=iTri,
                edge_index=edge_index, A=A, Cw=Cw, Nx=Nx, Ny=Ny, Nxx=Nxx, Nxy=Nxy, Nyy=Nyy, iAci=iAci, Aci=Aci, VAc=VAc,
                Nx_Ac=Nx_Ac, Ny_Ac=Ny_Ac, Np_Ac=Np_Ac, No_Ac=No_Ac, edge_index_Ac=eiAc, nVAaAc=M, VAaAc=VAaAc,
                nCAaAc=nCAaAc, CAaAc=CAaAc, Nx_AaAc=NxA, Ny_AaAc=NyA, Nxx_AaAc=NxxA, Nxy_AaAc=NxyA, Nyy_AaAc=NyyA,
                colour=colour, colour_vi=colour_vi, colour_nV=colour_nV)


class PrimaryMesh:
    """Only the PRIMARY mesh data (V, Tri, nC, C, niTri, iTri, edge_index + the domain bounds): what the reference holds right after mesh
    generation / a mesh update / reading a restart file, before it derives the secondary data (``src/mesh_creation_module.f90:1724-1737``).
    ``IceModelGPU(mesh, primary_only=True)`` hands exactly this to ``ufm_mesh_upload_primary``, which derives the rest in the library.
    ``nAc`` / ``nVAaAc`` are Euler's counts (nV + nTri - 1 edges), so that field shapes are known without the secondary data."""

    def __init__(self, points, xmin, xmax, ymin, ymax, nC_mem=NC_MEM):
        from scipy.spatial import Delaunay

        lib = _load()
        pts = np.ascontiguousarray(points, dtype=np.float64)
        tri = Delaunay(pts).simplices.astype(np.int64)
        a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
        cross = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
        good = np.abs(cross) > 1e-12 * (xmax - xmin) * (ymax - ymin)
        tri, cross = tri[good], cross[good]
        flip = cross < 0
        tri[flip] = tri[flip][:, [0, 2, 1]]
        self.xmin, self.xmax, self.ymin, self.ymax = float(xmin), float(xmax), float(ymin), float(ymax)
        self.nV, self.nTri, self.nC_mem = len(pts), len(tri), nC_mem
        self.V = np.asfortranarray(pts)
        self.Tri = np.asfortranarray((tri + 1).astype(np.int32))
        x, y = pts[:, 0], pts[:, 1]
        N, E, S, W = y == ymax, x == xmax, y == ymin, x == xmin
        ei = np.zeros(self.nV, np.int32)
        ei[N] = 1; ei[E] = 3; ei[S] = 5; ei[W] = 7; ei[N & E] = 2; ei[S & E] = 4; ei[S & W] = 6; ei[N & W] = 8
        self.edge_index = ei
        self.nC = np.zeros(self.nV, np.int32)
        self.C = np.zeros((self.nV, nC_mem), np.int32, order="F")
        self.niTri = np.zeros(self.nV, np.int32)
        self.iTri = np.zeros((self.nV, nC_mem), np.int32, order="F")
        rc = lib.ufm_mesh_connectivity(self.nV, self.nTri, _p(self.Tri), nC_mem, _p(self.nC), _p(self.C), _p(self.niTri), _p(self.iTri))
        if rc:
            raise RuntimeError(f"ufm_mesh_connectivity failed rc={rc}")
        self.nAc = self.nV + self.nTri - 1          # Euler: V - E + F = 1 for a triangulated disc
        self.nVAaAc = self.nV + self.nAc


def primary_mesh_with_nv(half_width, nv_target, seed=20211103, order="random") -> PrimaryMesh:
    area = (2.0 * half_width) ** 2
    h = np.sqrt(2.0 * area / (np.sqrt(3.0) * nv_target))
    pts = make_points(-half_width, half_width, -half_width, half_width, h, seed=seed, order=order)
    return PrimaryMesh(pts, -half_width, half_width, -half_width, half_width)


def make_mesh(xmin, xmax, ymin, ymax, h, seed=20211103, order="random", warp=None, nC_mem=NC_MEM) -> Mesh:
    pts = make_points(xmin, xmax, ymin, ymax, h, seed=seed, order=order, warp=warp)
    return build_mesh(pts, xmin, xmax, ymin, ymax, nC_mem=nC_mem)


def square_mesh_with_nv(half_width, nv_target, seed=20211103, order="random") -> Mesh:
    """Square domain [-hw, hw]^2 with about ``nv_target`` vertices (hex lattice: nV ~ 2 A / (sqrt3 h^2))."""
    area = (2.0 * half_width) ** 2
    h = np.sqrt(2.0 * area / (np.sqrt(3.0) * nv_target))
    return make_mesh(-half_width, half_width, -half_width, half_width, h, seed=seed, order=order)


def check_mesh(m: Mesh) -> None:
    """Structural invariants in the spirit of ``check_mesh`` (``src/mesh_help_functions_module.f90:2777-3195``)
    and ``check_solution`` (``src/mesh_five_colour_module.f90:318-343``)."""
    nV = m.nV
    assert list(m.edge_index[:4]) == [6, 4, 2, 8], "vertices 1..4 must be the SW, SE, NE, NW corners"
    # C symmetric
    rows = np.repeat(np.arange(1, nV + 1), m.nC_mem)
    cols = m.C.ravel(order="C")
    mask = cols > 0
    e = set(zip(rows[mask].tolist(), cols[mask].tolist())) if nV < 20000 else None
    if e is not None:
        assert all((b, a) in e for a, b in e), "C not symmetric"
    assert int(m.nC.sum()) == 2 * m.nAc
    assert abs(m.A.sum() / ((m.xmax - m.xmin) * (m.ymax - m.ymin)) - 1.0) < 1e-4, "Voronoi areas do not tile the domain"
    assert (m.Aci[:, 0] < m.Aci[:, 1]).all()
    M = m.nVAaAc
    for ai in range(M) if M < 50000 else range(0, M, max(1, M // 50000)):
        for c in range(m.nCAaAc[ai]):
            assert m.colour[m.CAaAc[ai, c] - 1] != m.colour[ai]
    assert int(m.colour_nV.sum()) == M
