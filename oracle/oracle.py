"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- see ufm_oracle.h, "PARITY PIN").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The ctypes structures are generated from the X-macro field lists in
``ufm_oracle.h`` so the two cannot drift apart.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libufm_oracle.so")
_HDR = os.path.join(_HERE, "ufm_oracle.h")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ufm_oracle.c")
    stale = (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < max(os.path.getmtime(src), os.path.getmtime(_HDR))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libufm_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    return _LIB


def _fields(macro: str):
    txt = open(_HDR).read()
    body = txt[txt.index(f"#define {macro}(X)"):]
    body = body[: body.index("\n\n")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return re.findall(r"X\((\w+),\s*(\w+),\s*(\w+),\s*(\w+)\)", body)


MESH_FIELDS = _fields("ORA_MESH_FIELDS")
ICE_FIELDS = _fields("ORA_ICE_FIELDS")


class OraMesh(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("nV", "nAc", "nVAaAc", "nC_mem", "nTri")] + [(n, ctypes.c_void_p) for _, n, _, _ in MESH_FIELDS]


class OraIce(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for _, n, _, _ in ICE_FIELDS]


class OraConfig(ctypes.Structure):
    _fields_ = [("nZ", ctypes.c_int), ("zeta", ctypes.c_double * 32), ("m_enh_sia", ctypes.c_double), ("m_enh_ssa", ctypes.c_double),
                ("use_analytical_GL_flux", ctypes.c_int), ("SSA_RN_tol", ctypes.c_double), ("SSA_max_outer_loops", ctypes.c_int),
                ("SSA_max_residual_UV", ctypes.c_double), ("SSA_SOR_omega", ctypes.c_double), ("SSA_max_inner_loops", ctypes.c_int),
                ("dt_max", ctypes.c_double), ("benchmark", ctypes.c_int), ("nthreads", ctypes.c_int), ("dt_thermo", ctypes.c_double),
                ("thermo", ctypes.c_int)]


class OraSsaStats(ctypes.Structure):
    _fields_ = [("n_outer", ctypes.c_int), ("n_inner_total", ctypes.c_int), ("n_inner_last", ctypes.c_int), ("did_reset", ctypes.c_int),
                ("rc", ctypes.c_int), ("last_max_residual", ctypes.c_double), ("last_RN", ctypes.c_double)]


NT = 8
T_SIA, T_SSA, T_THERMO, T_CLIMATE, T_SMB, T_BMB, T_ELRA, T_OUTPUT = range(8)


class OraRegion(ctypes.Structure):
    _fields_ = [("time", ctypes.c_double), ("dt", ctypes.c_double), ("dt_prev", ctypes.c_double),
                ("t0", ctypes.c_double * NT), ("t1", ctypes.c_double * NT), ("dtc", ctypes.c_double * NT), ("do_", ctypes.c_int * NT),
                ("H0", ctypes.c_double), ("R0", ctypes.c_double), ("lam", ctypes.c_double),
                ("n_steps", ctypes.c_long), ("n_sia", ctypes.c_long), ("n_ssa", ctypes.c_long), ("n_sor_total", ctypes.c_long),
                ("n_outer_total", ctypes.c_long), ("dt_crit_last", ctypes.c_double * 3)]


BENCHMARKS = {"none": 0, "EISMINT_1": 1, "EISMINT_2": 2, "EISMINT_3": 3, "EISMINT_4": 4, "EISMINT_5": 5, "EISMINT_6": 6,
              "Halfar": 7, "Bueler": 8, "MISMIP_mod": 9, "mesh_generation_test": 10, "SSA_icestream": 11}


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        d, i, p = ctypes.c_double, ctypes.c_int, ctypes.c_void_p
        L.ora_calculate_ice_thickness_change.argtypes = [p, p, p, d]
        L.ora_update_general_ice_model_data.argtypes = [p, p, p, d]
        L.ora_solve_SIA.argtypes = [p, p, p]
        L.ora_solve_SIA_3D_UV.argtypes = [p, p, p]
        L.ora_solve_SIA_3D.argtypes = [p, p, p]
        L.ora_update_ice_temperature.argtypes = [p, p, p, p]
        L.ora_replace_Ti_with_robin_solution.argtypes = [p, p, p, i]
        L.ora_dgtsv.argtypes = [i, p, p, p, p, p]
        L.ora_solve_SSA.argtypes = [p, p, p, p]
        L.ora_determine_timesteps.argtypes = [p, p, p, p]
        for f in ("ora_basal_yield_stress", "ora_calculate_GL_flux", "ora_SSA_gather_AaAc", "ora_SSA_effective_viscosity", "ora_SSA_sliding_term"):
            getattr(L, f).argtypes = [p, p, p]
        L.ora_solve_SSA_linearised.argtypes = [p, p, p, i, i, p, p, p]
        L.ora_apply_Neumann_boundary_AaAc.argtypes = [p, p, p]
        L.ora_get_mesh_derivatives.argtypes = [p, p, p, p, p]
        L.ora_map_Ac_to_Aa.argtypes = [p, p, p, p]
        L.ora_remap_cons_2D.argtypes = [i, i, p, p, p, p, p, p, p, p, p, p]
        L.ora_Halfar_solution.argtypes = [d] * 5
        L.ora_Halfar_solution.restype = d
        L.ora_Bueler_solution.argtypes = [d] * 6
        L.ora_Bueler_solution.restype = d
        L.ora_Bueler_solution_MB.argtypes = [d] * 6
        L.ora_Bueler_solution_MB.restype = d
        L.ora_run_SMB_benchmark.argtypes = [p, p, p, d, d, d, d]
        L.ora_region_init.argtypes = [p, d]
        L.ora_run_model.argtypes = [p, p, p, p, d, ctypes.c_long]
        L.ora_config_defaults.argtypes = [p]
        _lib = L
    return _lib


def _dim(tag, mesh, nZ):
    return {"NV": mesh.nV, "NAC": mesh.nAc, "NAA": mesh.nVAaAc, "NZ": nZ, "NCM": mesh.nC_mem, "NCM1": mesh.nC_mem + 1,
            "N1": 1, "N2": 2, "N3": 3, "N4": 4, "N5": 5, "N12": 12, "NTRI": mesh.nTri}[tag]


class Oracle:
    """Reference-layout state (numpy, Fortran order) + calls into libufm_oracle.so."""

    def __init__(self, mesh, benchmark="Halfar", nthreads=1, **cfg):
        self.L = lib()
        self.mesh = mesh
        self.cfg = OraConfig()
        self.L.ora_config_defaults(ctypes.byref(self.cfg))
        self.cfg.benchmark = BENCHMARKS[benchmark]
        self.cfg.nthreads = nthreads
        for k, v in cfg.items():
            setattr(self.cfg, k, v)
        self.nZ = self.cfg.nZ
        self._keep = []
        self.cm = OraMesh(nV=mesh.nV, nAc=mesh.nAc, nVAaAc=mesh.nVAaAc, nC_mem=mesh.nC_mem, nTri=mesh.nTri)
        for t, n, r, c in MESH_FIELDS:
            a = getattr(mesh, n)
            dt = np.float64 if t == "double" else np.int32
            a = np.asfortranarray(a, dtype=dt)
            assert a.size == _dim(r, mesh, self.nZ) * _dim(c, mesh, self.nZ), (n, a.shape)
            self._keep.append(a)
            setattr(self.cm, n, a.ctypes.data)
        self.ice = OraIce()
        self.f = {}
        for t, n, r, c in ICE_FIELDS:
            dt = np.float64 if t == "double" else np.int32
            shape = (_dim(r, mesh, self.nZ), _dim(c, mesh, self.nZ))
            a = np.zeros(shape if shape[1] > 1 else shape[0], dtype=dt, order="F")
            self.f[n] = a
            setattr(self.ice, n, a.ctypes.data)

    def __getitem__(self, k):
        return self.f[k]

    def _a(self):
        return ctypes.byref(self.cm), ctypes.byref(self.ice), ctypes.byref(self.cfg)

    def calculate_ice_thickness_change(self, dt):
        self.L.ora_calculate_ice_thickness_change(*self._a(), float(dt))

    def update_general_ice_model_data(self, time=0.0):
        self.L.ora_update_general_ice_model_data(*self._a(), float(time))

    def solve_SIA(self):
        self.L.ora_solve_SIA(*self._a())

    def solve_SIA_3D(self, with_W=False):
        (self.L.ora_solve_SIA_3D if with_W else self.L.ora_solve_SIA_3D_UV)(*self._a())

    def update_ice_temperature(self):
        """Returns (rc, n_unstable); rc as documented at ora_update_ice_temperature."""
        nu = ctypes.c_int(0)
        rc = self.L.ora_update_ice_temperature(*self._a(), ctypes.byref(nu))
        return rc, nu.value

    def replace_Ti_with_robin_solution(self, vi):
        self.L.ora_replace_Ti_with_robin_solution(*self._a(), int(vi))

    def solve_SSA(self):
        st = OraSsaStats()
        self.L.ora_solve_SSA(*self._a(), ctypes.byref(st))
        return st

    def determine_timesteps(self):
        out = (ctypes.c_double * 3)()
        self.L.ora_determine_timesteps(*self._a(), out)
        return list(out)

    def basal_yield_stress(self):
        self.L.ora_basal_yield_stress(*self._a())

    def calculate_GL_flux(self):
        self.L.ora_calculate_GL_flux(*self._a())

    def SSA_gather_AaAc(self):
        self.L.ora_SSA_gather_AaAc(*self._a())

    def SSA_effective_viscosity(self):
        self.L.ora_SSA_effective_viscosity(*self._a())

    def SSA_sliding_term(self):
        self.L.ora_SSA_sliding_term(*self._a())

    def solve_SSA_linearised(self, max_inner=0, force_iters=False):
        n, r, rs = ctypes.c_int(0), ctypes.c_double(0), ctypes.c_int(0)
        warn = self.L.ora_solve_SSA_linearised(*self._a(), int(max_inner), int(force_iters), ctypes.byref(n), ctypes.byref(r), ctypes.byref(rs))
        return n.value, r.value, rs.value, warn

    def get_mesh_derivatives(self, d):
        ddx, ddy = np.zeros(self.mesh.nV), np.zeros(self.mesh.nV)
        d = np.ascontiguousarray(d, np.float64)
        self.L.ora_get_mesh_derivatives(ctypes.byref(self.cm), ctypes.byref(self.cfg), d.ctypes.data, ddx.ctypes.data, ddy.ctypes.data)
        return ddx, ddy

    def remap_cons_2D(self, order, vli1, vli2, vi, w0, w1x, w1y, d_src):
        """Apply a conservative remapping from this oracle's mesh to a destination mesh with len(vli1) vertices."""
        ddx, ddy = self.get_mesh_derivatives(d_src)
        a = [np.ascontiguousarray(x, np.int32) for x in (vli1, vli2, vi)] + [np.ascontiguousarray(x, np.float64) for x in (w0, w1x if w1x is not None else w0, w1y if w1y is not None else w0)]
        d_src = np.ascontiguousarray(d_src, np.float64)
        out = np.zeros(len(a[0]))
        self.L.ora_remap_cons_2D(int(order), len(a[0]), *[x.ctypes.data for x in a], d_src.ctypes.data, ddx.ctypes.data, ddy.ctypes.data, out.ctypes.data)
        return out

    def apply_Neumann_boundary_AaAc(self, d):
        assert d.dtype == np.float64 and d.size == self.mesh.nVAaAc
        self.L.ora_apply_Neumann_boundary_AaAc(ctypes.byref(self.cm), ctypes.byref(self.cfg), d.ctypes.data)

    def run_SMB_benchmark(self, time, H0=5000.0, R0=300000.0, lam=5.0):
        self.L.ora_run_SMB_benchmark(*self._a(), float(time), H0, R0, lam)

    def region(self, start_time=0.0):
        r = OraRegion()
        self.L.ora_region_init(ctypes.byref(r), float(start_time))
        return r

    def run_model(self, region, t_end, max_steps=0):
        return self.L.ora_run_model(*self._a(), ctypes.byref(region), float(t_end), int(max_steps))


def dgtsv(dl, d, du, b):
    """The oracle's restatement of LAPACK DGTSV (one right-hand side); returns (x, info)."""
    dl, d, du, b = (np.array(v, np.float64) for v in (dl, d, du, b))
    info = ctypes.c_int(0)
    lib().ora_dgtsv(len(d), dl.ctypes.data, d.ctypes.data, du.ctypes.data, b.ctypes.data, ctypes.byref(info))
    return b, info.value


def halfar_solution(H0, R0, x, y, t):
    L = lib()
    return np.array([L.ora_Halfar_solution(H0, R0, float(a), float(b), float(t)) for a, b in zip(np.ravel(x), np.ravel(y))])


def bueler_solution(H0, R0, lam, x, y, t):
    L = lib()
    return np.array([L.ora_Bueler_solution(H0, R0, lam, float(a), float(b), float(t)) for a, b in zip(np.ravel(x), np.ravel(y))])
