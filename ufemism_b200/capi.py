"""ctypes binding of the C ABI (``include/ufemism_b200.h``) -- the same calls the Fortran shim
(``fortran/ufemism_b200_shim.f90``) makes through ISO_C_BINDING, with Fortran-ordered numpy arrays
standing in for the reference's shared-memory windows.

``IceModelGPU`` mirrors the reference's call surface for this path
(``calculate_ice_thickness_change``, ``update_general_ice_model_data``, ``solve_SIA``, ``solve_SSA``:
src/UFEMISM_main_model.f90:90,115,124,132).  There is no CPU fallback: if the CUDA library is
missing or no B200 is visible the constructor raises.
"""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UFM_B200_LIB", os.path.join(_HERE, "libufemism_b200.so"))  # env override: kernel-tuning builds only
HEADER = os.path.join(_HERE, "..", "include", "ufemism_b200.h")
UFM_MAX_NZ = 32
UFM_NT = 8
T_SIA, T_SSA, T_THERMO, T_CLIMATE, T_SMB, T_BMB, T_ELRA, T_OUTPUT = range(8)

BENCHMARKS = {"none": 0, "EISMINT_1": 1, "EISMINT_2": 2, "EISMINT_3": 3, "EISMINT_4": 4, "EISMINT_5": 5, "EISMINT_6": 6,
              "Halfar": 7, "Bueler": 8, "MISMIP_mod": 9, "mesh_generation_test": 10, "SSA_icestream": 11}


class UfmError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"ufemism_b200 rc={rc}: {msg}")
        self.rc = rc


class Params(ctypes.Structure):
    _fields_ = [("nZ", ctypes.c_int), ("zeta", ctypes.c_double * UFM_MAX_NZ), ("m_enh_sia", ctypes.c_double), ("m_enh_ssa", ctypes.c_double),
                ("use_analytical_GL_flux", ctypes.c_int), ("SSA_RN_tol", ctypes.c_double), ("SSA_max_outer_loops", ctypes.c_int),
                ("SSA_max_residual_UV", ctypes.c_double), ("SSA_SOR_omega", ctypes.c_double), ("SSA_max_inner_loops", ctypes.c_int),
                ("dt_max", ctypes.c_double), ("benchmark", ctypes.c_int), ("exact_xy", ctypes.c_int), ("dt_thermo", ctypes.c_double)]


_MESH_PTRS = ["V", "A", "nC", "C", "Cw", "edge_index", "Nx", "Ny", "Aci", "iAci", "edge_index_Ac", "Nx_Ac", "Ny_Ac", "No_Ac", "Np_Ac",
              "nCAaAc", "CAaAc", "Nx_AaAc", "Ny_AaAc", "Nxx_AaAc", "Nxy_AaAc", "Nyy_AaAc", "colour_vi", "colour_nV"]
_INT_FIELDS = {"nC", "C", "edge_index", "Aci", "iAci", "edge_index_Ac", "nCAaAc", "CAaAc", "colour_vi", "colour_nV"}


_THERMO_PTRS = ["Tri", "niTri", "iTri", "R", "NxTri", "NyTri"]
_INT_FIELDS |= {"Tri", "niTri", "iTri"}


class MeshDesc(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in ("nV", "nAc", "nC_mem", "ldV", "ldAc", "ldAaAc")] + [(n, ctypes.c_void_p) for n in _MESH_PTRS] +
                [("nTri", ctypes.c_int), ("ldTri", ctypes.c_int)] + [(n, ctypes.c_void_p) for n in _THERMO_PTRS])


class MeshPrimary(ctypes.Structure):
    """ufm_mesh_primary"""
    _fields_ = ([(n, ctypes.c_int) for n in ("nV", "nTri", "nC_mem", "ldV", "ldTri")] + [(n, ctypes.c_double) for n in ("xmin", "xmax", "ymin", "ymax")] +
                [(n, ctypes.c_void_p) for n in ("V", "nC", "C", "niTri", "iTri", "edge_index", "Tri")] + [("thermo", ctypes.c_int)])


class ThermoStats(ctypes.Structure):
    _fields_ = [("n_unstable", ctypes.c_int), ("rc", ctypes.c_int)]


class SsaStats(ctypes.Structure):
    _fields_ = [("n_outer", ctypes.c_int), ("n_inner_total", ctypes.c_int), ("n_inner_last", ctypes.c_int), ("did_reset", ctypes.c_int),
                ("rc", ctypes.c_int), ("last_max_residual", ctypes.c_double), ("last_RN", ctypes.c_double)]


class Counters(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_longlong), ("sor_iterations", ctypes.c_longlong), ("sor_ms", ctypes.c_double),
                ("sor_launches", ctypes.c_longlong), ("sor_bytes_per_iteration", ctypes.c_double), ("h2d_bytes", ctypes.c_double),
                ("d2h_bytes", ctypes.c_double)]


class Region(ctypes.Structure):
    _fields_ = [("time", ctypes.c_double), ("dt", ctypes.c_double), ("dt_prev", ctypes.c_double),
                ("t0", ctypes.c_double * UFM_NT), ("t1", ctypes.c_double * UFM_NT), ("dtc", ctypes.c_double * UFM_NT), ("do_", ctypes.c_int * UFM_NT),
                ("H0", ctypes.c_double), ("R0", ctypes.c_double), ("lam", ctypes.c_double),
                ("n_steps", ctypes.c_long), ("n_sia", ctypes.c_long), ("n_ssa", ctypes.c_long), ("n_sor_total", ctypes.c_long),
                ("n_outer_total", ctypes.c_long), ("dt_crit_last", ctypes.c_double * 3)]


class RemapCons(ctypes.Structure):
    _fields_ = [("nV_dst", ctypes.c_int), ("n_tot", ctypes.c_int)] + [(n, ctypes.c_void_p) for n in ("vli1", "vli2", "vi", "w0", "w1x", "w1y")]


class HostIce(ctypes.Structure):
    """ufm_host_ice: host arrays (reference vertex order) moved every step in drop-in mode."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("Hi", "Hb", "SL", "dHb_dt", "SMB_year", "BMB", "mask_noice", "Hi_out", "Hi_prev", "dHi_dt", "Hs",
                                               "U_SSA", "V_SSA", "U_SIA", "V_SIA", "D_SIA", "mask")]


def _parse_fields():
    txt = open(HEADER).read()
    body = txt[txt.index("enum ufm_field {"):]
    body = body[: body.index("};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"UFM_F_(\w+)", body)
    out, i = {}, 0
    for n in names:
        if n == "COUNT":
            break
        out[n] = i
        i += 1
    return out


FIELD_IDS = _parse_fields()
# reference field name -> (id, kind, dtype); kind in {"Aa","Ac","AaAc","3D"}
_REF_NAMES = {}
for _n, _i in FIELD_IDS.items():
    if _n.endswith("_AAAC"):
        kind = "AaAc"
    elif _n.endswith("_AC"):
        kind = "Ac"
    elif _n in ("U_3D", "V_3D", "TI", "W_3D"):
        kind = "3D"
    elif _n == "T2M":
        kind = "12"
    else:
        kind = "Aa"
    _REF_NAMES[_n] = (_i, kind, np.int32 if _n.startswith("MASK") else np.float64)

EXPORTED = ["ufm_create", "ufm_destroy", "ufm_set_params", "ufm_set_stream", "ufm_synchronize", "ufm_last_error", "ufm_abi_version",
            "ufm_mesh_upload", "ufm_mesh_free", "ufm_partition_set", "ufm_partition_owners", "ufm_comm_export", "ufm_comm_connect", "ufm_state_upload", "ufm_state_download", "ufm_host_register", "ufm_host_unregister", "ufm_remap_stash", "ufm_remap_apply", "ufm_thickness_update", "ufm_update_general",
            "ufm_solve_SIA", "ufm_solve_SIA_3D", "ufm_solve_SSA", "ufm_cfl", "ufm_ssa_prepare", "ufm_ssa_viscosity", "ufm_ssa_sliding_and_setup", "ufm_ssa_sor",
            "ufm_ssa_finish", "ufm_region_init", "ufm_run_model", "ufm_run_model_host", "ufm_counters_get", "ufm_counters_reset", "ufm_sor_trace_get",
            "ufm_update_ice_temperature", "ufm_thermo_w3d", "ufm_thermo_heat", "ufm_field_resident", "ufm_resident_dims", "ufm_pow_mode", "ufm_pow_host", "ufm_tan_host", "ufm_powtab_status", "ufm_div_small_host", "ufm_partition_owner_of", "ufm_partition_halo_counts",
            "ufm_restart_create", "ufm_restart_append", "ufm_restart_write", "ufm_restart_inquire_mesh", "ufm_restart_read_mesh",
            "ufm_restart_inquire_init", "ufm_restart_read_init", "ufm_restart_load", "ufm_help_fields_create", "ufm_help_fields_write",
            "ufm_output_filename", "ufm_mesh_upload_primary", "ufm_mesh_derive_secondary", "ufm_mesh_derive_secondary_reuse", "ufm_mesh_derived_get", "ufm_mesh_derived_free",
            "ufm_mesh_secondary_get"]

_lib = None


def load_library():
    """Load libufemism_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UfmError(-5, f"{LIB_PATH} not built; run `python -m ufemism_b200.build` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        p, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.ufm_create.argtypes = [i, p, p]
        L.ufm_destroy.argtypes = [p]
        L.ufm_set_params.argtypes = [p, p]
        L.ufm_set_stream.argtypes = [p, p]
        L.ufm_synchronize.argtypes = [p]
        L.ufm_last_error.restype = ctypes.c_char_p
        L.ufm_mesh_upload.argtypes = [p, p]
        L.ufm_mesh_free.argtypes = [p]
        L.ufm_partition_set.argtypes = [p, i, i]
        L.ufm_partition_owners.argtypes = [p, i, p]
        L.ufm_comm_export.argtypes = [p, p]
        L.ufm_comm_connect.argtypes = [p, p]
        L.ufm_state_upload.argtypes = [p, i, p]
        L.ufm_state_download.argtypes = [p, i, p]
        L.ufm_host_register.argtypes = [p, p, ctypes.c_ulonglong]
        L.ufm_host_unregister.argtypes = [p, p]
        L.ufm_remap_stash.argtypes = [p, i]
        L.ufm_remap_apply.argtypes = [p, i, p, i]
        L.ufm_thickness_update.argtypes = [p, d]
        L.ufm_update_general.argtypes = [p, d]
        L.ufm_solve_SIA.argtypes = [p]
        L.ufm_solve_SIA_3D.argtypes = [p]
        L.ufm_solve_SSA.argtypes = [p, p]
        L.ufm_cfl.argtypes = [p, p]
        L.ufm_ssa_prepare.argtypes = [p]
        L.ufm_ssa_viscosity.argtypes = [p, p]
        L.ufm_ssa_sliding_and_setup.argtypes = [p]
        L.ufm_ssa_sor.argtypes = [p, i, i, p]
        L.ufm_ssa_finish.argtypes = [p]
        L.ufm_region_init.argtypes = [p, d]
        L.ufm_run_model.argtypes = [p, p, d, ctypes.c_long]
        L.ufm_run_model_host.argtypes = [p, p, d, ctypes.c_long, p]
        L.ufm_counters_get.argtypes = [p, p]
        L.ufm_sor_trace_get.argtypes = [p, p, ctypes.c_int]
        L.ufm_counters_reset.argtypes = [p]
        L.ufm_update_ice_temperature.argtypes = [p, p]
        L.ufm_thermo_w3d.argtypes = [p]
        L.ufm_thermo_heat.argtypes = [p, p]
        L.ufm_field_resident.argtypes = [p, i]
        L.ufm_resident_dims.argtypes = [p, p]
        L.ufm_pow_mode.argtypes = [p]
        L.ufm_partition_owner_of.argtypes = [p, p]
        L.ufm_pow_host.argtypes = [d, d]
        L.ufm_pow_host.restype = d
        L.ufm_tan_host.argtypes = [d]
        L.ufm_tan_host.restype = d
        L.ufm_powtab_status.argtypes = []
        L.ufm_powtab_status.restype = ctypes.c_int
        L.ufm_div_small_host.argtypes = [d, ctypes.c_int]
        L.ufm_div_small_host.restype = d
        s = ctypes.c_char_p
        L.ufm_restart_create.argtypes = [s, p, i, p]
        L.ufm_restart_append.argtypes = [s, d, p]
        L.ufm_restart_write.argtypes = [p, s, d, p, p]
        L.ufm_restart_inquire_mesh.argtypes = [s, p, p, p]
        L.ufm_restart_read_mesh.argtypes = [s] + [p] * 10
        L.ufm_restart_inquire_init.argtypes = [s, i, p, p]
        L.ufm_restart_read_init.argtypes = [s, d, p, p]
        L.ufm_restart_load.argtypes = [p, s, d, p, p]
        L.ufm_help_fields_create.argtypes = [s, p, i, p, i, p]
        L.ufm_help_fields_write.argtypes = [p, s, d, i, p, p]
        L.ufm_output_filename.argtypes = [s, s, i, p, i]
        L.ufm_mesh_upload_primary.argtypes = [p, p]
        L.ufm_mesh_derive_secondary.argtypes = [p, p]
        L.ufm_mesh_derive_secondary_reuse.argtypes = [p, p]
        L.ufm_mesh_derived_get.argtypes = [p] * 7
        L.ufm_mesh_derived_free.argtypes = [p]
        L.ufm_mesh_derived_free.restype = None
        L.ufm_mesh_secondary_get.argtypes = [p] * 7
        _lib = L
    return _lib


def default_params(benchmark="Halfar", **kw) -> Params:
    """Defaults of src/configuration_module.f90:37,124-126,169-184 with the benchmark configs' m_enh = 1."""
    P = Params()
    P.nZ = 15
    for k, z in enumerate([0.00, 0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 0.925, 0.95, 0.975, 0.99, 1.00]):
        P.zeta[k] = z
    P.m_enh_sia = 1.0
    P.m_enh_ssa = 1.0
    P.use_analytical_GL_flux = 0
    P.SSA_RN_tol = 1e-5
    P.SSA_max_outer_loops = 50
    P.SSA_max_residual_UV = 2.5
    P.SSA_SOR_omega = 1.2
    P.SSA_max_inner_loops = 10000
    P.dt_max = 10.0
    P.benchmark = BENCHMARKS[benchmark]
    P.exact_xy = 1
    P.dt_thermo = 10.0
    for k, v in kw.items():
        setattr(P, k, v)
    return P


_NF_AAAC = ("Nx_AaAc", "Ny_AaAc", "Nxx_AaAc", "Nxy_AaAc", "Nyy_AaAc", "Nx", "Ny", "Nx_Ac", "Ny_Ac", "No_Ac", "Np_Ac")


def mesh_desc(mesh, thermo=False, derive_nf=False):
    """ufm_mesh_desc over the Fortran-ordered arrays of a Mesh; returns (desc, keepalive).  ``thermo`` adds the triangle
    data that only update_ice_temperature reads."""
    d = MeshDesc(nV=mesh.nV, nAc=mesh.nAc, nC_mem=mesh.nC_mem, ldV=mesh.nV, ldAc=mesh.nAc, ldAaAc=mesh.nVAaAc)
    keep = []
    if thermo:
        d.nTri, d.ldTri = mesh.nTri, mesh.nTri
    for n in _MESH_PTRS + (_THERMO_PTRS if thermo else []):
        if derive_nf and n in _NF_AAAC:
            continue          # NULL: the library derives the Aa, Ac and AaAc neighbour functions on the device
        a = np.asfortranarray(getattr(mesh, n), dtype=np.int32 if n in _INT_FIELDS else np.float64)
        keep.append(a)
        setattr(d, n, a.ctypes.data)
    return d, keep


def mesh_primary(mesh, thermo=False):
    """ufm_mesh_primary over the primary arrays of a Mesh (or of a dict as ``restart.read_restart_mesh`` returns, plus the domain
    bounds); returns (struct, keepalive)."""
    get = (lambda n: mesh[n]) if isinstance(mesh, dict) else (lambda n: getattr(mesh, n))
    V = np.asfortranarray(get("V"), np.float64)
    C = np.asfortranarray(get("C"), np.int32)
    Tri = np.asfortranarray(get("Tri"), np.int32)
    keep = [V, C, Tri, np.ascontiguousarray(get("nC"), np.int32), np.ascontiguousarray(get("niTri"), np.int32),
            np.asfortranarray(get("iTri"), np.int32), np.ascontiguousarray(get("edge_index"), np.int32)]
    p = MeshPrimary(nV=V.shape[0], nTri=Tri.shape[0], nC_mem=C.shape[1], ldV=V.shape[0], ldTri=Tri.shape[0],
                    xmin=float(get("xmin")), xmax=float(get("xmax")), ymin=float(get("ymin")), ymax=float(get("ymax")),
                    V=V.ctypes.data, C=C.ctypes.data, Tri=Tri.ctypes.data, nC=keep[3].ctypes.data, niTri=keep[4].ctypes.data,
                    iTri=keep[5].ctypes.data, edge_index=keep[6].ctypes.data, thermo=int(bool(thermo)))
    return p, keep


def _derived_to_dict(L, getter, obj, nV, nTri, W):
    """Copy the arrays behind ufm_mesh_derived_get / ufm_mesh_secondary_get into numpy (Fortran order, compacted)."""
    d = MeshDesc()
    ptrs = [ctypes.c_void_p() for _ in range(5)]
    rc = getter(obj, ctypes.byref(d), *[ctypes.byref(q) for q in ptrs])
    if rc:
        raise UfmError(rc, L.ufm_last_error().decode())
    nAc, ldAc, M = d.nAc, d.ldAc, d.nV + d.nAc

    def arr(ptr, ct, rows, cols=None, ld=None):
        ld = rows if ld is None else ld
        n = ld * (cols or 1)
        a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,)).copy()
        return a if cols is None else np.asfortranarray(a.reshape(cols, ld).T[:rows])

    I, D = ctypes.c_int, ctypes.c_double
    out = {"nV": d.nV, "nAc": nAc, "nC_mem": d.nC_mem, "nVAaAc": M, "ldAc": ldAc,
           "A": arr(d.A, D, nV), "Cw": arr(d.Cw, D, nV, W), "Aci": arr(d.Aci, I, nAc, 4, ldAc), "iAci": arr(d.iAci, I, nV, W),
           "edge_index_Ac": arr(d.edge_index_Ac, I, ldAc)[:nAc], "nCAaAc": arr(d.nCAaAc, I, M), "CAaAc": arr(d.CAaAc, I, M, W),
           "colour_vi": arr(d.colour_vi, I, M, 5), "colour_nV": arr(d.colour_nV, I, 5),
           "Tricc": arr(ptrs[0], D, nTri, 2), "Tri_edge_index": arr(ptrs[1], I, nTri), "VAc": arr(ptrs[2], D, nAc, 2, ldAc),
           "VAaAc": arr(ptrs[3], D, M, 2), "colour": arr(ptrs[4], I, M),
           "nf_pointers_null": all(not getattr(d, n) for n in _NF_AAAC)}
    if d.R:
        out.update(R=arr(d.R, D, nV), NxTri=arr(d.NxTri, D, nTri, 3), NyTri=arr(d.NyTri, D, nTri, 3))
    return out


def derive_secondary(mesh, thermo=False, repeat=1):
    """Host-only: what ufm_mesh_upload_primary derives from the primary mesh data, as a dict of numpy arrays.  repeat > 1 derives that
    many times into the same object (buffers reused, as a re-upload does) and returns the last result."""
    L = load_library()
    p, keep = mesh_primary(mesh, thermo)
    obj = ctypes.c_void_p()
    rc = L.ufm_mesh_derive_secondary(ctypes.byref(p), ctypes.byref(obj))
    for _ in range(repeat - 1):
        if rc:
            break
        rc = L.ufm_mesh_derive_secondary_reuse(ctypes.byref(p), ctypes.byref(obj))
    if rc:
        raise UfmError(rc, L.ufm_last_error().decode())
    try:
        return _derived_to_dict(L, L.ufm_mesh_derived_get, obj, p.nV, p.nTri, p.nC_mem)
    finally:
        L.ufm_mesh_derived_free(obj)


def partition_owners(mesh, nranks):
    """Owner rank of every AaAc vertex (reference order) in an nranks-way vertex partition (host-only call)."""
    L = load_library()
    d, keep = mesh_desc(mesh)
    out = np.zeros(mesh.nVAaAc, np.uint8)
    rc = L.ufm_partition_owners(ctypes.byref(d), int(nranks), out.ctypes.data)
    if rc:
        raise UfmError(rc, L.ufm_last_error().decode())
    return out


def partition_halo_counts(mesh, nranks):
    """(cnt_aa, cnt_ac)[s, q]: values rank s sends to rank q per halo exchange of the partitioned per-step kernels (host-only call)."""
    L = load_library()
    d, keep = mesh_desc(mesh)
    a, c = np.zeros((nranks, nranks), np.int32), np.zeros((nranks, nranks), np.int32)
    rc = L.ufm_partition_halo_counts(ctypes.byref(d), int(nranks), a.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise UfmError(rc, L.ufm_last_error().decode())
    return a, c


def exchange_blobs(dist, blob: bytes, nranks: int, device=None):
    """All-gather one fixed-size byte blob per rank with torch.distributed (any backend); returns them in rank order."""
    import torch

    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(nranks)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def max_over_ranks(dist, value: float, device=None) -> float:
    """The time every multi-rank number is based on: the slowest rank's."""
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class IceModelGPU:
    """One model region resident on one B200."""

    def __init__(self, mesh, benchmark="Halfar", device=0, rank=0, nranks=1, thermo=False, derive_nf=False, primary_only=False, **params):
        self.L = load_library()
        self.mesh = mesh
        self.thermo = bool(thermo)
        self.derive_nf = bool(derive_nf)
        self.primary_only = bool(primary_only)   # upload with ufm_mesh_upload_primary: the library derives all secondary mesh data
        self.P = default_params(benchmark, **params)
        self.h = ctypes.c_void_p()
        self.rank, self.nranks = int(rank), int(nranks)
        self._ck(self.L.ufm_create(int(device), ctypes.byref(self.P), ctypes.byref(self.h)))
        if self.nranks > 1:
            self._ck(self.L.ufm_partition_set(self.h, self.rank, self.nranks))
        self.upload_mesh(mesh)

    def _ck(self, rc, allow_warning=False):
        if rc < 0 or (rc > 0 and not allow_warning):
            raise UfmError(rc, self.L.ufm_last_error().decode())
        return rc

    def close(self):
        if self.h:
            self.L.ufm_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.P, k, BENCHMARKS[v] if k == "benchmark" and isinstance(v, str) else v)
        self._ck(self.L.ufm_set_params(self.h, ctypes.byref(self.P)))

    def upload_mesh(self, mesh):
        if self.primary_only:
            p, keep = mesh_primary(mesh, thermo=self.thermo)
            self._ck(self.L.ufm_mesh_upload_primary(self.h, ctypes.byref(p)))
        else:
            d, keep = mesh_desc(mesh, thermo=self.thermo, derive_nf=self.derive_nf)
            self._ck(self.L.ufm_mesh_upload(self.h, ctypes.byref(d)))
        self.mesh = mesh

    def secondary(self):
        """The host arrays ufm_mesh_upload_primary derived for the resident mesh (copies)."""
        m = self.mesh
        return _derived_to_dict(self.L, self.L.ufm_mesh_secondary_get, self.h, m.nV, m.nTri, m.nC_mem)

    COMM_BLOB_BYTES = 256

    def comm_export(self) -> bytes:
        buf = ctypes.create_string_buffer(self.COMM_BLOB_BYTES)
        self._ck(self.L.ufm_comm_export(self.h, buf))
        return buf.raw

    def comm_connect(self, blobs):
        """``blobs``: the COMM_BLOB_BYTES export of every rank, in rank order."""
        assert len(blobs) == self.nranks and all(len(b) == self.COMM_BLOB_BYTES for b in blobs)
        self._ck(self.L.ufm_comm_connect(self.h, b"".join(blobs)))

    def connect(self, dist, device=None):
        """All-gather the IPC blobs with ``torch.distributed`` (any backend) and connect; collective."""
        self.comm_connect(exchange_blobs(dist, self.comm_export(), self.nranks, device))
        dist.barrier()

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.ufm_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(self.L.ufm_synchronize(self.h))

    # ---- state ----
    def _shape(self, kind):
        m = self.mesh
        return {"Aa": (m.nV,), "Ac": (m.nAc,), "AaAc": (m.nVAaAc,), "3D": (m.nV, self.P.nZ), "12": (m.nV, 12)}[kind]

    def upload(self, name, arr):
        fid, kind, dt = _REF_NAMES[name.upper()]
        a = np.asfortranarray(arr, dtype=dt)
        assert a.shape == self._shape(kind), (name, a.shape, self._shape(kind))
        self._ck(self.L.ufm_state_upload(self.h, fid, a.ctypes.data))

    def host_register(self, arr):
        """Page-lock a numpy array so uploads/downloads of it are DMA'd directly."""
        self._ck(self.L.ufm_host_register(self.h, arr.ctypes.data, arr.nbytes))

    def host_unregister(self, arr):
        self._ck(self.L.ufm_host_unregister(self.h, arr.ctypes.data))

    def remap_stash(self, name):
        self._ck(self.L.ufm_remap_stash(self.h, _REF_NAMES[name.upper()][0]))

    def remap_apply(self, name, vli1, vli2, vi, w0, w1x=None, w1y=None):
        order = 1 if w1x is None else 2
        arrs = [np.ascontiguousarray(vli1, np.int32), np.ascontiguousarray(vli2, np.int32), np.ascontiguousarray(vi, np.int32), np.ascontiguousarray(w0, np.float64)]
        arrs += [np.ascontiguousarray(w1x, np.float64), np.ascontiguousarray(w1y, np.float64)] if order == 2 else []
        mp = RemapCons(nV_dst=len(arrs[0]), n_tot=len(arrs[2]), vli1=arrs[0].ctypes.data, vli2=arrs[1].ctypes.data, vi=arrs[2].ctypes.data, w0=arrs[3].ctypes.data,
                       w1x=arrs[4].ctypes.data if order == 2 else None, w1y=arrs[5].ctypes.data if order == 2 else None)
        self._ck(self.L.ufm_remap_apply(self.h, _REF_NAMES[name.upper()][0], ctypes.byref(mp), order))

    def download(self, name, out=None):
        fid, kind, dt = _REF_NAMES[name.upper()]
        if out is None:
            out = np.zeros(self._shape(kind), dtype=dt, order="F")
        self._ck(self.L.ufm_state_download(self.h, fid, out.ctypes.data))
        return out

    # ---- the reference's call surface ----
    def calculate_ice_thickness_change(self, dt):
        self._ck(self.L.ufm_thickness_update(self.h, float(dt)))

    def update_general_ice_model_data(self, time=0.0):
        self._ck(self.L.ufm_update_general(self.h, float(time)))

    def solve_SIA(self):
        self._ck(self.L.ufm_solve_SIA(self.h))

    def solve_SIA_3D(self):
        self._ck(self.L.ufm_solve_SIA_3D(self.h))

    def update_ice_temperature(self):
        """update_ice_temperature; returns ThermoStats (raises UfmError for rc -8 / -9 / -10)."""
        st = ThermoStats()
        self._ck(self.L.ufm_update_ice_temperature(self.h, ctypes.byref(st)))
        return st

    def thermo_w3d(self):
        self._ck(self.L.ufm_thermo_w3d(self.h))

    def thermo_heat(self):
        st = ThermoStats()
        self._ck(self.L.ufm_thermo_heat(self.h, ctypes.byref(st)))
        return st

    def solve_SSA(self):
        st = SsaStats()
        self._ck(self.L.ufm_solve_SSA(self.h, ctypes.byref(st)), allow_warning=True)
        return st

    def determine_timesteps(self):
        out = (ctypes.c_double * 3)()
        self._ck(self.L.ufm_cfl(self.h, out))
        return list(out)

    # ---- pieces of solve_SSA ----
    def ssa_prepare(self):
        self._ck(self.L.ufm_ssa_prepare(self.h))

    def ssa_viscosity(self):
        out = (ctypes.c_double * 2)()
        self._ck(self.L.ufm_ssa_viscosity(self.h, out))
        return list(out)

    def ssa_sliding_and_setup(self):
        self._ck(self.L.ufm_ssa_sliding_and_setup(self.h))

    def ssa_sor(self, max_inner=0, force_iters=False):
        st = SsaStats()
        self._ck(self.L.ufm_ssa_sor(self.h, int(max_inner), int(force_iters), ctypes.byref(st)))
        return st

    def ssa_finish(self):
        self._ck(self.L.ufm_ssa_finish(self.h))

    # ---- region loop ----
    def region(self, start_time=0.0):
        r = Region()
        self._ck(self.L.ufm_region_init(ctypes.byref(r), float(start_time)))
        return r

    def run_model(self, region, t_end, max_steps=0):
        return self._ck(self.L.ufm_run_model(self.h, ctypes.byref(region), float(t_end), int(max_steps)))

    def run_model_host(self, region, t_end, max_steps, host):
        return self._ck(self.L.ufm_run_model_host(self.h, ctypes.byref(region), float(t_end), int(max_steps), ctypes.byref(host)))

    def counters(self):
        c = Counters()
        self._ck(self.L.ufm_counters_get(self.h, ctypes.byref(c)))
        return c

    def sor_trace(self, raw=False):
        """Phase timestamps of the SOR kernel's fourth iteration, shape (n_ctas, 6, 4) in ns (needs UFM_SOR_TRACE=1); raw: the whole
        tuning buffer (tools/df_stats_probe.py)."""
        buf = np.zeros(4096 * 24, np.uint64)
        n = self.L.ufm_sor_trace_get(self.h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
        if n < 0:
            self._ck(n)
        return buf if raw else buf[: n * 24].reshape(n, 6, 4)

    def reset_counters(self):
        self._ck(self.L.ufm_counters_reset(self.h))

    # ---- restart / help_fields files (reference format; ufemism_b200/restart.py holds the host-only half) ----
    def owners(self):
        """(owner rank of every AaAc vertex in reference order, per-step kernels partitioned?) -- see ufm_partition_owner_of."""
        out = np.zeros(self.mesh.nVAaAc, np.uint8)
        rc = self._ck(self.L.ufm_partition_owner_of(self.h, out.ctypes.data_as(ctypes.c_void_p)), allow_warning=True)
        return out, rc == 1

    def download_global(self, dist, name, device=None):
        """Collective: the complete field on every rank of a partitioned run.  With partitioned per-step kernels a rank's download is valid
        for the elements it owns; the pieces are put together with an all-gather over ``torch.distributed`` (what the reference's shared
        MPI window does for free)."""
        import torch

        a = self.download(name)
        own, part = self.owners()
        if self.nranks <= 1 or not part:
            return a
        _, kind, _ = _REF_NAMES[name.upper()]
        m = self.mesh
        own_k = {"Aa": own[: m.nV], "Ac": own[m.nV:], "AaAc": own, "3D": own[: m.nV], "12": own[: m.nV]}[kind]
        t = torch.from_numpy(np.ascontiguousarray(a))
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(self.nranks)]
        dist.all_gather(parts, t)
        out = np.empty_like(a)
        for q in range(self.nranks):
            sel = own_k == q
            out[sel] = parts[q].cpu().numpy()[sel]
        return out

    def pow_mode(self) -> int:
        """bit 0: device pow has the bits of the host's libm (ufm_pow.cuh), bit 1: so has tan on the yield-stress range; 0: CUDA's pow / tan."""
        return int(self.L.ufm_pow_mode(self.h))

    def field_resident(self, name) -> bool:
        return self._ck(self.L.ufm_field_resident(self.h, _REF_NAMES[name.upper()][0]), allow_warning=True) == 1

    def write_restart(self, filename, time, FirnDepth=None, MeltPreviousYear=None) -> int:
        """write_to_restart_file_mesh with the device's fields; returns the 1-based time-frame index."""
        fd = None if FirnDepth is None else np.asfortranarray(FirnDepth, np.float64)
        mp = None if MeltPreviousYear is None else np.ascontiguousarray(MeltPreviousYear, np.float64)
        return self._ck(self.L.ufm_restart_write(self.h, os.fsencode(filename), float(time), None if fd is None else fd.ctypes.data,
                                                 None if mp is None else mp.ctypes.data), allow_warning=True)

    def load_restart(self, filename, time_to_restart_from) -> int:
        """read_restart_file_init straight onto the device (Hi, Hb, U_SSA, V_SSA, Ti); returns the frame index read."""
        return self._ck(self.L.ufm_restart_load(self.h, os.fsencode(filename), float(time_to_restart_from), None, None), allow_warning=True)

    def write_help_fields(self, filename, time, names, host=None) -> int:
        """write_to_help_fields_file_mesh: ``names`` as in C%help_field_01..50; ``host`` maps a name to a host array for the
        fields the device does not hold."""
        host = host or {}
        keep = [np.asfortranarray(host[n]) if n in host else None for n in names]
        arr = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
        ptrs = (ctypes.c_void_p * len(names))(*[None if k is None else k.ctypes.data for k in keep])
        return self._ck(self.L.ufm_help_fields_write(self.h, os.fsencode(filename), float(time), len(names), arr, ptrs), allow_warning=True)
