"""Host-only half of the restart / help_fields file support: ctypes wrappers of the ``ufm_restart_*`` and
``ufm_help_fields_create`` entry points of ``include/ufemism_b200.h``.

The files are NetCDF classic, laid out exactly as the reference's ``create_restart_file_mesh`` /
``create_help_fields_file_mesh`` write them (``src/netcdf_module.f90:489-820``) and as ``read_mesh_from_restart_file`` /
``read_init_data_from_restart_file`` (``src/restart_module.f90:31-144``) expect them.  Nothing here needs a GPU; the two calls
that move fields from / to the device are methods of ``capi.IceModelGPU`` (``write_restart``, ``load_restart``,
``write_help_fields``).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .capi import UfmError, load_library

_NC_MESH_INT = ("Tri", "nC", "C", "niTri", "iTri", "edge_index", "TriC", "Tri_edge_index", "Aci", "iAci", "TriAaAc", "vi_transect")
_NC_MESH_PTRS = ("V", "Tri", "nC", "C", "niTri", "iTri", "edge_index", "Tricc", "TriC", "Tri_edge_index", "VAc", "Aci", "iAci", "VAaAc",
                 "TriAaAc", "A", "R", "vi_transect", "w_transect")


class NcMesh(ctypes.Structure):
    """ufm_nc_mesh"""
    _fields_ = ([(n, ctypes.c_int) for n in ("nV", "nTri", "nC_mem", "nAc", "nV_transect", "nVAaAc", "nTriAaAc")] +
                [(n, ctypes.c_void_p) for n in _NC_MESH_PTRS])


class RestartFrame(ctypes.Structure):
    """ufm_restart_frame"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("Hi", "Hb", "Hs", "U_SIA", "V_SIA", "U_SSA", "V_SSA", "Ti", "FirnDepth", "MeltPreviousYear")]


class RestartFrameOut(ctypes.Structure):
    """ufm_restart_frame_out"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("Hi", "Hb", "Hs", "Ti", "U_SSA", "V_SSA", "MeltPreviousYear", "FirnDepth")]


def _ck(rc, allow_warning=False):
    if rc < 0 or (rc > 0 and not allow_warning):
        raise UfmError(rc, load_library().ufm_last_error().decode())
    return rc


def nc_mesh(mesh, extra=None):
    """ufm_nc_mesh over a ``mesh.Mesh``.  Arrays the substrate does not carry (TriC, TriAaAc, the transect; ``extra`` may supply
    them) are written as the NetCDF fill value; their dimensions default to the sizes the reference would have
    (nTriAaAc = 4 nTri: every triangle of the Aa mesh is split in four on the combined mesh, src/mesh_ArakawaC_module.f90:431-540)."""
    extra = dict(extra or {})
    d = NcMesh(nV=mesh.nV, nTri=mesh.nTri, nC_mem=mesh.nC_mem, nAc=mesh.nAc, nVAaAc=mesh.nVAaAc)
    d.nV_transect = int(extra.get("nV_transect", 1))
    d.nTriAaAc = int(extra.get("nTriAaAc", 4 * mesh.nTri))
    keep = []
    for n in _NC_MESH_PTRS:
        a = extra.get(n)
        if a is None:
            a = getattr(mesh, n, None)
        if a is None:
            continue
        a = np.asfortranarray(a, dtype=np.int32 if n in _NC_MESH_INT else np.float64)
        keep.append(a)
        setattr(d, n, a.ctypes.data)
    return d, keep


def create_restart(filename, mesh, zeta, extra=None):
    """create_restart_file_mesh"""
    d, keep = nc_mesh(mesh, extra)
    z = np.ascontiguousarray(zeta, np.float64)
    _ck(load_library().ufm_restart_create(os.fsencode(filename), ctypes.byref(d), len(z), z.ctypes.data))


def append_restart(filename, time, **fields) -> int:
    """write_to_restart_file_mesh from host arrays; returns the 1-based frame index."""
    fr = RestartFrame()
    keep = []
    for n, a in fields.items():
        a = np.asfortranarray(a, np.float64)
        keep.append(a)
        setattr(fr, n, a.ctypes.data)
    return _ck(load_library().ufm_restart_append(os.fsencode(filename), float(time), ctypes.byref(fr)), allow_warning=True)


def inquire_restart_mesh(filename):
    """inquire_restart_file_mesh -> (nV, nTri, nC_mem)"""
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _ck(load_library().ufm_restart_inquire_mesh(os.fsencode(filename), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def read_restart_mesh(filename):
    """read_restart_file_mesh -> dict of the primary mesh arrays (Fortran order, 1-based indices)."""
    nV, nTri, W = inquire_restart_mesh(filename)
    out = {"V": np.zeros((nV, 2), np.float64, order="F"), "nC": np.zeros(nV, np.int32), "C": np.zeros((nV, W), np.int32, order="F"),
           "niTri": np.zeros(nV, np.int32), "iTri": np.zeros((nV, W), np.int32, order="F"), "edge_index": np.zeros(nV, np.int32),
           "Tri": np.zeros((nTri, 3), np.int32, order="F"), "Tricc": np.zeros((nTri, 2), np.float64, order="F"),
           "TriC": np.zeros((nTri, 3), np.int32, order="F"), "Tri_edge_index": np.zeros(nTri, np.int32)}
    order = ("V", "nC", "C", "niTri", "iTri", "edge_index", "Tri", "Tricc", "TriC", "Tri_edge_index")
    _ck(load_library().ufm_restart_read_mesh(os.fsencode(filename), *[out[n].ctypes.data for n in order]))
    return out


def inquire_restart_init(filename, zeta):
    """inquire_restart_file_init -> (number of time frames, zeta_matches)"""
    z = np.ascontiguousarray(zeta, np.float64)
    nt = ctypes.c_int()
    rc = _ck(load_library().ufm_restart_inquire_init(os.fsencode(filename), len(z), z.ctypes.data, ctypes.byref(nt)), allow_warning=True)
    return nt.value, rc == 0


def read_restart_init(filename, time_to_restart_from, nV, nZ):
    """read_restart_file_init -> (dict of fields, 1-based frame index)"""
    out = {"Hi": np.zeros(nV), "Hb": np.zeros(nV), "Hs": np.zeros(nV), "Ti": np.zeros((nV, nZ), order="F"), "U_SSA": np.zeros(nV),
           "V_SSA": np.zeros(nV), "MeltPreviousYear": np.zeros(nV), "FirnDepth": np.zeros((nV, 12), order="F")}
    o = RestartFrameOut(**{n: a.ctypes.data for n, a in out.items()})
    ti = ctypes.c_int()
    _ck(load_library().ufm_restart_read_init(os.fsencode(filename), float(time_to_restart_from), ctypes.byref(o), ctypes.byref(ti)))
    return out, ti.value


def create_help_fields(filename, mesh, zeta, names, extra=None):
    """create_help_fields_file_mesh with the fields of C%help_field_01..50"""
    d, keep = nc_mesh(mesh, extra)
    z = np.ascontiguousarray(zeta, np.float64)
    arr = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
    _ck(load_library().ufm_help_fields_create(os.fsencode(filename), ctypes.byref(d), len(z), z.ctypes.data, len(names), arr))


def write_help_fields_host(filename, time, names, host) -> int:
    """write_to_help_fields_file_mesh with every field taken from host arrays (no device involved)."""
    keep = [np.asfortranarray(host[n]) for n in names]
    arr = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
    ptrs = (ctypes.c_void_p * len(names))(*[k.ctypes.data for k in keep])
    return _ck(load_library().ufm_help_fields_write(None, os.fsencode(filename), float(time), len(names), arr, ptrs), allow_warning=True)


def output_filename(output_dir, region_name, kind="restart"):
    """get_output_filenames: first free restart_<NAM>_0000n.nc / help_fields_<NAM>_0000n.nc in ``output_dir`` (must end in '/')."""
    buf = ctypes.create_string_buffer(1024)
    _ck(load_library().ufm_output_filename(os.fsencode(output_dir), region_name.encode(), 0 if kind == "restart" else 1, buf, 1024), allow_warning=True)
    return buf.value.decode()
