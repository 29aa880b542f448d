"""Row N4: restart / help_fields files in the reference's on-disk format.

No NetCDF library ships with the product, so the checker here is an independent implementation of the classic format:
``scipy.io.netcdf_file``.  Files written by ``ufm_restart_*`` / ``ufm_help_fields_*`` must read back through scipy with the
dimension / variable names, order, types and attributes that ``create_restart_file_mesh`` / ``create_help_fields_file_mesh``
define (src/netcdf_module.f90:489-820), and files written by scipy in that layout must read through ``ufm_restart_read_*``
as ``read_mesh_from_restart_file`` / ``read_init_data_from_restart_file`` (src/restart_module.f90:31-144) would.
None of these tests needs a GPU (the device-facing calls are covered in test_gpu_parity.py).
"""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

from ufemism_b200 import mesh as M
from ufemism_b200 import restart as R
from ufemism_b200.capi import UfmError

ZETA = [0.00, 0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 0.925, 0.95, 0.975, 0.99, 1.00]

# create_restart_file_mesh, src/netcdf_module.f90:521-600, in definition order; dims in the FORTRAN order the reference lists
MESH_DIMS = ["vi", "ti", "ci", "aci", "ciplusone", "two", "three", "four", "vii", "ai", "tai"]
MESH_VARS = [("V", "d", ["vi", "two"], "Vertex coordinates", "m"), ("Tri", "i", ["ti", "three"], "Vertex indices", None),
             ("nC", "i", ["vi"], "Number of connected vertices", None), ("C", "i", ["vi", "ci"], "Indices of connected vertices", None),
             ("niTri", "i", ["vi"], "Number of inverse triangles", None), ("iTri", "i", ["vi", "ci"], "Indices of inverse triangles", None),
             ("edge_index", "i", ["vi"], "Edge index", None), ("Tricc", "d", ["ti", "two"], "Triangle circumcenter", "m"),
             ("TriC", "i", ["ti", "three"], "Triangle neighbours", None), ("Tri_edge_index", "i", ["ti"], "Triangle edge index", None),
             ("VAc", "d", ["aci", "two"], "Staggered vertex coordinates", "m"),
             ("Aci", "i", ["aci", "four"], "Staggered to regular vertex indices", None),
             ("iAci", "i", ["vi", "ci"], "Regular to staggered vertex indices", None), ("VAaAc", "d", ["ai", "two"], "Aa/Ac vertex coordinates", "m"),
             ("TriAaAc", "i", ["tai", "three"], "Aa/Ac vertex indices", None), ("A", "d", ["vi"], "Vertex Voronoi cell area", "m^2"),
             ("R", "d", ["vi"], "Vertex resolution", "m"), ("vi_transect", "i", ["vii", "two"], "Transect vertex pairs", None),
             ("w_transect", "d", ["vii", "two"], "Transect interpolation weights", None),
             ("time", "d", ["time"], "Time", "years"),
             ("zeta", "d", ["zeta"], "Vertical scaled coordinate", "unitless (0 = ice surface, 1 = bedrock)"),
             ("month", "d", ["month"], "Month", "1-12")]
RESTART_VARS = [("Hi", "d", ["vi", "time"], "Ice Thickness", "m"), ("Hb", "d", ["vi", "time"], "Bedrock Height", "m"),
                ("Hs", "d", ["vi", "time"], "Surface Height", "m"), ("U_SIA", "d", ["vi", "time"], "SIA ice x-velocity", "m/yr"),
                ("V_SIA", "d", ["vi", "time"], "SIA ice y-velocity", "m/yr"), ("U_SSA", "d", ["vi", "time"], "SSA ice x-velocity", "m/yr"),
                ("V_SSA", "d", ["vi", "time"], "SSA ice y-velocity", "m/yr"), ("Ti", "d", ["vi", "zeta", "time"], "Ice temperature", "K"),
                ("FirnDepth", "d", ["vi", "month", "time"], "Firn depth", "m"),
                ("MeltPreviousYear", "d", ["vi", "time"], "Melt during previous year", "mie")]


@pytest.fixture(scope="module")
def mesh():
    return M.square_mesh_with_nv(750e3, 600)


def _check_layout(f, expected_vars):
    assert list(f.dimensions)[:11] == MESH_DIMS
    assert list(f.dimensions)[11:] == ["zeta", "month", "time"]
    assert f.dimensions["time"] is None, "time must be the unlimited dimension"
    assert list(f.variables) == [v[0] for v in expected_vars]
    for name, ty, fdims, long_name, units in expected_vars:
        v = f.variables[name]
        assert v.dimensions == tuple(reversed(fdims)), name          # Fortran order is fastest-first, the file slowest-first
        assert v.data.dtype == (np.dtype(">f8") if ty == "d" else np.dtype(">i4")), name
        assert v.long_name == long_name.encode(), name
        assert getattr(v, "units", None) == (units.encode() if units else None), name
        assert set(v._attributes) <= {"long_name", "units"}, name


@pytest.mark.parametrize("force64", [False, True])
def test_restart_file_layout_and_contents(mesh, tmp_path, monkeypatch, force64):
    if force64:
        monkeypatch.setenv("UFM_NC_FORCE_64BIT_OFFSET", "1")
    fn = R.output_filename(str(tmp_path) + "/", "ANT")
    assert os.path.basename(fn) == "restart_ANT_00001.nc"            # get_output_filenames, src/netcdf_module.f90:82
    extra = {"TriC": mesh.TriC, "R": mesh.R, "nV_transect": 3, "vi_transect": np.array([[1, 2], [3, 4], [5, 6]], np.int32),
             "w_transect": np.array([[0.25, 0.75], [0.5, 0.5], [1.0, 0.0]])}
    R.create_restart(fn, mesh, ZETA, extra)
    rng = np.random.default_rng(1)
    frames = []
    for k, t in enumerate([-120000.0, -119000.0, -118000.0]):
        fr = {n: rng.standard_normal(mesh.nV) for n in ("Hi", "Hb", "Hs", "U_SIA", "V_SIA", "U_SSA", "V_SSA", "MeltPreviousYear")}
        fr["Ti"] = np.asfortranarray(250.0 + rng.standard_normal((mesh.nV, 15)))
        fr["FirnDepth"] = np.asfortranarray(rng.random((mesh.nV, 12)))
        assert R.append_restart(fn, t, **fr) == k + 1                # netcdf%ti
        frames.append(fr)
    f = netcdf_file(fn, "r", mmap=False)
    assert f.version_byte == (2 if force64 else 1)                   # nf90_clobber without nf90_64bit_offset: classic
    _check_layout(f, MESH_VARS + RESTART_VARS)
    dims = f.dimensions
    assert (dims["vi"], dims["ti"], dims["ci"], dims["aci"], dims["ciplusone"], dims["vii"], dims["ai"], dims["zeta"], dims["month"]) == \
        (mesh.nV, mesh.nTri, mesh.nC_mem, mesh.nAc, mesh.nC_mem + 1, 3, mesh.nVAaAc, 15, 12)
    for name in ("V", "Tri", "nC", "C", "niTri", "iTri", "edge_index", "Tricc", "Tri_edge_index", "VAc", "Aci", "iAci", "VAaAc", "A"):
        assert np.array_equal(f.variables[name][:], np.asarray(getattr(mesh, name)).T), name
    assert np.array_equal(f.variables["TriC"][:], mesh.TriC.T) and np.array_equal(f.variables["R"][:], mesh.R)
    assert np.array_equal(f.variables["vi_transect"][:], extra["vi_transect"].T) and np.array_equal(f.variables["w_transect"][:], extra["w_transect"].T)
    assert np.array_equal(f.variables["zeta"][:], ZETA) and np.array_equal(f.variables["month"][:], np.arange(1, 13))
    assert np.array_equal(f.variables["time"][:], [-120000.0, -119000.0, -118000.0])
    for k, fr in enumerate(frames):
        for n, a in fr.items():
            assert np.array_equal(f.variables[n][k], a.T), (n, k)
    assert (f.variables["TriAaAc"][:] == -2147483647).all(), "arrays the caller does not have are written as NC_FILL_INT"
    f.close()
    # a second file in the same directory gets the next number; an existing file is never overwritten (:508-512)
    assert os.path.basename(R.output_filename(str(tmp_path) + "/", "ANT")) == "restart_ANT_00002.nc"
    with pytest.raises(UfmError) as e:
        R.create_restart(fn, mesh, ZETA)
    assert e.value.rc == -13 and "already exists" in str(e.value)


def test_restart_read_back_and_frame_selection(mesh, tmp_path):
    fn = str(tmp_path / "restart_GRL_00001.nc")
    R.create_restart(fn, mesh, ZETA, {"TriC": mesh.TriC})
    times = [0.0, 100.0, 250.0, 1000.0]
    for t in times:
        R.append_restart(fn, t, Hi=np.full(mesh.nV, t), Hb=np.full(mesh.nV, -t), Hs=np.zeros(mesh.nV), U_SSA=np.full(mesh.nV, 2 * t),
                         V_SSA=np.full(mesh.nV, 3 * t), Ti=np.full((mesh.nV, 15), 200.0 + t), FirnDepth=np.full((mesh.nV, 12), t / 7),
                         MeltPreviousYear=np.full(mesh.nV, t / 3), U_SIA=np.zeros(mesh.nV), V_SIA=np.zeros(mesh.nV))
    assert R.inquire_restart_mesh(fn) == (mesh.nV, mesh.nTri, mesh.nC_mem)
    prim = R.read_restart_mesh(fn)
    for n in ("V", "nC", "C", "niTri", "iTri", "edge_index", "Tri", "Tricc", "TriC", "Tri_edge_index"):
        assert np.array_equal(prim[n], np.asarray(getattr(mesh, n))), n
    assert R.inquire_restart_init(fn, ZETA) == (4, True)
    z2 = list(ZETA); z2[3] += 0.01
    assert R.inquire_restart_init(fn, z2) == (4, False)              # the reference only warns (:3084-3088)
    with pytest.raises(UfmError) as e:
        R.inquire_restart_init(fn, ZETA[:10])
    assert e.value.rc == -14 and "nZ in restart file doesnt match" in str(e.value)
    # read_restart_file_init: nearest frame; ties go to the earlier one (strict `dt < dt_min`, :3163)
    for t_req, ti_want in [(0.0, 1), (40.0, 1), (60.0, 2), (175.0, 2), (176.0, 3), (1000.0, 4), (700.0, 4), (624.0, 3)]:
        out, ti = R.read_restart_init(fn, t_req, mesh.nV, 15)
        t = times[ti_want - 1]
        assert ti == ti_want, t_req
        assert (out["Hi"] == t).all() and (out["Hb"] == -t).all() and (out["U_SSA"] == 2 * t).all() and (out["V_SSA"] == 3 * t).all()
        assert (out["Ti"] == 200.0 + t).all() and (out["FirnDepth"] == t / 7).all() and (out["MeltPreviousYear"] == t / 3).all()
    for bad in (-1.0, 1000.5):
        with pytest.raises(UfmError) as e:
            R.read_restart_init(fn, bad, mesh.nV, 15)
        assert e.value.rc == -15 and "outside range of restart file" in str(e.value)
    with pytest.raises(UfmError) as e:
        R.inquire_restart_mesh(str(tmp_path / "nope.nc"))
    assert e.value.rc == -11
    junk = tmp_path / "junk.nc"
    junk.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)               # a NetCDF-4 / HDF5 signature
    with pytest.raises(UfmError) as e:
        R.inquire_restart_mesh(str(junk))
    assert e.value.rc == -12


@pytest.mark.parametrize("version", [1, 2])
def test_reader_accepts_files_from_an_independent_writer(mesh, tmp_path, version):
    """A restart file produced by another NetCDF implementation (scipy), with extra variables, attributes and a different
    variable order than ours, reads correctly: the reader goes by names, as nf90_inq_varid does."""
    fn = str(tmp_path / "restart_scipy.nc")
    f = netcdf_file(fn, "w", version=version)
    f.history = "written by scipy"
    nV, nTri, W = mesh.nV, mesh.nTri, mesh.nC_mem
    for n, l in [("time", None), ("vi", nV), ("ti", nTri), ("ci", W), ("two", 2), ("three", 3), ("zeta", 15), ("month", 12)]:
        f.createDimension(n, l)
    rng = np.random.default_rng(5)
    data = {"Hi": rng.random((2, nV)), "Hb": rng.random((2, nV)), "Hs": rng.random((2, nV)), "U_SSA": rng.random((2, nV)),
            "V_SSA": rng.random((2, nV)), "MeltPreviousYear": rng.random((2, nV)), "Ti": rng.random((2, 15, nV)), "FirnDepth": rng.random((2, 12, nV))}
    shapes = {"Ti": ("time", "zeta", "vi"), "FirnDepth": ("time", "month", "vi")}
    for n in ("FirnDepth", "Hi", "extra_field", "Ti", "Hb", "Hs", "U_SSA", "V_SIA", "V_SSA", "U_SIA", "MeltPreviousYear"):
        v = f.createVariable(n, "d", shapes.get(n, ("time", "vi")))
        v.units = "whatever"
        v[:] = data.get(n, np.zeros((2, nV)))
    f.createVariable("time", "d", ("time",))[:] = [10.0, 20.0]
    f.createVariable("zeta", "d", ("zeta",))[:] = ZETA
    f.createVariable("month", "d", ("month",))[:] = np.arange(1, 13)
    for n, ty, dims in [("Tri_edge_index", "i", ("ti",)), ("V", "d", ("two", "vi")), ("nC", "i", ("vi",)), ("C", "i", ("ci", "vi")),
                        ("niTri", "i", ("vi",)), ("iTri", "i", ("ci", "vi")), ("edge_index", "i", ("vi",)), ("Tri", "i", ("three", "ti")),
                        ("Tricc", "d", ("two", "ti")), ("TriC", "i", ("three", "ti"))]:
        f.createVariable(n, ty, dims)[:] = np.asarray(getattr(mesh, n)).T
    f.close()
    assert R.inquire_restart_mesh(fn) == (nV, nTri, W)
    prim = R.read_restart_mesh(fn)
    for n in ("V", "nC", "C", "niTri", "iTri", "edge_index", "Tri", "Tricc", "TriC", "Tri_edge_index"):
        assert np.array_equal(prim[n], np.asarray(getattr(mesh, n))), n
    assert R.inquire_restart_init(fn, ZETA) == (2, True)
    out, ti = R.read_restart_init(fn, 16.0, nV, 15)
    assert ti == 2
    for n, a in data.items():
        assert np.array_equal(out[n], a[1].T), n
    # appending to a file made by another writer keeps it valid for that writer's reader
    assert R.append_restart(fn, 30.0, Hi=np.full(nV, 7.0)) == 3
    g = netcdf_file(fn, "r", mmap=False)
    assert np.array_equal(g.variables["time"][:], [10.0, 20.0, 30.0]) and (g.variables["Hi"][2] == 7.0).all()
    assert np.array_equal(g.variables["Hi"][1], data["Hi"][1]) and np.array_equal(g.variables["extra_field"][1], np.zeros(nV))
    g.close()


def test_restart_inquire_rejects_wrong_types_and_dimensions(mesh, tmp_path):
    """inquire_double_var / inquire_int_var (src/netcdf_module.f90:3694-3795) STOP on a type or dimension mismatch."""
    def make(path, v_type="d", c_dims=("ci", "vi"), drop=None):
        f = netcdf_file(path, "w")
        for n, l in [("vi", mesh.nV), ("ti", mesh.nTri), ("ci", mesh.nC_mem), ("two", 2), ("three", 3)]:
            f.createDimension(n, l)
        for n, ty, dims in [("V", v_type, ("two", "vi")), ("nC", "i", ("vi",)), ("C", "i", c_dims), ("niTri", "i", ("vi",)), ("iTri", "i", ("ci", "vi")),
                            ("edge_index", "i", ("vi",)), ("Tri", "i", ("three", "ti")), ("Tricc", "d", ("two", "ti")), ("TriC", "i", ("three", "ti")),
                            ("Tri_edge_index", "i", ("ti",))]:
            if n != drop:
                f.createVariable(n, ty, dims)
        f.close()

    for k, (kw, msg) in enumerate([({"v_type": "f"}, "is not nf90_DOUBLE"), ({"c_dims": ("vi", "ci")}, "does not match required dimensions"),
                                   ({"c_dims": ("vi",)}, "does not match required number of dimensions"), ({"drop": "TriC"}, "Variable not found")]):
        p = str(tmp_path / f"bad{k}.nc")
        make(p, **kw)
        with pytest.raises(UfmError) as e:
            R.inquire_restart_mesh(p)
        assert e.value.rc == -12 and msg in str(e.value), (kw, str(e.value))


def test_help_fields_file(mesh, tmp_path):
    fn = R.output_filename(str(tmp_path) + "/", "NAM", kind="help_fields")
    assert os.path.basename(fn) == "help_fields_NAM_00001.nc"
    names = ["lat", "lon", "none", "resolution", "Hi", "Hs", "Ti", "T2m", "mask", "mask_gl", "GHF", "U_SSA", "D_SIA_3D", "dHb", "phi_fric", "none"]
    R.create_help_fields(fn, mesh, ZETA, names, {"TriC": mesh.TriC})
    f = netcdf_file(fn, "r", mmap=False)
    want = MESH_VARS + [("lat", "d", ["vi"], "Latitude", "degrees north"), ("lon", "d", ["vi"], "Longitude", "degrees east"),
                        ("Hi", "d", ["vi", "time"], "Ice thickness", "m"), ("Hs", "d", ["vi", "time"], "Surface elevation", "m w.r.t PD sealevel"),
                        ("Ti", "d", ["vi", "zeta", "time"], "Englacial temperature", "K"),
                        ("T2m", "d", ["vi", "month", "time"], "Monthly mean 2-m air temperature", "K"), ("mask", "i", ["vi", "time"], "mask", None),
                        ("mask_gl", "i", ["vi", "time"], "grounding-line mask", None), ("GHF", "d", ["vi"], "Geothermal heat flux", "J m^-2 yr^-1"),
                        ("U_SSA", "d", ["vi", "time"], "Vertically averaged SSA ice x-velocity", "m/yr"),
                        ("D_SIA_3D", "d", ["vi", "zeta", "time"], "3D SIA ice diffusivity", None),
                        ("dHb", "d", ["vi", "time"], "Change in bedrock elevation w.r.t. PD", "m"),
                        ("phi_fric", "d", ["vi", "time"], "till friction angle", "degrees")]
    _check_layout(f, want)
    f.close()
    rng = np.random.default_rng(2)
    host = {"lat": rng.random(mesh.nV), "lon": rng.random(mesh.nV), "Hi": rng.random(mesh.nV), "Hs": rng.random(mesh.nV),
            "Ti": np.asfortranarray(rng.random((mesh.nV, 15))), "T2m": np.asfortranarray(rng.random((mesh.nV, 12))),
            "mask": rng.integers(0, 8, mesh.nV).astype(np.int32), "mask_gl": rng.integers(0, 2, mesh.nV).astype(np.int32),
            "GHF": rng.random(mesh.nV), "U_SSA": rng.random(mesh.nV), "D_SIA_3D": np.asfortranarray(rng.random((mesh.nV, 15))),
            "dHb": rng.random(mesh.nV), "phi_fric": rng.random(mesh.nV)}
    live = [n for n in names if n not in ("none", "resolution")]
    assert R.write_help_fields_host(fn, 5.0, names, {**host, "none": np.zeros(1), "resolution": np.zeros(1)}) == 1
    # second frame writes only two of the fields: the others keep the fill value in that frame
    assert R.write_help_fields_host(fn, 6.0, ["Hi", "mask"], {"Hi": host["Hi"] * 2, "mask": host["mask"] + 1}) == 2
    f = netcdf_file(fn, "r", mmap=False)
    assert np.array_equal(f.variables["time"][:], [5.0, 6.0])
    for n in live:
        v = f.variables[n]
        got = v[:] if "time" not in v.dimensions else v[0]
        assert np.array_equal(got, host[n].T), n
    assert np.array_equal(f.variables["Hi"][1], host["Hi"] * 2) and np.array_equal(f.variables["mask"][1], host["mask"] + 1)
    assert (f.variables["Hs"][1] == np.float64(9.9692099683868690e+36)).all() and (f.variables["mask_gl"][1] == -2147483647).all()
    f.close()
    with pytest.raises(UfmError) as e:
        R.create_help_fields(str(tmp_path / "x.nc"), mesh, ZETA, ["Hi", "no_such_field"])
    assert e.value.rc == -16 and "not implemented in create_help_field_mesh" in str(e.value)
    assert not os.path.exists(tmp_path / "x.nc")
    with pytest.raises(UfmError) as e:                                # a field that is not in the file
        R.write_help_fields_host(fn, 7.0, ["Hb"], {"Hb": host["Hi"]})
    assert e.value.rc == -12
    with pytest.raises(UfmError) as e:                                # device-resident field without a handle or a host array
        import ctypes
        from ufemism_b200.capi import load_library
        arr = (ctypes.c_char_p * 1)(b"Hi")
        rc = load_library().ufm_help_fields_write(None, os.fsencode(fn), 8.0, 1, arr, None)
        raise UfmError(rc, load_library().ufm_last_error().decode())
    assert e.value.rc == -2


def test_every_help_field_name_of_the_reference_is_known(mesh, tmp_path):
    """The 74 names create_help_field_mesh accepts (src/netcdf_module.f90:850-1033); 50 per file at most (C%help_field_01..50)."""
    names = ["lat", "lon", "GHF", "Hi", "Hb", "Hs", "SL", "dHs_dx", "dHs_dy", "Ti", "Cpi", "Ki", "Ti_basal", "Ti_pmp", "A_flow", "A_flow_mean",
             "U_SIA", "V_SIA", "U_SSA", "V_SSA", "U_vav", "V_vav", "U_surf", "V_surf", "U_base", "V_base", "U_3D", "V_3D", "W_3D", "D_SIA", "D_SIA_3D",
             "T2m", "T2m_year", "Precip", "Precip_year", "Wind_WE", "Wind_WE_year", "Wind_SN", "Wind_SN_year", "SMB", "SMB_year", "BMB_sheet",
             "BMB_shelf", "BMB", "Snowfall", "Snowfall_year", "Rainfall", "Rainfall_year", "AddedFirn", "AddedFirn_year", "Refreezing",
             "Refreezing_year", "Runoff", "Runoff_year", "Albedo", "Albedo_year", "FirnDepth", "FirnDepth_year", "mask", "mask_land", "mask_ocean",
             "mask_lake", "mask_ice", "mask_sheet", "mask_shelf", "mask_coast", "mask_margin", "mask_gl", "mask_cf", "phi_fric", "tau_yield",
             "iso_ice", "iso_surf", "dHb"]
    assert len(names) == 74
    fn = str(tmp_path / "all.nc")
    R.create_help_fields(fn, mesh, ZETA, names)
    f = netcdf_file(fn, "r", mmap=False)
    assert list(f.variables)[len(MESH_VARS):] == names
    ints = {n for n in names if n.startswith("mask")}
    for n in names:
        v = f.variables[n]
        assert v.data.dtype == (np.dtype(">i4") if n in ints else np.dtype(">f8")), n
        assert v.dimensions[-1] == "vi" and (v.dimensions[0] == "time" or n in ("lat", "lon", "GHF")), n
    f.close()
    with pytest.raises(UfmError):                                     # the same name twice: NetCDF "name in use"
        R.create_help_fields(str(tmp_path / "dup.nc"), mesh, ZETA, ["Hi", "Hi"])


# ---- row N3: secondary mesh data derived inside the library from the primary data a restart file (or a mesh update) holds ----
def test_secondary_mesh_data_derived_from_a_restart_file(mesh, tmp_path):
    """restart file -> ufm_restart_read_mesh -> ufm_mesh_derive_secondary reproduces every array of the mesh the file was written
    from (the host-only half of ufm_mesh_upload_primary; the reference does the same with read_mesh_from_restart_file,
    src/restart_module.f90:31-116)."""
    from ufemism_b200 import capi

    fn = str(tmp_path / "restart_ANT_00001.nc")
    R.create_restart(fn, mesh, ZETA, {"TriC": mesh.TriC})
    prim = R.read_restart_mesh(fn)
    prim.update(xmin=mesh.xmin, xmax=mesh.xmax, ymin=mesh.ymin, ymax=mesh.ymax)
    d = capi.derive_secondary(prim, thermo=True)
    assert (d["nV"], d["nAc"], d["nVAaAc"]) == (mesh.nV, mesh.nAc, mesh.nVAaAc)
    assert d["ldAc"] == mesh.nV + mesh.nTri and mesh.nAc == mesh.nV + mesh.nTri - 1, "Euler: a triangulated disc has nV + nTri - 1 edges"
    for n in ("A", "Cw", "Aci", "iAci", "edge_index_Ac", "nCAaAc", "CAaAc", "colour", "colour_vi", "colour_nV", "Tricc", "Tri_edge_index", "VAc",
              "VAaAc", "R", "NxTri", "NyTri"):
        assert np.array_equal(d[n], np.asarray(getattr(mesh, n))), n
    assert d["nf_pointers_null"], "neighbour functions are left to the device"
    # leading dimensions larger than the mesh (the reference's nV_mem-sized arrays before crop_mesh_primary)
    pad = 7
    big = dict(prim)
    for n in ("V", "C", "iTri"):
        a = np.zeros((mesh.nV + pad, prim[n].shape[1]), prim[n].dtype, order="F")
        a[: mesh.nV] = prim[n]
        big[n] = a
    p, keep = capi.mesh_primary(big)
    p.nV = mesh.nV                                                   # mesh_primary took nV from the padded array
    import ctypes
    L = capi.load_library()
    obj = ctypes.c_void_p()
    assert L.ufm_mesh_derive_secondary(ctypes.byref(p), ctypes.byref(obj)) == 0, L.ufm_last_error()
    d2 = capi._derived_to_dict(L, L.ufm_mesh_derived_get, obj, mesh.nV, mesh.nTri, mesh.nC_mem)
    L.ufm_mesh_derived_free(obj)
    for n in ("A", "Cw", "Aci", "CAaAc", "colour_vi"):
        assert np.array_equal(d2[n], d[n]), n


def test_derived_arrays_are_the_same_when_the_previous_meshs_buffers_are_reused():
    """ufm_mesh_upload_primary hands the previous mesh's derived object back to the derivation, which reuses its buffers (no fresh
    allocations, nothing zero-filled that will be overwritten).  Whatever mesh was there before -- larger, smaller, the same -- every
    derived array, the colouring included, equals that of a derivation into a fresh object; a failed derivation leaves nothing behind."""
    import ctypes

    from ufemism_b200 import capi
    from ufemism_b200 import mesh as M

    L = capi.load_library()
    meshes = [M.square_mesh_with_nv(750e3, n, seed=s) for n, s in ((1500, 3), (2600, 4), (900, 5), (900, 5))]
    obj = ctypes.c_void_p()
    names = ("A", "Cw", "Aci", "iAci", "edge_index_Ac", "nCAaAc", "CAaAc", "colour", "colour_vi", "colour_nV", "Tricc", "Tri_edge_index", "VAc", "VAaAc", "R", "NxTri", "NyTri")
    for m in meshes:
        p, keep = capi.mesh_primary(m, True)
        assert L.ufm_mesh_derive_secondary_reuse(ctypes.byref(p), ctypes.byref(obj)) == 0, L.ufm_last_error()
        d = capi._derived_to_dict(L, L.ufm_mesh_derived_get, obj, m.nV, m.nTri, m.nC_mem)
        fresh = capi.derive_secondary(m, thermo=True)
        assert (d["nV"], d["nAc"], d["ldAc"]) == (fresh["nV"], fresh["nAc"], fresh["ldAc"])
        for n in names:
            assert np.array_equal(d[n], fresh[n]), (m.nV, n)
            assert np.array_equal(d[n], np.asarray(getattr(m, n))), (m.nV, n)
    # without the thermodynamics arrays after a mesh that had them: the descriptor must not point at the old ones
    p, keep = capi.mesh_primary(meshes[0], False)
    assert L.ufm_mesh_derive_secondary_reuse(ctypes.byref(p), ctypes.byref(obj)) == 0
    d = capi._derived_to_dict(L, L.ufm_mesh_derived_get, obj, meshes[0].nV, meshes[0].nTri, meshes[0].nC_mem)
    assert "R" not in d and np.array_equal(d["CAaAc"], np.asarray(meshes[0].CAaAc))
    # a broken mesh: error, and the object is gone (nothing stale to hand out)
    bad = {n: np.array(getattr(meshes[0], n), order="F") for n in ("V", "nC", "C", "niTri", "iTri", "edge_index", "Tri")}
    bad.update(xmin=meshes[0].xmin, xmax=meshes[0].xmax, ymin=meshes[0].ymin, ymax=meshes[0].ymax)
    bad["C"][5, 0] = 0
    p, keep = capi.mesh_primary(bad)
    assert L.ufm_mesh_derive_secondary_reuse(ctypes.byref(p), ctypes.byref(obj)) != 0
    assert not obj.value


def test_derive_secondary_rejects_broken_primary_data(mesh):
    from ufemism_b200 import capi

    base = {n: np.array(getattr(mesh, n), order="F") for n in ("V", "nC", "C", "niTri", "iTri", "edge_index", "Tri")}
    base.update(xmin=mesh.xmin, xmax=mesh.xmax, ymin=mesh.ymin, ymax=mesh.ymax)

    def broken(**kw):
        d = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in base.items()}
        for k, f in kw.items():
            d[k] = f(d[k])
        return d

    def poke(i, j, val):
        def f(a):
            if a.ndim == 1:
                a[i] = val
            else:
                a[i, j] = val
            return a
        return f

    cases = [broken(C=poke(10, 0, mesh.nV + 5)), broken(C=poke(10, 0, 0)), broken(nC=poke(3, None, mesh.nC_mem + 1)), broken(nC=poke(3, None, 1)),
             broken(iTri=poke(7, 0, mesh.nTri + 1)), broken(Tri=poke(2, 1, 0)), broken(edge_index=poke(5, None, 9)), broken(xmax=lambda x: mesh.xmin),
             broken(Tri=poke(2, 1, int(mesh.Tri[2, 0])))]           # the last one: a degenerate triangle -> an edge without its triangle
    for k, d in enumerate(cases):
        with pytest.raises(UfmError) as e:
            capi.derive_secondary(d)
        assert e.value.rc == -2, (k, str(e.value))


def test_relabelled_five_colouring_gives_the_same_colouring(mesh):
    """ufm_mesh_upload_primary runs calculate_five_colouring_AaAc on a Morton relabelling of the graph (cache locality).  The
    relabelled run must take the decisions of the plain run: same colour for every vertex, for any permutation."""
    lib = M._load()
    Mv, p = mesh.nVAaAc, M._p
    rng = np.random.default_rng(11)
    by_x = np.empty(Mv, np.int32)
    by_x[np.argsort(mesh.VAaAc[:, 0], kind="stable")] = np.arange(1, Mv + 1, dtype=np.int32)
    for label in (None, (rng.permutation(Mv) + 1).astype(np.int32), by_x, np.arange(Mv, 0, -1, dtype=np.int32)):
        colour, cvi, cn = np.zeros(Mv, np.int32), np.zeros((Mv, 5), np.int32, order="F"), np.zeros(5, np.int32)
        rc = lib.ufm_mesh_five_colouring_labelled(Mv, mesh.nC_mem, p(mesh.nCAaAc), p(mesh.CAaAc), None if label is None else p(label), p(colour), p(cvi), p(cn))
        assert rc == 0
        assert np.array_equal(colour, mesh.colour) and np.array_equal(cvi, mesh.colour_vi) and np.array_equal(cn, mesh.colour_nV)


# ---- live cross-check of the file layouts against the reference's source text (where it is mounted) ----
REF_SRC = "/root/reference/src"


def _reference_layout(routine, type_name):
    """(dimension list, variable list) as `routine` of netcdf_module.f90 defines them, names resolved through the
    `name_dim_* / name_var_*` defaults of `type_name` in data_types_netcdf_module.f90."""
    import re

    from oracle import f90py as F
    st = [s for _, s in F.extract_unit(open(os.path.join(REF_SRC, "netcdf_module.f90")).read(), routine)]
    types = open(os.path.join(REF_SRC, "data_types_netcdf_module.f90")).read()
    body = types[types.index(f"TYPE {type_name}"):types.index(f"END TYPE {type_name}")]
    names = {m.group(1): m.group(2).strip() for m in re.finditer(r"::\s*(name_(?:dim|var)_\w+)\s*=\s*'([^']*)'", body)}
    dims, variables, alias = [], [], {}
    for s in st:
        m = re.match(r"CALL create_dim\(\s*netcdf%ncid,\s*netcdf%(name_dim_\w+)\s*,\s*(.+?),\s*netcdf%(id_dim_\w+)\s*\)", s)
        if m:
            dims.append(names[m.group(1)])
            alias["netcdf%" + m.group(3)] = names[m.group(1)]
            continue
        m = re.match(r"(\w+)\s*=\s*netcdf%(id_dim_\w+)$", s)          # "vi = netcdf%id_dim_vi": short aliases used in the definitions
        if m:
            alias[m.group(1)] = alias["netcdf%" + m.group(2)]
            continue
        m = re.match(r"CALL create_(double|int)_var\(\s*netcdf%ncid,\s*netcdf%(name_var_\w+)\s*,\s*\[([^\]]*)\],\s*netcdf%id_var_\w+\s*,\s*long_name='([^']*)'(?:,\s*units='([^']*)')?\s*\)", s)
        if m:
            variables.append((names[m.group(2)], m.group(1)[0], [alias[d.strip()] for d in m.group(3).split(",")], m.group(4), m.group(5)))
    return dims, variables


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="/root/reference is not mounted here")
def test_live_restart_layout_is_the_reference_sources(mesh, tmp_path):
    """Dimensions and variables of a file written by ufm_restart_create, read back with scipy, against what create_restart_file_mesh
    defines in the reference's source: same names, order, types, dimension order, long_name and units."""
    dims, variables = _reference_layout("create_restart_file_mesh", "type_netcdf_restart")
    assert len(dims) == 14 and len(variables) == 32
    fn = str(tmp_path / "restart_ANT_00001.nc")
    R.create_restart(fn, mesh, ZETA)
    f = netcdf_file(fn, "r", mmap=False)
    assert list(f.dimensions) == dims
    assert list(f.variables) == [v[0] for v in variables]
    for name, ty, fdims, long_name, units in variables:
        v = f.variables[name]
        assert v.dimensions == tuple(reversed(fdims)), name
        assert v.data.dtype == (np.dtype(">f8") if ty == "d" else np.dtype(">i4")), name
        assert v.long_name == long_name.encode() and getattr(v, "units", None) == (units.encode() if units else None), name
    f.close()
    # the hard-coded expectation of the non-live tests above is that very list
    assert [(n, t, d, l, u) for n, t, d, l, u in MESH_VARS + RESTART_VARS] == variables


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="/root/reference is not mounted here")
def test_live_help_field_table_is_the_reference_sources(mesh, tmp_path):
    """Every field create_help_field_mesh knows (name, type, dimensions, long_name, units), parsed from the reference's source,
    against the file ufm_help_fields_create writes when asked for all of them."""
    import re

    from oracle import f90py as F
    st = [s for _, s in F.extract_unit(open(os.path.join(REF_SRC, "netcdf_module.f90")).read(), "create_help_field_mesh")]
    table = []
    for s in st:
        m = re.search(r"CALL create_(double|int)_var\(\s*netcdf%ncid,\s*'(\w+)',\s*\[([^\]]*)\],\s*id_var,\s*long_name='([^']*)'(?:,\s*units='([^']*)')?", s)
        if m:
            fd = [{"vi": "vi", "t": "time", "z": "zeta", "m": "month"}[d.strip()] for d in m.group(3).split(",")]
            table.append((m.group(2), m.group(1)[0], fd, m.group(4), m.group(5)))
    assert len(table) == 74
    fn = str(tmp_path / "help_fields_ANT_00001.nc")
    R.create_help_fields(fn, mesh, ZETA, [t[0] for t in table])
    f = netcdf_file(fn, "r", mmap=False)
    assert list(f.variables)[len(MESH_VARS):] == [t[0] for t in table]
    for name, ty, fdims, long_name, units in table:
        v = f.variables[name]
        assert v.dimensions == tuple(reversed(fdims)), name
        assert v.data.dtype == (np.dtype(">f8") if ty == "d" else np.dtype(">i4")), name
        assert v.long_name == long_name.encode() and getattr(v, "units", None) == (units.encode() if units else None), name
    f.close()
    # and the mesh part of the help_fields file is defined exactly like the restart file's
    d1, v1 = _reference_layout("create_restart_file_mesh", "type_netcdf_restart")
    d2, v2 = _reference_layout("create_help_fields_file_mesh", "type_netcdf_help_fields")
    assert d1 == d2 and v1[:len(MESH_VARS)] == v2 == [(n, t, d, l, u) for n, t, d, l, u in MESH_VARS]
