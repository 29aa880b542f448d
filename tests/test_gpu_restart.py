"""GPU half of rows N3 / N4: device upload from primary mesh data, restart and help_fields files written from and
loaded onto the device.  Bit-exact throughout: files carry the fp64 fields unchanged, and a mesh whose secondary data were
derived inside the library must behave exactly like one whose secondary data came from the host."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from tests.util import assert_bits_equal, make_gpu
from ufemism_b200 import restart as R
from ufemism_b200 import scenarios as S

pytestmark = pytest.mark.gpu

ZETA = [0.00, 0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 0.925, 0.95, 0.975, 0.99, 1.00]
FIELDS = ("Hi", "Hs", "U_SIA", "V_SIA", "D_SIA", "U_SSA", "V_SSA", "dHs_dx", "dHs_dy", "mask", "Hi_Ac", "dHs_dp_Ac", "Up_SIA_Ac", "Up_SSA_Ac",
          "U_SSA_AaAc", "tau_c_AaAc", "eta_AaAc")


def test_upload_from_primary_mesh_data_is_bit_identical(mesh_2k):
    """ufm_mesh_upload_primary (areas, Cw, Ac / AaAc meshes, colouring derived on the host inside the library, every neighbour
    function on the device) vs ufm_mesh_upload with all arrays built by the host: the same hybrid SIA/SSA run, bit for bit."""
    from tests.test_gpu_parity import scenario

    st = scenario(mesh_2k, "mismip")
    a, b = make_gpu(mesh_2k, st), make_gpu(mesh_2k, st, primary_only=True)
    ra, rb = a.region(0.0), b.region(0.0)
    a.run_model(ra, 2.0)
    b.run_model(rb, 2.0)
    assert (ra.n_steps, ra.n_sor_total, ra.n_outer_total) == (rb.n_steps, rb.n_sor_total, rb.n_outer_total) and ra.n_sor_total > 0
    for f in FIELDS:
        assert_bits_equal(b.download(f), a.download(f), f)
    sec = b.secondary()
    for n in ("A", "Cw", "Aci", "iAci", "nCAaAc", "CAaAc", "colour_vi", "colour_nV", "edge_index_Ac"):
        assert np.array_equal(sec[n], np.asarray(getattr(mesh_2k, n))), n
    # a plain upload afterwards drops the derived arrays (they described the replaced mesh)
    from ufemism_b200.capi import UfmError
    b.primary_only = False
    b.upload_mesh(mesh_2k)
    with pytest.raises(UfmError):
        b.secondary()


def test_primary_upload_with_thermodynamics(mesh_2k):
    """thermo = 1: R, NxTri, NyTri derived inside the library; update_ice_temperature bit-identical to the host-built mesh."""
    from tests.test_gpu_parity import THERMO_IN
    from ufemism_b200.capi import IceModelGPU

    st = S.state_thermo_dome(mesh_2k, benchmark="none")
    g, g2 = IceModelGPU(mesh_2k, benchmark="none", thermo=True), IceModelGPU(mesh_2k, benchmark="none", thermo=True, primary_only=True)
    for f in THERMO_IN:
        g.upload(f, st[f]); g2.upload(f, st[f])
    for m in (g, g2):
        m.update_general_ice_model_data(0.0); m.solve_SIA(); m.solve_SSA()
        m.update_ice_temperature()
    assert np.abs(g.download("W_3D")).max() > 0.0 and np.ptp(g.download("Ti")) > 1.0
    for f in ("Ti", "W_3D", "U_3D", "frictional_heating"):
        assert_bits_equal(g2.download(f), g.download(f), f)


def test_restart_and_help_fields_written_from_the_device(mesh_2k, tmp_path):
    from tests.test_gpu_parity import scenario

    st = scenario(mesh_2k, "mismip")
    g = make_gpu(mesh_2k, st)
    fn = R.output_filename(str(tmp_path) + "/", "ANT")
    hf = R.output_filename(str(tmp_path) + "/", "ANT", kind="help_fields")
    names = ["Hi", "Hb", "Hs", "SL", "dHs_dx", "dHs_dy", "U_SIA", "V_SIA", "U_SSA", "V_SSA", "D_SIA", "A_flow_mean", "SMB_year", "BMB", "mask",
             "mask_land", "mask_ocean", "mask_ice", "mask_sheet", "mask_shelf", "mask_gl", "mask_cf", "mask_margin", "mask_coast", "mask_lake",
             "phi_fric", "tau_yield", "U_3D", "V_3D", "resolution", "none", "lat"]
    R.create_restart(fn, mesh_2k, ZETA, {"TriC": mesh_2k.TriC})
    R.create_help_fields(hf, mesh_2k, ZETA, names, {"TriC": mesh_2k.TriC})
    assert not g.field_resident("Ti") and g.field_resident("Hi") and g.field_resident("A_flow_mean")
    r = g.region(0.0)
    snaps = []
    lat = np.linspace(-80.0, -60.0, mesh_2k.nV)
    for k, t_end in enumerate([1.0, 2.0, 3.0]):
        g.run_model(r, t_end)
        melt = np.full(mesh_2k.nV, 0.25 * k)
        assert g.write_restart(fn, r.time, MeltPreviousYear=melt) == k + 1
        assert g.write_help_fields(hf, r.time, names, host={"lat": lat}) == k + 1
        snaps.append({n: g.download(n) for n in ("Hi", "Hb", "Hs", "U_SIA", "V_SIA", "U_SSA", "V_SSA", "SL", "dHs_dx", "D_SIA", "SMB_year", "mask",
                                                 "mask_gl", "mask_shelf", "U_3D", "tau_c_AaAc", "phi_fric_AaAc", "A_flow_mean")} | {"time": r.time, "melt": melt})
    f = netcdf_file(fn, "r", mmap=False)
    h = netcdf_file(hf, "r", mmap=False)
    assert np.array_equal(f.variables["time"][:], [s["time"] for s in snaps]) and np.array_equal(h.variables["time"][:], f.variables["time"][:])
    for k, s in enumerate(snaps):
        for n in ("Hi", "Hb", "Hs", "U_SIA", "V_SIA", "U_SSA", "V_SSA"):
            assert_bits_equal(f.variables[n][k], s[n], f"restart {n}[{k}]")
        assert_bits_equal(f.variables["MeltPreviousYear"][k], s["melt"], "MeltPreviousYear")
        assert (f.variables["Ti"][k] == np.float64(9.9692099683868690e+36)).all(), "Ti is not resident for this benchmark: fill value"
        for n in ("Hi", "Hs", "SL", "dHs_dx", "D_SIA", "SMB_year", "mask", "mask_gl", "mask_shelf", "U_SSA", "A_flow_mean"):
            assert_bits_equal(h.variables[n][k], s[n], f"help_fields {n}[{k}]")
        assert_bits_equal(h.variables["U_3D"][k], s["U_3D"].T, "U_3D")
        assert_bits_equal(h.variables["tau_yield"][k], s["tau_c_AaAc"][: mesh_2k.nV], "tau_yield")
        assert_bits_equal(h.variables["phi_fric"][k], s["phi_fric_AaAc"][: mesh_2k.nV], "phi_fric")
    assert_bits_equal(h.variables["lat"][:], lat, "lat")
    assert snaps[2]["U_SSA"].any() and snaps[2]["mask_shelf"].any() and not np.array_equal(snaps[0]["Hi"], snaps[2]["Hi"])
    f.close(); h.close()

    # restart: mesh from the file's primary data, state from the frame nearest to the requested time
    prim = R.read_restart_mesh(fn)
    prim.update(xmin=mesh_2k.xmin, xmax=mesh_2k.xmax, ymin=mesh_2k.ymin, ymax=mesh_2k.ymax)

    class PrimMesh:                                                   # what IceModelGPU needs of a mesh when it uploads primary data
        nV, nTri, nC_mem, nAc, nVAaAc = mesh_2k.nV, mesh_2k.nTri, mesh_2k.nC_mem, mesh_2k.nAc, mesh_2k.nVAaAc
        xmin, xmax, ymin, ymax = mesh_2k.xmin, mesh_2k.xmax, mesh_2k.ymin, mesh_2k.ymax
    for n, a in prim.items():
        setattr(PrimMesh, n, a)
    from ufemism_b200.capi import IceModelGPU, UfmError
    g2 = IceModelGPU(PrimMesh, benchmark=st["benchmark"], primary_only=True)
    with pytest.raises(UfmError) as e:
        g2.load_restart(fn, snaps[2]["time"] + 1.0)
    assert e.value.rc == -15
    assert g2.load_restart(fn, 0.5 * (snaps[1]["time"] + snaps[2]["time"]) + 1e-6) == 3
    for n in ("Hi", "Hb", "U_SSA", "V_SSA"):
        assert_bits_equal(g2.download(n), snaps[2][n], f"loaded {n}")
    # what the reference recomputes after a restart is then identical to the original run's
    for n in ("SL", "SMB_year", "BMB"):
        g2.upload(n, st[n])
    g2.update_general_ice_model_data(r.time); g.update_general_ice_model_data(r.time)
    g2.solve_SIA(); g.solve_SIA()
    for n in ("Hs", "dHs_dx", "mask", "mask_gl", "Hi_Ac", "U_SIA", "V_SIA", "D_SIA"):
        assert_bits_equal(g2.download(n), g.download(n), f"after restart {n}")


def test_files_of_another_mesh_or_vertical_grid_are_refused(mesh_2k, tmp_path):
    """A restart / help_fields file whose `vi` or `zeta` dimension differs from the resident mesh / config (a file of the mesh before a mesh
    update, another nZ) is refused with an error before anything is moved, as the reference stops in NetCDF on a shape mismatch -- the frame
    buffers are sized from the file."""
    from tests.conftest import get_mesh
    from tests.test_gpu_parity import scenario
    from ufemism_b200.capi import UfmError

    other = get_mesh(3000)
    g = make_gpu(mesh_2k, scenario(mesh_2k, "mismip"))
    fn, hf = str(tmp_path / "restart_other.nc"), str(tmp_path / "help_other.nc")
    R.create_restart(fn, other, ZETA, {"TriC": other.TriC})
    R.create_help_fields(hf, other, ZETA, ["Hi", "lat"], {"TriC": other.TriC})
    hi = g.download("Hi")
    for call in (lambda: g.write_restart(fn, 1.0), lambda: g.load_restart(fn, 1.0), lambda: g.write_help_fields(hf, 1.0, ["Hi"])):
        with pytest.raises(UfmError) as e:
            call()
        assert e.value.rc == -12 and "vertices" in str(e.value)
    assert g.write_help_fields(hf, 1.0, ["lat"], host={"lat": np.zeros(other.nV)}) == 1     # host-only fields do not involve the resident mesh
    fz = str(tmp_path / "restart_nz.nc")
    R.create_restart(fz, mesh_2k, ZETA[:-1], {"TriC": mesh_2k.TriC})
    with pytest.raises(UfmError) as e:
        g.write_restart(fz, 1.0)
    assert e.value.rc == -14
    assert_bits_equal(g.download("Hi"), hi, "Hi untouched")


# ---- drop-in loop: ufm_run_model_host moves the host's fields every step; copies overlap each other and the SSA solve ----
@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("overlap", ["1", "0"])
def test_run_model_host_matches_device_resident_run(mesh_2k, monkeypatch, pinned, overlap):
    """Same trajectory as ufm_run_model, every downloaded field bit-identical to a download after the step -- with page-locked
    and with pageable host arrays, with the overlapped copies (default) and with one synchronous copy per field."""
    from tests.test_gpu_parity import scenario
    from ufemism_b200.capi import HostIce

    monkeypatch.setenv("UFM_XFER_OVERLAP", overlap)
    st = scenario(mesh_2k, "mismip")
    a, b = make_gpu(mesh_2k, st), make_gpu(mesh_2k, st)
    nV = mesh_2k.nV
    hb = {n: np.zeros(nV) for n in ("Hi", "Hb", "SL", "dHb_dt", "SMB_year", "BMB", "Hi_prev", "dHi_dt", "Hs", "U_SSA", "V_SSA", "U_SIA", "V_SIA", "D_SIA")}
    hb["mask_noice"] = np.zeros(nV, np.int32); hb["mask"] = np.zeros(nV, np.int32)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        hb[k][:] = st[k]
    host = HostIce(**{n: hb[n].ctypes.data for n in hb if n != "Hi"}, Hi=hb["Hi"].ctypes.data, Hi_out=hb["Hi"].ctypes.data)
    if pinned:
        for v in hb.values():
            b.host_register(v)
    ra, rb = a.region(0.0), b.region(0.0)
    for step in range(4):
        a.run_model(ra, 1e12, max_steps=1)
        b.run_model_host(rb, 1e12, 1, host)
        assert (ra.time, ra.dt, ra.n_sor_total, ra.n_outer_total) == (rb.time, rb.dt, rb.n_sor_total, rb.n_outer_total)
        for hname, fname in [("Hi", "Hi"), ("Hi_prev", "Hi_prev"), ("dHi_dt", "dHi_dt"), ("Hs", "Hs"), ("U_SSA", "U_SSA"), ("V_SSA", "V_SSA"),
                             ("U_SIA", "U_SIA"), ("V_SIA", "V_SIA"), ("D_SIA", "D_SIA"), ("mask", "mask")]:
            assert_bits_equal(hb[hname], a.download(fname), f"step {step} {hname}")
    assert ra.n_sor_total > 0 and hb["U_SSA"].any() and np.ptp(hb["Hi"]) > 0
    # the host edits a field between steps (what ELRA / SMB components do): the next step must see it
    hb["Hb"][:] = hb["Hb"] + 5.0
    a.upload("Hb", hb["Hb"])
    a.run_model(ra, 1e12, max_steps=1); b.run_model_host(rb, 1e12, 1, host)
    a.run_model(ra, 1e12, max_steps=1); b.run_model_host(rb, 1e12, 1, host)
    assert_bits_equal(hb["Hi"], a.download("Hi"), "Hi after the host changed Hb")
    assert_bits_equal(hb["Hs"], a.download("Hs"), "Hs after the host changed Hb")
    cnt = b.counters()
    # 5 set-up uploads, then per step 6 double + 1 int fields in, 9 double + 1 int fields out
    assert cnt.h2d_bytes == (5 * 8 + 6 * (6 * 8 + 4)) * nV and cnt.d2h_bytes == 6 * (9 * 8 + 4) * nV


def test_thermodynamics_fields_in_restart_and_help_fields(mesh_2k, tmp_path):
    """With a thermodynamics mesh Ti is resident: it goes into the restart frame and comes back with ufm_restart_load; the
    help fields derived from device arrays (Ti_basal = Ti(:,nZ), T2m_year = SUM(T2m,2)/12, src/netcdf_module.f90:348,388)
    equal the same expressions evaluated on downloads."""
    from tests.test_gpu_parity import THERMO_IN
    from ufemism_b200.capi import IceModelGPU

    st = S.state_thermo_dome(mesh_2k, benchmark="none")
    g = IceModelGPU(mesh_2k, benchmark="none", thermo=True)
    for f in THERMO_IN:
        g.upload(f, st[f])
    g.update_general_ice_model_data(0.0); g.solve_SIA(); g.solve_SSA(); g.update_ice_temperature()
    fn, hf = str(tmp_path / "restart_GRL_00001.nc"), str(tmp_path / "help_fields_GRL_00001.nc")
    names = ["Ti", "Ti_basal", "T2m", "T2m_year", "GHF", "W_3D", "A_flow_mean", "U_3D"]
    R.create_restart(fn, mesh_2k, ZETA, {"TriC": mesh_2k.TriC})
    R.create_help_fields(hf, mesh_2k, ZETA, names, {"TriC": mesh_2k.TriC})
    assert g.field_resident("Ti") and g.field_resident("W_3D")
    firn = np.asfortranarray(np.random.default_rng(3).random((mesh_2k.nV, 12)))
    assert g.write_restart(fn, 10.0, FirnDepth=firn) == 1
    assert g.write_help_fields(hf, 10.0, names) == 1
    Ti, T2m = g.download("Ti"), g.download("T2m")
    f, h = netcdf_file(fn, "r", mmap=False), netcdf_file(hf, "r", mmap=False)
    assert_bits_equal(f.variables["Ti"][0], Ti.T, "restart Ti")
    assert_bits_equal(f.variables["FirnDepth"][0], firn.T, "restart FirnDepth")
    assert_bits_equal(h.variables["Ti"][0], Ti.T, "Ti")
    assert_bits_equal(h.variables["Ti_basal"][0], Ti[:, -1], "Ti_basal")
    t2y = np.zeros(mesh_2k.nV)
    for m in range(12):                                               # SUM over the month dimension in index order, then / 12
        t2y = t2y + T2m[:, m]
    assert_bits_equal(h.variables["T2m_year"][0], t2y / 12.0, "T2m_year")
    assert_bits_equal(h.variables["T2m"][0], T2m.T, "T2m")
    assert_bits_equal(h.variables["GHF"][:], g.download("GHF"), "GHF")
    assert_bits_equal(h.variables["W_3D"][0], g.download("W_3D").T, "W_3D")
    assert_bits_equal(h.variables["A_flow_mean"][0], g.download("A_flow_mean"), "A_flow_mean")
    assert np.ptp(Ti) > 1.0 and np.abs(h.variables["W_3D"][0]).max() > 0.0
    f.close(); h.close()
    g2 = IceModelGPU(mesh_2k, benchmark="none", thermo=True, primary_only=True)
    firn2, melt2 = np.zeros((mesh_2k.nV, 12), order="F"), np.zeros(mesh_2k.nV)
    assert g2._ck(g2.L.ufm_restart_load(g2.h, fn.encode(), 10.0, firn2.ctypes.data, melt2.ctypes.data), allow_warning=True) == 1
    assert_bits_equal(g2.download("Ti"), Ti, "Ti loaded from the restart file")
    assert_bits_equal(firn2, firn, "FirnDepth handed to the host")
    assert (melt2 == np.float64(9.9692099683868690e+36)).all()      # never written: the fill value comes back, as with netcdf-fortran
