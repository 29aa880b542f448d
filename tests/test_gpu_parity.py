"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle on the same inputs.

Bar: BIT-EXACT wherever the arithmetic is add/mul/div/sqrt/compare (stencils, masks, thickness update,
SOR sweep, Neumann pass, CFL) and -- with the host libm's pow / tan re-stated on the device (csrc/ufm_pow.cuh, the default when the
running libm is the glibc it knows) -- behind those two functions as well; |rel| <= 1e-13 is what the tests ask for behind a pow()/tan()
so that they also hold with CUDA's own functions (UFM_POW_EXACT=0, <= 2 ulp vs glibc); north_star tolerances for composed results: velocities <= 1e-10 rel-L2 after the same SOR iteration
count, ice thickness <= 1e-8 rel after N model years.
"""
import numpy as np
import pytest

from tests.conftest import get_mesh
from tests.util import assert_bits_equal, make_gpu, make_oracle, rel_l2
from ufemism_b200 import scenarios as S

pytestmark = pytest.mark.gpu

AA_EXACT = ["Hs", "dHs_dt", "dHi_dx", "dHi_dy", "dHs_dx", "dHs_dy", "dHs_dx_shelf", "dHs_dy_shelf"]
AC_EXACT = ["Hi_Ac", "Hb_Ac", "Hs_Ac", "SL_Ac"] + [f"d{f}_d{c}_Ac" for f in ("Hi", "Hb", "Hs", "SL") for c in "xypo"] + ["dHs_dx_shelf_Ac", "dHs_dy_shelf_Ac"]
MASKS = ["mask_land", "mask_ocean", "mask_lake", "mask_ice", "mask_sheet", "mask_shelf", "mask_coast", "mask_margin", "mask_gl", "mask_cf", "mask"]


def scenario(mesh, name):
    if name == "halfar":
        return S.state_halfar(mesh)
    if name == "icestream":
        return S.state_ssa_icestream(mesh, scale=750e3 / 1800e3)
    if name == "mismip":
        st = S.state_mismip(mesh)
        r = np.hypot(mesh.V[:, 0], mesh.V[:, 1])
        st["Hi"] = np.where(r < 600e3, 800.0 - 600.0 * r / 600e3, 0.0) + np.where((r >= 600e3) & (r < 680e3), 150.0, 0.0)
        st["Hi"][mesh.edge_index > 0] = 0.0
        return st
    raise ValueError(name)


@pytest.mark.parametrize("name", ["halfar", "icestream", "mismip"])
def test_update_general_bit_exact(mesh_10k, name):
    st = scenario(mesh_10k, name)
    o, g = make_oracle(mesh_10k, st), make_gpu(mesh_10k, st)
    o.update_general_ice_model_data(0.0)
    g.update_general_ice_model_data(0.0)
    for f in AA_EXACT + AC_EXACT:
        assert_bits_equal(g.download(f), o[f], f)
    for mname in MASKS:
        assert_bits_equal(g.download(mname), o[mname], mname)
        assert_bits_equal(g.download(mname + "_Ac"), o[mname + "_Ac"], mname + "_Ac")
    assert_bits_equal(g.download("A_flow_mean"), o["A_flow_mean"], "A_flow_mean")


@pytest.mark.parametrize("name", ["halfar", "mismip"])
def test_solve_SIA(mesh_10k, name):
    st = scenario(mesh_10k, name)
    o, g = make_oracle(mesh_10k, st), make_gpu(mesh_10k, st)
    o.update_general_ice_model_data(0.0); o.solve_SIA()
    g.update_general_ice_model_data(0.0); g.solve_SIA()
    for f in ["D_SIA_Ac", "Ux_SIA_Ac", "Uy_SIA_Ac", "Up_SIA_Ac", "Uo_SIA_Ac", "U_SIA", "V_SIA", "D_SIA"]:
        a, b = g.download(f), o[f]
        assert np.abs(b).max() > 0, f
        # one CUDA pow(x, 3.0) per column: <= 2 ulp
        np.testing.assert_allclose(a, b, rtol=1e-14, atol=1e-14 * np.abs(b).max(), err_msg=f)
    assert (o["D_SIA_Ac"] <= 0).all()  # negative by construction


@pytest.mark.parametrize("name,dt", [("halfar", 0.05), ("halfar", 0.0), ("mismip", 0.5), ("mismip", 40.0), ("icestream_as_coded", 0.5)])
@pytest.mark.parametrize("edge_pass", ["1", "0"])
def test_thickness_update_bit_exact(mesh_10k, name, dt, edge_pass, monkeypatch):
    monkeypatch.setenv("UFM_THK_EDGE", edge_pass)   # 1: flux once per staggered vertex (k_thk_flux), 0: from the velocities in both vertex passes
    if name == "icestream_as_coded":
        # 'SSA_icestream' is the one benchmark without a thickness boundary condition (ice_dynamics_module.f90:206): edge vertices keep their ice
        st = dict(scenario(mesh_10k, "icestream"), benchmark="SSA_icestream")
    else:
        st = scenario(mesh_10k, name)
    if name == "mismip":
        st["SMB_year"] = np.where(np.hypot(mesh_10k.V[:, 0], mesh_10k.V[:, 1]) > 400e3, -3.0, 0.3)  # melt-all + limiter branches
    o, g = make_oracle(mesh_10k, st), make_gpu(mesh_10k, st)
    o.update_general_ice_model_data(0.0); o.solve_SIA()
    g.update_general_ice_model_data(0.0)
    # identical velocities on both sides so the flux arithmetic itself is compared bit for bit
    g.upload("Up_SIA_Ac", o["Up_SIA_Ac"])
    rng = np.random.default_rng(1)
    up_ssa = rng.normal(0, 50.0, mesh_10k.nAc) * (o["mask_ice_Ac"] > 0)
    o["Up_SSA_Ac"][:] = up_ssa
    g.upload("Up_SSA_Ac", up_ssa)
    noice = (rng.random(mesh_10k.nV) < 0.01).astype(np.int32)
    o["mask_noice"][:] = noice
    g.upload("mask_noice", noice)
    o.calculate_ice_thickness_change(dt)
    g.calculate_ice_thickness_change(dt)
    for f in ["Hi", "dHi_dt", "Hi_prev"]:
        assert_bits_equal(g.download(f), o[f], f)
    if dt > 0 and name == "mismip":
        assert (o["Hi"] >= -1e-9).all() and (o["Hi"] == 0).any()  # limiter keeps H >= 0 up to rounding
    edge = (mesh_10k.edge_index > 0) & (noice == 0)
    assert (o["Hi"][edge] > 0).any() if name == "icestream_as_coded" else (o["Hi"][edge] == 0).all()


def _ssa_setup_pair(mesh, nthreads=1, **params):
    st = scenario(mesh, "icestream")
    o, g = make_oracle(mesh, st, nthreads=nthreads, **{k: v for k, v in params.items() if k != "exact_xy"}), make_gpu(mesh, st, **params)
    o.update_general_ice_model_data(0.0)
    g.update_general_ice_model_data(0.0)
    return o, g


def test_ssa_prepare_viscosity_setup(mesh_10k):
    o, g = _ssa_setup_pair(mesh_10k)
    o.basal_yield_stress(); o.SSA_gather_AaAc()
    g.ssa_prepare()
    assert_bits_equal(g.download("phi_fric_AaAc"), o["phi_fric_AaAc"], "phi_fric")
    np.testing.assert_allclose(g.download("tau_c_AaAc"), o["tau_c_AaAc"], rtol=1e-14)  # tan()
    # a smooth non-trivial velocity field, identical on both sides
    x, y = mesh_10k.VAaAc[:, 0], mesh_10k.VAaAc[:, 1]
    U = 100.0 * np.sin(x / 2e5) * np.cos(y / 3e5); V = -80.0 * np.cos(x / 2.5e5) * np.sin(y / 2e5)
    o["U_SSA_AaAc"][:] = U; o["V_SSA_AaAc"][:] = V
    g.upload("U_SSA_AaAc", U); g.upload("V_SSA_AaAc", V)
    g.upload("tau_c_AaAc", o["tau_c_AaAc"])
    o.SSA_effective_viscosity()
    sums = g.ssa_viscosity()
    for f, of in [("dU_dx_AaAc", "dU_SSA_dx_AaAc"), ("dU_dy_AaAc", "dU_SSA_dy_AaAc"), ("dV_dx_AaAc", "dV_SSA_dx_AaAc"), ("dV_dy_AaAc", "dV_SSA_dy_AaAc")]:
        assert_bits_equal(g.download(f), o[of], f)
    np.testing.assert_allclose(g.download("eta_AaAc"), o["eta_AaAc"], rtol=1e-14)  # pow()
    n = o["N_AaAc"]
    np.testing.assert_allclose(sums[1], float((n * n).sum()), rtol=1e-12)
    # identical eta on both sides -> sliding term within pow() accuracy, RHS / centre coefficients bit-exact given S
    g.upload("eta_AaAc", o["eta_AaAc"])
    o.SSA_sliding_term()
    o.solve_SSA_linearised(max_inner=1, force_iters=True)  # fills RHS, eu, ev (and does one sweep we do not look at)
    g.ssa_sliding_and_setup()
    np.testing.assert_allclose(g.download("S_AaAc"), o["S_AaAc"], rtol=1e-14)
    assert_bits_equal(g.download("RHSx_AaAc"), o["RHSx_AaAc"], "RHSx")
    assert_bits_equal(g.download("RHSy_AaAc"), o["RHSy_AaAc"], "RHSy")
    fin = np.isfinite(o["eu_i_AaAc"])
    np.testing.assert_allclose(g.download("eu_i_AaAc")[fin], o["eu_i_AaAc"][fin], rtol=1e-13)
    np.testing.assert_allclose(g.download("ev_i_AaAc")[fin], o["ev_i_AaAc"][fin], rtol=1e-13)


@pytest.mark.parametrize("k", [1, 10, 100])
@pytest.mark.parametrize("nthreads", [1, 8])
def test_sor_bit_exact_at_prescribed_iterations(mesh_10k, k, nthreads):
    """SURVEY section 7 acceptance: identical linear system on both sides, k forced SOR iterations."""
    o, g = _ssa_setup_pair(mesh_10k, nthreads=nthreads)
    o.basal_yield_stress(); o.SSA_gather_AaAc(); o.SSA_effective_viscosity(); o.SSA_sliding_term()
    g.ssa_prepare()
    # the oracle's eta / tau_c define the system on both sides
    g.upload("tau_c_AaAc", o["tau_c_AaAc"]); g.ssa_viscosity(); g.upload("eta_AaAc", o["eta_AaAc"])
    g.ssa_sliding_and_setup()
    g.upload("S_AaAc", o["S_AaAc"])
    n, res, _, _ = o.solve_SSA_linearised(max_inner=k, force_iters=True)
    for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):
        g.upload(f, o[f])
    st = g.ssa_sor(max_inner=k, force_iters=True)
    assert st.n_inner_last == k == n
    assert_bits_equal(g.download("U_SSA_AaAc"), o["U_SSA_AaAc"], "U_SSA_AaAc")
    assert_bits_equal(g.download("V_SSA_AaAc"), o["V_SSA_AaAc"], "V_SSA_AaAc")
    assert st.last_max_residual == res


@pytest.mark.parametrize("variant", [
    {"UFM_SOR_DATAFLOW": "0", "UFM_SOR_CHUNK": "0", "UFM_SOR_FUSE_BC": "0", "UFM_SOR_BAR": "0"},
    {"UFM_SOR_DATAFLOW": "0", "UFM_SOR_CHUNK": "1", "UFM_SOR_FUSE_BC": "0", "UFM_SOR_BAR": "0"},
    {"UFM_SOR_DATAFLOW": "0", "UFM_SOR_CHUNK": "0", "UFM_SOR_FUSE_BC": "1", "UFM_SOR_BAR": "0"},
    {"UFM_SOR_DATAFLOW": "0", "UFM_SOR_CHUNK": "0", "UFM_SOR_FUSE_BC": "0", "UFM_SOR_BAR": "1"},
    {"UFM_SOR_DATAFLOW": "0", "UFM_SOR_CHUNK": "1", "UFM_SOR_FUSE_BC": "1", "UFM_SOR_BAR": "1"},
    {"UFM_SOR_DATAFLOW": "1"},
    {"UFM_SOR_DATAFLOW": "1", "UFM_SOR_GRID": "2"},
    {"UFM_SOR_DATAFLOW": "1", "UFM_SOR_GRID": "1"},
    {"UFM_SOR_DATAFLOW": "1", "UFM_SOR_GRID": "7", "UFM_ROW_ORDER": "bands:5:64"},
], ids=["plain", "equal_share", "fused_neumann", "release_barrier", "barrier_default", "dataflow", "dataflow_2_ctas", "dataflow_1_cta", "dataflow_7_ctas_5_bands"])
def test_sor_schedule_variants_bit_exact(mesh_10k, monkeypatch, variant):
    """The SOR kernels only reorder work: the barrier kernel's scheduling switches (equal slice shares per warp, Neumann pass inside the
    fifth colour phase, release/acquire grid barrier) inside a colour phase, the dataflow kernel (opt-in; no grid barrier
    between the colours, every slice waits for the stage holding its lower-coloured neighbours) across the colours of one iteration.
    UFM_SOR_GRID shrinks the grid so that this small mesh is swept in several rounds per colour (at 2 CTAs: 4-5 rounds).  Every variant
    must give the oracle's bits."""
    for k_, v_ in variant.items():
        monkeypatch.setenv(k_, v_)
    o, g = _ssa_setup_pair(mesh_10k, nthreads=8)
    o.basal_yield_stress(); o.SSA_gather_AaAc(); o.SSA_effective_viscosity(); o.SSA_sliding_term()
    g.ssa_prepare(); g.upload("tau_c_AaAc", o["tau_c_AaAc"]); g.ssa_viscosity(); g.upload("eta_AaAc", o["eta_AaAc"]); g.ssa_sliding_and_setup()
    n, res, _, _ = o.solve_SSA_linearised(max_inner=57, force_iters=True)
    for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):
        g.upload(f, o[f])
    st = g.ssa_sor(max_inner=57, force_iters=True)
    assert st.n_inner_last == 57 == n
    assert_bits_equal(g.download("U_SSA_AaAc"), o["U_SSA_AaAc"], "U_SSA_AaAc")
    assert_bits_equal(g.download("V_SSA_AaAc"), o["V_SSA_AaAc"], "V_SSA_AaAc")
    assert st.last_max_residual == res


@pytest.mark.skipif(not __import__("os").environ.get("UFM_TEST_EXPERIMENTAL"), reason="experimental device row order, not measured yet: run with UFM_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("order", ["bands:16:256", "bands:64", "bands:3:32"])
def test_experimental_band_row_order_bit_exact(mesh_10k, monkeypatch, order):
    """UFM_ROW_ORDER=bands:n[:window] (x-bands, Morton inside a band, degree sorted inside windows; groundwork for a sweep without grid
    barriers between colours) only moves rows inside their colour block: geometry, the SOR sweep at a prescribed
    count and the whole solve must give the oracle's bits / counts exactly as the default order does."""
    monkeypatch.setenv("UFM_ROW_ORDER", order)
    o, g = _ssa_setup_pair(mesh_10k, nthreads=8)
    o.basal_yield_stress(); o.SSA_gather_AaAc(); o.SSA_effective_viscosity(); o.SSA_sliding_term()
    g.ssa_prepare(); g.upload("tau_c_AaAc", o["tau_c_AaAc"]); g.ssa_viscosity(); g.upload("eta_AaAc", o["eta_AaAc"]); g.ssa_sliding_and_setup()
    n, res, _, _ = o.solve_SSA_linearised(max_inner=57, force_iters=True)
    for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):
        g.upload(f, o[f])
    st = g.ssa_sor(max_inner=57, force_iters=True)
    assert st.n_inner_last == 57 == n
    assert_bits_equal(g.download("U_SSA_AaAc"), o["U_SSA_AaAc"], "U_SSA_AaAc")
    assert_bits_equal(g.download("V_SSA_AaAc"), o["V_SSA_AaAc"], "V_SSA_AaAc")
    assert st.last_max_residual == res
    so, sg = o.solve_SSA(), g.solve_SSA()
    assert (sg.n_outer, sg.n_inner_total) == (so.n_outer, so.n_inner_total)
    assert rel_l2(g.download("U_SSA"), o["U_SSA"]) <= 1e-10 and rel_l2(g.download("V_SSA"), o["V_SSA"]) <= 1e-10


@pytest.mark.parametrize("dataflow", ["0", "1"])
def test_dataflow_sor_at_250k_vertices_bit_exact(monkeypatch, dataflow):
    """Both SOR kernels at a size where the full grid (148 x 32 warps) sweeps every colour in two rounds: forced iterations and the
    whole solve_SSA (with the analytical grounding-line flux) against the oracle -- bit-exact sweep, identical iteration counts."""
    monkeypatch.setenv("UFM_SOR_DATAFLOW", dataflow)
    m = get_mesh(250000, half_width=S.CONFIG3["half_width"])
    st = S.state_ssa_icestream(m, scale=1.0, Hb=S.CONFIG3["Hb"], H_shelf=S.CONFIG3["H_shelf"])   # the bench workload at a quarter of its size
    o, g = make_oracle(m, st, nthreads=16, use_analytical_GL_flux=1), make_gpu(m, st, use_analytical_GL_flux=1)
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    o.basal_yield_stress(); o.calculate_GL_flux(); o.SSA_gather_AaAc(); o.SSA_effective_viscosity(); o.SSA_sliding_term()
    g.ssa_prepare(); g.upload("tau_c_AaAc", o["tau_c_AaAc"])
    # the grounding-line rows are held at the analytical flux velocity, which sits behind a pow() and a tan(): start both sides from the oracle's bits
    g.upload("U_SSA_AaAc", o["U_SSA_AaAc"]); g.upload("V_SSA_AaAc", o["V_SSA_AaAc"])
    g.ssa_viscosity(); g.upload("eta_AaAc", o["eta_AaAc"]); g.ssa_sliding_and_setup()
    n, res, _, _ = o.solve_SSA_linearised(max_inner=12, force_iters=True)
    for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):
        g.upload(f, o[f])
    st_ = g.ssa_sor(max_inner=12, force_iters=True)
    assert st_.n_inner_last == 12 == n
    assert_bits_equal(g.download("U_SSA_AaAc"), o["U_SSA_AaAc"], "U_SSA_AaAc")
    assert_bits_equal(g.download("V_SSA_AaAc"), o["V_SSA_AaAc"], "V_SSA_AaAc")
    assert st_.last_max_residual == res
    o.cfg.SSA_max_outer_loops = 6; g.set_params(SSA_max_outer_loops=6)
    so, sg = o.solve_SSA(), g.solve_SSA()
    assert (sg.n_outer, sg.n_inner_total) == (so.n_outer, so.n_inner_total)
    assert rel_l2(g.download("U_SSA"), o["U_SSA"]) <= 1e-10 and rel_l2(g.download("V_SSA"), o["V_SSA"]) <= 1e-10


def test_device_pow_has_the_host_libms_bits(mesh_10k):
    """ufm_pow.cuh: with the host libm's tables found and validated (pow_mode 1) everything behind a pow() -- SIA diffusivity and velocities,
    effective viscosity, sliding term -- is bit-identical with the oracle, not merely within an ulp or two."""
    st = scenario(mesh_10k, "icestream")
    o, g = make_oracle(mesh_10k, st), make_gpu(mesh_10k, st)
    if not g.pow_mode() & 1:
        pytest.skip("pow tables of this host's libm not usable: CUDA pow, tolerance-level parity only")
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    o.solve_SIA(); g.solve_SIA()
    for f in ("D_SIA_Ac", "Ux_SIA_Ac", "Up_SIA_Ac", "U_SIA", "D_SIA"):
        assert_bits_equal(g.download(f), o[f], f)
    o.basal_yield_stress(); o.SSA_gather_AaAc(); g.ssa_prepare()
    if g.pow_mode() & 2:          # tan of the friction angle too
        assert_bits_equal(g.download("tau_c_AaAc"), o["tau_c_AaAc"], "tau_c_AaAc")
    x, y = mesh_10k.VAaAc[:, 0], mesh_10k.VAaAc[:, 1]
    U = 100.0 * np.sin(x / 2e5) * np.cos(y / 3e5); V = -80.0 * np.cos(x / 2.5e5) * np.sin(y / 2e5)
    o["U_SSA_AaAc"][:] = U; o["V_SSA_AaAc"][:] = V
    g.upload("U_SSA_AaAc", U); g.upload("V_SSA_AaAc", V); g.upload("tau_c_AaAc", o["tau_c_AaAc"])
    o.SSA_effective_viscosity(); o.SSA_sliding_term(); g.ssa_viscosity()
    for f in ("eta_AaAc", "N_AaAc", "S_AaAc"):
        assert_bits_equal(g.download(f), o[f], f)


def test_whole_trajectory_bit_identical_with_exact_pow_and_tan(mesh_10k):
    """With pow and tan evaluated as the host's libm does, a hybrid SIA / SSA run on the MISMIP bed with the analytical grounding-line flux
    (pow and tan inside calculate_GL_flux, a spatially varying friction angle) is bit-identical with the oracle step after step: thickness,
    velocities, masks, time steps, iteration counts."""
    st = scenario(mesh_10k, "mismip")
    o = make_oracle(mesh_10k, st, nthreads=8, use_analytical_GL_flux=1)
    g = make_gpu(mesh_10k, st, use_analytical_GL_flux=1)
    if g.pow_mode() != 3:
        pytest.skip("pow / tan tables of this host's libm not usable")
    ro, rg = o.region(0.0), g.region(0.0)
    for step in range(12):
        o.run_model(ro, 1e12, max_steps=1); g.run_model(rg, 1e12, max_steps=1)
        assert (rg.time, rg.dt, rg.n_sor_total, rg.n_outer_total) == (ro.time, ro.dt, ro.n_sor_total, ro.n_outer_total), step
        for f in ("Hi", "U_SSA", "V_SSA", "U_SIA", "Hs", "mask", "mask_gl_Ac", "Qabs_GL_Ac"):
            assert_bits_equal(g.download(f), o[f], f"step {step} {f}")
    assert ro.n_ssa >= 3 and np.abs(o["U_SSA"]).max() > 1.0


def test_sor_presummed_xy_within_tolerance(mesh_10k):
    o, g = _ssa_setup_pair(mesh_10k, exact_xy=0)
    o.basal_yield_stress(); o.SSA_gather_AaAc(); o.SSA_effective_viscosity(); o.SSA_sliding_term()
    g.ssa_prepare(); g.upload("tau_c_AaAc", o["tau_c_AaAc"]); g.ssa_viscosity(); g.upload("eta_AaAc", o["eta_AaAc"]); g.ssa_sliding_and_setup()
    o.solve_SSA_linearised(max_inner=100, force_iters=True)
    for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):
        g.upload(f, o[f])
    g.ssa_sor(max_inner=100, force_iters=True)
    assert rel_l2(g.download("U_SSA_AaAc"), o["U_SSA_AaAc"]) <= 1e-10
    assert rel_l2(g.download("V_SSA_AaAc"), o["V_SSA_AaAc"]) <= 1e-10


def test_sor_natural_stop_and_quirk(mesh_10k):
    """Stop test on the un-relaxed residual (ice_dynamics_module.f90:651-659,676)."""
    o, g = _ssa_setup_pair(mesh_10k)
    o.basal_yield_stress(); o.SSA_gather_AaAc(); o.SSA_effective_viscosity(); o.SSA_sliding_term()
    g.ssa_prepare(); g.upload("tau_c_AaAc", o["tau_c_AaAc"]); g.ssa_viscosity(); g.upload("eta_AaAc", o["eta_AaAc"]); g.ssa_sliding_and_setup()
    n, res, did_reset, warn = o.solve_SSA_linearised()
    for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):
        g.upload(f, o[f])
    st = g.ssa_sor()
    assert (st.n_inner_last, st.did_reset) == (n, did_reset)
    assert st.last_max_residual == res and res < 2.5
    assert_bits_equal(g.download("U_SSA_AaAc"), o["U_SSA_AaAc"], "U")


@pytest.mark.parametrize("gl_flux", [0, 1])
def test_solve_SSA_full(mesh_10k, gl_flux):
    """north_star gate: velocities within 1e-10 relative L2 after the same SOR iteration count."""
    st = scenario(mesh_10k, "icestream")
    o = make_oracle(mesh_10k, st, nthreads=8, use_analytical_GL_flux=gl_flux)
    g = make_gpu(mesh_10k, st, use_analytical_GL_flux=gl_flux)
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    so, sg = o.solve_SSA(), g.solve_SSA()
    assert (sg.n_outer, sg.n_inner_total, sg.did_reset) == (so.n_outer, so.n_inner_total, so.did_reset)
    assert so.n_inner_total > 10
    for f in ("U_SSA", "V_SSA", "Ux_SSA_Ac", "Uy_SSA_Ac", "Up_SSA_Ac", "Uo_SSA_Ac"):
        assert rel_l2(g.download(f), o[f]) <= 1e-10, f
    np.testing.assert_allclose(sg.last_RN, so.last_RN, rtol=1e-9)
    # second call starts from the previous solution and the left-over N (ice_dynamics_module.f90:505)
    so2, sg2 = o.solve_SSA(), g.solve_SSA()
    assert (sg2.n_outer, sg2.n_inner_total) == (so2.n_outer, so2.n_inner_total)
    assert rel_l2(g.download("U_SSA"), o["U_SSA"]) <= 1e-10


def test_solve_SSA_zero_branches(mesh_2k):
    st = S.state_halfar(mesh_2k)
    g = make_gpu(mesh_2k, st)
    g.update_general_ice_model_data(0.0)
    g.upload("U_SSA", np.ones(mesh_2k.nV))
    s = g.solve_SSA()  # Halfar: SSA not solved, velocities set to zero (ice_dynamics_module.f90:431-465)
    assert s.n_outer == 0 and not g.download("U_SSA").any()
    st2 = S.state_mismip(mesh_2k); st2["Hi"][:] = 0.0
    g2 = make_gpu(mesh_2k, st2)
    g2.update_general_ice_model_data(0.0)
    assert g2.solve_SSA().n_outer == 0  # no grounded ice anywhere


def test_cfl_bit_exact(mesh_10k):
    st = scenario(mesh_10k, "halfar")
    o, g = make_oracle(mesh_10k, st), make_gpu(mesh_10k, st)
    o.update_general_ice_model_data(0.0); o.solve_SIA()
    g.update_general_ice_model_data(0.0)
    rng = np.random.default_rng(3)
    for f, n in (("D_SIA_Ac", mesh_10k.nAc), ("U_SSA", mesh_10k.nV), ("V_SSA", mesh_10k.nV)):
        a = o[f] if f == "D_SIA_Ac" else rng.normal(0, 300.0, n)
        o[f][:] = a; g.upload(f, a)
    u3 = rng.normal(0, 100.0, (mesh_10k.nV, 15))
    o["U_3D"][:] = u3; g.upload("U_3D", u3)
    assert g.determine_timesteps() == o.determine_timesteps()


def test_run_model_halfar(mesh_2k):
    """north_star gate: ice thickness after N model years within 1e-8 relative (same dt sequence)."""
    from oracle.oracle import T_THERMO
    st = S.state_halfar(mesh_2k)
    o, g = make_oracle(mesh_2k, st, nthreads=4), make_gpu(mesh_2k, st)
    ro, rg = o.region(0.0), g.region(0.0)
    ro.dtc[T_THERMO] = 5.0; rg.dtc[T_THERMO] = 5.0
    assert o.run_model(ro, 20.0) == 0
    g.run_model(rg, 20.0)
    assert rg.n_steps == ro.n_steps and rg.time == ro.time == 20.0
    assert rel_l2(g.download("Hi"), o["Hi"]) <= 1e-8
    assert np.abs(g.download("Hi") - o["Hi"]).max() <= 1e-8 * o["Hi"].max()


@pytest.mark.parametrize("graph", ["1", "0"])
@pytest.mark.parametrize("name", ["halfar", "eismint1"])
def test_device_loop_matches_host_loop(mesh_2k, monkeypatch, name, graph):
    """run_model_device (the step's kernels gated and timed by a control block on the device, 64 steps enqueued per synchronisation) against
    the host-driven loop: same region state after every call (time, dt, timers, flags, counters), same fields bit for bit -- with max_steps
    that end inside a batch, with an end time, and across repeated calls."""
    st = S.state_halfar(mesh_2k) if name == "halfar" else S.state_eismint1(mesh_2k)
    gs = []
    for loop in ("0", "1"):
        monkeypatch.setenv("UFM_DEVICE_LOOP", loop)
        monkeypatch.setenv("UFM_DEVICE_GRAPH", graph)          # pairs of steps as one CUDA graph / plain launches
        g = make_gpu(mesh_2k, st)
        r = g.region(0.0)
        snaps = []
        for t_end, ms in ((1e12, 1), (1e12, 7), (1e12, 100), (12.5, 0), (30.0, 0), (30.0, 5)):
            g.run_model(r, t_end, max_steps=ms)
            snaps.append((r.time, r.dt, r.dt_prev, list(r.t0), list(r.t1), list(r.dtc), list(r.do_), r.n_steps, r.n_sia, r.n_ssa, list(r.dt_crit_last),
                          g.download("Hi").copy(), g.download("Hi_prev").copy(), g.download("U_SIA").copy(), g.download("U_3D").copy(), g.determine_timesteps()))
        gs.append(snaps)
    assert gs[0][-1][7] > 100                                              # more than one batch of 64 steps
    for a, b in zip(*gs):
        assert a[:11] == b[:11]
        for x, y in zip(a[11:15], b[11:15]):
            assert_bits_equal(y, x)
        assert a[15] == b[15]


def test_run_model_mismip_hybrid(mesh_2k):
    """Hybrid SIA/SSA (config 4 physics) for a few model years: thickness and velocities track the oracle."""
    st = scenario(mesh_2k, "mismip")
    o, g = make_oracle(mesh_2k, st, nthreads=4), make_gpu(mesh_2k, st)
    ro, rg = o.region(0.0), g.region(0.0)
    assert o.run_model(ro, 3.0) == 0
    g.run_model(rg, 3.0)
    assert rg.n_steps == ro.n_steps and rg.n_sor_total == ro.n_sor_total
    assert rel_l2(g.download("Hi"), o["Hi"]) <= 1e-8
    assert rel_l2(g.download("U_SSA"), o["U_SSA"]) <= 1e-8


def test_mesh_reupload_and_errors(mesh_2k):
    from ufemism_b200.capi import UfmError
    st = S.state_halfar(mesh_2k)
    g = make_gpu(mesh_2k, st)
    m2 = get_mesh(3000, seed=7)
    g.upload_mesh(m2)  # device re-upload after a CPU mesh update: state is reallocated and zero
    assert not g.download("Hi").any() and g.download("Hi").shape == (m2.nV,)
    with pytest.raises(UfmError):
        g.upload("mask_ice", np.zeros(m2.nV, np.int32))  # masks are outputs
    bad = get_mesh(2000)
    col = bad.colour_vi.copy()
    import copy
    b2 = copy.copy(bad); b2.colour_vi = col.copy()
    # break the colouring: move one vertex into a neighbour's colour
    a = int(col[0, 0]); nb = int(bad.CAaAc[a - 1, 0]); cn = int(bad.colour[nb - 1])
    b2.colour_vi[0, 0] = col[bad.colour_nV[cn - 1] - 1, cn - 1]; b2.colour_vi[bad.colour_nV[cn - 1] - 1, cn - 1] = a
    with pytest.raises(UfmError):
        g.upload_mesh(b2)


def test_partitioned_multi_gpu_matches_single_gpu():
    """vertex-partitioned solve over NVLink == single-GPU solve, bit for bit (needs >= 2 GPUs on the box)."""
    import os
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(root, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def _realistic_pair(mesh, nthreads=4):
    """do_benchmark_experiment = .FALSE.: temperature-dependent (Arrhenius) flow factor, Ti uploaded by the host."""
    st = scenario(mesh, "icestream")
    st["benchmark"] = "none"
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    zeta = np.array([0.00, 0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 0.925, 0.95, 0.975, 0.99, 1.00])
    Ts = 240.0 + 15.0 * np.sin(x / 4e5) * np.cos(y / 3e5)            # surface temperature
    Ti = Ts[:, None] + (272.0 - Ts[:, None]) * zeta[None, :] ** 2      # warming towards the bed, crosses 263.15 K
    o, g = make_oracle(mesh, st, nthreads=nthreads), make_gpu(mesh, st)
    o["Ti"][:] = Ti
    g.upload("Ti", Ti)
    return o, g


def test_realistic_flow_factor_general_and_sia(mesh_10k):
    o, g = _realistic_pair(mesh_10k)
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    for f in AA_EXACT + AC_EXACT:
        assert_bits_equal(g.download(f), o[f], f)
    # exp() per layer: CUDA libm vs glibc
    np.testing.assert_allclose(g.download("A_flow_mean"), o["A_flow_mean"], rtol=1e-13)
    np.testing.assert_allclose(g.download("A_flow_mean_Ac"), o["A_flow_mean_Ac"], rtol=1e-13)
    assert o["A_flow_mean"].max() / o["A_flow_mean"].min() > 3.0
    o.solve_SIA(); g.solve_SIA()
    for f in ["D_SIA_Ac", "Up_SIA_Ac", "Uo_SIA_Ac", "U_SIA", "D_SIA"]:
        b = o[f]
        np.testing.assert_allclose(g.download(f), b, rtol=1e-12, atol=1e-13 * np.abs(b).max(), err_msg=f)


@pytest.mark.parametrize("gl_flux", [0, 1])
def test_realistic_flow_factor_solve_SSA(mesh_10k, gl_flux):
    o, g = _realistic_pair(mesh_10k, nthreads=8)
    o.cfg.use_analytical_GL_flux = gl_flux
    g.set_params(use_analytical_GL_flux=gl_flux)
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    so, sg = o.solve_SSA(), g.solve_SSA()
    assert (sg.n_outer, sg.n_inner_total, sg.did_reset) == (so.n_outer, so.n_inner_total, so.did_reset)
    for f in ("U_SSA", "V_SSA", "Up_SSA_Ac"):
        assert rel_l2(g.download(f), o[f]) <= 1e-10, f
    # and one thickness step on top (mass continuity with SIA + SSA velocities)
    o.solve_SIA(); g.solve_SIA()
    o.calculate_ice_thickness_change(0.5); g.calculate_ice_thickness_change(0.5)
    assert rel_l2(g.download("Hi"), o["Hi"]) <= 1e-8


def test_host_registered_buffers(mesh_2k):
    st = S.state_halfar(mesh_2k)
    g = make_gpu(mesh_2k, st)
    a = st["Hi"].copy()
    out = np.zeros_like(a)
    g.host_register(a); g.host_register(out)
    g.upload("Hi", a)           # DMA straight from the registered array
    g.download("Hi", out)
    assert np.array_equal(out, st["Hi"])
    g.host_unregister(a); g.host_unregister(out)
    g.upload("Hi", a)
    assert np.array_equal(g.download("Hi"), st["Hi"])


def test_solve_SIA_3D_uv(mesh_10k):
    st = scenario(mesh_10k, "mismip")
    o, g = make_oracle(mesh_10k, st), make_gpu(mesh_10k, st)
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    rng = np.random.default_rng(2)
    us, vs = rng.normal(0, 20.0, mesh_10k.nV), rng.normal(0, 20.0, mesh_10k.nV)
    o["U_SSA"][:] = us; o["V_SSA"][:] = vs
    g.upload("U_SSA", us); g.upload("V_SSA", vs)
    o.solve_SIA_3D(); g.solve_SIA_3D()
    for f in ("U_3D", "V_3D"):
        b = o[f]
        assert np.abs(b - us[:, None] if f == "U_3D" else b - vs[:, None]).max() > 1.0   # SIA part is there
        np.testing.assert_allclose(g.download(f), b, rtol=1e-13, atol=1e-13 * np.abs(b).max(), err_msg=f)
    assert g.determine_timesteps()[2] == pytest.approx(o.determine_timesteps()[2], rel=1e-12)


def test_run_model_eismint1(mesh_2k):
    """BASELINE config 1: EISMINT-1 moving margin from an ice-free start; thermodynamics timer refreshes U_3D, which
    limits dt: same step sequence and thickness within 1e-8 after 400 model years."""
    st = S.state_eismint1(mesh_2k)
    o, g = make_oracle(mesh_2k, st, nthreads=4), make_gpu(mesh_2k, st)
    ro, rg = o.region(0.0), g.region(0.0)
    assert o.run_model(ro, 400.0) == 0
    g.run_model(rg, 400.0)
    assert rg.n_steps == ro.n_steps and rg.time == ro.time == 400.0
    assert o["Hi"].max() > 100.0 and np.abs(o["U_3D"]).max() > 0.0
    assert rel_l2(g.download("Hi"), o["Hi"]) <= 1e-8
    assert rel_l2(g.download("U_3D"), o["U_3D"]) <= 1e-8


@pytest.fixture(scope="module")
def big_case():
    """BASELINE config-2 size (~250 k vertices, 1 M AaAc rows): too large for the scalar oracle to be quick, so the
    checks below are size-independent PROPERTIES of the CUDA path."""
    m = get_mesh(250000, half_width=1800e3)
    st = S.state_ssa_icestream(m, Hb=-250.0, H_shelf=150.0)
    return m, st


def test_full_size_sor_linearity_determinism_neumann(big_case):
    m, st = big_case
    g = make_gpu(m, st, use_analytical_GL_flux=0)
    g.update_general_ice_model_data(0.0)
    g.ssa_prepare(); g.ssa_viscosity()
    rng = np.random.default_rng(11)
    U0 = rng.normal(0, 50.0, m.nVAaAc); V0 = rng.normal(0, 50.0, m.nVAaAc)
    rx, ry = g.download("RHSx_AaAc"), g.download("RHSy_AaAc")

    def run(scale, k=20):
        g.upload("U_SSA_AaAc", scale * U0); g.upload("V_SSA_AaAc", scale * V0)
        g.upload("RHSx_AaAc", scale * rx); g.upload("RHSy_AaAc", scale * ry)
        s_ = g.ssa_sor(max_inner=k, force_iters=True)
        return g.download("U_SSA_AaAc"), g.download("V_SSA_AaAc"), s_.last_max_residual

    u1, v1, r1 = run(1.0)
    u1b, v1b, r1b = run(1.0)
    assert np.array_equal(u1, u1b) and np.array_equal(v1, v1b) and r1 == r1b          # deterministic
    u2, v2, r2 = run(2.0)
    # the sweep is linear in (U, V, RHS); scaling by a power of two is exact in fp64 -> bit-identical
    assert np.array_equal(u2, 2.0 * u1) and np.array_equal(v2, 2.0 * v1) and r2 == 2.0 * r1
    # Neumann pass: every domain-edge row (not a corner) holds the mean of its non-edge neighbours
    is_edge = np.concatenate([m.edge_index, m.edge_index_Ac]) > 0
    for ai in np.flatnonzero(is_edge)[4:400]:
        nb = m.CAaAc[ai, : m.nCAaAc[ai]] - 1
        vals = [u1[j] for j in nb if not is_edge[j]]
        s_ = 0.0
        for x in vals:
            s_ += x
        assert u1[ai] == s_ / len(vals)


def test_full_size_thickness_update_conserves_mass(big_case):
    m, st = big_case
    g = make_gpu(m, st, use_analytical_GL_flux=1)
    r = g.region(0.0)
    g.run_model(r, 1e12, max_steps=2)      # geometry, SIA, SSA once; velocities now non-trivial
    H0 = g.download("Hi")
    dt = 0.05
    g.calculate_ice_thickness_change(dt)
    H1 = g.download("Hi")
    assert np.array_equal(g.download("Hi_prev"), H0)
    inner = m.edge_index == 0
    dV = float((m.A[inner] * (H1[inner] - H0[inner])).sum())
    smb = float((m.A[inner] * st["SMB_year"][inner] * dt).sum())
    # flux form: what leaves one cell enters its neighbour; only fluxes into domain-edge cells (reset to 0) are lost
    lost_to_edge = float((m.A[~inner] * H0[~inner]).sum())
    assert abs(dV - smb) <= 1e-9 * abs(smb) + lost_to_edge + 1e-6 * float((m.A * H0).sum()) * dt
    assert (H1 >= -1e-9).all()


def test_solve_SSA_warning_and_instability_paths(mesh_2k):
    """Error behaviour of solve_SSA (src/ice_dynamics_module.f90:679-689, 533-540): hitting SSA_max_inner_loops only warns
    (rc 1, run continues); max_res > 1e6 resets the velocities once; a second reset aborts (rc -1).  Same on both sides."""
    from ufemism_b200.capi import UfmError
    st = S.state_ssa_icestream(mesh_2k, scale=750e3 / 1800e3, Hb=-250.0, H_shelf=150.0)
    # (1) too few inner loops -> warning, identical counts and velocities
    o = make_oracle(mesh_2k, st, nthreads=2, use_analytical_GL_flux=1, SSA_max_inner_loops=3, SSA_max_outer_loops=6)
    g = make_gpu(mesh_2k, st, use_analytical_GL_flux=1, SSA_max_inner_loops=3, SSA_max_outer_loops=6)
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    so, sg = o.solve_SSA(), g.solve_SSA()
    assert so.rc == 1 and sg.rc == 1
    assert (sg.n_outer, sg.n_inner_total) == (so.n_outer, so.n_inner_total) == (6, 18)
    assert rel_l2(g.download("U_SSA"), o["U_SSA"]) <= 1e-10
    # (2) over-relaxation far beyond 2 -> SOR diverges -> reset -> diverges again -> abort
    o = make_oracle(mesh_2k, st, nthreads=2, use_analytical_GL_flux=1, SSA_SOR_omega=3.5)
    g = make_gpu(mesh_2k, st, use_analytical_GL_flux=1, SSA_SOR_omega=3.5)
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    so = o.solve_SSA()
    assert so.rc == -1 and so.did_reset == 1
    with pytest.raises(UfmError) as ei:
        g.solve_SSA()
    assert ei.value.rc == -1 and "unstable" in str(ei.value)


def test_cpp_host_mirror(mesh_2k, tmp_path):
    """The reference's call sequence (src/UFEMISM_main_model.f90:90-135) issued from COMPILED host code through
    host/ufemism_host.hpp (the C++ twin of the Fortran shim) reproduces the oracle running the same sequence."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "run_steps")
    libdir = os.path.join(root, "ufemism_b200")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(root, "include"), os.path.join(root, "host", "run_steps.cpp"), "-o", exe,
                    "-L", libdir, "-lufemism_b200", f"-Wl,-rpath,{libdir}"], check=True)
    m = mesh_2k
    st = S.state_ssa_icestream(m, scale=750e3 / 1800e3, Hb=-250.0, H_shelf=150.0)
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        np.array([m.nV, m.nAc, m.nC_mem], np.int32).tofile(f)
        for name, dt in (("V", "f8"), ("A", "f8"), ("nC", "i4"), ("C", "i4"), ("Cw", "f8"), ("edge_index", "i4"), ("Nx", "f8"), ("Ny", "f8"), ("Aci", "i4"),
                         ("iAci", "i4"), ("edge_index_Ac", "i4"), ("Nx_Ac", "f8"), ("Ny_Ac", "f8"), ("No_Ac", "f8"), ("Np_Ac", "f8"), ("nCAaAc", "i4"), ("CAaAc", "i4"),
                         ("Nx_AaAc", "f8"), ("Ny_AaAc", "f8"), ("Nxx_AaAc", "f8"), ("Nxy_AaAc", "f8"), ("Nyy_AaAc", "f8"), ("colour_vi", "i4"), ("colour_nV", "i4")):
            np.asfortranarray(getattr(m, name), dtype=dt).ravel(order="F").tofile(f)
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
            np.asarray(st[k], "f8").tofile(f)
    n_steps = 3
    subprocess.run([exe, inp, outp, str(n_steps), "1"], check=True, timeout=300)
    raw = open(outp, "rb").read()
    N = m.nV
    Hi, U, V, Usia = (np.frombuffer(raw, "f8", N, i * 8 * N) for i in range(4))
    mask = np.frombuffer(raw, "i4", N, 32 * N)
    n_outer, n_inner = np.frombuffer(raw, "i4", 2, 36 * N)
    time = np.frombuffer(raw, "f8", 1, 36 * N + 8)[0]
    # the oracle, same sequence
    o = make_oracle(m, st, nthreads=2, use_analytical_GL_flux=1)
    t, dt, no, ni = 0.0, 0.0, 0, 0
    for _ in range(n_steps):
        o.calculate_ice_thickness_change(dt); o.update_general_ice_model_data(t); o.solve_SIA()
        s_ = o.solve_SSA(); no += s_.n_outer; ni += s_.n_inner_total
        dt = min(min(o.determine_timesteps()), 10.0); t += dt
    assert (n_outer, n_inner) == (no, ni) and abs(time - t) <= 1e-12 * t
    assert rel_l2(Hi, o["Hi"]) <= 1e-8 and rel_l2(U, o["U_SSA"]) <= 1e-10 and rel_l2(V, o["V_SSA"]) <= 1e-10
    np.testing.assert_allclose(Usia, o["U_SIA"], rtol=1e-12, atol=1e-12 * np.abs(o["U_SIA"]).max())
    assert np.array_equal(mask, o["mask"])


def test_hybrid_run_with_mesh_update(mesh_2k):
    """BASELINE config 4 in miniature: hybrid SIA/SSA steps, then a CPU mesh update (new mesh, Hi remapped on the host,
    everything else reallocated: src/UFEMISM_main_model.f90:240-315, src/ice_dynamics_module.f90:1208-1218) -> device
    re-upload -> more steps.  U_SSA restarts from zero after the update on both sides."""
    from scipy.spatial import cKDTree
    st = scenario(mesh_2k, "mismip")
    o, g = make_oracle(mesh_2k, st, nthreads=2, use_analytical_GL_flux=1), make_gpu(mesh_2k, st, use_analytical_GL_flux=1)
    ro, rg = o.region(0.0), g.region(0.0)
    o.run_model(ro, 1e12, max_steps=2); g.run_model(rg, 1e12, max_steps=2)
    assert rel_l2(g.download("Hi"), o["Hi"]) <= 1e-8
    # "run_model_update_mesh": a finer mesh; the host remaps Hi (nearest vertex here), Hb/SL/SMB are re-evaluated
    m2 = get_mesh(3500, seed=99)
    idx = cKDTree(mesh_2k.V).query(m2.V)[1]
    st2 = scenario(m2, "mismip")
    Hi_o, Hi_g = o["Hi"][idx].copy(), g.download("Hi")[idx].copy()
    Hi_o[m2.edge_index > 0] = 0.0; Hi_g[m2.edge_index > 0] = 0.0
    o2 = make_oracle(m2, dict(st2, Hi=Hi_o), nthreads=2, use_analytical_GL_flux=1)
    g.upload_mesh(m2)
    for k in ("Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st2[k])
    g.upload("Hi", Hi_g)
    assert not g.download("U_SSA").any()
    # all "do" flags true after a mesh update (:307-313), dt carried over
    ro2, rg2 = o2.region(ro.time), g.region(rg.time)
    ro2.dt = ro.dt; rg2.dt = rg.dt
    o2.run_model(ro2, 1e12, max_steps=3); g.run_model(rg2, 1e12, max_steps=3)
    assert (rg2.n_steps, rg2.n_sor_total, rg2.n_outer_total) == (ro2.n_steps, ro2.n_sor_total, ro2.n_outer_total)
    assert rel_l2(g.download("Hi"), o2["Hi"]) <= 1e-8
    assert rel_l2(g.download("U_SSA"), o2["U_SSA"]) <= 1e-8


def test_halfar_250k_vs_analytic_and_oracle():
    """BASELINE config 2: Halfar dome on a ~250 k-vertex mesh (h ~ 3.2 km), started from Halfar_solution(t = 1000 yr) where the
    reference's diffusivity clip is inactive.  (a) the first steps against the oracle; (b) 150 more model years (CFL-limited
    steps, all on the device) against Halfar_solution(t) and volume conservation."""
    m = get_mesh(250000, half_width=750e3)
    st = S.state_halfar(m, t=1000.0)
    g = make_gpu(m, st)
    o = make_oracle(m, st, nthreads=8)
    ro, rg = o.region(0.0), g.region(0.0)
    o.run_model(ro, 1e12, max_steps=4); g.run_model(rg, 1e12, max_steps=4)
    assert rg.time == ro.time and rel_l2(g.download("Hi"), o["Hi"]) <= 1e-12
    vol0 = float((st["Hi"] * m.A).sum())
    g.run_model(rg, 150.0)
    assert rg.time == 150.0 and rg.n_steps > 100
    H = g.download("Hi")
    Han = S.halfar_H(5000.0, 300000.0, m.V[:, 0], m.V[:, 1], 1150.0)
    assert rel_l2(st["Hi"], Han) > 0.01            # the solution moved ...
    assert rel_l2(H, Han) < 0.004                  # ... and the run followed it
    assert abs(H.max() / Han.max() - 1.0) < 1e-3
    assert abs(float((H * m.A).sum()) / vol0 - 1.0) < 1e-9


@pytest.mark.parametrize("order", [1, 2])
def test_conservative_remap_application(mesh_2k, order):
    """Row N3: remap_cons_{1st,2nd}_order_2D applied on the device across a mesh update, bit-exact against the oracle.  The
    weights are synthetic (3 nearest source vertices, inverse-distance w0, w1 = w0 * offset); building real conservative
    weights is CPU work outside the path."""
    from scipy.spatial import cKDTree
    st = scenario(mesh_2k, "mismip")
    o, g = make_oracle(mesh_2k, st), make_gpu(mesh_2k, st)
    m2 = get_mesh(3500, seed=99)
    dist, idx = cKDTree(mesh_2k.V).query(m2.V, k=3)
    w = 1.0 / (dist + 1.0); w /= w.sum(1, keepdims=True)
    n = m2.nV
    vli1 = 1 + 3 * np.arange(n); vli2 = vli1 + 2
    vli2[::7] = vli1[::7] + 1                      # ragged: some destination vertices use only two entries
    vli1[5::11] = 1; vli2[5::11] = 0               # ... and some none (empty range -> 0)
    vi = (idx + 1).ravel(); w0 = w.ravel()
    w1x = (w * (m2.V[:, None, 0] - mesh_2k.V[idx, 0])).ravel(); w1y = (w * (m2.V[:, None, 1] - mesh_2k.V[idx, 1])).ravel()
    want = o.remap_cons_2D(order, vli1, vli2, vi, w0, w1x if order == 2 else None, w1y if order == 2 else None, st["Hi"])
    g.remap_stash("Hi")
    g.upload_mesh(m2)
    g.remap_apply("Hi", vli1, vli2, vi, w0, *( (w1x, w1y) if order == 2 else ()))
    got = g.download("Hi")
    assert_bits_equal(got, want, "remapped Hi")
    assert got[5] == 0.0 and np.abs(got).max() > 100.0
    from ufemism_b200.capi import UfmError
    with pytest.raises(UfmError):
        g.remap_apply("Hi", vli1, vli2, vi, w0)    # nothing stashed any more


# ---------------------------------------------------------------- thermodynamics
THERMO_IN = ("Hi", "Hb", "SL", "SMB_year", "BMB", "T2m", "GHF", "Ti")


def _thermo_pair(mesh, benchmark, nthreads=4, **params):
    from oracle.oracle import Oracle
    from ufemism_b200.capi import IceModelGPU

    st = S.state_thermo_dome(mesh, benchmark=benchmark)
    o = Oracle(mesh, benchmark=benchmark, nthreads=nthreads)
    g = IceModelGPU(mesh, benchmark=benchmark, thermo=True, **params)
    for k in THERMO_IN:
        o[k][:] = st[k]
        g.upload(k, st[k])
    o.update_general_ice_model_data(0.0)
    g.update_general_ice_model_data(0.0)
    return st, o, g


def test_thermo_w3d_and_heat_equation_bit_exact(mesh_10k):
    """Vertical velocity, heat equation (upwind advection, DGTSV, pressure-melting clamps), 3-D Neumann pass: with the EISMINT
    ice properties and no sliding every operation is +,-,*,/,sqrt, so from identical U_3D / V_3D the device must give the
    oracle's bits, step after step."""
    st, o, g = _thermo_pair(mesh_10k, "EISMINT_1")
    o.solve_SIA_3D(with_W=True)
    g.upload("U_3D", o["U_3D"]); g.upload("V_3D", o["V_3D"])
    g.thermo_w3d()
    assert np.abs(o["W_3D"]).max() > 0.1
    assert_bits_equal(g.download("W_3D"), o["W_3D"], "W_3D")
    for step in range(3):
        rc, n_unstable = o.update_ice_temperature()
        assert rc == 0 and n_unstable == 0
        ts = g.thermo_heat()
        assert ts.n_unstable == 0
        assert_bits_equal(g.download("Ti"), o["Ti"], f"Ti after step {step + 1}")
    assert_bits_equal(g.download("frictional_heating"), o["frictional_heating"], "frictional_heating")
    assert np.abs(o["Ti"] - st["Ti"]).max() > 0.05          # the field did move


@pytest.mark.parametrize("benchmark", ["EISMINT_1", "none"])
def test_update_ice_temperature_whole_routine(mesh_10k, benchmark):
    """ufm_update_ice_temperature end to end (solve_SIA_3D, frictional heating, heat equation, safety net) against the
    oracle; the realistic branch adds sliding (pow), temperature-dependent conductivity (exp) and Arrhenius flow factors."""
    st, o, g = _thermo_pair(mesh_10k, benchmark)
    if benchmark == "none":
        so, sg = o.solve_SSA(), g.solve_SSA()
        assert (so.n_outer, so.n_inner_total) == (sg.n_outer, sg.n_inner_total)
    for step in range(2):
        rc, n_unstable = o.update_ice_temperature()
        ts = g.update_ice_temperature()
        assert rc == 0 and (ts.rc, ts.n_unstable) == (0, n_unstable)
        o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    np.testing.assert_allclose(g.download("U_3D"), o["U_3D"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g.download("W_3D"), o["W_3D"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(g.download("Ti"), o["Ti"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(g.download("frictional_heating"), o["frictional_heating"], rtol=1e-12, atol=1e-12)
    if benchmark == "none":
        assert o["frictional_heating"].max() > 1.0


def test_thermo_safety_net_and_abort(mesh_2k):
    """Columns colder than 150 K are replaced by the Robin solution (erf: <= 2 ulp vs glibc); more than 1 % of them is fatal
    in the reference (STOP, src/thermodynamics_module.f90:195-199) -> rc -8."""
    from ufemism_b200.capi import UfmError

    st, o, g = _thermo_pair(mesh_2k, "EISMINT_1")
    Ti = st["Ti"].copy()
    cold = np.flatnonzero((mesh_2k.edge_index == 0) & (st["Hi"] > 1500.0))[:5]
    Ti[cold, 7] = -4000.0        # a wildly wrong layer: the implicit step leaves the column below 150 K
    o["Ti"][:] = Ti; g.upload("Ti", Ti)
    rc, n_unstable = o.update_ice_temperature()
    ts = g.update_ice_temperature()
    assert rc == 0 and n_unstable >= 5 and ts.n_unstable == n_unstable
    np.testing.assert_allclose(g.download("Ti"), o["Ti"], rtol=1e-12)
    Ti[:] = 100.0
    o["Ti"][:] = Ti; g.upload("Ti", Ti)
    rc, _ = o.update_ice_temperature()
    assert rc == -1
    with pytest.raises(UfmError) as e:
        g.update_ice_temperature()
    assert e.value.rc == -8


def test_thermo_needs_triangle_data(mesh_2k):
    from ufemism_b200.capi import IceModelGPU, UfmError

    g = IceModelGPU(mesh_2k, benchmark="EISMINT_1")
    with pytest.raises(UfmError) as e:
        g.update_ice_temperature()
    assert e.value.rc == -2
    g = IceModelGPU(mesh_2k, benchmark="Halfar", thermo=True)      # thermodynamics_module.f90:52-57: not included -> no-op
    assert g.update_ice_temperature().rc == 0


def test_run_model_eismint1_with_thermodynamics(mesh_2k):
    """The region loop with the whole update_ice_temperature on the thermodynamics timer (dt_thermo = 10 yr)."""
    from oracle.oracle import Oracle
    from ufemism_b200.capi import IceModelGPU

    st = S.state_thermo_dome(mesh_2k, benchmark="EISMINT_1")
    o = Oracle(mesh_2k, benchmark="EISMINT_1", nthreads=4, thermo=1)
    g = IceModelGPU(mesh_2k, benchmark="EISMINT_1", thermo=True)
    for k in THERMO_IN:
        o[k][:] = st[k]
        g.upload(k, st[k])
    ro, rg = o.region(0.0), g.region(0.0)
    assert o.run_model(ro, 25.0) == 0
    g.run_model(rg, 25.0)
    assert (rg.n_steps, rg.n_sia) == (ro.n_steps, ro.n_sia) and rg.time == ro.time
    assert rel_l2(g.download("Hi"), o["Hi"]) <= 1e-8
    np.testing.assert_allclose(g.download("Ti"), o["Ti"], rtol=1e-11)
    assert np.abs(o["Ti"] - st["Ti"]).max() > 0.05


# ---------------------------------------------------------------- mesh data derived on the device
@pytest.mark.parametrize("exact_xy", [1, 0])
def test_device_derived_neighbour_functions(mesh_10k, exact_xy):
    """With no Nx_AaAc ... Nyy_AaAc in the mesh descriptor the library evaluates get_neighbour_functions_vertex_gr on the
    device.  Everything downstream of those coefficients (viscosity gradients, centre coefficients, both cross-term modes of
    the sweep, the whole solve) must have the bits of a run on host-built coefficients."""
    st = scenario(mesh_10k, "icestream")
    ga = make_gpu(mesh_10k, st, use_analytical_GL_flux=1, exact_xy=exact_xy)
    gb = make_gpu(mesh_10k, st, use_analytical_GL_flux=1, exact_xy=exact_xy, derive_nf=True)
    for g in (ga, gb):
        g.update_general_ice_model_data(0.0)
    for f in AA_EXACT + AC_EXACT:      # Aa (Nx, Ny) and Ac (Nx_Ac, Ny_Ac, No_Ac, Np_Ac) neighbour functions
        assert_bits_equal(gb.download(f), ga.download(f), f)
    sa, sb = ga.solve_SSA(), gb.solve_SSA()
    assert (sa.n_outer, sa.n_inner_total, sa.last_max_residual, sa.last_RN) == (sb.n_outer, sb.n_inner_total, sb.last_max_residual, sb.last_RN)
    assert sa.n_inner_total > 20
    for f in ("U_SSA", "V_SSA", "U_SSA_AaAc", "V_SSA_AaAc", "eta_AaAc", "eu_i_AaAc", "ev_i_AaAc", "RHSx_AaAc", "dU_dx_AaAc", "dV_dy_AaAc", "dU_dy_AaAc"):
        assert_bits_equal(gb.download(f), ga.download(f), f)


# ---------------------------------------------------------------- rows wider than the unrolled kernels (degree 9..16)
def test_high_degree_rows_take_the_generic_paths(mesh_fan):
    """Degree-10, -14 and -16 vertices (16 = nC_mem): geometry, thickness update, SOR sweep (bit-exact at 40 iterations), the
    whole SSA solve, device-derived neighbour functions and thermodynamics on a mesh whose widest slices use the generic
    row loops instead of the width-templated ones."""
    m = mesh_fan
    st = scenario(m, "icestream")
    o = make_oracle(m, st, nthreads=4)
    g = make_gpu(m, st)
    gd = make_gpu(m, st, derive_nf=True)
    o.update_general_ice_model_data(0.0)
    for q in (g, gd):
        q.update_general_ice_model_data(0.0)
    for f in AA_EXACT + AC_EXACT:
        assert_bits_equal(g.download(f), o[f], f)
        assert_bits_equal(gd.download(f), o[f], f + " (device-derived neighbour functions)")
    # SOR on the oracle's system
    o.basal_yield_stress(); o.SSA_gather_AaAc(); o.SSA_effective_viscosity(); o.SSA_sliding_term()
    for q in (g, gd):
        q.ssa_prepare(); q.upload("tau_c_AaAc", o["tau_c_AaAc"]); q.ssa_viscosity(); q.upload("eta_AaAc", o["eta_AaAc"]); q.ssa_sliding_and_setup()
    n, res, _, _ = o.solve_SSA_linearised(max_inner=40, force_iters=True)
    for q in (g, gd):
        for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):
            q.upload(f, o[f])
        stq = q.ssa_sor(max_inner=40, force_iters=True)
        assert stq.n_inner_last == 40 == n and stq.last_max_residual == res
        assert_bits_equal(q.download("U_SSA_AaAc"), o["U_SSA_AaAc"], "U_SSA_AaAc")
        assert_bits_equal(q.download("V_SSA_AaAc"), o["V_SSA_AaAc"], "V_SSA_AaAc")
    # whole solve + a thickness step
    o2 = make_oracle(m, st, nthreads=4, use_analytical_GL_flux=1)
    g2 = make_gpu(m, st, use_analytical_GL_flux=1)
    o2.update_general_ice_model_data(0.0); g2.update_general_ice_model_data(0.0)
    so, sg = o2.solve_SSA(), g2.solve_SSA()
    assert (so.n_outer, so.n_inner_total) == (sg.n_outer, sg.n_inner_total)
    assert rel_l2(g2.download("U_SSA"), o2["U_SSA"]) <= 1e-10 and rel_l2(g2.download("V_SSA"), o2["V_SSA"]) <= 1e-10
    o2.solve_SIA(); g2.solve_SIA()
    o2.calculate_ice_thickness_change(0.5); g2.calculate_ice_thickness_change(0.5)
    assert rel_l2(g2.download("Hi"), o2["Hi"]) <= 1e-12


def test_high_degree_thermodynamics_bit_exact(mesh_fan):
    st, o, g = _thermo_pair(mesh_fan, "EISMINT_1")
    o.solve_SIA_3D(with_W=True)
    g.upload("U_3D", o["U_3D"]); g.upload("V_3D", o["V_3D"])
    g.thermo_w3d()
    assert_bits_equal(g.download("W_3D"), o["W_3D"], "W_3D")
    rc, n_unstable = o.update_ice_temperature()
    ts = g.thermo_heat()
    assert rc == 0 and (ts.n_unstable, n_unstable) == (0, 0)
    assert_bits_equal(g.download("Ti"), o["Ti"], "Ti")


# ---------------------------------------------------------------- closed-form benchmark SMB on the device (row N1)
@pytest.mark.parametrize("benchmark", ["Bueler", "EISMINT_2", "EISMINT_5"])
def test_run_model_time_dependent_benchmark_smb(mesh_2k, benchmark):
    """run_SMB_model's benchmark branches are evaluated on the device on the SMB timer (src/SMB_module.f90:55-97, 172-283):
    Bueler's mass balance (two pow per vertex) and the sinusoidal EISMINT forcings; same step sequence as the oracle."""
    from oracle.oracle import Oracle, bueler_solution
    from ufemism_b200.capi import IceModelGPU

    m = mesh_2k
    x, y = m.V[:, 0], m.V[:, 1]
    if benchmark == "Bueler":
        H0, R0, lam, t0 = 3000.0, 500e3, 5.0, 10764.260159329711
        ts, te = 0.6 * t0, 0.6 * t0 + 120.0
        Hi = bueler_solution(H0, R0, lam, x, y, ts)
    else:
        H0, R0, lam = 5000.0, 300e3, 5.0
        ts, te = 4900.0, 5030.0          # a quarter period into the 20 kyr cycle: E resp. M_max well away from their means
        Hi = S.state_thermo_dome(m)["Hi"]
    o = Oracle(m, benchmark=benchmark, nthreads=4)
    g = IceModelGPU(m, benchmark=benchmark)
    o["Hi"][:] = Hi; o["SL"][:] = -10000.0
    g.upload("Hi", Hi); g.upload("SL", np.full(m.nV, -10000.0))
    ro, rg = o.region(ts), g.region(ts)
    ro.H0, ro.R0, ro.lam = H0, R0, lam
    rg.H0, rg.R0, rg.lam = H0, R0, lam
    assert o.run_model(ro, te) == 0
    g.run_model(rg, te)
    assert (rg.n_steps, rg.n_sia) == (ro.n_steps, ro.n_sia) and abs(rg.time - ro.time) <= 1e-9 * te
    np.testing.assert_allclose(g.download("SMB_year"), o["SMB_year"], rtol=1e-13, atol=1e-13)   # hypot: <= 1 ulp of ~1e5 m, times S_b = 1e-5
    assert np.ptp(o["SMB_year"]) > 0.1
    assert rel_l2(g.download("Hi"), o["Hi"]) <= 1e-8


def test_ssa_solve_after_cfl_with_3d_velocities(mesh_10k):
    """Regression: the CFL minima, the RN reduction of the viscosity loop and the thermodynamics status share one control
    block on the device.  With non-trivial U_3D / V_3D the third CFL key has arbitrary low bits; a solve_SSA issued after it
    must still count its outer iterations exactly like the oracle."""
    st, o, g = _thermo_pair(mesh_10k, "none")
    so, sg = o.solve_SSA(), g.solve_SSA()
    assert (so.n_outer, so.n_inner_total) == (sg.n_outer, sg.n_inner_total)
    o.update_ice_temperature(); g.update_ice_temperature()
    do, dg = o.determine_timesteps(), g.determine_timesteps()
    np.testing.assert_allclose(dg, do, rtol=1e-12)
    assert do[2] < 900.0                      # the 3-D velocities do set a finite critical time step
    o["Hi"][:] = o["Hi"] * 0.98; g.upload("Hi", o["Hi"])
    o.update_general_ice_model_data(0.0); g.update_general_ice_model_data(0.0)
    so, sg = o.solve_SSA(), g.solve_SSA()
    assert so.n_outer >= 2 and (so.n_outer, so.n_inner_total) == (sg.n_outer, sg.n_inner_total)
    assert rel_l2(g.download("U_SSA"), o["U_SSA"]) <= 1e-10
