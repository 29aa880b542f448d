"""The oracle, the mesh substrate and the CUDA path against the REFERENCE'S OWN SOURCE.

The Fortran reference cannot be compiled here, but its source can be read: oracle/f90py.py translates the text of 78 hot-path
routines (src/ice_dynamics_module.f90, general_ice_model_data_module.f90, mesh_ArakawaC_module.f90, mesh_derivatives_module.f90,
mesh_five_colour_module.f90, mesh_help_functions_module.f90 (Voronoi cell areas, connection widths), zeta_module.f90,
thermodynamics_module.f90, SMB_module.f90, reference_fields_module.f90, mesh_mapping_module.f90, UFEMISM_main_model.f90) statement by statement into
Python (single MPI rank, IEEE double arithmetic in source order, libm for the transcendental intrinsics) and this module runs them
on the golden 600-vertex mesh:

* `test_live_*` (only where /root/reference is mounted): translated reference == oracle / mesh substrate, **bit for bit**, for
  update_general_ice_model_data (all 22 masks, every gradient), calculate_ice_thickness_change, solve_SIA, basal_yield_stress,
  SSA_effective_viscosity, SSA_sliding_term, the five-colour SOR sweep with its Neumann pass, the whole solve_SSA with the
  analytical grounding-line flux, the critical time steps, update_ice_temperature (one heat-equation step with DGTSV from a real LAPACK,
  EISMINT and temperature-dependent ice properties), run_SMB_model (every benchmark branch), the Halfar / Bueler solutions, the
  application of a conservative remapping (1st and 2nd order), find_Voronoi_cell_areas, find_connection_widths, get_neighbour_functions, make_Ac_mesh (+ the combined AaAc mesh and
  its neighbour functions) and calculate_five_colouring_AaAc (the SOR order).
* `test_golden_*` (everywhere): the same comparison against the committed outputs of the translated reference
  (tests/golden/reference_source_600.npz, written by tests/golden/make_reference_source_golden.py), so the pin travels to machines
  without the reference -- including the B200 box, where `test_gpu_matches_reference_source` holds the CUDA path to them.
"""
import hashlib
import os

import numpy as np
import pytest

from tests import ref_cases as RC
from tests import ref_source as RS
from tests.util import assert_bits_equal, make_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_source_600.npz")


def _arr(x):
    from oracle import f90py as F
    return x.a if isinstance(x, F.FArray) else np.asarray(x)


# ---------------------------------------------------------------------------------------------------------------------------
# the cases, once for the translated reference source ...
# ---------------------------------------------------------------------------------------------------------------------------
def run_reference_source(mesh):
    """Every case through the translated reference; returns {case__field: array}."""
    from oracle import f90py as F

    np.seterr(all="ignore")
    st = RC.start_state(mesh)
    o = make_oracle(mesh, st, nthreads=1, use_analytical_GL_flux=1)   # only a container of correctly shaped, zeroed reference arrays
    P = RS.program(o.cfg)
    P.C.choice_benchmark_experiment = st["benchmark"]
    out = {}

    def grab(case, ns, fields):
        for f in fields:
            out[f"{case}__{f}"] = np.array(_arr(getattr(ns, f)))

    # mesh operators from the primary mesh data
    out.update(reference_mesh_operators(P, mesh))
    # geometry, masks, gradients
    mref = RS.mesh_ns(mesh)
    ice = RS.ice_ns(o)
    P.update_general_ice_model_data(mref, ice, np.float64(0.0))
    grab("general", ice, RC.GENERAL_FIELDS)
    # SIA
    P.solve_sia(mref, ice)
    grab("sia", ice, RC.SIA_FIELDS)
    # pieces of solve_SSA from prescribed velocities
    U, V = RC.random_velocities(mesh)
    P.basal_yield_stress(mref, ice)
    _gather_aaac(mesh, ice)
    ice.u_ssa_aaac.a[:] = U; ice.v_ssa_aaac.a[:] = V
    P.ssa_effective_viscosity(mref, ice)
    P.ssa_sliding_term(mref, ice)
    grab("pieces", ice, RC.PIECES_FIELDS)
    # SOR: SOR_ITERS sweeps incl. the Neumann pass
    P.C.ssa_max_inner_loops = RC.SOR_ITERS
    P.solve_ssa_linearised(mref, ice, False)
    grab("sor", ice, RC.SOR_FIELDS)
    P.C.ssa_max_inner_loops = 10000
    # the whole solve_SSA (analytical GL flux on), a few viscosity iterations, from rest
    ice2 = RS.ice_ns(o)
    P.update_general_ice_model_data(mref, ice2, np.float64(0.0))
    P.solve_sia(mref, ice2)
    P.C.ssa_max_outer_loops, P.C.ssa_max_inner_loops = RC.SSA_OUTER, RC.SSA_INNER
    P.solve_ssa(mref, ice2)
    P.C.ssa_max_outer_loops, P.C.ssa_max_inner_loops = 50, 10000
    grab("ssa", ice2, RC.SSA_FIELDS)
    # mass continuity with those velocities, then the critical time steps
    smb, bmb = F.NS(smb_year=np.array(st["SMB_year"])), F.NS(bmb=np.array(st["BMB"]))
    P.calculate_ice_thickness_change(mref, ice2, smb, bmb, np.float64(0.5), np.zeros(mesh.nV, np.int32))
    grab("thk", ice2, RC.THK_FIELDS)
    P.C.dt_max, P.C.dt_thermo, P.C.dt_climate, P.C.dt_smb, P.C.dt_bmb, P.C.dt_bedrock_elra, P.C.dt_output = (np.float64(1e9),) * 7
    region = F.NS(mesh=mref, ice=ice2, time=np.float64(0.0), dt=np.float64(0.0), dt_prev=np.float64(1.0))
    for t in ("sia", "ssa", "thermo", "climate", "smb", "bmb", "elra", "output"):
        setattr(region, "t0_" + t, np.float64(0.0))
    P.determine_timesteps_and_actions(region, np.float64(1e12))
    out["cfl__dt_SIA_dt_SSA"] = np.array([region.dt_sia, region.dt_ssa])
    # thermodynamics: one implicit step of the heat equation (solve_SIA_3D incl. W, frictional heating, zeta Jacobians, upwind
    # advection, DGTSV through a real LAPACK, 3-D Neumann pass), EISMINT and temperature-dependent ice properties
    from oracle.oracle import Oracle
    from ufemism_b200 import scenarios as S
    for bm in RC.THERMO_BENCHMARKS:
        stt = S.state_thermo_dome(mesh, benchmark=bm)
        ot = Oracle(mesh, benchmark=bm, nthreads=1)               # container of zeroed reference arrays
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB", "T2m", "GHF", "Ti"):
            ot[k][:] = stt[k]
        Pt = RS.program(ot.cfg)
        Pt.C.do_benchmark_experiment, Pt.C.choice_benchmark_experiment = (bm != "none"), bm
        icet = RS.ice_ns(ot)
        Pt.update_general_ice_model_data(mref, icet, np.float64(0.0))
        if bm == "none":
            Pt.C.ssa_max_outer_loops, Pt.C.ssa_max_inner_loops = RC.THERMO_SSA_OUTER, RC.THERMO_SSA_INNER
            Pt.solve_ssa(mref, icet)
        Pt.initialize_zeta_discretization()
        Pt.update_ice_temperature(mref, icet, F.NS(applied=F.NS(t2m=np.array(stt["T2m"], order="F"))), F.NS(smb_year=np.array(stt["SMB_year"])))
        grab("thermo_" + bm, icet, RC.THERMO_FIELDS)
    # run_SMB_model, benchmark branches (row N1), and the analytic solutions the benchmarks start from
    P.C.halfar_solution_h0, P.C.halfar_solution_r0, P.C.bueler_solution_lambda = np.float64(RC.H0), np.float64(RC.R0), np.float64(RC.LAMBDA)
    for bm, t in RC.SMB_CASES:
        P.C.do_benchmark_experiment, P.C.choice_benchmark_experiment = True, bm
        smbm = F.NS(smb_year=np.full(mesh.nV, -9.0), smb=np.zeros((mesh.nV, 12), order="F"))
        P.run_smb_model(mref, None, None, np.float64(t), smbm, None)
        out[f"smb__{bm}_{t:g}"] = np.array(smbm.smb_year.a)
    for t in RC.ANALYTIC_TIMES:
        out[f"analytic__halfar_{t:g}"] = np.array([P.halfar_solution(np.float64(RC.H0), np.float64(RC.R0), x, y, np.float64(t)) for x, y in np.asarray(mesh.V)])
        out[f"analytic__bueler_{t:g}"] = np.array([P.bueler_solution(np.float64(RC.H0), np.float64(RC.R0), np.float64(RC.LAMBDA), x, y, np.float64(t)) for x, y in np.asarray(mesh.V)])
    # application of a conservative remapping (row N3)
    mp = RC.remap_map(mesh)
    mapns = F.NS(**{k: np.array(v) for k, v in mp.items() if k != "d_src"})
    dst = F.NS(v1=1, v2=len(mp["vli1"]))
    for order in (1, 2):
        d_dst = np.zeros(len(mp["vli1"]))
        if order == 1:
            P.remap_cons_1st_order_2d(dst, mapns, np.array(mp["d_src"]), d_dst)
        else:
            P.remap_cons_2nd_order_2d(mref, dst, mapns, np.array(mp["d_src"]), d_dst)
        out[f"remap__order{order}"] = d_dst
    return out


def reference_mesh_operators(P, mesh):
    """Secondary mesh data as the translated reference derives them from the primary data: find_Voronoi_cell_areas,
    find_connection_widths, get_neighbour_functions, make_Ac_mesh (incl. find_Ac_edge_indices, make_combined_AaAc_mesh and its
    neighbour functions), calculate_five_colouring_AaAc."""
    out = {}
    mm = RS.mesh_ns(mesh)
    for f in ("nx", "ny", "nxx", "nxy", "nyy", "nxtri", "nytri"):
        getattr(mm, f).a[...] = -7.0
    mm.a.a[...] = -7.0; mm.cw.a[...] = 0.0
    P.find_voronoi_cell_areas(mm)
    P.find_connection_widths(mm)            # the reference leaves Cw beyond nC(vi) untouched: compared as zeros on both sides
    mm.r.a[...] = -7.0
    P.determine_mesh_resolution(mm)
    P.get_neighbour_functions(mm)
    P.make_ac_mesh(mm)
    P.calculate_five_colouring_aaac(mm)
    assert int(mm.nac) == mesh.nAc
    for f in RC.MESH_FIELDS:
        a = np.array(_arr(getattr(mm, f)))
        want = np.asarray(getattr(mesh, f)).shape
        out[f"mesh__{f}"] = a[: want[0]] if a.ndim >= 1 and a.shape[0] > want[0] else a
    return out


def _gather_aaac(mesh, ice):
    """The inline gather of solve_SSA (src/ice_dynamics_module.f90:468-496): Aa and Ac fields side by side on the combined mesh."""
    nV = mesh.nV
    for f in ("Hi", "Hb", "SL", "dHs_dx_shelf", "dHs_dy_shelf", "A_flow_mean"):
        getattr(ice, f + "_AaAc").a[:nV] = getattr(ice, f).a
        getattr(ice, f + "_AaAc").a[nV:] = getattr(ice, f + "_Ac").a


# ... and once for the oracle + mesh substrate
def run_oracle(mesh):
    from ufemism_b200 import mesh as M

    st = RC.start_state(mesh)
    out = {}
    fresh = M.build_mesh(np.array(mesh.V), mesh.xmin, mesh.xmax, mesh.ymin, mesh.ymax)   # the substrate, from the primary data again
    assert np.array_equal(fresh.Tri, mesh.Tri)
    for f in RC.MESH_FIELDS:
        out[f"mesh__{f}"] = np.asarray(getattr(fresh, f))
    out["mesh__Cw"] = np.where(np.arange(mesh.nC_mem)[None, :] < fresh.nC[:, None], fresh.Cw, 0.0)

    def grab(case, o, fields):
        for f in fields:
            out[f"{case}__{f}"] = o[f].copy()

    o = make_oracle(mesh, st, nthreads=1, use_analytical_GL_flux=1)
    o.update_general_ice_model_data(0.0)
    grab("general", o, RC.GENERAL_FIELDS)
    o.solve_SIA()
    grab("sia", o, RC.SIA_FIELDS)
    U, V = RC.random_velocities(mesh)
    o.basal_yield_stress(); o.SSA_gather_AaAc()
    o["U_SSA_AaAc"][:] = U; o["V_SSA_AaAc"][:] = V
    o.SSA_effective_viscosity(); o.SSA_sliding_term()
    grab("pieces", o, RC.PIECES_FIELDS)
    n, res, _, _ = o.solve_SSA_linearised(max_inner=RC.SOR_ITERS)
    assert n == RC.SOR_ITERS
    grab("sor", o, RC.SOR_FIELDS)
    o2 = make_oracle(mesh, st, nthreads=1, use_analytical_GL_flux=1, SSA_max_outer_loops=RC.SSA_OUTER, SSA_max_inner_loops=RC.SSA_INNER)
    o2.update_general_ice_model_data(0.0); o2.solve_SIA()
    s = o2.solve_SSA()
    assert s.n_outer == RC.SSA_OUTER
    grab("ssa", o2, RC.SSA_FIELDS)
    o2.calculate_ice_thickness_change(0.5)
    grab("thk", o2, RC.THK_FIELDS)
    d = o2.determine_timesteps()
    out["cfl__dt_SIA_dt_SSA"] = np.array([min(d[0], d[2]), d[1]])
    from oracle.oracle import Oracle
    from ufemism_b200 import scenarios as S
    for bm in RC.THERMO_BENCHMARKS:
        stt = S.state_thermo_dome(mesh, benchmark=bm)
        ot = Oracle(mesh, benchmark=bm, nthreads=1, SSA_max_outer_loops=RC.THERMO_SSA_OUTER, SSA_max_inner_loops=RC.THERMO_SSA_INNER)
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB", "T2m", "GHF", "Ti"):
            ot[k][:] = stt[k]
        ot.update_general_ice_model_data(0.0)
        if bm == "none":
            ot.solve_SSA()
        rc, nu = ot.update_ice_temperature()
        assert (rc, nu) == (0, 0)
        grab("thermo_" + bm, ot, RC.THERMO_FIELDS)
    from oracle.oracle import bueler_solution, halfar_solution
    for bm, t in RC.SMB_CASES:
        ob = Oracle(mesh, benchmark=bm, nthreads=1)
        ob["SMB_year"][:] = -9.0
        ob.run_SMB_benchmark(t, RC.H0, RC.R0, RC.LAMBDA)
        out[f"smb__{bm}_{t:g}"] = ob["SMB_year"].copy()
    V = np.asarray(mesh.V)
    for t in RC.ANALYTIC_TIMES:
        out[f"analytic__halfar_{t:g}"] = halfar_solution(RC.H0, RC.R0, V[:, 0], V[:, 1], t)
        out[f"analytic__bueler_{t:g}"] = bueler_solution(RC.H0, RC.R0, RC.LAMBDA, V[:, 0], V[:, 1], t)
    mp = RC.remap_map(mesh)
    for order in (1, 2):
        out[f"remap__order{order}"] = o.remap_cons_2D(order, mp["vli1"], mp["vli2"], mp["vi"], mp["w0"], mp["w1x"] if order == 2 else None,
                                                     mp["w1y"] if order == 2 else None, mp["d_src"])
    return out


def _compare(got, want, keys=None):
    bad = []
    for k in sorted(keys or want):
        try:
            assert_bits_equal(np.asarray(got[k]), np.asarray(want[k]), k)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, "\n".join(bad)


@pytest.fixture(scope="module")
def oracle_out():
    return run_oracle(RC.golden_mesh())


@pytest.fixture(scope="module")
def golden():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


@pytest.mark.skipif(not RS.available(), reason="/root/reference is not mounted here; the golden vectors stand in (test_golden_*)")
def test_live_translated_reference_source_equals_oracle_and_golden(oracle_out, golden):
    ref = run_reference_source(RC.golden_mesh())
    assert len(ref) >= 120
    _compare(oracle_out, ref)                               # every output of every routine, bit for bit
    for k, a in ref.items():                                # and the committed golden file is what the reference source gives today
        assert str(golden["sha256__" + k]) == hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest(), k
    # the cases are not trivial
    assert np.abs(ref["ssa__U_SSA"]).max() > 10.0 and np.abs(ref["sor__resU_AaAc"]).max() > 1.0 and ref["general__mask_shelf"].sum() > 50
    assert ref["general__mask_gl"].sum() > 10 and np.abs(ref["thk__dHi_dt"]).max() > 0.1 and np.abs(ref["ssa__Qabs_GL_Ac"]).max() > 0.0


def test_golden_reference_source_vectors_pin_the_oracle_and_the_mesh_substrate(oracle_out, golden):
    names = [k[len("sha256__"):] for k in golden if k.startswith("sha256__")]
    assert len(names) >= 120 and set(names) == set(oracle_out)
    for k in names:
        a = np.ascontiguousarray(oracle_out[k])
        assert str(golden["sha256__" + k]) == hashlib.sha256(a.tobytes()).hexdigest(), f"{k}: the oracle no longer reproduces the reference source"
    _compare(oracle_out, golden, [k for k in golden if not k.startswith("sha256__")])


@pytest.mark.gpu
def test_gpu_matches_reference_source(golden):
    """The CUDA path against the outputs of the translated reference source: bit-exact wherever only + - * / sqrt are involved
    (geometry, masks, gradients, thickness update, SOR sweeps at a prescribed count), <= 1e-13 behind one pow / tan."""
    from tests.util import make_gpu

    mesh = RC.golden_mesh()
    st = RC.start_state(mesh)
    g = make_gpu(mesh, st, use_analytical_GL_flux=1)
    g.update_general_ice_model_data(0.0)
    for f in RC.GENERAL_FIELDS:
        assert_bits_equal(g.download(f), golden["general__" + f], f)
    g.solve_SIA()
    for f in ("D_SIA_Ac", "Up_SIA_Ac", "U_SIA", "V_SIA", "D_SIA"):
        np.testing.assert_allclose(g.download(f), golden["sia__" + f], rtol=1e-13, atol=1e-13 * np.abs(golden["sia__" + f]).max(), err_msg=f)
    U, V = RC.random_velocities(mesh)
    g.ssa_prepare()
    np.testing.assert_allclose(g.download("tau_c_AaAc"), golden["pieces__tau_c_AaAc"], rtol=1e-13)
    g.upload("tau_c_AaAc", golden["pieces__tau_c_AaAc"])
    g.upload("U_SSA_AaAc", U); g.upload("V_SSA_AaAc", V)
    g.ssa_viscosity()
    np.testing.assert_allclose(g.download("eta_AaAc"), golden["pieces__eta_AaAc"], rtol=1e-13)
    assert_bits_equal(g.download("dU_dx_AaAc"), golden["pieces__dU_SSA_dx_AaAc"], "dU_SSA_dx_AaAc")
    g.upload("eta_AaAc", golden["pieces__eta_AaAc"])
    g.ssa_sliding_and_setup()
    np.testing.assert_allclose(g.download("S_AaAc"), golden["pieces__S_AaAc"], rtol=1e-13)
    for f in ("RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc"):     # identical linear system -> the sweep itself must give the reference's bits
        g.upload(f, golden["sor__" + f])
    s = g.ssa_sor(max_inner=RC.SOR_ITERS)
    assert s.n_inner_last == RC.SOR_ITERS
    assert_bits_equal(g.download("U_SSA_AaAc"), golden["sor__U_SSA_AaAc"], "U_SSA_AaAc after the reference's SOR sweeps")
    assert_bits_equal(g.download("V_SSA_AaAc"), golden["sor__V_SSA_AaAc"], "V_SSA_AaAc after the reference's SOR sweeps")
    # whole solve_SSA, then mass continuity and the critical time steps
    g2 = make_gpu(mesh, st, use_analytical_GL_flux=1, SSA_max_outer_loops=RC.SSA_OUTER, SSA_max_inner_loops=RC.SSA_INNER)
    g2.update_general_ice_model_data(0.0); g2.solve_SIA()
    s = g2.solve_SSA()
    assert s.n_outer == RC.SSA_OUTER
    for f in ("U_SSA", "V_SSA", "Up_SSA_Ac", "Qabs_GL_Ac"):
        a, b = g2.download(f), golden["ssa__" + f]
        assert np.linalg.norm(a - b) <= 1e-10 * np.linalg.norm(b), f
    g2.calculate_ice_thickness_change(0.5)
    for f in ("Hi", "dHi_dt"):
        a, b = g2.download(f), golden["thk__" + f]
        assert np.abs(a - b).max() <= 1e-8 * np.abs(b).max(), f
    # thermodynamics: the reference's update_ice_temperature (translated) vs ufm_update_ice_temperature
    from ufemism_b200 import scenarios as S
    from ufemism_b200.capi import IceModelGPU
    for bm in RC.THERMO_BENCHMARKS:
        stt = S.state_thermo_dome(mesh, benchmark=bm)
        gt = IceModelGPU(mesh, benchmark=bm, thermo=True, SSA_max_outer_loops=RC.THERMO_SSA_OUTER, SSA_max_inner_loops=RC.THERMO_SSA_INNER)
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB", "T2m", "GHF", "Ti"):
            gt.upload(k, stt[k])
        gt.update_general_ice_model_data(0.0)
        if bm == "none":
            gt.solve_SSA()
        assert gt.update_ice_temperature().n_unstable == 0
        for f in ("Ti", "W_3D", "U_3D", "frictional_heating"):
            b = golden[f"thermo_{bm}__{f}"]
            np.testing.assert_allclose(gt.download(f), b, rtol=1e-12, atol=1e-12 * max(np.abs(b).max(), 1e-300), err_msg=f"{bm} {f}")


@pytest.mark.skipif(not RS.available(), reason="/root/reference is not mounted here")
def test_live_partition_list_as_coded():
    """partition_list (src/mesh_help_functions_module.f90:1475-1496), the rank ranges behind every loop of the reference: the
    oracle's restatement against the translated source for 2700 (ntot, i, n) triples, including lists longer than 2**24, where the
    reference's single-precision REAL() shifts the ranges (and can drop the last element -- restated as coded, see ufm_oracle.c)."""
    import ctypes

    from oracle import f90py as F
    from oracle.oracle import lib

    P = F.Program({"dp": 8})
    P.add(open(os.path.join(RS.REF_SRC, "mesh_help_functions_module.f90")).read(), ["partition_list"])
    L = lib()
    rng = np.random.default_rng(1)
    cases = [(10, 4), (7, 8), (5, 2), (1000003, 16), (3999413, 8), (16001557, 16), (16777217, 8), (20000001, 16), (33554433, 3), (9, 4), (8, 4), (1, 1), (2, 1)]
    cases += [(int(rng.integers(1, 40_000_000)), int(rng.integers(1, 17))) for _ in range(300)]
    lost = 0
    for ntot, n in cases:
        covered = 0
        for i in range(n):
            a, b = ctypes.c_int(), ctypes.c_int()
            L.ora_partition_list(ntot, i, n, ctypes.byref(a), ctypes.byref(b))
            assert (a.value, b.value) == tuple(int(x) for x in P.partition_list(ntot, i, n, 0, 0)), (ntot, i, n)
            covered += max(0, b.value - a.value + 1)
        if ntot <= 1 << 24:
            assert covered == ntot, (ntot, n)            # below 2**24 the ranges tile the list
        lost += covered != ntot
    assert lost > 0                                       # above it they need not (the reference's latent defect, as coded)


@pytest.mark.skipif(not RS.available(), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("which", ["fan_degree_10_14_16", "seed_3_lattice_order", "seed_5"])
def test_live_mesh_operators_on_other_meshes(which):
    """The mesh substrate against the translated reference on meshes with other shapes than the golden one: vertices of degree
    10, 14 and 16 (= nC_mem), lattice (non-shuffled) vertex order, another seed.  Areas, connection widths, every neighbour
    function, the Ac numbering / orientation, the AaAc connectivity and the five-colouring must be the reference's, bit for bit."""
    from oracle.oracle import OraConfig, lib
    from tests.conftest import fan_mesh
    from ufemism_b200 import mesh as M
    import ctypes

    mesh = {"fan_degree_10_14_16": lambda: fan_mesh(nv=500), "seed_3_lattice_order": lambda: M.square_mesh_with_nv(750e3, 350, seed=3, order="lattice"),
            "seed_5": lambda: M.square_mesh_with_nv(400e3, 300, seed=5)}[which]()
    if which.startswith("fan"):
        assert sorted(set(mesh.nC.tolist()))[-1] == mesh.nC_mem
    c = OraConfig()
    lib().ora_config_defaults(ctypes.byref(c))
    np.seterr(all="ignore")
    ref = reference_mesh_operators(RS.program(c), mesh)
    got = {f"mesh__{f}": np.asarray(getattr(mesh, f)) for f in RC.MESH_FIELDS}
    got["mesh__Cw"] = np.where(np.arange(mesh.nC_mem)[None, :] < mesh.nC[:, None], mesh.Cw, 0.0)
    _compare(got, ref)


@pytest.mark.skipif(not RS.available(), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("scenario_name", ["halfar", "mismip", "eismint_ice_free", "thermo_dome_realistic"])
def test_live_per_step_routines_on_other_states(scenario_name):
    """update_general_ice_model_data (masks!), solve_SIA and calculate_ice_thickness_change of the translated reference vs the oracle on
    states that exercise other branches than the golden ice-stream case: a land-based dome (no ocean), the MISMIP sloping bed (coast,
    grounding line, thin shelf), an ice-free start (every mask empty, zero diffusivity) and a dome with the Arrhenius flow factor."""
    from oracle import f90py as F
    from oracle.oracle import Oracle
    from ufemism_b200 import scenarios as S

    np.seterr(all="ignore")
    mesh = RC.golden_mesh()
    bm = {"halfar": "Halfar", "mismip": "MISMIP_mod", "eismint_ice_free": "EISMINT_1", "thermo_dome_realistic": "none"}[scenario_name]
    st = {"halfar": lambda: S.state_halfar(mesh), "mismip": lambda: S.state_mismip(mesh), "eismint_ice_free": lambda: S.state_eismint1(mesh),
          "thermo_dome_realistic": lambda: S.state_thermo_dome(mesh, benchmark="none")}[scenario_name]()
    o = Oracle(mesh, benchmark=bm, nthreads=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB") + (("Ti",) if bm == "none" else ()):
        o[k][:] = st[k]
    P = RS.program(o.cfg)
    P.C.do_benchmark_experiment, P.C.choice_benchmark_experiment = (bm != "none"), bm
    mref, ice = RS.mesh_ns(mesh), RS.ice_ns(o)
    fields = RC.GENERAL_FIELDS + RC.SIA_FIELDS + RC.THK_FIELDS + (["A_flow", "A_flow_Ac", "Ti_Ac", "Ti_pmp", "Cpi", "Ki"] if bm == "none" else [])
    for step, dt in enumerate((0.5, 2.0)):
        P.update_general_ice_model_data(mref, ice, np.float64(0.0))
        P.solve_sia(mref, ice)
        P.calculate_ice_thickness_change(mref, ice, F.NS(smb_year=np.array(st["SMB_year"])), F.NS(bmb=np.array(st["BMB"])), np.float64(dt), np.zeros(mesh.nV, np.int32))
        o.update_general_ice_model_data(0.0); o.solve_SIA(); o.calculate_ice_thickness_change(dt)
        _compare({f: o[f] for f in fields}, {f: _arr(getattr(ice, f)) for f in fields})
    if scenario_name == "eismint_ice_free":
        assert not st["Hi"].any() and o["mask_ice"].any() and o["Hi"].max() > 0.0     # ice appears only through the mass balance
    elif scenario_name == "mismip":
        assert o["mask_gl"].any() and o["mask_shelf"].any() and o["mask_coast"].any()
    else:
        assert o["mask_sheet"].any() and np.abs(o["dHi_dt"]).max() > 0.0


@pytest.mark.skipif(not RS.available(), reason="/root/reference is not mounted here")
def test_live_thickness_update_of_the_SSA_icestream_experiment_keeps_the_domain_edge():
    """calculate_ice_thickness_change sets the thickness of domain-edge vertices to zero for every experiment but 'SSA_icestream'
    (src/ice_dynamics_module.f90:189-206: its branch is empty).  The translated reference and the oracle, switched to that experiment for
    this one routine (geometry and SIA, which do not know the name, are evaluated as MISMIP_mod), agree bit for bit and leave ice on the edge."""
    from oracle import f90py as F
    from oracle.oracle import BENCHMARKS, Oracle
    from ufemism_b200 import scenarios as S

    np.seterr(all="ignore")
    mesh = RC.golden_mesh()
    st = S.state_ssa_icestream(mesh, scale=750e3 / 1800e3, Hb=-250.0, H_shelf=150.0)
    o = Oracle(mesh, benchmark="MISMIP_mod", nthreads=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o[k][:] = st[k]
    o.update_general_ice_model_data(0.0); o.solve_SIA()
    P = RS.program(o.cfg)
    P.C.do_benchmark_experiment, P.C.choice_benchmark_experiment = True, "SSA_icestream"
    mref, ice = RS.mesh_ns(mesh), RS.ice_ns(o)
    P.calculate_ice_thickness_change(mref, ice, F.NS(smb_year=np.array(st["SMB_year"])), F.NS(bmb=np.array(st["BMB"])), np.float64(0.5), np.zeros(mesh.nV, np.int32))
    o.cfg.benchmark = BENCHMARKS["SSA_icestream"]
    o.calculate_ice_thickness_change(0.5)
    _compare({f: o[f] for f in RC.THK_FIELDS}, {f: _arr(getattr(ice, f)) for f in RC.THK_FIELDS})
    edge = mesh.edge_index > 0
    assert (o["Hi"][edge] > 0).any(), "the edge keeps its ice in this experiment"
    # ... and with any other experiment name the same update clears it
    o2 = Oracle(mesh, benchmark="MISMIP_mod", nthreads=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o2[k][:] = st[k]
    o2.update_general_ice_model_data(0.0); o2.solve_SIA(); o2.calculate_ice_thickness_change(0.5)
    assert not o2["Hi"][edge].any()


@pytest.mark.skipif(not RS.available(), reason="/root/reference is not mounted here")
def test_live_region_loop_scheduling():
    """Three steps of the region loop with the reference's own determine_timesteps_and_actions (critical time steps, the eight
    timers, the do_* flags, src/UFEMISM_main_model.f90:708-843) and its own per-step routines, called in run_model's order (:78-214;
    the CPU components between them are the benchmark no-ops) vs ora_run_model: same dt, time, timers and flags after every step,
    and the same thickness, bit for bit."""
    from oracle import f90py as F
    from oracle.oracle import T_BMB, T_CLIMATE, T_ELRA, T_OUTPUT, T_SIA, T_SMB, T_SSA, T_THERMO

    np.seterr(all="ignore")
    mesh = RC.golden_mesh()
    st = RC.start_state(mesh)
    # dt_max and the fixed timers are set long, so that the critical time steps of the dynamics decide the step (the golden mesh is coarse)
    cfg = dict(use_analytical_GL_flux=1, SSA_max_outer_loops=2, SSA_max_inner_loops=5, dt_max=1000.0, dt_thermo=700.0)   # C%dt_thermo sets the thermodynamics timer
    o = make_oracle(mesh, st, nthreads=1, **cfg)
    P = RS.program(o.cfg)
    P.C.choice_benchmark_experiment = st["benchmark"]
    C = P.C
    C.dt_max, C.dt_thermo, C.dt_climate, C.dt_smb, C.dt_bmb, C.dt_bedrock_elra, C.dt_output = (np.float64(v) for v in (1000.0, 700.0, 650.0, 600.0, 550.0, 800.0, 5000.0))
    mref, ice = RS.mesh_ns(mesh), RS.ice_ns(o)
    smb, bmb = F.NS(smb_year=np.array(st["SMB_year"]), smb=np.zeros((mesh.nV, 12), order="F")), F.NS(bmb=np.array(st["BMB"]))
    # initialise_model, src/UFEMISM_main_model.f90:352-390
    reg = F.NS(mesh=mref, ice=ice, time=np.float64(0.0), dt=np.float64(0.0), dt_prev=np.float64(1000.0), dt_sia=np.float64(0.0), dt_ssa=np.float64(0.0),
               do_solve_sia=True, do_solve_ssa=True, do_thermodynamics=False, do_climate=True, do_smb=True, do_bmb=True, do_elra=True, do_write_output=True)
    for t in ("sia", "ssa", "thermo", "climate", "smb", "bmb", "elra", "output"):
        setattr(reg, "t0_" + t, np.float64(0.0))
    r = o.region(0.0)
    for k, v in ((T_THERMO, 700.0), (T_CLIMATE, 650.0), (T_SMB, 600.0), (T_BMB, 550.0), (T_ELRA, 800.0), (T_OUTPUT, 5000.0)):
        r.dtc[k] = v
    r.t1[T_THERMO] = 700.0
    t_end = np.float64(1e12)
    dts = []
    for step in range(3):
        reg.t0_elra = reg.time                                                         # run_ELRA_model, benchmark branch
        P.calculate_ice_thickness_change(mref, ice, smb, bmb, reg.dt, np.zeros(mesh.nV, np.int32))
        P.update_general_ice_model_data(mref, ice, reg.time)
        if reg.do_solve_sia:
            P.solve_sia(mref, ice); reg.t0_sia = reg.time
        if reg.do_solve_ssa:
            P.solve_ssa(mref, ice); reg.t0_ssa = reg.time
        if reg.do_climate:
            reg.t0_climate = reg.time
        if reg.do_smb:
            P.run_smb_model(mref, ice, None, reg.time, smb, None); reg.t0_smb = reg.time
        if reg.do_bmb:
            reg.t0_bmb = reg.time
        if reg.do_thermodynamics:
            P.update_ice_temperature(mref, ice, None, smb); reg.t0_thermo = reg.time    # MISMIP_mod: returns at once (thermodynamics_module.f90:58)
        if reg.do_write_output:
            reg.t0_output = reg.time
        P.determine_timesteps_and_actions(reg, t_end)
        assert o.run_model(r, float(t_end), max_steps=1) == 0
        got = [reg.dt, reg.time, reg.dt_prev, reg.dt_sia, reg.dt_ssa] + [getattr(reg, "t1_" + t) for t in ("sia", "ssa", "thermo", "climate", "smb", "bmb", "elra", "output")]
        want = [r.dt, r.time, r.dt_prev, r.dtc[T_SIA], r.dtc[T_SSA]] + [r.t1[k] for k in (T_SIA, T_SSA, T_THERMO, T_CLIMATE, T_SMB, T_BMB, T_ELRA, T_OUTPUT)]
        assert [float(x) for x in got] == [float(x) for x in want], (step, got, want)
        flags = [reg.do_solve_sia, reg.do_solve_ssa, reg.do_thermodynamics, reg.do_climate, reg.do_smb, reg.do_bmb, reg.do_elra, reg.do_write_output]
        assert [bool(x) for x in flags] == [bool(r.do_[k]) for k in (T_SIA, T_SSA, T_THERMO, T_CLIMATE, T_SMB, T_BMB, T_ELRA, T_OUTPUT)], step
        for f in ("Hi", "U_SSA", "D_SIA_Ac", "SMB_year"):
            assert_bits_equal(o[f] if f != "SMB_year" else o["SMB_year"], _arr(getattr(ice, f)) if f != "SMB_year" else smb.smb_year.a, f"step {step} {f}")
        dts.append(float(r.dt))
    assert r.n_steps == 3 and r.n_ssa >= 2 and min(dts) < 500.0 and len(set(dts)) == 3, dts     # every step is set by a critical time step of the dynamics
