"""The cases run through BOTH the translated reference source (tests/ref_source.py, where /root/reference is mounted) and the
oracle / mesh substrate / CUDA path.  One definition, so that the golden vectors (tests/golden/reference_source_600.npz), the
live comparison (tests/test_reference_source.py) and the GPU comparison use identical inputs."""
import numpy as np

SOR_ITERS = 4          # C%SSA_max_inner_loops of the SOR case (the sweep starts far from convergence: all of them run)
SSA_OUTER = 4          # C%SSA_max_outer_loops of the solve_SSA case
SSA_INNER = 12         # C%SSA_max_inner_loops of the solve_SSA case (the reference warns and carries on; keeps the Python run short)

GENERAL_FIELDS = (["mask_land", "mask_ocean", "mask_lake", "mask_ice", "mask_sheet", "mask_shelf", "mask_coast", "mask_margin", "mask_gl", "mask_cf", "mask"] +
                  [f + "_Ac" for f in ("mask_land", "mask_ocean", "mask_lake", "mask_ice", "mask_sheet", "mask_shelf", "mask_coast", "mask_margin", "mask_gl", "mask_cf", "mask")] +
                  ["Hs", "dHs_dt", "dHi_dx", "dHi_dy", "dHs_dx", "dHs_dy", "dHs_dx_shelf", "dHs_dy_shelf", "Hi_Ac", "Hb_Ac", "Hs_Ac", "SL_Ac"] +
                  [f"d{f}_d{c}_Ac" for f in ("Hi", "Hb", "Hs", "SL") for c in "xypo"] +
                  ["dHs_dx_shelf_Ac", "dHs_dy_shelf_Ac", "A_flow_mean", "A_flow_mean_Ac"])
THK_FIELDS = ["Hi", "dHi_dt", "Hi_prev", "dVi_in"]
SIA_FIELDS = ["D_SIA_Ac", "Ux_SIA_Ac", "Uy_SIA_Ac", "Up_SIA_Ac", "Uo_SIA_Ac", "U_SIA", "V_SIA", "D_SIA", "D_SIA_3D_Ac"]
PIECES_FIELDS = ["tau_c_AaAc", "phi_fric_AaAc", "dU_SSA_dx_AaAc", "dU_SSA_dy_AaAc", "dV_SSA_dx_AaAc", "dV_SSA_dy_AaAc", "eta_AaAc", "N_AaAc", "S_AaAc"]
SOR_FIELDS = ["RHSx_AaAc", "RHSy_AaAc", "eu_i_AaAc", "ev_i_AaAc", "U_SSA_AaAc", "V_SSA_AaAc", "resU_AaAc", "resV_AaAc"]
SSA_FIELDS = ["U_SSA", "V_SSA", "Ux_SSA_Ac", "Uy_SSA_Ac", "Up_SSA_Ac", "Uo_SSA_Ac", "eta_AaAc", "N_AaAc", "S_AaAc", "tau_c_AaAc", "Qabs_GL_Ac", "Qp_GL_Ac", "U_SSA_AaAc", "V_SSA_AaAc"]
THERMO_FIELDS = ["Ti", "W_3D", "U_3D", "V_3D", "frictional_heating", "Ki", "Cpi", "Ti_pmp", "dzeta_dx", "dzeta_dy", "dzeta_dz", "A_flow_mean"]
THERMO_BENCHMARKS = ("EISMINT_1", "none")   # EISMINT ice properties / temperature-dependent properties with sliding
THERMO_SSA_OUTER = 2
THERMO_SSA_INNER = 8
MESH_FIELDS = ["A", "Cw", "R", "Nx", "Ny", "Nxx", "Nxy", "Nyy", "NxTri", "NyTri", "Aci", "iAci", "VAc", "Nx_Ac", "Ny_Ac", "Np_Ac", "No_Ac", "edge_index_Ac", "nCAaAc", "CAaAc", "VAaAc",
               "Nx_AaAc", "Ny_AaAc", "Nxx_AaAc", "Nxy_AaAc", "Nyy_AaAc", "colour", "colour_vi", "colour_nV"]


def golden_mesh():
    import os

    from ufemism_b200 import mesh as M
    return M.Mesh.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_600.npz"))


def start_state(mesh):
    """The hybrid SIA/SSA state every case starts from: an ice stream / shelf geometry (grounded sheet, grounding line, floating shelf,
    ice-free ocean), with the velocities of two model steps of the oracle -- only used as INPUT, identical for both sides."""
    from ufemism_b200 import scenarios as S
    st = dict(S.state_ssa_icestream(mesh, scale=750e3 / 1800e3, Hb=-250.0, H_shelf=150.0))
    st["benchmark"] = "MISMIP_mod"   # a choice_benchmark_experiment the reference itself knows ('SSA_icestream' is accepted by some routines only)
    return st


def random_velocities(mesh):
    rng = np.random.default_rng(20211103)
    return rng.normal(0.0, 50.0, mesh.nVAaAc), rng.normal(0.0, 50.0, mesh.nVAaAc)


# closed-form benchmark mass balance (run_SMB_model's benchmark branches) and the analytic solutions, at these (benchmark, time) pairs
SMB_CASES = [("EISMINT_1", 0.0), ("EISMINT_2", 3000.0), ("EISMINT_2", -5.0), ("EISMINT_3", 7000.0), ("EISMINT_4", 0.0), ("EISMINT_5", 4000.0),
             ("EISMINT_6", 33000.0), ("Halfar", 10.0), ("Bueler", 500.0), ("MISMIP_mod", 1.0), ("mesh_generation_test", 0.0)]
H0, R0, LAMBDA = 3000.0, 500000.0, 5.0      # config_Bueler_*: halfar_solution_H0 / _R0, bueler_solution_lambda
ANALYTIC_TIMES = (0.0, 1000.0)


def remap_map(mesh, n_dst=300):
    """A synthetic type_remapping_conservative (src/data_types_module.f90:544-556): for every destination vertex 1..6 source
    vertices with weights; the weights are arbitrary -- only their APPLICATION is on the path."""
    rng = np.random.default_rng(99)
    cnt = rng.integers(1, 7, n_dst)
    vli2 = np.cumsum(cnt).astype(np.int32)
    vli1 = (vli2 - cnt + 1).astype(np.int32)
    n = int(vli2[-1])
    return dict(vli1=vli1, vli2=vli2, vi=rng.integers(1, mesh.nV + 1, n).astype(np.int32), w0=rng.random(n), w1x=rng.normal(0, 1e3, n), w1y=rng.normal(0, 1e3, n),
                d_src=rng.normal(1000.0, 300.0, mesh.nV))
