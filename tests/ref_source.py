"""Harness around oracle/f90py.py: translate the reference's hot-path routines from /root/reference/src and run them on the oracle's
arrays.  Only usable where the reference is mounted (this container); the golden vectors it produces travel instead."""
import os

import numpy as np

REF_SRC = "/root/reference/src"

# parameters_module.f90:9-21 (PARAMETERs; module-level names the routines USE)
CONSTANTS = dict(dp=8, pi=np.float64(3.141592653589793), sec_per_year=np.float64(31556943.36), grav=np.float64(9.81), n_flow=np.float64(3.0),
                 ice_density=np.float64(910.0), seawater_density=np.float64(1028.0), smt=np.float64(271.15), t0=np.float64(273.16),
                 cc=np.float64(8.7e-04), l_fusion=np.float64(3.335e5))

# (file, units).  Interfaces of all units are declared first, so the order of translation does not matter.
UNITS = [
    ("zeta_module.f90", ["vertical_integrate", "vertical_average", "initialize_zeta_discretization", "calculate_zeta_derivatives"]),
    ("mesh_help_functions_module.f90", ["is_boundary_segment", "is_in_triangle", "cross2", "find_triangle_area", "find_connection_widths", "determine_mesh_resolution",
                                        "find_Voronoi_cell_areas", "find_Voronoi_cell_vertices", "find_Voronoi_cell_vertices_free",
                                        "find_Voronoi_cell_vertices_corner", "find_Voronoi_cell_vertices_edge", "crop_circumcenter", "line_from_points",
                                        "line_line_intersection"]),
    ("general_ice_model_data_module.f90", ["is_floating", "determine_masks", "ice_physical_properties", "update_general_ice_model_data"]),
    ("mesh_derivatives_module.f90", ["get_neighbour_functions_vertex_gr", "get_neighbour_functions", "get_mesh_derivatives_vertex", "get_mesh_derivatives",
                                     "get_mesh_derivatives_vertex_3D", "get_mesh_derivatives_3D", "apply_Neumann_boundary", "apply_Neumann_boundary_3D",
                                     "get_upwind_derivative_vertex_3D"]),
    ("mesh_ArakawaC_module.f90", ["make_Ac_mesh", "find_Ac_edge_indices", "make_combined_AaAc_mesh", "get_mesh_derivatives_vertex_Ac", "get_mesh_derivatives_Ac",
                                  "get_mesh_derivatives_vertex_AaAc", "get_mesh_derivatives_AaAc", "get_mesh_curvatures_vertex_AaAc", "apply_Neumann_boundary_AaAc",
                                  "map_Aa_to_Ac", "map_Aa_to_Ac_3D", "map_Ac_to_Aa", "map_Ac_to_Aa_3D", "rotate_xy_to_po"]),
    ("ice_dynamics_module.f90", ["calculate_ice_thickness_change", "solve_SIA", "solve_SIA_3D", "SSA_effective_viscosity", "SSA_sliding_term", "basal_yield_stress",
                                 "calculate_GL_flux", "solve_SSA_linearised", "solve_SSA"]),
    ("UFEMISM_main_model.f90", ["determine_timesteps_and_actions"]),
    ("thermodynamics_module.f90", ["bottom_frictional_heating", "tridiagonal_solve", "replace_Ti_with_robin_solution", "update_ice_temperature"]),
    ("SMB_module.f90", ["run_SMB_model", "EISMINT_SMB", "Bueler_solution_MB"]),
    ("reference_fields_module.f90", ["Halfar_solution", "Bueler_solution"]),
    ("mesh_mapping_module.f90", ["remap_cons_1st_order_2D", "remap_cons_2nd_order_2D"]),
    ("mesh_five_colour_module.f90", None),   # None = every unit of the file
]


def available():
    return os.path.isdir(REF_SRC)


def program(config):
    """Translate every unit of UNITS; `config` = the oracle's OraConfig (C%... values)."""
    from oracle import f90py as F

    C = F.NS(nz=int(config.nZ), zeta=np.array([config.zeta[k] for k in range(config.nZ)]), m_enh_sia=np.float64(config.m_enh_sia),
             m_enh_ssa=np.float64(config.m_enh_ssa), use_analytical_gl_flux=bool(config.use_analytical_GL_flux),
             ssa_rn_tol=np.float64(config.SSA_RN_tol), ssa_max_outer_loops=int(config.SSA_max_outer_loops),
             ssa_max_residual_uv=np.float64(config.SSA_max_residual_UV), ssa_sor_omega=np.float64(config.SSA_SOR_omega),
             ssa_max_inner_loops=int(config.SSA_max_inner_loops), choice_sliding_law="Coulomb_regularised", do_benchmark_experiment=True,
             choice_benchmark_experiment="", nconmax=16, dt_thermo=np.float64(config.dt_thermo),
             c_sliding=np.float64(1.0e7), m_sliding=np.float64(1.0) / np.float64(3.0))   # configuration_module.f90:172-173
    consts = dict(CONSTANTS)
    consts.update(c=C, p_zeta=F.NS(), par=F.NS(master=True, i=0, n=1, mem=F.NS(n=0)), mpi_in_place=None, mpi_double_precision=None, mpi_max=None, mpi_comm_world=None, ierr=0, cerr=0)
    P = F.Program(consts)
    texts = {fn: open(os.path.join(REF_SRC, fn)).read() for fn, _ in UNITS}
    units = {fn: (u if u is not None else F.unit_names(texts[fn])) for fn, u in UNITS}
    for fn, _ in UNITS:
        P.declare(texts[fn], units[fn])
    for fn, _ in UNITS:
        P.add(texts[fn], units[fn])
    P.C = C
    return P


def mesh_ns(mesh):
    """type_mesh stand-in over copies of a mesh.Mesh's arrays, single-rank ranges."""
    from oracle import f90py as F

    m = F.NS()
    for k, v in mesh.__dict__.items():
        if k == "extra":
            continue
        setattr(m, k, np.array(v, order="F") if isinstance(v, np.ndarray) else v)
    M = mesh.nVAaAc
    m.nxtri, m.nytri, m.r, m.tric = np.array(mesh.NxTri, order="F"), np.array(mesh.NyTri, order="F"), np.array(mesh.R), np.array(mesh.TriC, order="F")
    m.t1, m.t2, m.ntriaaac = 1, mesh.nTri, 0
    m.itri, m.nitri, m.tri = np.array(mesh.iTri, order="F"), np.array(mesh.niTri), np.array(mesh.Tri, order="F")
    # src/restart_module.f90:81 / mesh_creation_module.f90: tol_dist = ((xmax - xmin) + (ymax - ymin)) * tol / 2 with tol = 1E-9_dp
    m.tol_dist = ((np.float64(mesh.xmax) - mesh.xmin) + (np.float64(mesh.ymax) - mesh.ymin)) * np.float64(1e-9) / np.float64(2.0)
    m.v1, m.v2, m.ac1, m.ac2, m.a1, m.a2 = 1, mesh.nV, 1, mesh.nAc, 1, M
    m.colour_v1 = np.ones(5, np.int32)
    m.colour_v2 = np.array(mesh.colour_nV, np.int32)
    return m


def ice_ns(oracle, extra=()):
    """type_ice_model stand-in over COPIES of the oracle's fields (the translated routines run independently of the oracle)."""
    from oracle import f90py as F

    ice = F.NS()
    for k, v in oracle.f.items():
        setattr(ice, k, np.array(v, order="F"))
    ice.dvi_out = np.zeros_like(oracle.f["dVi_in"])
    for k, shape in extra:
        setattr(ice, k, np.zeros(shape, order="F"))
    return ice
