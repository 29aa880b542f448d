// ufemism_host.hpp -- compiled-language host side above the C ABI (include/ufemism_b200.h).
//
// The reference host is Fortran; this image has no Fortran compiler, so besides the ISO_C_BINDING shim delivered as
// source (ufemism_b200/fortran/ufemism_b200_shim.f90) the same drop-in layer is given here in C++ and is exercised by
// tests (host/run_steps.cpp, tests/test_gpu_parity.py::test_cpp_host_mirror).  It mirrors the reference's interface for
// this path: the derived types that cross the boundary (type_mesh, type_ice_model, type_SMB_model, type_BMB_model:
// src/data_types_module.f90:15-345) reduced to the members the path touches, as non-owning views of host arrays in the
// reference layout (column-major, 1-based indices in the arrays), and the four routines with the reference's names and
// argument lists (src/ice_dynamics_module.f90:31,240,408; src/general_ice_model_data_module.f90:23).  Each routine does
// what the Fortran wrapper does: upload what CPU components may have changed, compute on the GPU, download what CPU
// components read.  Errors follow the reference: message on stderr, then abort (MPI_ABORT in the Fortran host);
// the SSA non-convergence warning is printed and the run continues.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../include/ufemism_b200.h"

namespace ufemism {

struct type_mesh {                 // src/data_types_module.f90:216-345 (members used by the path)
  int nV = 0, nAc = 0, nC_mem = 16;
  const double *V = nullptr, *A = nullptr, *Cw = nullptr, *Nx = nullptr, *Ny = nullptr;
  const int *nC = nullptr, *C = nullptr, *edge_index = nullptr;
  const int *Aci = nullptr, *iAci = nullptr, *edge_index_Ac = nullptr;
  const double *Nx_Ac = nullptr, *Ny_Ac = nullptr, *No_Ac = nullptr, *Np_Ac = nullptr;
  const int *nCAaAc = nullptr, *CAaAc = nullptr;
  const double *Nx_AaAc = nullptr, *Ny_AaAc = nullptr, *Nxx_AaAc = nullptr, *Nxy_AaAc = nullptr, *Nyy_AaAc = nullptr;
  const int *colour_vi = nullptr, *colour_nV = nullptr;
  // read only by update_ice_temperature (upwind advection); leave Tri null to keep thermodynamics off the device
  int nTri = 0;
  const int *Tri = nullptr, *niTri = nullptr, *iTri = nullptr;
  const double *R = nullptr, *NxTri = nullptr, *NyTri = nullptr;
};

struct type_ice_model {            // src/data_types_module.f90:15-214 (members the wrappers move)
  double *Hi = nullptr, *Hb = nullptr, *SL = nullptr, *dHb_dt = nullptr;                       // inputs owned by CPU components
  double *Hs = nullptr, *Hi_prev = nullptr, *dHi_dt = nullptr, *dHs_dt = nullptr;              // outputs
  double *dHi_dx = nullptr, *dHi_dy = nullptr, *dHs_dx = nullptr, *dHs_dy = nullptr;
  double *U_SIA = nullptr, *V_SIA = nullptr, *D_SIA = nullptr, *U_SSA = nullptr, *V_SSA = nullptr;
  int *mask = nullptr, *mask_land = nullptr, *mask_ocean = nullptr, *mask_ice = nullptr, *mask_sheet = nullptr, *mask_shelf = nullptr,
      *mask_coast = nullptr, *mask_margin = nullptr, *mask_gl = nullptr, *mask_cf = nullptr;
  double *Ti = nullptr, *U_3D = nullptr, *V_3D = nullptr, *W_3D = nullptr;   // (nV,nZ) thermodynamics
  const double *GHF = nullptr;
};
struct type_climate_model { const double *T2m = nullptr; };   // climate%applied%T2m (nV,12)
struct type_SMB_model { const double *SMB_year = nullptr; };
struct type_BMB_model { const double *BMB = nullptr; };

class B200IceDynamics {
 public:
  bool throw_on_error = false;   // tests: throw instead of abort
  ufm_ssa_stats last_ssa_stats{};

  // after initialize_main_constants (src/UFEMISM_program.f90:103)
  void initialise(int device, const ufm_params &params) { check(ufm_create(device, &params, &h_), "ufm_create"); }
  ~B200IceDynamics() { if (h_) ufm_destroy(h_); }

  // end of create_final_mesh_from_merged_submesh / mesh swap (src/mesh_creation_module.f90:1737, src/UFEMISM_main_model.f90:294)
  void upload_mesh(const type_mesh &m)
  {
    ufm_mesh_desc d{};
    d.nV = m.nV; d.nAc = m.nAc; d.nC_mem = m.nC_mem; d.ldV = m.nV; d.ldAc = m.nAc; d.ldAaAc = m.nV + m.nAc;
    d.V = m.V; d.A = m.A; d.nC = m.nC; d.C = m.C; d.Cw = m.Cw; d.edge_index = m.edge_index; d.Nx = m.Nx; d.Ny = m.Ny;
    d.Aci = m.Aci; d.iAci = m.iAci; d.edge_index_Ac = m.edge_index_Ac; d.Nx_Ac = m.Nx_Ac; d.Ny_Ac = m.Ny_Ac; d.No_Ac = m.No_Ac; d.Np_Ac = m.Np_Ac;
    d.nCAaAc = m.nCAaAc; d.CAaAc = m.CAaAc; d.Nx_AaAc = m.Nx_AaAc; d.Ny_AaAc = m.Ny_AaAc; d.Nxx_AaAc = m.Nxx_AaAc; d.Nxy_AaAc = m.Nxy_AaAc;
    d.Nyy_AaAc = m.Nyy_AaAc; d.colour_vi = m.colour_vi; d.colour_nV = m.colour_nV;
    d.nTri = m.nTri; d.ldTri = m.nTri; d.Tri = m.Tri; d.niTri = m.niTri; d.iTri = m.iTri; d.R = m.R; d.NxTri = m.NxTri; d.NyTri = m.NyTri;
    check(ufm_mesh_upload(h_, &d), "ufm_mesh_upload");
  }

  // calculate_ice_thickness_change( mesh, ice, SMB, BMB, dt, mask_noice)   src/ice_dynamics_module.f90:31
  void calculate_ice_thickness_change(const type_mesh &, type_ice_model &ice, const type_SMB_model &SMB, const type_BMB_model &BMB, double dt, const int *mask_noice)
  {
    up(UFM_F_SMB_YEAR, SMB.SMB_year); up(UFM_F_BMB, BMB.BMB); up(UFM_F_MASK_NOICE, mask_noice); up(UFM_F_HI, ice.Hi);
    check(ufm_thickness_update(h_, dt), "calculate_ice_thickness_change");
    down(UFM_F_HI, ice.Hi); down(UFM_F_HI_PREV, ice.Hi_prev); down(UFM_F_DHI_DT, ice.dHi_dt);
  }
  // update_general_ice_model_data( mesh, ice, time)   src/general_ice_model_data_module.f90:23
  void update_general_ice_model_data(const type_mesh &, type_ice_model &ice, double time)
  {
    up(UFM_F_HB, ice.Hb); up(UFM_F_SL, ice.SL); up(UFM_F_DHB_DT, ice.dHb_dt);
    check(ufm_update_general(h_, time), "update_general_ice_model_data");
    down(UFM_F_HS, ice.Hs); down(UFM_F_DHS_DT, ice.dHs_dt); down(UFM_F_DHI_DX, ice.dHi_dx); down(UFM_F_DHI_DY, ice.dHi_dy);
    down(UFM_F_DHS_DX, ice.dHs_dx); down(UFM_F_DHS_DY, ice.dHs_dy);
    down(UFM_F_MASK, ice.mask); down(UFM_F_MASK_LAND, ice.mask_land); down(UFM_F_MASK_OCEAN, ice.mask_ocean); down(UFM_F_MASK_ICE, ice.mask_ice);
    down(UFM_F_MASK_SHEET, ice.mask_sheet); down(UFM_F_MASK_SHELF, ice.mask_shelf); down(UFM_F_MASK_COAST, ice.mask_coast);
    down(UFM_F_MASK_MARGIN, ice.mask_margin); down(UFM_F_MASK_GL, ice.mask_gl); down(UFM_F_MASK_CF, ice.mask_cf);
  }
  // solve_SIA( mesh, ice)   src/ice_dynamics_module.f90:240
  void solve_SIA(const type_mesh &, type_ice_model &ice)
  {
    check(ufm_solve_SIA(h_), "solve_SIA");
    down(UFM_F_U_SIA, ice.U_SIA); down(UFM_F_V_SIA, ice.V_SIA); down(UFM_F_D_SIA, ice.D_SIA);
  }
  // solve_SSA( mesh, ice)   src/ice_dynamics_module.f90:408
  void solve_SSA(const type_mesh &, type_ice_model &ice)
  {
    check(ufm_solve_SSA(h_, &last_ssa_stats), "solve_SSA");
    down(UFM_F_U_SSA, ice.U_SSA); down(UFM_F_V_SSA, ice.V_SSA);
  }
  // update_ice_temperature( mesh, ice, climate, SMB)   src/thermodynamics_module.f90:23
  ufm_thermo_stats last_thermo_stats{};
  void update_ice_temperature(const type_mesh &, type_ice_model &ice, const type_climate_model &climate, const type_SMB_model &SMB)
  {
    up(UFM_F_T2M, climate.T2m); up(UFM_F_SMB_YEAR, SMB.SMB_year); up(UFM_F_GHF, ice.GHF);
    check(ufm_update_ice_temperature(h_, &last_thermo_stats), "update_ice_temperature");
    down(UFM_F_TI, ice.Ti); down(UFM_F_U_3D, ice.U_3D); down(UFM_F_V_3D, ice.V_3D); down(UFM_F_W_3D, ice.W_3D);
  }
  // the three loops + MPI_ALLREDUCE MIN of determine_timesteps_and_actions (src/UFEMISM_main_model.f90:747-778)
  void critical_timesteps(double &dt_D_2D_min, double &dt_V_2D_SSA_min, double &dt_V_3D_SIA_min)
  {
    double o[3];
    check(ufm_cfl(h_, o), "determine_timesteps_and_actions");
    dt_D_2D_min = o[0]; dt_V_2D_SSA_min = o[1]; dt_V_3D_SIA_min = o[2];
  }
  ufm_handle *handle() { return h_; }

 private:
  ufm_handle *h_ = nullptr;
  void up(int f, const void *p) { if (p) check(ufm_state_upload(h_, f, p), "ufm_state_upload"); }
  void down(int f, void *p) { if (p) check(ufm_state_download(h_, f, p), "ufm_state_download"); }
  void check(int rc, const char *where)
  {
    if (rc == 0) return;
    if (rc > 0) { std::fprintf(stderr, "%s\n", ufm_last_error()); return; }   // WRITE(0,*) ' WARNING - ...' and carry on
    std::fprintf(stderr, "  ERROR in %s: %s\n", where, ufm_last_error());
    if (throw_on_error) throw std::runtime_error(std::string(where) + ": " + ufm_last_error());
    std::abort();   // CALL MPI_ABORT( MPI_COMM_WORLD, cerr, ierr)
  }
};

}  // namespace ufemism
