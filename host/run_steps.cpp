// run_steps.cpp -- test driver for host/ufemism_host.hpp: the run_model call sequence (src/UFEMISM_main_model.f90:90-135)
// issued from compiled code through the drop-in layer, on a mesh + state dumped by tests/test_gpu_parity.py.
//   run_steps <in.bin> <out.bin> <n_steps> <use_analytical_GL_flux>
// in.bin : int32 nV, nAc, nC_mem; then the arrays in the order read below (raw, column-major)
// out.bin: Hi, U_SSA, V_SSA, U_SIA (double, nV each), mask (int32, nV), then n_outer, n_inner_total (int32) and time (double)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ufemism_host.hpp"

template <class T> static std::vector<T> rd(FILE *f, size_t n) { std::vector<T> v(n); if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }

int main(int argc, char **argv)
{
  if (argc < 5) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 2;
  const size_t N = hdr[0], E = hdr[1], W = hdr[2], M = N + E, W1 = W + 1;
  auto V = rd<double>(f, N * 2); auto A = rd<double>(f, N); auto nC = rd<int>(f, N); auto C = rd<int>(f, N * W); auto Cw = rd<double>(f, N * W);
  auto ei = rd<int>(f, N); auto Nx = rd<double>(f, N * W1); auto Ny = rd<double>(f, N * W1);
  auto Aci = rd<int>(f, E * 4); auto iAci = rd<int>(f, N * W); auto eiAc = rd<int>(f, E);
  auto NxAc = rd<double>(f, E * 4); auto NyAc = rd<double>(f, E * 4); auto NoAc = rd<double>(f, E * 4); auto NpAc = rd<double>(f, E);
  auto nCA = rd<int>(f, M); auto CA = rd<int>(f, M * W);
  auto NxA = rd<double>(f, M * W1); auto NyA = rd<double>(f, M * W1); auto NxxA = rd<double>(f, M * W1); auto NxyA = rd<double>(f, M * W1); auto NyyA = rd<double>(f, M * W1);
  auto cvi = rd<int>(f, M * 5); auto cnV = rd<int>(f, 5);
  auto Hi = rd<double>(f, N); auto Hb = rd<double>(f, N); auto SL = rd<double>(f, N); auto SMB = rd<double>(f, N); auto BMB = rd<double>(f, N);
  fclose(f);

  ufemism::type_mesh mesh;
  mesh.nV = (int)N; mesh.nAc = (int)E; mesh.nC_mem = (int)W;
  mesh.V = V.data(); mesh.A = A.data(); mesh.nC = nC.data(); mesh.C = C.data(); mesh.Cw = Cw.data(); mesh.edge_index = ei.data(); mesh.Nx = Nx.data(); mesh.Ny = Ny.data();
  mesh.Aci = Aci.data(); mesh.iAci = iAci.data(); mesh.edge_index_Ac = eiAc.data(); mesh.Nx_Ac = NxAc.data(); mesh.Ny_Ac = NyAc.data(); mesh.No_Ac = NoAc.data(); mesh.Np_Ac = NpAc.data();
  mesh.nCAaAc = nCA.data(); mesh.CAaAc = CA.data(); mesh.Nx_AaAc = NxA.data(); mesh.Ny_AaAc = NyA.data(); mesh.Nxx_AaAc = NxxA.data(); mesh.Nxy_AaAc = NxyA.data(); mesh.Nyy_AaAc = NyyA.data();
  mesh.colour_vi = cvi.data(); mesh.colour_nV = cnV.data();

  std::vector<double> z(N, 0.0), Hs(N), Hp(N), dH(N), dHs(N), U_SIA(N), V_SIA(N), D_SIA(N), U_SSA(N), V_SSA(N);
  std::vector<int> mask(N), noice(N, 0);
  ufemism::type_ice_model ice;
  ice.Hi = Hi.data(); ice.Hb = Hb.data(); ice.SL = SL.data(); ice.dHb_dt = z.data();
  ice.Hs = Hs.data(); ice.Hi_prev = Hp.data(); ice.dHi_dt = dH.data(); ice.dHs_dt = dHs.data();
  ice.U_SIA = U_SIA.data(); ice.V_SIA = V_SIA.data(); ice.D_SIA = D_SIA.data(); ice.U_SSA = U_SSA.data(); ice.V_SSA = V_SSA.data(); ice.mask = mask.data();
  ufemism::type_SMB_model smb{SMB.data()};
  ufemism::type_BMB_model bmb{BMB.data()};

  ufm_params P;
  memset(&P, 0, sizeof(P));
  const double zeta[15] = {0.00, 0.10, 0.20, 0.30, 0.40, 0.50, 0.60, 0.70, 0.80, 0.90, 0.925, 0.95, 0.975, 0.99, 1.00};
  P.nZ = 15; for (int k = 0; k < 15; k++) P.zeta[k] = zeta[k];
  P.m_enh_sia = 1.0; P.m_enh_ssa = 1.0; P.use_analytical_GL_flux = atoi(argv[4]); P.SSA_RN_tol = 1e-5; P.SSA_max_outer_loops = 50;
  P.SSA_max_residual_UV = 2.5; P.SSA_SOR_omega = 1.2; P.SSA_max_inner_loops = 10000; P.dt_max = 10.0; P.benchmark = UFM_BM_MISMIP_MOD; P.exact_xy = 1;

  ufemism::B200IceDynamics dyn;
  dyn.initialise(0, P);
  dyn.upload_mesh(mesh);
  // run_model with both solvers due every step and dt from the critical time steps (a reduced determine_timesteps_and_actions)
  const int n_steps = atoi(argv[3]);
  double time = 0.0, dt = 0.0;
  int n_outer = 0, n_inner = 0;
  for (int s = 0; s < n_steps; s++) {
    dyn.calculate_ice_thickness_change(mesh, ice, smb, bmb, dt, noice.data());
    dyn.update_general_ice_model_data(mesh, ice, time);
    dyn.solve_SIA(mesh, ice);
    dyn.solve_SSA(mesh, ice);
    n_outer += dyn.last_ssa_stats.n_outer; n_inner += dyn.last_ssa_stats.n_inner_total;
    double a, b, c;
    dyn.critical_timesteps(a, b, c);
    dt = std::fmin(std::fmin(std::fmin(a, b), c), P.dt_max);
    time += dt;
  }
  FILE *o = fopen(argv[2], "wb");
  fwrite(Hi.data(), 8, N, o); fwrite(U_SSA.data(), 8, N, o); fwrite(V_SSA.data(), 8, N, o); fwrite(U_SIA.data(), 8, N, o); fwrite(mask.data(), 4, N, o);
  int cnt[2] = {n_outer, n_inner};
  fwrite(cnt, 4, 2, o); fwrite(&time, 8, 1, o);
  fclose(o);
  return 0;
}
