// Mirror of test/decompose_domain_loop.cpp: four blocks A {L, L/4}, B {L/4, L/2}, C {L, L/4}, D {L/4, L/2} that close a
// square channel.  Every block is its own lbm_domain with its own wall rules (:171-230); the "Bind the domains" lines
// (:232-261) are eight lbm_link_face calls; the body force on rows L/4+5 .. L/4+55 of A (:66-69,151-158) is
// lbm_set_force_region.  usage: decompose_domain_loop [L = 512] [T = 50000]
#include "common.hpp"

namespace
{
struct Pair { int q, qs; };
const Pair TOP[3] = {{8, 6}, {1, 3}, {5, 7}}, BOTTOM[3] = {{7, 5}, {3, 1}, {6, 8}};
const Pair LEFT[3] = {{2, 4}, {5, 7}, {6, 8}}, RIGHT[3] = {{4, 2}, {7, 5}, {8, 6}};

void wall(lbm_domain* d, int xb, int xe, int yb, int ye, const Pair (&p)[3])
{
  for (const Pair& w : p)
  {
    lbm_bc_op op;
    lbm_bc_op_default(&op);
    op.kind = LBM_BC_LINEAR; op.lattice = 0;
    op.x_begin = xb; op.x_end = xe; op.y_begin = yb; op.y_end = ye;
    op.dst_q = w.q; op.src_q = w.qs; op.coef = 1.0;
    DRV_CHECK(lbm_bc_add(d, &op));
  }
}
}  // namespace

int main(int argc, char* argv[])
{
  const int L = argc > 1 ? std::atoi(argv[1]) : 512, T = argc > 2 ? std::atoi(argv[2]) : 50000;
  const int snapshot_period = 50, L2 = L / 2, L4 = L / 4, Ts = T / snapshot_period;
  std::cout << "T=" << T << std::endl;
  const double tau = std::sqrt(3.0 / 16.0) + 0.5, omega = 1.0 / tau, u_max = 0.1, nu = (2.0 * tau - 1.0) / 6.0;
  std::cout << "omega=" << omega << "\nnu=" << nu << "\nRe=" << L4 * u_max / nu << std::endl;
  const double F[2] = {3e-3, 0.0};

  const char* names = "ABCD";
  const int R[4] = {L, L4, L, L4}, C[4] = {L4, L2, L4, L2};
  lbm_domain* dom[4];
  for (int k = 0; k < 4; k++)
  {
    lbm_config cfg;
    lbm_config_default(&cfg);
    cfg.X = R[k]; cfg.Y = C[k]; cfg.x1 = R[k];
    cfg.omega = omega;
    cfg.equilibrium = LBM_EQ_COMPRESSIBLE;
    cfg.force = k == 0 ? LBM_FORCE_IBM : LBM_FORCE_NONE;  // block A reads the force-field slot
    DRV_CHECK(lbm_create(&cfg, &dom[k]));
    DRV_CHECK(lbm_bc_clear(dom[k]));
    wall(dom[k], 0, 1, 0, LBM_END, TOP);
    wall(dom[k], -1, LBM_END, 0, LBM_END, BOTTOM);
  }
  lbm_domain *A = dom[0], *B = dom[1], *Cb = dom[2], *D = dom[3];
  wall(A, L4, -L4, 0, 1, LEFT);
  wall(A, 1, -1, -1, LBM_END, RIGHT);
  wall(Cb, 1, -1, 0, 1, LEFT);
  wall(Cb, L4, -L4, -1, LBM_END, RIGHT);
  DRV_CHECK(lbm_link_face(A, 0, L - L4, L4, B, 0));   DRV_CHECK(lbm_link_face(B, 1, 0, L4, A, L - L4));   // A-B
  DRV_CHECK(lbm_link_face(B, 0, 0, L4, Cb, L - L4));  DRV_CHECK(lbm_link_face(Cb, 1, L - L4, L4, B, 0));  // B-C
  DRV_CHECK(lbm_link_face(Cb, 1, 0, L4, D, 0));       DRV_CHECK(lbm_link_face(D, 0, 0, L4, Cb, 0));       // C-D
  DRV_CHECK(lbm_link_face(D, 1, 0, L4, A, 0));        DRV_CHECK(lbm_link_face(A, 0, 0, L4, D, 0));        // D-A
  DRV_CHECK(lbm_set_force_region(A, L4 + 5, L4 + 55, 0, LBM_END, F[0], F[1], 3.0, 9.0));

  std::vector<drv::Series> ux, uy, rhos;
  std::vector<std::vector<double>> u(4), rho(4);
  for (int k = 0; k < 4; k++)
  {
    DRV_CHECK(lbm_bc_commit(dom[k]));
    const size_t N = (size_t)R[k] * C[k];
    u[k].assign(2 * N, 0.0);   // m_1 = 0, m_0 = 1 (:73-76); adve_f = equilibrium(m_1, m_0) (:108-111)
    rho[k].assign(N, 1.0);
    DRV_CHECK(lbm_init_equilibrium(dom[k], 0, LBM_EQ_COMPRESSIBLE, rho[k].data(), u[k].data()));
    ux.emplace_back(R[k], C[k], Ts); uy.emplace_back(R[k], C[k], Ts); rhos.emplace_back(R[k], C[k], Ts);
  }

  std::cout << "main loop starts" << std::endl;
  for (int t = 0; t < T; t++)
  {
    if (t % snapshot_period == 0)
    {
      // the snapshot shows m_0, m_1 of the iteration before (the initial values at t = 0), after `m_1[force_idx] += F` (:116)
      const int ts = t / snapshot_period;
      for (int k = 0; k < 4; k++)
      {
        if (k == 0)
          for (int r = std::min(L4 + 5, L); r < std::min(L4 + 55, L); r++)
            for (int c = 0; c < C[0]; c++)
            {
              u[0][2 * ((size_t)r * C[0] + c)] += F[0];
              u[0][2 * ((size_t)r * C[0] + c) + 1] += F[1];
            }
        ux[k].put(ts, u[k], 2, 0); uy[k].put(ts, u[k], 2, 1); rhos[k].put(ts, rho[k], 1, 0);
      }
    }
    // m_0, m_1 of this iteration are the moments of the state it starts from: fetch them only when the next snapshot needs them
    if ((t + 1) % snapshot_period == 0)
      for (int k = 0; k < 4; k++) DRV_CHECK(lbm_get_moments(dom[k], 0, rho[k].data(), u[k].data()));
    DRV_CHECK(lbm_step_group(dom, 4, 1));
  }
  for (int k = 0; k < 4; k++) DRV_CHECK(lbm_synchronize(dom[k]));

  std::cout << "saving results into files" << std::endl;
  for (int k = 0; k < 4; k++)
  {
    const std::string pre = std::string(1, names[k]) + "-domain-decomp-";
    ux[k].save(pre + "hpt-ux.pt"); uy[k].save(pre + "hpt-uy.pt"); rhos[k].save(pre + "hpt-rho.pt");
    lbm_destroy(dom[k]);
  }
  return 0;
}
