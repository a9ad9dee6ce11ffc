"""Builders that configure an lbm_b200.Domain like each reference driver does (test helpers)."""
import os

import numpy as np

import lbm_b200 as L

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def channel_constants(H, W, u_max):
    """test/horizontal_poiseuille_test.cpp:50-66"""
    tau = np.sqrt(3.0 / 16.0) + 0.5
    omega = 1.0 / tau
    nu = (2.0 * tau - 1.0) / 6.0
    p_grad = 8.0 * nu * u_max / (W * W)
    rho_out = 1.0
    rho_in = 3.0 * (H - 1) * p_grad + rho_out
    return omega, rho_in, rho_out


def poiseuille(H=21, W=21, u_max=1.030985714e-1, **slab):
    omega, rho_in, rho_out = channel_constants(H, W, u_max)
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=H, Y=W, omega=omega, equilibrium=L.EQ_INCOMPRESSIBLE, **slab))
    d.preset_poiseuille(rho_in, rho_out)
    return d, (omega, rho_in, rho_out)


def specular(H=51, W=51, u_max=0.1):
    omega, rho_in, rho_out = channel_constants(H, W, u_max)
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=H, Y=W, omega=omega, equilibrium=L.EQ_COMPRESSIBLE))
    d.preset_specular_channel(rho_in, rho_out)
    return d, (omega, rho_in, rho_out)


def gravity(H=21, W=21, Fg=(-0.0003, 0.0)):
    omega, _, _ = channel_constants(H, W, 0.1)
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=H, Y=W, omega=omega, equilibrium=L.EQ_INCOMPRESSIBLE,
                                  force=L.FORCE_UNIFORM, Fg=Fg))
    d.preset_poiseuille(1.0, 1.0)
    return d, omega


def free_stream(X, Y, omega, uwx):
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=X, Y=Y, omega=omega, equilibrium=L.EQ_INCOMPRESSIBLE))
    d.preset_free_stream(uwx, 0.0)
    return d


def cylinder(X, Y, omega, u_lb, xs, ys):
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=X, Y=Y, omega=omega, equilibrium=L.EQ_COMPRESSIBLE,
                                  force=L.FORCE_IBM))
    d.preset_free_stream(u_lb, 0.0)
    d.ibm_set_markers(xs, ys)
    return d


def sedimentation(X, Y, omega, u_lb, w_s, C_w, walls, ibm=False, **slab):
    d = L.Domain(L.default_config(model=L.MODEL_BGK_ADE, X=X, Y=Y, omega=omega, omega_g=omega / 1.0,
                                  equilibrium=L.EQ_COMPRESSIBLE, w_s=w_s, force=L.FORCE_IBM if ibm else L.FORCE_NONE, **slab))
    d.preset_sedimentation(u_lb, C_w, int(walls[0]), int(walls[1]), int(walls[2]))
    return d


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


RED = dict(rho_0=3.0, alpha=0.7, A=0.5, nu=0.04, beta=0.7)     # configs/mrtcg-rayleigh-taylor-gamma3.toml
BLUE = dict(rho_0=1.0, alpha=0.1, A=0.5, nu=0.04, beta=-0.7)
RK_RED = dict(rho_0=1.2, alpha=1.0 / 3.0, A=1e-4, nu=0.16, beta=0.7)   # rk_static_droplet_test.cpp:504-506
RK_BLUE = dict(rho_0=1.0, alpha=0.2, A=1e-4, nu=0.14, beta=-0.7)


def mrtcg(R, C, Fg, add_force, sigma=0.1, **slab):
    d = L.Domain(L.default_config(model=L.MODEL_MRTCG, X=R, Y=C, red=RED, blue=BLUE, sigma=sigma, delta=0.1,
                                  Fg=Fg, add_force=add_force, **slab))
    d.preset_mrtcg()
    return d


def rk(Ln, **slab):
    d = L.Domain(L.default_config(model=L.MODEL_RK, X=Ln, Y=Ln, red=RK_RED, blue=RK_BLUE, delta=0.98, **slab))
    d.preset_rk()
    return d


def kbc(X, Y, s2, poiseuille=None, **slab):
    """ulbm::d2q9::kbc: fully periodic (test/ulbm_double_shear_flow.cpp) or, with poiseuille = (rho_in, rho_out),
    pressure rows + bounce-back columns (test/ulbm_poiseuille.cpp)"""
    d = L.Domain(L.default_config(model=L.MODEL_KBC, X=X, Y=Y, omega=s2, **slab))
    if poiseuille is None:
        d.preset_periodic()
    else:
        d.preset_poiseuille(*poiseuille)
    return d


def double_shear_fields(R, C, u_max=0.02, alpha=80.0, delta=0.05):
    """set_initial_conditions of test/ulbm_double_shear_flow.cpp:44-67"""
    r = np.arange(R)[:, None] + 0.0 * np.arange(C)[None, :]
    c = np.arange(C)[None, :] + 0.0 * np.arange(R)[:, None]
    u = np.zeros((R, C, 2))
    u[..., 0] = u_max * np.tanh(alpha * (0.25 * R - np.abs(c - 0.5 * R)))
    u[..., 1] = u_max * delta * np.sin(6.2832 * (r + 0.25 * R) / R)
    return np.ones((R, C)), u


def csf(R, C, Fg=(6.25e-6, 0.0), sigma=0.1, **slab):
    """test/mrt_rayleigh_taylor.cpp: MRT colour gradient with the continuum-surface-force perturbation"""
    d = L.Domain(L.default_config(model=L.MODEL_MRT_CSF, X=R, Y=C, red=RED, blue=BLUE, sigma=sigma, delta=0.1, Fg=Fg,
                                  add_force=1, **slab))
    d.preset_mrtcg()
    return d


def loop_blocks(Ln, omega, F=(3e-3, 0.0), device=0):
    """test/decompose_domain_loop.cpp: four blocks A {L, L/4}, B {L/4, L/2}, C {L, L/4}, D {L/4, L/2} closing a square
    channel; each block keeps its own coordinates and wall rules (:171-230), the column faces are bound pairwise
    (:232-261), block A carries the body force on rows L/4+5 .. L/4+55 (:66-69)."""
    L4, L2, END = Ln // 4, Ln // 2, L.LBM_END
    dom = {}
    for k, (R, Cc) in (("A", (Ln, L4)), ("B", (L4, L2)), ("C", (Ln, L4)), ("D", (L4, L2))):
        dom[k] = L.Domain(L.default_config(model=L.MODEL_BGK, X=R, Y=Cc, omega=omega, equilibrium=L.EQ_COMPRESSIBLE,
                                           force=L.FORCE_IBM if k == "A" else L.FORCE_NONE, device=device))

    def wall(d, xb, xe, yb, ye, pairs):
        for q, qs in pairs:
            d.bc_add(kind=L.BC_LINEAR, lattice=0, x_begin=xb, x_end=xe, y_begin=yb, y_end=ye, dst_q=q, src_q=qs, coef=1.0)

    top, bottom = [(8, 6), (1, 3), (5, 7)], [(7, 5), (3, 1), (6, 8)]
    left, right = [(2, 4), (5, 7), (6, 8)], [(4, 2), (7, 5), (8, 6)]
    for k, d in dom.items():
        d.bc_clear()
        wall(d, 0, 1, 0, END, top)
        wall(d, -1, END, 0, END, bottom)
    wall(dom["A"], L4, -L4, 0, 1, left)
    wall(dom["A"], 1, -1, -1, END, right)
    wall(dom["C"], 1, -1, 0, 1, left)
    wall(dom["C"], L4, -L4, -1, END, right)
    A, B, Cb, D = dom["A"], dom["B"], dom["C"], dom["D"]
    A.link_face(0, Ln - L4, L4, B, 0); B.link_face(1, 0, L4, A, Ln - L4)      # A-B
    B.link_face(0, 0, L4, Cb, Ln - L4); Cb.link_face(1, Ln - L4, L4, B, 0)    # B-C
    Cb.link_face(1, 0, L4, D, 0); D.link_face(0, 0, L4, Cb, 0)                # C-D
    D.link_face(1, 0, L4, A, 0); A.link_face(0, 0, L4, D, 0)                  # D-A
    A.set_force_region(L4 + 5, L4 + 55, 0, END, F[0], F[1], 3.0, 9.0)
    for d in dom.values():
        d.bc_commit()
    return dom


LOOP_BLOCKS = ("A", "B", "C", "D")


def loop_block_on_rank(Ln, omega, rank, unique_id, F=(3e-3, 0.0), device=0):
    """The same channel with ONE block per process (rank 0..3 = A..D): the column faces are bound across ranks
    (lbm_comm_init_blocks, lbm_link_face_rank; every rank then calls comm_faces_commit() and bc_commit())."""
    L4, L2, END = Ln // 4, Ln // 2, L.LBM_END
    shapes = {"A": (Ln, L4), "B": (L4, L2), "C": (Ln, L4), "D": (L4, L2)}
    k = LOOP_BLOCKS[rank]
    R, Cc = shapes[k]
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=R, Y=Cc, omega=omega, equilibrium=L.EQ_COMPRESSIBLE,
                                  force=L.FORCE_IBM if k == "A" else L.FORCE_NONE, device=device))
    d.comm_init_blocks(unique_id, 4, rank)

    def wall(xb, xe, yb, ye, pairs):
        for q, qs in pairs:
            d.bc_add(kind=L.BC_LINEAR, lattice=0, x_begin=xb, x_end=xe, y_begin=yb, y_end=ye, dst_q=q, src_q=qs, coef=1.0)

    top, bottom = [(8, 6), (1, 3), (5, 7)], [(7, 5), (3, 1), (6, 8)]
    left, right = [(2, 4), (5, 7), (6, 8)], [(4, 2), (7, 5), (8, 6)]
    d.bc_clear()
    wall(0, 1, 0, END, top)
    wall(-1, END, 0, END, bottom)
    A, B, Cb, D = 0, 1, 2, 3
    if k == "A":
        wall(L4, -L4, 0, 1, left)
        wall(1, -1, -1, END, right)
        d.link_face_rank(0, Ln - L4, L4, B, 0)     # A-B
        d.link_face_rank(0, 0, L4, D, 0)           # D-A
        d.set_force_region(L4 + 5, L4 + 55, 0, END, F[0], F[1], 3.0, 9.0)
    elif k == "B":
        d.link_face_rank(1, 0, L4, A, Ln - L4)     # A-B
        d.link_face_rank(0, 0, L4, Cb, Ln - L4)    # B-C
    elif k == "C":
        wall(1, -1, 0, 1, left)
        wall(L4, -L4, -1, END, right)
        d.link_face_rank(1, Ln - L4, L4, B, 0)     # B-C
        d.link_face_rank(1, 0, L4, D, 0)           # C-D
    else:
        d.link_face_rank(0, 0, L4, Cb, 0)          # C-D
        d.link_face_rank(1, 0, L4, A, 0)           # D-A
    return d
