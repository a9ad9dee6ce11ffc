"""ctypes bindings for the parity checkers (TEST INFRASTRUCTURE).

* ``Oracle``  -> oracle/_build/liblbm_oracle.so, the plain-C restatement (oracle/lbm_oracle.c)
* ``Ref``     -> oracle/_ref/libref_harness.so, the unmodified reference compiled for CPU libtorch

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liblbm_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

dp = C.POINTER(C.c_double)
lp = C.POINTER(C.c_long)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(ROOT, "oracle", "lbm_oracle.c")
    ):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class MrtcgParams(C.Structure):
    _fields_ = [
        ("R", C.c_int), ("C", C.c_int),
        ("r_rho0", C.c_double), ("r_alpha", C.c_double), ("r_nu", C.c_double), ("r_beta", C.c_double),
        ("b_rho0", C.c_double), ("b_alpha", C.c_double), ("b_nu", C.c_double), ("b_beta", C.c_double),
        ("sigma", C.c_double), ("delta", C.c_double), ("Fg", C.c_double * 2), ("add_force", C.c_int),
    ]


class CsfParams(C.Structure):
    _fields_ = [
        ("R", C.c_int), ("C", C.c_int),
        ("r_rho0", C.c_double), ("r_alpha", C.c_double), ("r_nu", C.c_double), ("r_beta", C.c_double), ("r_A", C.c_double),
        ("b_rho0", C.c_double), ("b_alpha", C.c_double), ("b_nu", C.c_double), ("b_beta", C.c_double), ("b_A", C.c_double),
        ("sigma", C.c_double), ("delta", C.c_double), ("Fg", C.c_double * 2),
    ]


class RkParams(C.Structure):
    _fields_ = [
        ("L", C.c_int), ("radius", C.c_double),
        ("r_rho0", C.c_double), ("r_alpha", C.c_double), ("r_A", C.c_double), ("r_nu", C.c_double),
        ("b_rho0", C.c_double), ("b_alpha", C.c_double), ("b_A", C.c_double), ("b_nu", C.c_double),
        ("delta", C.c_double),
    ]


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_ibm_create.restype = C.c_void_p
        L.orc_ibm_create.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_ibm_destroy.argtypes = [C.c_void_p]
        L.orc_ibm_roi.argtypes = [C.c_void_p, lp]
        L.orc_ibm_force.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int, dp]
        L.orc_cylinder_step.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, dp]

    def num_threads(self, n=0):
        """OpenMP threads of the port's loops; n > 0 sets the count first"""
        return int(self.lib.orc_num_threads(int(n)))

    # ---- granular ops
    def constants(self):
        w = np.zeros(9); c = np.zeros((2, 9))
        self.lib.orc_constants(_p(w), _p(c))
        return w, c

    def calc_rho(self, f):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 1))
        self.lib.orc_calc_rho(_p(f), X, Y, _p(out))
        return out

    def calc_u(self, f, rho):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 2))
        self.lib.orc_calc_u(_p(f), _p(rho), X, Y, _p(out))
        return out

    def calc_incomp_u(self, f):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 2))
        self.lib.orc_calc_incomp_u(_p(f), X, Y, _p(out))
        return out

    def equilibrium(self, u, rho):
        X, Y, _ = u.shape
        out = np.zeros((X, Y, 9))
        self.lib.orc_equilibrium(_p(u), _p(rho), X, Y, _p(out))
        return out

    def incomp_equilibrium(self, u, rho):
        X, Y, _ = u.shape
        out = np.zeros((X, Y, 9))
        self.lib.orc_incomp_equilibrium(_p(u), _p(rho), X, Y, _p(out))
        return out

    def collision(self, f, feq, omega):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 9))
        self.lib.orc_collision(_p(f), _p(feq), C.c_double(omega), X, Y, _p(out))
        return out

    def advect(self, f):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 9))
        self.lib.orc_advect(_p(f), X, Y, _p(out))
        return out

    def diff5(self, psi):
        R, Cc = psi.shape
        dx = np.zeros((R, Cc)); dy = np.zeros((R, Cc))
        self.lib.orc_diff5(_p(psi), R, Cc, _p(dx), _p(dy))
        return dx, dy

    def diff3(self, psi):
        R, Cc = psi.shape
        dx = np.zeros((R, Cc)); dy = np.zeros((R, Cc))
        self.lib.orc_diff3(_p(psi), R, Cc, _p(dx), _p(dy))
        return dx, dy

    # ---- host parameters
    def params_lattice(self, rho0, nu, u, l, tau, dx, xm, ym):
        i = np.array([rho0, nu, u, l, tau, dx, xm, ym], dtype=np.float64)
        o = np.zeros(9)
        self.lib.orc_params_lattice(_p(i), _p(o))
        return dict(Re=o[0], omega=o[1], nu=o[2], l=int(o[3]), dt=o[4], T=int(o[5]), u=o[6], X=int(o[7]), Y=int(o[8]))

    def params_simulation(self, stop_time, period, T):
        o = np.zeros(3)
        self.lib.orc_params_simulation(C.c_double(stop_time), C.c_double(period), int(T), _p(o))
        return dict(total_steps=int(o[0]), snapshot_steps=int(o[1]), total_snapshots=int(o[2]))

    def colour_params(self, rho0, alpha, nu):
        o = np.zeros(22)
        self.lib.orc_colour_params(C.c_double(rho0), C.c_double(alpha), C.c_double(nu), _p(o))
        return dict(mu=o[0], cs2=o[1], ics2=o[2], rlx=o[3], phi=o[4:13].copy(), eta=o[13:22].copy())

    # ---- drivers (in-place on f, u, rho)
    def poiseuille_step(self, f, u, rho, omega, rho_in, rho_out):
        X, Y, _ = f.shape
        self.lib.orc_poiseuille_step(_p(f), _p(u), _p(rho), X, Y, C.c_double(omega), C.c_double(rho_in), C.c_double(rho_out))

    def specular_step(self, f, u, rho, omega, rho_in, rho_out):
        X, Y, _ = f.shape
        self.lib.orc_specular_step(_p(f), _p(u), _p(rho), X, Y, C.c_double(omega), C.c_double(rho_in), C.c_double(rho_out))

    def gravity_step(self, f, u, rho, omega, rho_in, rho_out, Fg):
        X, Y, _ = f.shape
        Fg = np.ascontiguousarray(Fg, dtype=np.float64)
        self.lib.orc_gravity_step(_p(f), _p(u), _p(rho), X, Y, C.c_double(omega), C.c_double(rho_in), C.c_double(rho_out), _p(Fg))

    def free_stream_step(self, f, u, rho, omega, uwx):
        X, Y, _ = f.shape
        self.lib.orc_free_stream_step(_p(f), _p(u), _p(rho), X, Y, C.c_double(omega), C.c_double(uwx))

    def decompose_step(self, fA, uA, rhoA, fB, uB, rhoB, omega, rho_in, rho_out):
        X, Y, _ = fA.shape
        self.lib.orc_decompose_step(_p(fA), _p(uA), _p(rhoA), _p(fB), _p(uB), _p(rhoB), X, Y,
                                    C.c_double(omega), C.c_double(rho_in), C.c_double(rho_out))

    # ---- IBM / cylinder
    def ibm_create(self, xs, ys, m_max=5):
        xs = np.ascontiguousarray(xs, dtype=np.float64); ys = np.ascontiguousarray(ys, dtype=np.float64)
        return self.lib.orc_ibm_create(_p(xs), _p(ys), len(xs), m_max)

    def ibm_destroy(self, h):
        self.lib.orc_ibm_destroy(h)

    def ibm_roi(self, h):
        roi = (C.c_long * 4)()
        self.lib.orc_ibm_roi(h, roi)
        return tuple(int(v) for v in roi)

    def ibm_force(self, h, u, rho):
        X, Y, _ = u.shape
        r0, r1, c0, c1 = self.ibm_roi(h)
        F = np.zeros((r1 - r0, c1 - c0, 2))
        self.lib.orc_ibm_force(h, _p(u), _p(rho), X, Y, _p(F))
        return F

    def cylinder_step(self, f, u, rho, omega, u_lb, h):
        X, Y, _ = f.shape
        r0, r1, c0, c1 = self.ibm_roi(h)
        F = np.zeros((r1 - r0, c1 - c0, 2))
        self.lib.orc_cylinder_step(_p(f), _p(u), _p(rho), X, Y, C.c_double(omega), C.c_double(u_lb), h, _p(F))
        return F

    # ---- sedimentation
    def sedimentation_init(self, X, Y, u_lb, C_w):
        f = np.zeros((X, Y, 9)); g = np.zeros((X, Y, 9)); u = np.zeros((X, Y, 2))
        rho = np.zeros((X, Y, 1)); Cc = np.zeros((X, Y, 1))
        self.lib.orc_sedimentation_init(_p(f), _p(g), _p(u), _p(rho), _p(Cc), X, Y, C.c_double(u_lb), _p(C_w))
        return f, g, u, rho, Cc

    def sedimentation_step(self, f, g, u, rho, Cc, omega, u_lb, w_s, C_w, R23, C28, C38, ib=None):
        X, Y, _ = f.shape
        if ib is not None:  # with an immersed body (not a reference driver; see lbm_oracle.h)
            self.lib.orc_sedimentation_ibm_step.argtypes = [dp] * 5 + [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp,
                                                               C.c_int, C.c_int, C.c_int, C.c_void_p]
            self.lib.orc_sedimentation_ibm_step(_p(f), _p(g), _p(u), _p(rho), _p(Cc), X, Y, omega, u_lb, w_s, _p(C_w),
                                                int(R23), int(C28), int(C38), ib)
            return
        self.lib.orc_sedimentation_step(_p(f), _p(g), _p(u), _p(rho), _p(Cc), X, Y, C.c_double(omega),
                                        C.c_double(u_lb), C.c_double(w_s), _p(C_w), int(R23), int(C28), int(C38))

    # ---- MRT colour gradient
    def mrtcg_init(self, p, kind):
        N = (p.R, p.C)
        r_rho = np.zeros(N + (1,)); b_rho = np.zeros(N + (1,))
        if kind == "rt":
            self.lib.orc_mrtcg_init_rt(C.byref(p), _p(r_rho), _p(b_rho))
        else:
            self.lib.orc_mrtcg_init_droplet(C.byref(p), _p(r_rho), _p(b_rho))
        rho = np.zeros(N + (1,)); u = np.zeros(N + (2,))
        r_adv = np.zeros(N + (9,)); b_adv = np.zeros(N + (9,))
        self.lib.orc_mrtcg_init_state(C.byref(p), _p(r_rho), _p(b_rho), _p(rho), _p(u), _p(r_adv), _p(b_adv),
                                      1 if kind == "droplet" else 0)
        st = dict(r_adv=r_adv, b_adv=b_adv, r_rho=r_rho, b_rho=b_rho, rho=rho, u=u,
                  phase=np.zeros(N + (1,)), s_nu=np.zeros(N), grad=np.zeros(N + (2,)))
        return st

    def mrtcg_step(self, p, st):
        self.lib.orc_mrtcg_step(C.byref(p), _p(st["r_adv"]), _p(st["b_adv"]), _p(st["r_rho"]), _p(st["b_rho"]),
                                _p(st["rho"]), _p(st["u"]), _p(st["phase"]), _p(st["s_nu"]), _p(st["grad"]))

    # ---- MRT colour gradient with continuum surface force (test/mrt_rayleigh_taylor.cpp)
    def csf_init(self, p):
        N = (p.R, p.C)
        st = dict(r_rho=np.zeros(N + (1,)), b_rho=np.zeros(N + (1,)), rho=np.zeros(N + (1,)), u=np.zeros(N + (2,)),
                  r_adv=np.zeros(N + (9,)), b_adv=np.zeros(N + (9,)), phase=np.zeros(N + (1,)), s_nu=np.zeros(N),
                  Fs=np.zeros(N + (2,)))
        self.lib.orc_csf_init(C.byref(p), _p(st["r_rho"]), _p(st["b_rho"]), _p(st["rho"]), _p(st["u"]), _p(st["r_adv"]), _p(st["b_adv"]))
        return st

    def csf_step(self, p, st):
        self.lib.orc_csf_step(C.byref(p), _p(st["r_adv"]), _p(st["b_adv"]), _p(st["r_rho"]), _p(st["b_rho"]), _p(st["rho"]),
                              _p(st["u"]), _p(st["phase"]), _p(st["s_nu"]), _p(st["Fs"]))

    # ---- ulbm::d2q9::kbc
    def kbc_equilibrium(self, m0, m1, fresh_object=True):
        """fresh_object: kbc::eval_equilibrium as the driver calls it for its initial state (ux2 = uy2 = 0 still)"""
        X, Y = m0.shape
        out = np.zeros((X, Y, 9))
        self.lib.orc_kbc_equilibrium(_p(m0), _p(m1), X, Y, 1 if fresh_object else 0, _p(out))
        return out

    def kbc_step(self, f, m0, m1, s2, bc=0, rho_in=1.0, rho_out=1.0):
        X, Y, _ = f.shape
        self.lib.orc_kbc_step(_p(f), _p(m0), _p(m1), X, Y, C.c_double(s2), int(bc), C.c_double(rho_in), C.c_double(rho_out))

    # ---- RK droplet
    def rk_init(self, p, u0=None):
        L = p.L
        st = dict(r_adv=np.zeros((L, L, 9)), b_adv=np.zeros((L, L, 9)), r_rho=np.zeros((L, L)),
                  b_rho=np.zeros((L, L)), rho=np.zeros((L, L)),
                  u=np.zeros((L, L, 2)) if u0 is None else np.ascontiguousarray(u0, dtype=np.float64).copy(),
                  phase=np.zeros((L, L)), relax=np.zeros((L, L)), grad=np.zeros((L, L, 2)))
        self.lib.orc_rk_init(C.byref(p), _p(st["u"]), _p(st["r_adv"]), _p(st["b_adv"]), _p(st["r_rho"]),
                             _p(st["b_rho"]), _p(st["rho"]))
        return st

    def rk_step(self, p, st):
        self.lib.orc_rk_step(C.byref(p), _p(st["r_adv"]), _p(st["b_adv"]), _p(st["r_rho"]), _p(st["b_rho"]),
                             _p(st["rho"]), _p(st["u"]), _p(st["phase"]), _p(st["relax"]), _p(st["grad"]))

    def rk_diagnostics(self, p, st, sigma=5e-3):
        """the fields driver 17 snapshots per iteration, from the state BEFORE rk_step (relax is not advanced: a copy)"""
        L = p.L
        out = dict(phase=np.zeros((L, L)), grad=np.zeros((L, L, 2)), norm=np.zeros((L, L)), n=np.zeros((L, L, 2)),
                   K=np.zeros((L, L)), Fs=np.zeros((L, L, 2)), eta=np.zeros((L, L, 9)), kappa=np.zeros((L, L, 9)),
                   rparams=st["relax"].copy(), omega1=np.zeros((L, L, 9)), omega2=np.zeros((L, L, 9)))
        self.lib.orc_rk_diagnostics.argtypes = [C.c_void_p, C.c_double] + [dp] * 16
        self.lib.orc_rk_diagnostics(C.byref(p), sigma, _p(st["r_adv"]), _p(st["r_rho"]), _p(st["b_rho"]), _p(st["rho"]), _p(st["u"]),
                                    _p(out["phase"]), _p(out["grad"]), _p(out["norm"]), _p(out["n"]), _p(out["K"]), _p(out["Fs"]),
                                    _p(out["eta"]), _p(out["kappa"]), _p(out["rparams"]), _p(out["omega1"]), _p(out["omega2"]))
        out["omega3"] = out["omega1"] + out["omega2"]
        return out


def have_ref():
    return os.path.exists(REF_SO)


class Ref:
    """The unmodified reference library (CPU libtorch) behind oracle/ref_harness.cpp."""

    def __init__(self):
        import torch  # noqa: F401  (loads libtorch so the harness resolves its symbols)

        self.lib = C.CDLL(REF_SO)
        self.lib.ref_last_error.restype = C.c_char_p

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def constants(self):
        w = np.zeros(9); c = np.zeros((2, 9))
        self._chk(self.lib.ref_constants(_p(w), _p(c)))
        return w, c

    def calc_rho(self, f):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 1))
        self._chk(self.lib.ref_calc_rho(_p(f), X, Y, _p(out)))
        return out

    def calc_u(self, f, rho):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 2))
        self._chk(self.lib.ref_calc_u(_p(f), _p(rho), X, Y, _p(out)))
        return out

    def calc_incomp_u(self, f):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 2))
        self._chk(self.lib.ref_calc_incomp_u(_p(f), X, Y, _p(out)))
        return out

    def equilibrium(self, u, rho):
        X, Y, _ = u.shape
        out = np.zeros((X, Y, 9))
        self._chk(self.lib.ref_equilibrium(_p(u), _p(rho), X, Y, _p(out)))
        return out

    def incomp_equilibrium(self, u, rho):
        X, Y, _ = u.shape
        out = np.zeros((X, Y, 9))
        self._chk(self.lib.ref_incomp_equilibrium(_p(u), _p(rho), X, Y, _p(out)))
        return out

    def collision(self, f, feq, omega):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 9))
        self._chk(self.lib.ref_collision(_p(f), _p(feq), C.c_double(omega), X, Y, _p(out)))
        return out

    def advect(self, f):
        X, Y, _ = f.shape
        out = np.zeros((X, Y, 9))
        self._chk(self.lib.ref_advect(_p(f), X, Y, _p(out)))
        return out

    def differential(self, psi):
        R, Cc = psi.shape
        dx = np.zeros((R, Cc)); dy = np.zeros((R, Cc))
        self._chk(self.lib.ref_differential(_p(psi), R, Cc, _p(dx), _p(dy)))
        return dx, dy

    def params(self, path, with_simulation=True):
        o = np.zeros(21)
        self._chk(self.lib.ref_params(path.encode(), 1 if with_simulation else 0, _p(o)))
        keys = ["fp_nu", "fp_u", "fp_l", "fp_rho_0", "fp_Re", "tau", "omega", "Re", "nu", "l", "dx", "dt", "T",
                "u", "X", "Y", "stop_time", "snapshot_period", "total_steps", "snapshot_steps", "total_snapshots"]
        return dict(zip(keys, o))

    def colour(self, path, table):
        o = np.zeros(27)
        self._chk(self.lib.ref_colour(path.encode(), table.encode(), _p(o)))
        keys = ["rho_0", "alpha", "A", "nu", "mu", "beta", "cs2", "ics2", "rlx"]
        d = dict(zip(keys, o[:9]))
        d["phi"] = o[9:18].copy(); d["eta"] = o[18:27].copy()
        return d

    def ibm_force(self, path, name, u, rho):
        X, Y, _ = u.shape
        roi = (C.c_long * 4)()
        self._chk(self.lib.ref_ibm_force(path.encode(), name.encode(), None, None, X, Y, roi, None))
        r0, r1, c0, c1 = (int(v) for v in roi)
        F = np.zeros((r1 - r0, c1 - c0, 2))
        self._chk(self.lib.ref_ibm_force(path.encode(), name.encode(), _p(u), _p(rho), X, Y, roi, _p(F)))
        return (r0, r1, c0, c1), F

    def cylinder_loop(self, X, Y, omega, u_lb, markers_toml, warmup, steps):
        """seconds per step of the loop body of test/cylinder_test.cpp:100-163 on CPU libtorch"""
        sec = C.c_double(); chk = C.c_double()
        self._chk(self.lib.ref_cylinder_loop(X, Y, C.c_double(omega), C.c_double(u_lb), markers_toml.encode(),
                                             warmup, steps, C.byref(sec), C.byref(chk)))
        return sec.value, chk.value

    def poiseuille_loop(self, H, W, omega, rho_in, rho_out, warmup, steps):
        """seconds per step and checksum of the loop body of test/horizontal_poiseuille_test.cpp:100-153 on CPU libtorch"""
        if not hasattr(self.lib, "ref_poiseuille_loop"):
            raise RuntimeError("oracle/_ref/libref_harness.so predates ref_poiseuille_loop: rebuild it (make -C oracle ref)")
        sec = C.c_double(); chk = C.c_double()
        self._chk(self.lib.ref_poiseuille_loop(H, W, C.c_double(omega), C.c_double(rho_in), C.c_double(rho_out), warmup, steps,
                                               C.byref(sec), C.byref(chk)))
        return sec.value, chk.value

    def kbc_run(self, f, m0, m1, s2, steps, bc=0, rho_in=1.0, rho_out=1.0):
        """ulbm::d2q9::kbc stepped like its drivers (in place on f = adve_f, m0, m1)"""
        X, Y, _ = f.shape
        self._chk(self.lib.ref_kbc_run(X, Y, C.c_double(s2), _p(m0), _p(m1), _p(f), int(steps), int(bc),
                                       C.c_double(rho_in), C.c_double(rho_out)))

    def kbc_equilibrium(self, m0, m1):
        X, Y = m0.shape
        out = np.zeros((X, Y, 9))
        self._chk(self.lib.ref_kbc_equilibrium(X, Y, _p(m0), _p(m1), _p(out)))
        return out

    def num_threads(self):
        return int(self.lib.ref_get_num_threads())

    def domain_shapes(self, R, Cc):
        s = (C.c_long * 15)()
        self._chk(self.lib.ref_domain_shapes(R, Cc, s))
        return [tuple(s[3 * i:3 * i + 3]) for i in range(5)]


def run_ref_driver(name, args, cwd, stdin=None, timeout=3600):
    """Run one of the reference's own driver binaries (oracle/_ref/<name>) in `cwd`."""
    exe = os.path.join(REF_DIR, name)
    return subprocess.run([exe] + list(args), cwd=cwd, input=stdin, capture_output=True, text=True, timeout=timeout)


def load_pt(path):
    """Load a tensor written by the reference's torch::save (a TorchScript pickle archive)."""
    import torch

    m = torch.jit.load(path, map_location="cpu")
    ps = list(m.parameters())
    if ps:
        return ps[0].detach().numpy()
    bs = list(m.buffers())
    return bs[0].detach().numpy()
