"""TEST INFRASTRUCTURE — CPU restatement of test/decompose_domain_loop.cpp: four blocks A, B, C, D
forming a closed square channel, bound to one another across COLUMN faces, with a body force on part of block A.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.  The per-block arithmetic goes through an
`ops` object — the C port (tests/oracle_lib.Oracle) or the compiled reference (oracle_lib.Ref); this file adds the
driver's own loop body: which rows of which block are walls (decompose_domain_loop.cpp:171-230), the source term on block
A (:66-69,151-158) and the face bindings, transcribed line by line as data (:232-261).
"""
import numpy as np

CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1], dtype=np.float64)
CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1], dtype=np.float64)
W = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)

TOP = [(8, 6), (1, 3), (5, 7)]      # row 0   : f_adve[q] = f_coll[q']   (:173-176)
BOTTOM = [(7, 5), (3, 1), (6, 8)]   # row -1                            (:177-180)
LEFT = [(2, 4), (5, 7), (6, 8)]     # column 0                          (:181-184)
RIGHT = [(4, 2), (7, 5), (8, 6)]    # column -1                         (:185-188)


def shapes(L):
    return {"A": (L, L // 4), "B": (L // 4, L // 2), "C": (L, L // 4), "D": (L // 4, L // 2)}


def walls(L):
    """name -> list of (row slice, column index or slice, [(q, q')...])"""
    L4 = L // 4
    allc = slice(None)
    return {
        "A": [(slice(0, 1), allc, TOP), (slice(-1, None), allc, BOTTOM), (slice(L4, -L4), 0, LEFT), (slice(1, -1), -1, RIGHT)],
        "B": [(slice(0, 1), allc, TOP), (slice(-1, None), allc, BOTTOM)],
        "C": [(slice(0, 1), allc, TOP), (slice(-1, None), allc, BOTTOM), (slice(1, -1), 0, LEFT), (slice(L4, -L4), -1, RIGHT)],
        "D": [(slice(0, 1), allc, TOP), (slice(-1, None), allc, BOTTOM)],
    }


def bindings(L):
    """(dst block, dst rows, dst column, q, src block, src rows, src column), in the driver's order (:232-261)"""
    L4 = L // 4
    S = slice
    return [
        ("A", S(-L4, -1), 0, 6, "B", S(1, None), -1), ("A", S(-L4, None), 0, 2, "B", S(None), -1), ("A", S(-L4 + 1, None), 0, 5, "B", S(0, -1), -1),
        ("B", S(1, None), -1, 8, "A", S(-L4, -1), 0), ("B", S(None), -1, 4, "A", S(-L4, None), 0), ("B", S(0, -1), -1, 7, "A", S(-L4 + 1, None), 0),
        ("B", S(0, -1), 0, 6, "C", S(-L4 + 1, None), -1), ("B", S(None), 0, 2, "C", S(-L4, None), -1), ("B", S(1, None), 0, 5, "C", S(-L4, -1), -1),
        ("C", S(-L4, -1), -1, 7, "B", S(1, None), 0), ("C", S(-L4, None), -1, 4, "B", S(None), 0), ("C", S(-L4 + 1, None), -1, 8, "B", S(0, -1), 0),
        ("C", S(0, L4 - 1), -1, 7, "D", S(1, None), 0), ("C", S(0, L4), -1, 4, "D", S(None), 0), ("C", S(1, L4), -1, 8, "D", S(0, -1), 0),
        ("D", S(0, -1), 0, 6, "C", S(1, L4), -1), ("D", S(None), 0, 2, "C", S(0, L4), -1), ("D", S(1, None), 0, 5, "C", S(0, L4 - 1), -1),
        ("D", S(0, -1), -1, 7, "A", S(1, L4), 0), ("D", S(None), -1, 4, "A", S(0, L4), 0), ("D", S(1, None), -1, 8, "A", S(0, L4 - 1), 0),
        ("A", S(0, L4 - 1), 0, 6, "D", S(1, None), -1), ("A", S(0, L4), 0, 2, "D", S(None), -1), ("A", S(1, L4), 0, 5, "D", S(0, -1), -1),
    ]


def force_rows(L):
    return slice(L // 4 + 5, L // 4 + 55)   # force_idx (:67)


def init(ops, L):
    """solver::equilibrium(adve_f, m_1 = 0, m_0 = 1) on every block (:108-111)"""
    st = {}
    for k, (R, Cc) in shapes(L).items():
        u = np.zeros((R, Cc, 2)); rho = np.ones((R, Cc, 1))
        st[k] = {"f": ops.equilibrium(u, rho), "u": u, "rho": rho}
    return st


def step(ops, st, L, omega, F=(3e-3, 0.0), ics2=3.0, ics4=9.0):
    """one iteration of the main loop (:139-261); st[k]['u'], ['rho'] are left as the loop leaves m_1, m_0"""
    coll = {}
    for k, b in st.items():
        b["rho"] = ops.calc_rho(b["f"])
        b["u"] = ops.calc_u(b["f"], b["rho"])
        equi = ops.equilibrium(b["u"], b["rho"])
        if k == "A":
            c = b["f"] + (-omega * (b["f"] - equi))                      # :153-157
            fr = force_rows(L)
            u = b["u"][fr]
            cu = u[..., 0:1] * CX + u[..., 1:2] * CY
            cF = F[0] * CX + F[1] * CY
            uF = u[..., 0:1] * F[0] + u[..., 1:2] * F[1]
            c[fr] = c[fr] + ((1.0 - 0.5 * omega) * ((ics2 + ics4 * cu) * cF - ics2 * uF)) * W
            coll[k] = c
        else:
            coll[k] = ops.collision(b["f"], equi, omega)
    new = {k: ops.advect(coll[k]) for k in st}
    for k, rules in walls(L).items():
        for rows, cols, pairs in rules:
            for q, qs in pairs:
                new[k][rows, cols, q] = coll[k][rows, cols, qs]
    for dst, drows, dcol, q, src, srows, scol in bindings(L):
        new[dst][drows, dcol, q] = coll[src][srows, scol, q]
    for k in st:
        st[k]["f"] = new[k]
