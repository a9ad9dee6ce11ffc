"""GPU: the C ABI's error behaviour — status codes + lbm_last_error(), never an exception or a crash
(the reference throws std::runtime_error / c10::Error instead: src/params.cpp:13, src/ibm.cpp:90)."""
import numpy as np
import pytest

import cases
import lbm_b200 as L

pytestmark = pytest.mark.gpu


def status_of(fn):
    with pytest.raises(L.LbmError) as e:
        fn()
    return e.value.status, str(e.value)


def test_bad_geometry_and_models():
    assert status_of(lambda: L.Domain(L.default_config(model=L.MODEL_BGK, X=2, Y=10)))[0] == 1
    assert status_of(lambda: L.Domain(L.default_config(model=L.MODEL_BGK, X=10, Y=10, x0=4, x1=3)))[0] == 1
    assert status_of(lambda: L.Domain(L.default_config(model=9, X=10, Y=10)))[0] == 1
    assert status_of(lambda: L.Domain(L.default_config(model=L.MODEL_BGK, X=10, Y=10, device=99)))[0] == 1


def test_calls_out_of_order():
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=12, Y=12))
    st, msg = status_of(lambda: d.step(1))
    assert st == 1 and "no state" in msg
    d.set_f(np.zeros((12, 12, 9)))
    st, msg = status_of(lambda: d.step(1))
    assert st == 1 and "lbm_bc_commit" in msg
    d.preset_periodic()
    d.step(3)
    assert status_of(lambda: d.set_moments(np.ones((12, 12, 1)), np.zeros((12, 12, 2))))[0] == 4   # not a KBC domain
    assert status_of(lambda: d.get_phase())[0] == 1                                              # not two-phase
    assert status_of(lambda: d.init_equilibrium(np.ones((12, 12, 1)), np.zeros((12, 12, 2)), 7))[0] == 1


def test_rule_validation():
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=12, Y=12))
    d.bc_add(kind=L.BC_LINEAR, x_begin=0, x_end=1, dst_q=11, src_q=1)
    st, msg = status_of(lambda: d.bc_commit())
    assert st == 1 and "population index" in msg
    d.bc_clear()
    st, msg = status_of(lambda: d.bc_add(kind=L.BC_ADE_INLET, lattice=1, x_begin=0, x_end=1))   # needs per_row
    assert st == 1 and "per_row" in msg
    d.bc_add(kind=L.BC_LINEAR, lattice=1, x_begin=0, x_end=1, dst_q=1, src_q=3)                   # lattice 1 of a single-lattice model
    st, msg = status_of(lambda: d.bc_commit())
    assert st == 1 and "lattice" in msg


def test_ibm_roi_across_a_cut_needs_the_group_or_the_ring():
    """a body that crosses the slab's edge is accepted (the slabs share the solve), but such a slab cannot be stepped on
    its own; a body outside the grid or on the listed edge columns is refused; a body on other slabs is ignored"""
    d = L.Domain(L.default_config(model=L.MODEL_BGK, X=64, Y=64, x0=0, x1=32, force=L.FORCE_IBM))
    d.preset_free_stream(0.05)
    th = 2 * np.pi * np.arange(40) / 40
    d.ibm_set_markers(30.0 + 6 * np.cos(th), 32.0 + 6 * np.sin(th))
    d.set_f(np.full((32, 64, 9), 0.1))
    st, msg = status_of(lambda: d.step(1))
    assert st == 1 and "lbm_step_group" in msg
    st, msg = status_of(lambda: d.ibm_set_markers(60.0 + 6 * np.cos(th), 32.0 + 6 * np.sin(th)))
    assert st == 4 and "grid" in msg
    st, msg = status_of(lambda: d.ibm_set_markers(16.0 + 6 * np.cos(th), 5.0 + 6 * np.sin(th)))
    assert st == 4 and "interior columns" in msg
    d.ibm_set_markers(48.0 + 6 * np.cos(th), 32.0 + 6 * np.sin(th))   # rows 40..56: another slab's body
    st, msg = status_of(lambda: d.ibm_roi())
    assert st == 1 and "no immersed boundary" in msg


def test_linked_slab_refuses_plain_step():
    kw = dict(model=L.MODEL_BGK, X=20, Y=16, omega=1.2)
    a = L.Domain(L.default_config(x0=0, x1=10, **kw)); b = L.Domain(L.default_config(x0=10, x1=20, **kw))
    for d in (a, b):
        d.preset_periodic()
        d.set_f(np.full((10, 16, 9), 0.1))
    a.link(b, b); b.link(a, a)
    st, msg = status_of(lambda: a.step(1))
    assert st == 1 and "lbm_step_group" in msg
    L.step_group([a, b], 2)


def test_row_split_and_ring_calls_out_of_order():
    """lbm_row_split reports the step's early / bulk rows (single-phase family only); lbm_comm_share needs a ring member;
    the binding refuses arrays of the wrong shape before their pointers cross the C ABI"""
    d, _ = cases.poiseuille(40, 33)
    assert d.row_split() == (4, 36)          # rows 0, 1, X-2, X-1 feed the pressure-periodic stages
    tp = cases.mrtcg(24, 20, (0.0, 0.0), 0)
    st, msg = status_of(lambda: tp.row_split())
    assert st == 4 and "two-phase" in msg
    other = L.Domain(L.default_config(model=L.MODEL_BGK, X=12, Y=12))
    st, msg = status_of(lambda: d.comm_share(other))
    assert st == 1 and "has not joined a ring" in msg
    with pytest.raises(ValueError, match="expected shape"):
        d.set_f(np.zeros((40, 32, 9)))
    with pytest.raises(ValueError, match="expected shape"):
        tp.init_two_phase(np.ones((24, 20)), np.ones((24, 20)), np.zeros((24, 21, 2)))
    with pytest.raises(ValueError, match="one length"):
        cases.cylinder(40, 40, 1.2, 0.05, np.zeros(5), np.zeros(4))
