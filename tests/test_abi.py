"""CPU: the C-ABI library loads, exports every symbol include/lbm_b200.h declares, parses the
parameters.toml surface like the reference, and refuses to compute without a GPU."""
import os
import re

import numpy as np
import pytest

import lbm_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lbm_b200.h")
PARAMS = os.path.join(ROOT, "configs", "parameters.toml")
TWO_PHASE = os.path.join(ROOT, "configs", "mrtcg-rayleigh-taylor-gamma3.toml")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = L.load()
    names = declared_functions()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert missing == []
    assert sorted(L.EXPORTS) == names


def test_version_string():
    assert "sm_100a" in L.version()


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU error path")
def test_no_cpu_fallback():
    with pytest.raises(L.LbmError) as e:
        L.Domain(L.default_config(X=21, Y=21))
    assert e.value.status == L.ERR_CUDA and "no CPU fallback" in e.value.message
    with pytest.raises(L.LbmError) as e:
        L.calc_rho(np.ones((4, 4, 9)))
    assert e.value.status == L.ERR_CUDA


def test_params_from_toml_values():
    p = L.params_from_toml(PARAMS, True)
    # the derived values SURVEY §2c lists for the reference's parameters.toml
    assert (p.X, p.Y, p.l, p.T) == (2700, 2100, 300, 157995)
    assert abs(p.Re - 2848.19) < 0.01 and abs(p.omega - 1 / 0.55) < 1e-15 and abs(p.nu - 1 / 60) < 1e-15
    assert abs(p.u - 0.158233) < 1e-6 and abs(p.dt - 6.3293e-6) < 1e-9
    assert p.has_simulation == 1 and p.file_prefix == b"run-"
    assert p.total_steps == int(np.ceil(0.01 * p.T)) and p.snapshot_steps == int(np.ceil(0.001 * p.T))


def test_missing_keys_give_the_reference_messages(tmp_path):
    t = tmp_path / "p.toml"
    t.write_text("[flow]\ninitial_density = 1.0\nkinematic_viscosity = 1.0\ncharacteristic_velocity = 1.0\n")
    with pytest.raises(L.LbmError) as e:
        L.params_from_toml(str(t), False)
    assert e.value.status == L.ERR_CONFIG and e.value.message == "characteristic_length not defined in parameters file"
    t.write_text(open(PARAMS).read().rsplit("[simulation]", 1)[0])
    assert L.params_from_toml(str(t), False).has_simulation == 0
    with pytest.raises(L.LbmError) as e:
        L.params_from_toml(str(t), True)
    assert e.value.message == "stop_time not defined in parameters file"
    t.write_text("[red]\ninitial_density = 3.0\n")
    with pytest.raises(L.LbmError) as e:
        L.colour_from_toml(str(t), "red")
    assert e.value.message == "alphanot defined in parameters file"  # src/colour.cpp:45 glues them
    t.write_text("[flow\n")
    with pytest.raises(L.LbmError) as e:
        L.params_from_toml(str(t), False)
    assert e.value.status == L.ERR_CONFIG and e.value.message.startswith("Parsing failed")


def test_two_phase_and_colour_tables():
    tp = L.two_phase_from_toml(TWO_PHASE, True)
    assert (tp.rows, tp.columns, tp.time_steps, tp.nr_snapshots, tp.period_snapshots) == (256, 128, 100000, 1000, 100)
    assert tp.sigma == 0.1 and tp.gravity_magnitude == 6.25e-6 and tp.name == b"rt"
    r = L.colour_from_toml(TWO_PHASE, "red"); b = L.colour_from_toml(TWO_PHASE, "blue")
    assert abs(r.cs2 - 0.18) < 1e-15 and abs(r.rlx - 1.3846153846153846) < 1e-15
    assert abs(b.cs2 - 0.54) < 1e-15 and abs(b.rlx - 1.7419354838709677) < 1e-15
    assert r.phi[0] == 0.7 and abs(sum(r.phi) - 1.0) < 1e-15 and abs(sum(b.phi) - 1.0) < 1e-15


def test_decompose_rows_cover_the_grid():
    for X, P in [(21, 1), (42, 2), (8192, 8), (2700, 7), (10, 4)]:
        rows = [L.decompose_rows(X, P, r) for r in range(P)]
        assert rows[0][0] == 0 and rows[-1][1] == X
        assert all(rows[i][1] == rows[i + 1][0] for i in range(P - 1))
        sizes = [b - a for a, b in rows]
        assert max(sizes) - min(sizes) <= 1


def test_toml_reader_handles_the_reference_syntax(tmp_path):
    t = tmp_path / "b.toml"
    t.write_text('# comment\n[cylinder-a] # trailing\nx = [1.5, 2, 3.25e0,\n  4.0, # inner\n]\ny = [\n 0.5,\n 1_000.0, -2.5, +7\n]\n')
    xs, ys = L.markers_from_toml(str(t), "cylinder-a")
    assert list(xs) == [1.5, 2.0, 3.25, 4.0] and list(ys) == [0.5, 1000.0, -2.5, 7.0]
