"""CPU: lbm_save_pt writes the archive torch::save(tensor, path) writes in the reference drivers
(test/horizontal_poiseuille_test.cpp:157-160): loadable by torch.jit.load, bit-exact contents, the pickle
byte-identical to what libtorch 2.11 emits for the same tensor."""
import zipfile

import numpy as np
import pytest
import torch

import lbm_b200 as L


def load_pt(path):
    m = torch.jit.load(str(path), map_location="cpu")
    return list(m.parameters())[0].detach().numpy()


@pytest.mark.parametrize("shape", [(21, 21, 84), (5, 7, 9, 3), (30, 2), (1,), (3, 0, 4), (70000, 3)])
def test_round_trip_through_torch(tmp_path, shape):
    a = np.random.default_rng(0).standard_normal(shape)
    path = tmp_path / "hpt-ux.pt"
    L.save_pt(path, a)
    b = load_pt(path)
    assert b.dtype == np.float64 and b.shape == tuple(shape) and np.array_equal(a, b)
    with zipfile.ZipFile(path) as z:
        assert z.testzip() is None  # CRCs
        assert z.namelist()[0] == "hpt-ux/data/0" and "hpt-ux/data.pkl" in z.namelist()
        assert z.read("hpt-ux/version") == b"3\n" and z.read("hpt-ux/byteorder") == b"little"


def test_pickle_bytes_match_libtorch():
    # data.pkl of torch::save(torch::arange(24, kDouble).reshape({2,3,4}), "sample.pt") written by libtorch 2.11 (C++)
    want = (b"\x80\x02c__torch__\nModule\nq\x00)\x81}(X\x01\x00\x00\x000q\x01ctorch._utils\n_rebuild_tensor_v2\nq\x02((X\x07\x00\x00\x00storage"
            b"q\x03ctorch\nDoubleStorage\nq\x04h\x01X\x03\x00\x00\x00cpuq\x05K\x18tQq\x06K\x00(K\x02K\x03K\x04t(K\x0cK\x04K\x01t\x89ccollections\n"
            b"OrderedDict\nq\x07)RtRq\x08ubq\t.")
    import os
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "sample.pt")
        L.save_pt(p, np.arange(24, dtype=np.float64).reshape(2, 3, 4))
        with zipfile.ZipFile(p) as z:
            assert z.read("sample/data.pkl") == want
            assert z.read("sample/data/0") == np.arange(24, dtype=np.float64).tobytes()


def test_storage_is_64_byte_aligned(tmp_path):
    import struct

    path = tmp_path / "x.pt"
    L.save_pt(path, np.ones((3, 5)))
    raw = path.read_bytes()
    sig, _, _, _, _, _, _, cs, us, nl, el = struct.unpack("<IHHHHHIIIHH", raw[:30])
    assert sig == 0x04034B50 and cs == us == 120 and (30 + nl + el) % 64 == 0


def test_bad_arguments():
    with pytest.raises(L.LbmError):
        L.save_pt("/nonexistent-dir/x.pt", np.zeros(3))
