"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/vaevar.h declares; the product path refuses to run without a GPU (no fallback)."""
import ctypes
import pathlib
import re

import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from vaevar_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = (ROOT / "include" / "vaevar.h").read_text()
    declared = set(re.findall(r"VV_API\s+[\w\s\*]+?\b(vv_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vaevar.h but not exported"


def test_binding_covers_header(lib):
    from vaevar_b200 import _lib
    hdr = (ROOT / "include" / "vaevar.h").read_text()
    declared = set(re.findall(r"VV_API\s+[\w\s\*]+?\b(vv_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTED)


def test_config_struct_layout_matches_header():
    from vaevar_b200 import _lib
    # 3 + 8 + 8 + 3 + 2 + 2 + 1 + 8 + 8 + 1 ints
    assert ctypes.sizeof(_lib.NetConfigC) == 4 * 44
    assert ctypes.sizeof(_lib.ConfigC) == 2 * 4 * 44 + 6 * 4     # has_flow, T, recompute, use_graph, forward_fp16, no_ln_fold


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_fallback_without_gpu():
    from vaevar_b200.config import DECODER_FULL, small
    from vaevar_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(small(DECODER_FULL))


def test_product_does_not_import_oracle():
    for p in (ROOT / "vaevar_b200").glob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


@pytest.mark.parametrize("lr,hr", [((128, 256), (721, 1440)), ((32, 64), (181, 360)), ((8, 16), (16, 32)), ((7, 9), (7, 9)), ((5, 6), (3, 4))])
def test_seam_index_tables_follow_the_reference_rule(lib, gold, lr, hr):
    """Host-only entry point (no device call): the library's nearest-resampling tables against the oracle's restatement of ATen's
    rule and, at the reference's sizes, against the tables read off torch's own F.interpolate (tests/golden/seams.npz)."""
    import numpy as np
    from oracle.seams import nearest_index
    (H, W), (Hh, Wh) = lr, hr
    arr = lambda n: np.zeros(n, np.int32)
    ur, uc, dr, dc, sr, sc, srl, scl = arr(Hh), arr(Wh), arr(H), arr(W), arr(H), arr(W), arr(H + 1), arr(W + 1)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.vv_debug_seam_tables(H, W, Hh, Wh, P(ur), P(uc), P(dr), P(dc), P(sr), P(sc), P(srl), P(scl)) == 0
    assert np.array_equal(ur, nearest_index(Hh, H)) and np.array_equal(uc, nearest_index(Wh, W))
    assert np.array_equal(dr, nearest_index(H, Hh)) and np.array_equal(dc, nearest_index(W, Wh))
    if (lr, hr) == ((128, 256), (721, 1440)):
        g = gold("seams.npz")
        assert np.array_equal(ur, g["up_rows"]) and np.array_equal(uc, g["up_cols"])
        assert np.array_equal(dr, g["down_rows"]) and np.array_equal(dc, g["down_cols"])
        assert not np.array_equal(sr, np.arange(H))                 # the round trip is not the identity at the reference's sizes
    assert np.array_equal(sr, ur[dr]) and np.array_equal(sc, uc[dc])  # S = down o up
    for tab, lo, n in ((sr, srl, H), (sc, scl, W)):
        assert lo[0] == 0 and lo[n] == n and (np.diff(lo) >= 0).all()
        for r in range(n):
            assert (tab[lo[r]:lo[r + 1]] == r).all() and (tab[:lo[r]] < r).all() and (tab[lo[r + 1]:] > r).all()


def test_net1_config_struct_layout_matches_header():
    from vaevar_b200 import _lib
    # img_h, img_w, n_groups | in 8 | out 8 | enc_dim, embed_dim, win_h, win_w, n_levels | depth 4 | heads 4 | n_lg | 8 | 8 | keep_out
    assert ctypes.sizeof(_lib.Net1ConfigC) == 4 * (3 + 8 + 8 + 5 + 4 + 4 + 1 + 8 + 8 + 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_forecast_net_has_no_fallback_without_gpu():
    from vaevar_b200.config import FORECAST_MID
    from vaevar_b200.forecast import ForecastNet
    from vaevar_b200.modules import LGUnet_all_1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ForecastNet(FORECAST_MID)
    m = LGUnet_all_1(**FORECAST_MID.to_reference_kwargs())
    assert len(list(m.state_dict())) == 1079
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 69, 97, 192))
