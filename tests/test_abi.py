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
    assert ctypes.sizeof(_lib.ConfigC) == 2 * 4 * 44 + 5 * 4     # has_flow, T, recompute, use_graph, forward_fp16


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
