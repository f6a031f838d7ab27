"""CPU: the C-ABI library loads and exports every symbol include/mumpy_b200.h declares; the ctypes table matches
the header's prototypes (argument counts); no compute call is made (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_prototypes():
    src = open(os.path.join(ROOT, "include", "mumpy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:int|long|const char \*)\s*\*?(mumpy_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return protos


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from mumpy_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(lib):
    handle = lib.load()
    protos = header_prototypes()
    assert len(protos) >= 25
    for name in protos:
        assert hasattr(handle, name), name
    assert handle.mumpy_abi_version() == 1


def test_ctypes_table_matches_header(lib):
    protos = header_prototypes()
    protos.pop("mumpy_last_error")
    assert set(protos) == set(lib.SIGNATURES)
    for name, argtypes in lib.SIGNATURES.items():
        assert len(argtypes) == protos[name], name


def test_no_cpu_fallback():
    """CPU tensors must be rejected loudly, never computed on the host."""
    import torch
    import mumpy_b200
    from mumpy_b200._lib import MumpyError
    with pytest.raises(MumpyError):
        mumpy_b200.ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8))
    blk = mumpy_b200.models.modules.blocks.Block(32, 2, 64, 0.0, 0.0).eval()
    with pytest.raises(MumpyError):
        blk(torch.zeros(2, 3, 32))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the product package may import or execute it."""
    pkg = os.path.join(ROOT, "multilateral-temporal-view-pyramid-transformer-for-video-inpainting-detection_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), os.path.join(dp, f)
                assert "mumpy_oracle" not in text and "ref_harness" not in text, os.path.join(dp, f)


def test_workspace_size_queries(lib):
    """The caller-owned workspaces are sized by the library itself (no GPU call): statistics + per-chunk partial sums of GroupNorm,
    the five fp32 image sets of the fp32 FAF, the two ping-pong halves of the 16-bit FAF."""
    h = lib.load()
    assert h.mumpy_groupnorm_workspace_floats(2, 112 * 112, 128, 32) == 2 * 2 * 32 * (1 + (112 * 112 + 63) // 64)
    assert h.mumpy_groupnorm_workspace_floats(1, 10, 1024, 32) == 2 * 1 * 32 * (1 + 1)            # 12 pixels fit a chunk, HW = 10 -> one chunk
    assert h.mumpy_groupnorm_workspace_floats(0, 10, 64, 8) == 0
    assert h.mumpy_faf_workspace_floats(4, 224) == 5 * 4 * 3 * 224 * 224
    assert h.mumpy_faf16_workspace_bytes(4, 224) == 2 * 9 * 4 * 224 * 3 * 224 * 2
