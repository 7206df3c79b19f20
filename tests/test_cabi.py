"""CPU: the C-ABI shared library builds, loads and exports every symbol include/*.h declares;
the ctypes binding covers exactly those symbols; the product path has no CPU fallback."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "automoe_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amoe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/automoe_b200.h but not exported"


def test_ctypes_binding_matches_header(built_lib):
    from automoe_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == header_symbols()
    lib = _cabi.lib()
    assert lib.amoe_abi_version() == 1


def test_header_arity_matches_binding():
    """argument counts in the header == argtypes in the ctypes table."""
    from automoe_b200 import _cabi
    text = (ROOT / "include" / "automoe_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in _cabi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), (name, n, len(args))


def test_tc_supported_shapes(built_lib):
    from automoe_b200 import _cabi
    f = _cabi.lib().amoe_conv2d_tc_supported
    assert f(64, 64, 64, 64, 1, 1) == 1        # layer1
    assert f(64, 64, 64, 128, 2, 2) == 1       # layer2.0.conv1
    assert f(8, 8, 512, 512, 1, 1) == 1        # layer4
    assert f(256, 256, 4, 64, 2, 2) == 0       # stem: Cin=4 -> SIMT kernel
    assert f(45, 80, 256, 512, 2, 2) == 0      # odd height cannot use the parity view


def test_no_cpu_fallback():
    from automoe_b200 import _cabi
    from automoe_b200.models.automoe import create_automoe_model
    from oracle import synth
    m = create_automoe_model(synth.CONFIG_3EXPERT, "cpu").eval()
    batch = synth.synth_batch(1, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(batch)
    with pytest.raises(_cabi.AmoeError):
        _cabi.ctx(torch.device("cpu"))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from automoe_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_cabi.AmoeError, match="no fallback"):
        _cabi.lib()
