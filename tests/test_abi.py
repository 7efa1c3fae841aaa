"""The C-ABI library loads and exports every symbol include/bpgpu.h declares
(no compute call without a GPU), and fails loudly when no CUDA device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bpgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bpg_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from mpc_bulletproof_b200 import _lib

    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/bpgpu.h but not exported"
    # and the Python binding covers the same set
    assert sorted(_lib.exported_symbols()) == names


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from mpc_bulletproof_b200 import BpgError, Context
    from mpc_bulletproof_b200._lib import BPG_ERR_CUDA

    with pytest.raises(BpgError) as e:
        Context(0)
    assert e.value.code == BPG_ERR_CUDA


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "mpc_bulletproof_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "bp_oracle" not in txt, f


def test_rust_sys_crate_names_exist():
    """rust/bpgpu-sys declares a subset of include/bpgpu.h: every `pub fn bpg_*` it names must be
    an exported symbol of the library (the crate cannot be compiled here: no Rust toolchain)."""
    import re

    from mpc_bulletproof_b200 import _lib

    src = open(os.path.join(ROOT, "rust", "bpgpu-sys", "src", "lib.rs")).read()
    names = set(re.findall(r"pub fn (bpg_\w+)\s*\(", src))
    assert len(names) > 30
    missing = names - set(_lib.exported_symbols())
    assert not missing, f"declared in the Rust bindings but not exported: {sorted(missing)}"
    hdr = open(os.path.join(ROOT, "include", "bpgpu.h")).read()
    for code in re.findall(r"pub const (BPG_\w+): c_int = (-?\d+);", src):
        assert re.search(rf"#define {code[0]} {code[1]}\b", hdr), code
