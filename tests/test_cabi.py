"""The C-ABI library builds, loads, and exports every symbol include/aero_b200.h declares (no compute
calls: this runs without a GPU)."""
import os
import re

import pytest

from aero_b200 import build as build_mod
from aero_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_mod.build()
    return cabi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aero_b200.h")).read()
    return sorted(set(re.findall(r"\b(aero_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    names = declared_symbols()
    assert names, "no prototypes parsed"
    assert set(names) == set(cabi.SYMBOLS), set(names) ^ set(cabi.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aero_b200.h but not exported"


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "aero_b200.h")).read()
    assert "at::" not in text and "torch" not in text.replace("torch.", "").replace("PyTorch", "")


def test_version_and_error_channel(lib):
    assert lib.aero_abi_version() == 3
    # parameter validation happens before any device work: a null call must fail cleanly
    rc = lib.aero_stft_fwd(None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.aero_last_error()


def test_struct_sizes_match_header(lib):
    import ctypes
    assert ctypes.sizeof(cabi.StftParams) == 8 * 4 + 4 * 8 + 2 * 4
    assert ctypes.sizeof(cabi.IstftParams) == 8 * 4 + 4 * 8 + 2 * 4
    assert ctypes.sizeof(cabi.TapGemmParams) == 20 * 4 + 15 * 8 + 8
    assert ctypes.sizeof(cabi.NormActParams) == 11 * 4
    assert ctypes.sizeof(cabi.LstmParams) == 10 * 4
    assert ctypes.sizeof(cabi.AttnParams) == 7 * 4
    assert ctypes.sizeof(cabi.FtbLinParams) == 6 * 4 + 4 * 8


def test_oracle_is_not_imported_by_the_product():
    import subprocess, sys
    code = "import sys; import aero_b200, aero_b200.engine, aero_b200.spec, aero_b200.model; " \
           "assert not any(m.startswith('oracle') for m in sys.modules), 'product imports the oracle'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "aero_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("# oracle", "")
