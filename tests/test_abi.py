"""The C-ABI library: builds for sm_100a, loads, exports every symbol include/pagan2_b200.h declares,
struct layouts agree with the ctypes mirror, and -- without a GPU -- refuses to compute."""
import ctypes as C
import os
import re

import pytest

from pagan2_msa_b200 import abi, engine

HEADER = os.path.join(abi.REPO_ROOT, "include", "pagan2_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pg2_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()
    return engine.load_library()


def test_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "libpagan2_b200.so does not export %s" % n


def test_abi_version(lib):
    assert lib.pg2_abi_version() == abi.PG2_ABI_VERSION


def test_struct_sizes_match_header():
    # sizes computed by hand from the header's field lists (LP64)
    assert C.sizeof(abi.Graph) == 8 + 5 * 8
    assert C.sizeof(abi.ModelDesc) == 8 + 8 + 5 * 4 + 4
    assert C.sizeof(abi.Job) == 2 * C.sizeof(abi.Graph) + 8 + 16
    assert C.sizeof(abi.Result) == 40
    assert C.sizeof(abi.Step) == 32
    assert C.sizeof(abi.Stats) == 4 * 8 + 4 * 8 + 4 * 4 + 8 + 7 * 4 + 4  # (7 trailing int32 + padding to the double's alignment)


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for boxes without one")
    ctx = C.c_void_p()
    rc = lib.pg2_ctx_create(0, C.byref(ctx))
    assert rc == abi.PG2_ERR_NO_DEVICE
    assert b"no CPU" in lib.pg2_last_error()
    with pytest.raises(engine.Pg2Error):
        engine.Engine(0)


def test_product_library_does_not_link_the_oracle(lib):
    import subprocess

    out = subprocess.check_output(["nm", "-D", "--defined-only", engine.LIB_PATH]).decode()
    assert "pg2o_" not in out and "pagan2_ref" not in out
    ldd = subprocess.check_output(["ldd", engine.LIB_PATH]).decode()
    assert "oracle" not in ldd and "pagan2ref" not in ldd
