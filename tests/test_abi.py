"""CPU tests of the C-ABI library: it loads, exports every symbol include/cngp.h declares, its POD structs have the
layout the ctypes mirror assumes, the GPU-free entry points (kernel-expression parser, defaults) work, and creating a
context without a CUDA device fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from corenav_gp_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cngp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cngp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cngp.h but not exported by libcngp.so"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in corenav_gp_b200/_lib.py"


def test_version_and_struct_layout():
    lib = L.load()
    assert lib.cngp_version() == 100
    assert C.sizeof(L.Kernel) == 4 * (1 + L.MAX_OPS + 1)
    assert C.sizeof(L.Config) == 4 + 4 + 8 + 32
    assert C.sizeof(L.StopConfig) == 6 * 8 + 2 * 4 + 6 * 8


def test_stop_config_defaults_are_the_reference_constants():
    c = L.StopConfig()
    L.load().cngp_default_stop_config(C.byref(c))
    # gp_predictor.cpp:73-88,102 ; core_navigation/config/init_params.yaml:9-16
    assert (c.v_nom, c.floor_a, c.floor_b, c.track, c.scale, c.thresh, c.ratio) == (0.8, 0.03, 0.05, 0.685, 25.0, 3.0, 5)
    assert c.fix_h_packing == 0
    assert list(c.init_llh) == [0.693457963620326, -1.39498384275845, 334.993517334743]
    assert list(c.init_ecef) == [859153.0153, -4836303.7266, 4055378.501]


@pytest.mark.parametrize("text,ops,nparams", [
    ("rbf", [1], 2),
    ("rbf*brownian", [1, 6, 17], 3),                     # the deployed kernel, gp_slip_node.py:31
    ("se+periodic", [1, 5, 16], 5),
    ("RBF + Linear + Brownian", [1, 7, 16, 6, 16], 4),   # gp_slip_node.py:32-34 candidates
    ("rbf*linear", [1, 7, 17], 3),
    ("mat32+mat52", [2, 3, 16], 4),
    ("(rbf+linear)*brownian+white", [1, 7, 16, 6, 17, 9, 16], 5),
    ("rq+rbf*per", [4, 1, 5, 17, 16], 8),
])
def test_kernel_parse(text, ops, nparams):
    k = L.Kernel()
    assert L.load().cngp_kernel_parse(text.encode(), C.byref(k)) == 0
    assert list(k.ops[:k.n_ops]) == ops and k.n_params == nparams


@pytest.mark.parametrize("text", ["", "rbf+", "foo", "(rbf", "rbf)", "rbf**linear", "+"])
def test_kernel_parse_rejects(text):
    k = L.Kernel()
    assert L.load().cngp_kernel_parse(text.encode(), C.byref(k)) != 0


def test_parser_agrees_with_oracle_parser():
    from oracle import gp_oracle as go
    for text in ["rbf*brownian", "rbf+stdperiodic", "mat52*linear+white", "ratquad+stdperiodic*rbf",
                 "(rbf+linear)*brownian+white", "bias+rbf*(mat32+white)"]:
        k = L.Kernel()
        assert L.load().cngp_kernel_parse(text.encode(), C.byref(k)) == 0
        e = go.KernelExpr(text)
        assert list(k.ops[:k.n_ops]) == e.program and k.n_params == e.n_params


def test_kernel_finalize_validates_programs():
    lib = L.load()
    k = L.Kernel()
    k.n_ops = 2
    k.ops[0], k.ops[1] = 1, 16            # "rbf +" : operator without two operands
    assert lib.cngp_kernel_finalize(C.byref(k)) != 0
    k.n_ops = 3
    k.ops[0], k.ops[1], k.ops[2] = 1, 6, 17
    assert lib.cngp_kernel_finalize(C.byref(k)) == 0 and k.n_params == 3


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = L.load()
    h = C.c_void_p()
    rc = lib.cngp_create(None, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.cngp_last_error(None)
    from corenav_gp_b200.api import CngpError, GpContext
    with pytest.raises(CngpError):
        GpContext(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under corenav_gp_b200/ may reference it."""
    pkg = os.path.join(ROOT, "corenav_gp_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(root, f)
                assert "oracle/" not in src.replace("the C oracle", ""), os.path.join(root, f)
