"""CPU-side checks of the drop-in boundary: libmppi_b200.so loads and exports every symbol that
include/mppi_b200.h declares, the ctypes table covers the header, and without a GPU the library
fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from mppi_tf_b200 import _capi
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mppi_b200.h but not exported"
        assert n in _capi.SYMBOLS, f"{n} missing from the ctypes table"
    assert set(_capi.SYMBOLS) == set(names)
    assert lib.mppi_version().startswith(b"mppi_b200")


def test_config_struct_layout():
    """mppi_config_default fills the ctypes mirror the way the header documents."""
    from mppi_tf_b200 import _capi
    lib = _capi.load()
    cfg = _capi.MppiConfig()
    lib.mppi_config_default(C.byref(cfg), 1024, 20, 0.1, 1.5, 2, 1)
    assert (cfg.k, cfg.tau, cfg.s_dim, cfg.a_dim) == (1024, 20, 2, 1)
    assert abs(cfg.dt - 0.1) < 1e-7 and abs(cfg.mass - 1.5) < 1e-7 and cfg.lambda_ == 1.0
    assert cfg.seed == 1 and cfg.device == -1 and cfg.world == 1 and cfg.n_controllers == 1
    assert not cfg.sigma and not cfg.goal and not cfg.q and not cfg.stream


def test_host_only_stages_need_no_gpu():
    """blockDiag / getNew / shift / prepareAction are host-side constant plumbing in the new build
     and work anywhere."""
    from mppi_tf_b200 import blockDiag
    from tests.golden import kats
    dt = np.float32(kats.UTILE_DT)
    got = blockDiag(np.array([[1, dt], [0, 1]], np.float32), 3)
    np.testing.assert_array_equal(got, kats.blockdiag_expected(3)[0])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mppi_tf_b200 import ControllerBase, ModelBase, MppiError, _capi
    with pytest.raises(MppiError) as e:
        ControllerBase(1024, 20, 0.1, 1.0, 2, 1)
    assert e.value.code == _capi.MPPI_ERR_CUDA
    with pytest.raises(MppiError) as e:
        ModelBase(1.0, 0.01, 2, 1).predict([[0, 0]], [[1]])
    assert e.value.code == _capi.MPPI_ERR_CUDA


def test_product_never_imports_the_oracle():
    """No product source imports, includes, links or dlopens anything under oracle/."""
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|#\s*include\s*[<\"][^>\"]*oracle|"
                     r"libmppi_oracle|dlopen\([^)]*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mppi_tf_b200")):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), os.path.join(dirpath, f)
