"""CPU (no GPU) checks of the drop-in boundary: libse_b200.so builds, loads, and exports exactly the
symbols include/se_abi.h declares; without a device every entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from spark_ensemble_b200 import _native, build
    build.build()
    return _native.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "se_abi.h")).read()
    return sorted(set(re.findall(r"^SE_API\s+[\w\s\*]+?\b(se_\w+)\(", src, flags=re.M)))


def test_header_symbols_all_exported(lib):
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in se_abi.h but not exported"


def test_python_prototypes_cover_header(lib):
    from spark_ensemble_b200 import _native
    assert sorted(_native.PROTOTYPES) == _declared()


def test_abi_version(lib):
    assert lib.se_abi_version() == 1


def test_no_cpu_fallback_without_device(lib):
    from spark_ensemble_b200 import _native
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.se_ctx_create(0, ctypes.byref(h))
    assert rc == _native.SE_ERR_CUDA
    assert "no CPU fallback" in _native.last_error()
    from spark_ensemble_b200.context import Context
    with pytest.raises(_native.NativeError):
        Context(0)


def test_brent_host_optimizer_matches_oracle(lib, oracle):
    """se_brent_minimize (product, C++) and the oracle's Brent (C) are separate restatements of
    commons-math3 BrentOptimizer: same iterates => same abscissa and evaluation count."""
    from spark_ensemble_b200 import _native as N
    import math
    fns = [lambda x: (x - 2.5) ** 2, lambda x: math.cosh(x - 0.3) + 0.1 * x,
           lambda x: abs(x - 7.0) + 0.01 * x * x, lambda x: -math.exp(-(x - 40) ** 2 / 50.0),
           lambda x: x]
    for fn in fns:
        cb = N.FN1(lambda x, _u, fn=fn: float(fn(x)))
        x, f, ne = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        rc = lib.se_brent_minimize(cb, None, 0.0, 100.0, 1.0, 1e-6, 1e-6, 100, ctypes.byref(x),
                                   ctypes.byref(f), ctypes.byref(ne))
        assert rc == 0
        xo, neo, st = oracle.brent(fn)
        assert st == 0 and ne.value == neo
        assert x.value == xo
        assert f.value == fn(xo)
    # MaxEval exceeded -> SE_ERR_OPT (TooManyEvaluationsException in the reference)
    cb = N.FN1(lambda x, _u: (x - 37.123) ** 2)
    rc = lib.se_brent_minimize(cb, None, 0.0, 100.0, 1.0, 1e-6, 1e-6, 3, None, None, None)
    assert rc == N.SE_ERR_OPT
    # constructor checks of BrentOptimizer
    assert lib.se_brent_minimize(cb, None, 0.0, 100.0, 1.0, 1e-20, 1e-6, 10, None, None, None) == N.SE_ERR_ARG
    assert lib.se_brent_minimize(cb, None, 0.0, 100.0, 1.0, 1e-6, 0.0, 10, None, None, None) == N.SE_ERR_ARG
