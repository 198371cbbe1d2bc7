import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def random_state(n: int, seed: int = 7) -> np.ndarray:
    """Seeded random normalised state (normal re/im), the start state of most parity cases."""
    rng = np.random.Generator(np.random.MT19937(seed))
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    return v.astype(np.complex128)


def draws(count: int, seed: int = 42) -> np.ndarray:
    """Injected measurement draws in (0, 1], every one a multiple of 2^-53 like a real
    `1. - uniformZeroOne(rng)` (see oracle/ref_driver.cpp::injectDraw)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    k = rng.integers(1, 1 << 53, size=count, dtype=np.uint64)
    return k.astype(np.float64) / float(1 << 53)


@pytest.fixture(scope="session")
def have_ref():
    import oracle

    oracle.build_ref()
    return oracle.ref_available("sse2")


@pytest.fixture(autouse=True)
def _gpu_tests_check_against_the_reference_itself(request):
    """Every `-m gpu` parity test compares with the compiled reference (oracle/_ref, kind == "reference"), never
    silently with the C port: oracle.best_oracle raises under QCSIM_REQUIRE_REF when the reference build is missing."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import oracle

    assert oracle.ref_available("sse2"), "oracle/_ref/libqcsim_ref_sse2.so is missing: the GPU parity tests need the compiled reference"
    old = os.environ.get("QCSIM_REQUIRE_REF")
    os.environ["QCSIM_REQUIRE_REF"] = "1"
    try:
        yield
    finally:
        if old is None:
            os.environ.pop("QCSIM_REQUIRE_REF", None)
        else:
            os.environ["QCSIM_REQUIRE_REF"] = old
