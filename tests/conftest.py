import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "ray-tracer-s8_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle

    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def rt():
    import rt_b200

    return rt_b200


@pytest.fixture(scope="session")
def ctx(rt):
    """One GPU context for the whole session; fails loudly (no fallback) when there is no device."""
    c = rt.Context(0)
    yield c
    c.close()


def frame_compare(a, b):
    """north_star tolerance: every 8-bit channel within +-1 LSB on >= 99.9 % of pixels, PSNR >= 50 dB."""
    import numpy as np

    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    frac_ok = float((d.reshape(-1, 3).max(axis=-1) <= 1).mean()) if d.size else 1.0
    mse = float((d.astype(np.float64) ** 2).mean()) if d.size else 0.0
    psnr = float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)
    return {"frac_within_1lsb": frac_ok, "psnr_db": psnr, "n_diff": int((d > 0).sum()), "max_diff": int(d.max()) if d.size else 0}


def assert_parity(a, b, exact=True):
    c = frame_compare(a, b)
    # the tolerance BASELINE.json's north_star states
    assert c["frac_within_1lsb"] >= 0.999 and c["psnr_db"] >= 50.0, c
    if exact:
        # the design goal of this implementation: bit-exact with the oracle
        assert c["n_diff"] == 0, c
    return c
