import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def kabc():
    import kissabc_jl_b200 as k
    return k


@pytest.fixture(scope="session")
def ctx(kabc):
    """One device context for the whole GPU session.  Fails loudly (no fallback) if there is no GPU."""
    c = kabc.Context(device=0, seed=0x4B49535341424300)
    yield c
    c.close()


SEED = 0x4B49535341424300


def readme_prior(O):
    return O.make_priors([("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100)])


def readme_model(O, n=1000):
    return O.make_model(O.NORMAL_MEANSTD, n, target=(2.0, 0.04), param=(50.0,))
