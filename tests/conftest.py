import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(GOLDEN, "assets")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def voldata_ref():
    """The compiled, unmodified reference voldata (only where oracle/_ref was built)."""
    from oracle.binding import VoldataRef
    if not VoldataRef.available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return VoldataRef()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "voldata_golden.npz"))


@pytest.fixture(scope="session")
def smoke_golden():
    return np.load(os.path.join(GOLDEN, "smoke_brick_golden.npz"))


@pytest.fixture(scope="session")
def smoke_grid():
    from volren_b200 import formats
    return formats.load_brick(os.path.join(ASSETS, "smoke.brick"))


@pytest.fixture(scope="session")
def env_rgb():
    from volren_b200 import formats
    return formats.load_hdr(os.path.join(ASSETS, "table_mountain_2_puresky_1k.hdr"))


@pytest.fixture(scope="session")
def env_pyramid(oracle, env_rgb):
    return oracle.env_build(env_rgb)


@pytest.fixture(scope="session")
def lut_raw():
    from volren_b200 import formats
    return formats.load_lut_txt(os.path.join(ASSETS, "lut.txt"))


@pytest.fixture(scope="session")
def ctx():
    """One vrb_ctx on cuda:0 through the C ABI (fails loudly without the CUDA library / a GPU)."""
    from volren_b200 import Context
    c = Context(0)
    yield c
    c.close()
