import os
import sys

import pytest

# Virtual ranks (several ranks sharing one device in tests/test_gpu_multi.py) need their streams on distinct hardware
# queues, or a spinning consumer kernel could sit in front of its producer. Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ora():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """the reference's own sources built against the sequential StarPU stand-in (absent => skip)"""
    from oracle.oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    r = Reference()
    r.set_threads(1)
    return r


@pytest.fixture(scope="session")
def sn():
    import starneig_b200
    return starneig_b200


@pytest.fixture()
def node(sn):
    """initialised node with one GPU, finalised afterwards"""
    sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
    yield sn
    sn.starneig_node_finalize()


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("prand"))
