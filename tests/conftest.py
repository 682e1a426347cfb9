import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_dev():
    from oracle import oracle_device
    if not oracle_device.available():
        pytest.fail("oracle/_ref/liboracle_singleray.so missing: run `python oracle/build_ref.py` where /root/reference is mounted")
    d = oracle_device.open_oracle(num_threads=1)
    yield d


@pytest.fixture(scope="session")
def oracle_mt():
    from oracle import oracle_device
    d = oracle_device.open_oracle(num_threads=0)
    yield d


@pytest.fixture(scope="session")
def cuda_dev():
    from yulio_raytracer_b200 import Device
    d = Device.cuda(cfg="stats=0")      # raises if the CUDA library or the GPU is missing: no CPU fallback
    yield d
    d.close()
