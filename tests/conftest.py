import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pk_text():
    with open(os.path.join(ROOT, "tests", "golden", "powerspec.txt")) as f:
        return f.read()


@pytest.fixture(scope="session")
def ref_mod():
    """The compiled reference (oracle/_ref/libfastpm_ref.so); built in-container, shipped to the GPU box."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libfastpm_ref.so not built (run `make -C oracle ref` where /root/reference exists)")
    return ref
