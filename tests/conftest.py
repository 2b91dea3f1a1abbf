import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


EMUL_LIB = os.path.join(ROOT, "tests", "emul", "_build", "libfastpm_b200_emul.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    if os.environ.get("FASTPM_B200_TEST_EMUL"):
        # tests/test_cpu_full_emulation.py: the `gpu` cases of this process run on the CPU against the emulated build of the whole
        # library (tests/emul/emul_lib/: kernel sources on OS threads, fake CUDA runtime).  Test infrastructure only -- the product
        # loader (fastpm_b200/_lib.py) knows nothing about it.
        if not os.path.exists(EMUL_LIB):
            raise RuntimeError("FASTPM_B200_TEST_EMUL is set but %s is not built (python tests/emul/emul_lib/build.py)" % EMUL_LIB)
        from fastpm_b200 import _lib
        _lib.LIB_PATH = EMUL_LIB


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
