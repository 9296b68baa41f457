import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # A fresh checkout has no libcesm_b200.so (build artefacts are git-ignored): compile it once for the test
    # session.  The package itself never does this -- a missing library is a hard error there.
    from cesm_emulator_b200 import build as _build
    if not _build.LIB_PATH.exists():
        try:
            _build.build()
        except Exception as exc:  # no nvcc: the tests that need the library will say so
            sys.stderr.write(f"[conftest] could not build {_build.LIB_PATH.name}: {exc}\n")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")
