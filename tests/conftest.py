import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port_oracle():
    """The plain-C restatement (oracle/vega_port.c), built on demand with gcc."""
    from oracle import pyoracle

    if not pyoracle.available("port"):
        pyoracle.build("port")
    return pyoracle


@pytest.fixture(scope="session")
def ref_oracle():
    """The unmodified reference compiled in place (oracle/_ref); only where it has been built."""
    from oracle import pyoracle

    if not pyoracle.available("ref"):
        if os.path.isdir("/root/reference"):
            pyoracle.build("ref")
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return pyoracle


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
