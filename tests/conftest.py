import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Backend:
    """What the ports of the reference's tests are written against: a factory for Ruckig-like objects."""

    def __init__(self, name):
        self.name = name

    def ruckig(self, dofs, delta_time, handler):
        if self.name == "gpu":
            from rsruckig_b200 import Ruckig
            return Ruckig(dofs, delta_time, handler)
        from oracle.oracle import OracleRuckig
        return OracleRuckig(dofs, delta_time, handler, flavour=self.name.split("-")[1])


@pytest.fixture(params=["oracle-glibc", "oracle-portable", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "gpu":
        import torch
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
    return Backend(request.param)
