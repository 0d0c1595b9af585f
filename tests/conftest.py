import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["b20", "b30", "b40", "b50", "b60", "b70", "b80", "b90", "b100", "b200", "b512", "b1024"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def nb():
    """The product package (ctypes mirror of the C ABI)."""
    return importlib.import_module("nthu_ipc_nbody-simulation_b200")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding

    return oracle_binding


def case_path(case, ext="in"):
    return os.path.join(GOLDEN, "testcases", "%s.%s" % (case, ext))


def golden_lines(case):
    a, b, c = open(case_path(case, "out")).read().split("\n")[:3]
    dev, cost = c.split()
    return dict(min_dist=float(a), hit_time_step=int(b), gravity_device_id=int(dev), missile_cost=float(cost),
                text=open(case_path(case, "out")).read())
