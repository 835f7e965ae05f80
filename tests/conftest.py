import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CACHE = os.path.join(ROOT, ".table_cache")
# On a GPU box the fixtures carry the tables the reference main builds: the ion tabulated over its WHOLE energy grid
# (Analytical_IMFPs.f90:2242-2510), integrals evaluated by the GPU table builder.  In the GPU-less container the CPU tests
# tabulate the ion only around its own energy (the Monte-Carlo never looks elsewhere) to stay within minutes.
ON_GPU_BOX = os.path.exists("/dev/nvidiactl")


def table_options():
    return dict(shi_window_only=False, evaluator="gpu") if ON_GPU_BOX else dict(shi_window_only=True)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


def _case(tmp_path_factory, name, **kw):
    import trekis3_b200 as tk
    d = tk.make_run_dir(str(tmp_path_factory.mktemp("run_" + name)), name, **kw)
    c = tk.Case.load(d)
    c.build_tables(cache_dir=CACHE, **table_options())
    return c


@pytest.fixture(scope="session")
def case_c1(tmp_path_factory, _built):
    return _case(tmp_path_factory, "C1")


@pytest.fixture(scope="session")
def case_c2(tmp_path_factory, _built):
    return _case(tmp_path_factory, "C2")


@pytest.fixture(scope="session")
def case_c3(tmp_path_factory, _built):
    return _case(tmp_path_factory, "C3")


@pytest.fixture(scope="session")
def case_c4(tmp_path_factory, _built):
    return _case(tmp_path_factory, "C4")
