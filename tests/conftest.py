import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.pjob.gz")))


class _Golden(dict):
    """name -> list[FlatJob], loaded on first use (the BASELINE-size streams are large)."""

    def __missing__(self, name):
        from pagan2_msa_b200 import jobio

        path = os.path.join(ROOT, "tests", "golden", name + ".pjob.gz")
        if not os.path.exists(path):
            raise KeyError(name)
        self[name] = jobio.load_jobs(path)
        return self[name]


@pytest.fixture(scope="session")
def golden():
    return _Golden()
