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


@pytest.fixture(scope="session")
def golden():
    from pagan2_msa_b200 import jobio

    return {os.path.basename(p).split(".")[0]: jobio.load_jobs(p) for p in GOLDEN}
