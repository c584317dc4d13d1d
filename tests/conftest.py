import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _ensure(path, cmd, cwd):
    if not os.path.exists(path):
        subprocess.check_call(cmd, cwd=cwd)


@pytest.fixture(scope="session", autouse=True)
def _built_helpers():
    """The oracle / corpus / host-model helpers are tiny C files: (re)build them if missing."""
    _ensure(os.path.join(ROOT, "oracle", "libflate_oracle.so"), ["make", "-s", "-C", "oracle"], ROOT)
    _ensure(os.path.join(ROOT, "tools", "libfb_corpus.so"), ["make", "-s", "-C", "tools"], ROOT)
    yield
