import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
	"""The CPU checker (our C restatement); built on demand.  Test infrastructure only."""
	from oracle import pyoracle

	pyoracle.build()
	return pyoracle


@pytest.fixture(scope="session")
def built_lib():
	"""The product library, built in-tree if missing or stale (nvcc cross-compiles without a GPU)."""
	from cacaoengine_b200 import build as b

	b.build()
	from cacaoengine_b200 import _capi

	return _capi.lib()


@pytest.fixture(scope="session")
def golden():
	import numpy as np

	return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
