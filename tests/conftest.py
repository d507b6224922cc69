"""pytest configuration: ``gpu`` marker + shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
  try:
    import torch
    has_gpu = torch.cuda.is_available()
  except Exception:  # pylint: disable=broad-except
    has_gpu = False
  if has_gpu:
    return
  skip = pytest.mark.skip(reason="no CUDA device in this container")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


def load_golden(name):
  return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_roots():
  return load_golden("roots.npz")


@pytest.fixture(scope="session")
def golden_roots_f64():
  return load_golden("roots_f64.npz")


@pytest.fixture(scope="session")
def golden_quant():
  return load_golden("quant.npz")


@pytest.fixture(scope="session")
def golden_fd():
  return load_golden("fd.npz")


@pytest.fixture(scope="session")
def golden_optimizer():
  return load_golden("optimizer.npz")


@pytest.fixture(scope="session")
def golden_shapes():
  return load_golden("shapes.npz")
