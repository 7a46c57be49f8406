import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
  sys.path.insert(0, str(ROOT))
if str(ROOT / 'tests') not in sys.path:
  sys.path.insert(0, str(ROOT / 'tests'))


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


@pytest.fixture(autouse=True)
def _fresh_uuids():
  from embodied_b200 import elements
  elements.UUID.reset(debug=True)
  yield
  elements.UUID.reset(debug=False)
