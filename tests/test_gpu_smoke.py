"""GPU smoke: the driver-facing entry points run on a real device."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.gpu
def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
