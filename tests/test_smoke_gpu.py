"""GPU: the driver's smoke entry point (`__graft_entry__.smoke`) is part of the suite, so a change that breaks it shows up here."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_smoke_entry_point(capsys):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as g
    g.smoke()
    assert "smoke ok" in capsys.readouterr().out
