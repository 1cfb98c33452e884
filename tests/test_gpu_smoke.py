"""GPU tier: the driver's smoke() hook must keep passing (it is what runs first on a fresh box)."""
import pytest

pytestmark = pytest.mark.gpu


def test_smoke_hook():
    import __graft_entry__ as g
    g.smoke()
