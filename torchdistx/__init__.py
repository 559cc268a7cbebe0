"""Import-compatibility shim: ``import torchdistx.fake`` / ``torchdistx.deferred_init`` resolve to
the Blackwell-native implementation in :mod:`torchdistx_b200` (only the deferred-init path of the
reference is in scope)."""
from torchdistx_b200 import __version__  # noqa: F401
