"""B200-native successive-orders-of-scattering solver: host side (Python mirror of the reference
interface for the hot path) over the C-ABI library libsosgpu.so (hand-written CUDA for sm_100a).

The package directory name contains a hyphen; import it with
    importlib.import_module("radiativetransfer-sos_b200")
"""
from . import formats, synth  # noqa: F401
