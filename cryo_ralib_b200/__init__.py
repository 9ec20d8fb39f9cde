"""cryo_ralib_b200 -- B200-native engine for cryo-EM 2D multi-reference and
reference-free alignment (the one hot path of phonchi/Cryo-RAlib), behind the
reference's ctypes boundary.  See DESIGN.md / INTEGRATION.md."""
from .lib import Engine, CraConfig, CraSearch, CraResult, load_library, LibraryMissing  # noqa: F401

__all__ = ["Engine", "CraConfig", "CraSearch", "CraResult", "load_library", "LibraryMissing"]
