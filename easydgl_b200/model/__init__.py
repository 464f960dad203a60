from .Base import Sequential, layernorm  # noqa: F401
from .CTSMA import CTSMA  # noqa: F401
from .EasyDGL import EasyDGL  # noqa: F401
