"""jmac_b200 -- B200-native (sm_100a) implementation of JMAC's data-parallel hot path, a drop-in
behind the module interfaces of the reference's src/jmac_model.py and src/validate.py."""
from . import _lib  # noqa: F401
from .model import JMAC, BaseModel, RelationAwareLayer, get_param  # noqa: F401
from .validate import CompletionEvaluator  # noqa: F401



def check_errors():
    """Synchronise the device and raise IndexError if any kernel met an out-of-range index (see _lib.raise_if_flagged)."""
    _lib.raise_if_flagged(sync=True)


__all__ = ["JMAC", "BaseModel", "RelationAwareLayer", "CompletionEvaluator", "get_param", "check_errors"]
