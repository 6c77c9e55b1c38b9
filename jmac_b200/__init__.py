"""jmac_b200 -- B200-native (sm_100a) implementation of JMAC's data-parallel hot path, a drop-in
behind the module interfaces of the reference's src/jmac_model.py and src/validate.py."""
from . import _lib  # noqa: F401
from .model import JMAC, BaseModel, RelationAwareLayer, get_param  # noqa: F401
from .validate import CompletionEvaluator  # noqa: F401

__all__ = ["JMAC", "BaseModel", "RelationAwareLayer", "CompletionEvaluator", "get_param"]
