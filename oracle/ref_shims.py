"""Import the UNMODIFIED reference from /root/reference in the build container.  TEST
INFRASTRUCTURE ONLY (used by oracle/gen_golden.py and the optional `needs_reference` tests; the GPU
box has no /root/reference, nothing at run time may depend on this module).

Mechanical shims, none of which touches hot-path arithmetic except (1):
  1. ``torch_scatter`` is not installed -> a pure-torch module with ``scatter_add``, ``scatter``
     and ``scatter_softmax`` (written with explicit per-group loops-free torch ops, and checked
     against naive per-group Python loops in tests/test_oracle_golden.py).
  2. ``inspect.getargspec`` (removed in py3.11)  -> alias of ``getfullargspec(f)[:4]``.
  3. ``np.int/np.bool/np.float`` (removed in numpy 1.24) -> builtins.
``train.py`` is a SyntaxError as shipped (line 101) and is never imported; the two functions of it
on the hot path (``compute_alignment_quality`` :231-259) are exec'd from source with the one
character quote fix.
"""
import inspect
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("JMAC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src"))


def _scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device) if out is None else out
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return res.scatter_add(0, idx, src)


def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert reduce in ("sum", "add")
    return _scatter_add(src, index, dim, out, dim_size)


def _scatter_softmax(src, index, dim=0):
    assert dim == 0
    n = int(index.max()) + 1 if index.numel() else 0
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    mx = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    ex = (src - mx.gather(0, idx)).exp()
    den = torch.zeros_like(mx).scatter_add(0, idx, ex)
    return ex / den.gather(0, idx)


def install():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    if "torch_scatter" not in sys.modules:
        m = types.ModuleType("torch_scatter")
        m.scatter_add, m.scatter, m.scatter_softmax = _scatter_add, _scatter, _scatter_softmax
        sys.modules["torch_scatter"] = m
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = lambda f: inspect.getfullargspec(f)[:4]
    for name, typ in (("int", int), ("bool", bool), ("float", float)):
        if name not in np.__dict__:
            setattr(np, name, typ)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def load_train_functions():
    """exec the hot-path helpers of train.py (which does not parse as shipped)."""
    install()
    src = open(os.path.join(REF, "train.py")).read()
    src = src.replace("'--no_name_info, action=", "'--no_name_info', action=")
    start = src.index("def test_alignment_")
    end = src.index("def completion_data_processing")
    from modules.finding.evaluation import test
    from modules.utils.util import get_neg
    ns = {"torch": torch, "np": np, "test": test, "get_neg": get_neg}
    exec(compile(src[start:end], "train.py[test_alignment_ .. process_input_data]", "exec"), ns)
    return ns
