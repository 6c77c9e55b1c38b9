"""CPU checks of the boundary: the C-ABI library loads and exports every symbol that
include/jmac_b200.h declares (no compute calls without a GPU), the ctypes table covers them all, and
the product package neither imports the oracle nor has a CPU path."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "jmac_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jmac_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from jmac_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 27
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == syms, "ctypes signature table and header disagree"
    assert _lib.load().jmac_version().startswith(b"jmac_b200")


def test_argument_errors_do_not_need_a_gpu():
    from jmac_b200 import _lib
    lib = _lib.load()
    assert lib.jmac_edge_record_bytes(256) == 44 and lib.jmac_edge_record_bytes(300) == 76 and lib.jmac_edge_record_bytes(7) == 0
    assert lib.jmac_plan_info(None, None) == -1                       # JMAC_EINVAL
    assert b"bad argument" in lib.jmac_last_error()
    assert lib.jmac_plan_destroy(None) == 0


def test_product_never_imports_the_oracle_and_has_no_cpu_path():
    pkg = os.path.join(ROOT, "jmac_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn
    if not torch.cuda.is_available():
        from types import SimpleNamespace
        from jmac_b200 import RelationAwareLayer, _lib
        layer = RelationAwareLayer(8, 8, rel_dim=8, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub"))
        with pytest.raises(_lib.JmacError):                           # fails loudly on CPU tensors
            layer(torch.randn(4, 8), torch.randn(2, 8), torch.zeros(2, 3, dtype=torch.long), torch.zeros(3, dtype=torch.long))


def test_index_normalisation_quirks():
    import numpy as np
    from jmac_b200.functional import as_index
    for x in ([1, 2, 3], np.array([1.0, 2.0, 3.0]), np.array([1, 2, 3]), torch.tensor([1, 2, 3], dtype=torch.int32)):
        t = as_index(x, "cpu")
        assert t.dtype == torch.int64 and t.tolist() == [1, 2, 3]
