"""ctypes binding of libjmac_b200.so (the C ABI declared in include/jmac_b200.h).

There is no fallback: if the library has not been built, or no CUDA device is present when a kernel
is requested, the call raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C jmac_b200/csrc``.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjmac_b200.so")
_lib = None

P = c_void_p  # every device pointer crosses the ABI as a plain address

_SIGNATURES = {
    "jmac_version": (c_char_p, []),
    "jmac_last_error": (c_char_p, []),
    "jmac_launch_count": (ctypes.c_uint64, []),
    "jmac_status_poll": (c_int, []),
    "jmac_set_sm_reserve": (c_int, [c_int]),
    "jmac_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "jmac_peer_free": (c_int, [P]),
    "jmac_ipc_get_handle": (c_int, [P, ctypes.c_char_p]),
    "jmac_ipc_open": (c_int, [ctypes.c_char_p, POINTER(c_void_p)]),
    "jmac_ipc_close": (c_int, [P]),
    "jmac_memcpy_async": (c_int, [P, P, c_size_t, P]),
    "jmac_stream_write_u32": (c_int, [P, ctypes.c_uint32, P]),
    "jmac_stream_wait_geq_u32": (c_int, [P, ctypes.c_uint32, P]),
    "jmac_peer_sum": (c_int, [P, P, c_int, c_int64, c_int64, c_int, P]),
    "jmac_plan_create": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_int, P, POINTER(c_void_p)]),
    "jmac_plan_create_bipartite": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int, P, POINTER(c_void_p)]),
    "jmac_plan_destroy": (c_int, [P]),
    "jmac_plan_info": (c_int, [P, POINTER(c_int64)]),
    "jmac_plan_export": (c_int, [P, P, P, P, P, P, P]),
    "jmac_edge_fwd": (c_int, [P, P, P, P, P, P, P, c_float, c_int, P, P, P]),
    "jmac_edge_fwd_fused": (c_int, [P, P, P, P, P, P, P, c_float, c_int, P, P, P, P, P, P]),
    "jmac_edge_record_bytes": (c_size_t, [c_int]),
    "jmac_edge_bwd_dst": (c_int, [P, P, P, P, P, c_float, c_int, P, P, P, P, P, P, P, P]),
    "jmac_edge_bwd_src": (c_int, [P, P, P, P, c_float, c_int, P, P, P, P, P]),
    "jmac_edge_bwd_rel": (c_int, [P, P, P, P, c_float, c_int, P, P, P, P, P]),
    "jmac_bn_stats": (c_int, [P, c_int64, c_int, P, P]),
    "jmac_bn_tanh_apply": (c_int, [P, c_int64, c_int, P, c_int64, P, P, P, P, c_float, c_float, c_int, P, P, P, P]),
    "jmac_bn_tanh_bwd_stats": (c_int, [P, P, P, c_int64, c_int, P, P, P, P]),
    "jmac_bn_tanh_bwd_apply": (c_int, [P, P, P, c_int64, c_int, P, P, P, P, c_int64, c_int, P, P, P, P]),
    "jmac_row_normalize_fwd": (c_int, [P, c_int64, c_int, c_float, P, P, P]),
    "jmac_row_normalize_bwd": (c_int, [P, P, P, c_int64, c_int, P, P]),
    "jmac_split_tf32": (c_int, [P, c_int64, c_int64, c_int64, c_int, P, P, P]),
    "jmac_gemm_tf32x3": (c_int, [P, c_int64, P, P, c_int64, P, c_int64, c_int64, c_int, c_int, P]),
    "jmac_gemm_tf32x3_add": (c_int, [P, c_int64, P, P, c_int64, P, c_int64, c_int64, c_int, c_int, P, P, c_int64, P]),
    "jmac_gemm_tn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "jmac_gemm_tn_tf32x3": (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, P, P]),
    "jmac_transe_record_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "jmac_transe_margin_fwd": (c_int, [P, P, c_int, c_int, P, P, P, c_int64, c_int64, c_int64, c_int64, c_float, P, P, P, P]),
    "jmac_transe_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "jmac_transe_margin_bwd": (c_int, [c_int, c_int, P, P, P, c_int64, c_int64, c_int64, c_int64, P, P, P, P, P, c_size_t, P]),
    "jmac_sample_negatives": (c_int, [P, c_int64, P, P, c_int64, c_int64, c_int64, c_int, ctypes.c_uint64, P, P]),
    "jmac_pair_cos_fwd": (c_int, [P, P, c_int, P, P, c_int64, c_int64, c_int64, P, P]),
    "jmac_pair_cos_bwd": (c_int, [P, P, c_int, P, P, c_int64, c_int64, c_int64, P, P, P, P]),
    "jmac_l1_dist": (c_int, [P, P, c_int, c_int, P, P, c_int64, c_int64, c_int64, c_int, P, P]),
    "jmac_sm_count": (c_int, []),
    "jmac_sim_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "jmac_sim_grid": (c_int, [c_int64, c_int64, c_int]),
    "jmac_to_bf16_rows": (c_int, [P, c_int, P, c_int64, c_int, P, c_int, P]),
    "jmac_to_tf32_rows": (c_int, [P, c_int, P, c_int64, c_int, P, P, c_int, P]),
    "jmac_sim_topk": (c_int, [P, P, c_int64, c_int64, c_int, c_int, c_float, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "jmac_sim_topk_f32": (c_int, [P, P, P, P, c_int64, c_int64, c_int, c_int, c_float, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "jmac_csls_gold": (c_int, [c_float, P, P, P, P, c_int64, P, P]),
    "jmac_sim_softmax_stats": (c_int, [P, P, c_int64, c_int64, c_int, c_float, P, P, c_float, P, P, P, P, P, c_size_t, P]),
    "jmac_sim_softmax_stats_f32": (c_int, [P, P, P, P, c_int64, c_int64, c_int, c_float, P, P, c_float, P, P, P, P, P, c_size_t, P]),
    "jmac_l1_rank": (c_int, [P, P, c_int, c_int, P, P, P, c_int64, c_int64, c_int64, c_int, P, P, P, P]),
}

EXPORTS = tuple(_SIGNATURES)


class JmacError(RuntimeError):
    pass


def load():
    """Load the shared library once; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JmacError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run "
            "`python -c \"import __graft_entry__ as g; g.build()\"` (needs nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().jmac_last_error().decode(errors="replace")
        raise JmacError(f"{what or 'libjmac_b200'} failed with code {rc}: {msg}")


_STATUS_WHAT = {0: "batch_h / batch_r / batch_t of the completion scorer", 1: "an index list of the pair-cosine (alignment) loss",
                2: "e_index / r_index / target / filter list of link prediction",
                3: "the negative sampler ((h, r) out of range, or fewer than k entities that are not known tails)"}


def raise_if_flagged(sync: bool = False):
    """IndexError if a kernel met an index outside its table since the last poll (the reference raises IndexError on the
    CPU and a device-side assert on CUDA; here the kernel substitutes row 0, raises a flag in mapped host memory and
    carries on, so nothing is read or written out of bounds).  Called without synchronising at the entry of every
    index-consuming op -- an error surfaces at the next such call -- and with ``sync=True`` by ``jmac_b200.check_errors()``."""
    lib = load()
    if sync:
        import torch
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    bits = lib.jmac_status_poll()
    if bits:
        what = "; ".join(v for k, v in _STATUS_WHAT.items() if bits & (1 << k))
        raise IndexError(f"index out of range in {what} (detected on the device)")


def ptr(t):
    """Device address of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
