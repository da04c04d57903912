"""Loaders for the in-tree native libraries.  There is no fallback: a missing
library is an error that names the build command."""
from __future__ import annotations

import ctypes
import importlib.util
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "lib")
KERNELS_SO = os.path.join(LIB_DIR, "libgnn_b200.so")
EXT_SO = os.path.join(LIB_DIR, "spmm.so")

_lock = threading.Lock()
_ext = None
_cabi = None

_HINT = ("build it with `python -c 'import __graft_entry__ as g; g.build()'` (or `python gnn_b200/build.py`) "
         "from the repository root; there is no CPU or PyTorch fallback for this path")


def extension():
    """The pybind module named `spmm` (same name as the reference's JIT extension,
    reference custom_sparse_ops.py:8)."""
    global _ext
    with _lock:
        if _ext is None:
            if not (os.path.exists(EXT_SO) and os.path.exists(KERNELS_SO)):
                raise ImportError(f"gnn_b200 native extension not found at {EXT_SO}; {_HINT}")
            import torch  # noqa: F401  libtorch must be loaded first
            ctypes.CDLL(KERNELS_SO, mode=ctypes.RTLD_GLOBAL)
            spec = importlib.util.spec_from_file_location("spmm", EXT_SO)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _ext = mod
        return _ext


def cabi() -> ctypes.CDLL:
    """libgnn_b200.so through ctypes with the prototypes of include/gnn_b200.h."""
    global _cabi
    with _lock:
        if _cabi is None:
            if not os.path.exists(KERNELS_SO):
                raise ImportError(f"gnn_b200 kernel library not found at {KERNELS_SO}; {_HINT}")
            lib = ctypes.CDLL(KERNELS_SO)
            vp, i64, i32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_size_t
            protos = {
                "gnn_abi_version": (ctypes.c_int, []),
                "gnn_error_string": (ctypes.c_char_p, [ctypes.c_int]),
                "gnn_launch_count": (i64, []),
                "gnn_set_corunner_ctas": (ctypes.c_int, [ctypes.c_int]),
                "gnn_host_gather_ctas": (ctypes.c_int, []),
                "gnn_set_blocking_sync": (ctypes.c_int, [ctypes.c_int]),
                "gnn_build_adj": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, vp, i64, i64, i64, vp, vp, vp, vp, vp]),
                "gnn_coo_to_csr": (ctypes.c_int, [vp, i64, i64, vp, vp, vp]),
                "gnn_csr_spmm_workspace_bytes": (sz, [i64, i64, i64]),
                "gnn_csr_spmm_f32": (ctypes.c_int, [vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, sz, vp]),
                "gnn_gather_spmm_f32": (ctypes.c_int, [vp, vp, vp, i64, i64, i64, i64, vp, vp, i64, vp, sz, vp]),
                "gnn_csr_spmm_counter_bytes": (sz, [i64, i64, i64]),
                "gnn_csr_spmm_partial_bytes": (sz, [i64, i64, i64]),
                "gnn_csr_spmm_f32_ex": (ctypes.c_int, [vp, vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, vp, sz, ctypes.c_uint, vp]),
                "gnn_gather_spmm_f32_ex": (ctypes.c_int, [vp, vp, vp, vp, i64, i64, i64, i64, vp, vp, i64, vp, vp, sz, ctypes.c_uint, vp]),
                "gnn_csr_spmm_t_f32": (ctypes.c_int, [vp, vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp]),
                "gnn_set_transpose_budget": (i64, [i64]),
                "gnn_probe_row_gather_f32": (ctypes.c_int, [vp, i64, i64, vp, i64, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(i64), vp]),
                "gnn_csr_transpose_workspace_bytes": (sz, [i64, i64, i64]),
                "gnn_csr_transpose": (ctypes.c_int, [vp, vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, sz, vp]),
                "gnn_placement_remap": (ctypes.c_int, [vp, i64, vp, vp, vp, i64, vp, i64, i64, vp, vp, vp, vp, vp]),
                "gnn_gather_rows_f32": (ctypes.c_int, [vp, i64, i64, vp, i64, vp]),
                "gnn_gather_rows_src_f32": (ctypes.c_int, [vp, vp, i32, i64, i64, vp, i64, vp]),
                "gnn_index_rows_f32": (ctypes.c_int, [vp, i64, vp, i64, i64, vp, i64, vp]),
                "gnn_row_slice_count": (ctypes.c_int, [vp, vp, i64, vp, vp, vp]),
                "gnn_row_slice_fill": (ctypes.c_int, [vp, vp, vp, i64, vp, vp, vp, vp]),
                "gnn_member_set": (ctypes.c_int, [vp, vp, vp, i64, ctypes.c_int, vp]),
                "gnn_column_slice_chunks": (i64, [i64]),
                "gnn_column_slice_count": (ctypes.c_int, [vp, i64, vp, i64, vp, vp, vp, vp]),
                "gnn_column_slice_fill": (ctypes.c_int, [vp, i64, vp, vp, vp, vp, ctypes.c_int, vp]),
                "gnn_elu_rownorm_workspace_bytes": (sz, [i64]),
                "gnn_elu_rownorm_fwd_f32": (ctypes.c_int, [vp, i64, i64, i64, vp, vp, vp, i64, vp, vp, vp]),
                "gnn_elu_rownorm_bwd_f32": (ctypes.c_int, [vp, i64, vp, i64, i64, i64, vp, vp, vp, vp, i64, vp, vp, vp, sz, vp]),
                "gnn_linear_split_elems": (sz, [i64, i64]),
                "gnn_linear_split_weights_f32": (ctypes.c_int, [vp, i64, i64, i64, vp, vp, vp]),
                "gnn_linear_split_weights2_f32": (ctypes.c_int, [vp, i64, i64, i64, vp, vp, vp, i64, i64, i64, vp, vp, vp]),
                "gnn_linear_tf32x3_f32": (ctypes.c_int, [vp, i64, vp, i64, i64, vp, i64, vp, vp, i64, vp]),
                "gnn_linear_tf32x3_f32_ex": (ctypes.c_int, [vp, i64, vp, i64, i64, vp, i64, vp, vp, i64, vp, ctypes.c_uint, vp]),
                "gnn_linear_wgrad_workspace_bytes": (sz, [i64, i64, i64]),
                "gnn_linear_wgrad_tf32x3_f32": (ctypes.c_int, [vp, i64, vp, i64, vp, i64, i64, i64, vp, i64, vp, vp, sz, vp]),
                "gnn_legacy_choice_f64": (ctypes.c_int, [vp, vp, i64, i64, vp]),
                "gnn_ladies_layer_host": (i64, [vp, vp, vp, i64, vp, i64, ctypes.c_double, vp, i64, i64, vp, vp, vp, vp]),
                "gnn_ladies_layer_host_ex": (i64, [vp, vp, vp, i64, vp, i64, vp, i64, ctypes.c_double, vp, i64, i64, vp, vp, vp, vp]),
                "gnn_support_compact": (ctypes.c_int, [vp, i64, vp, vp, vp, vp, vp]),
                "gnn_ladies_layer_host_dense": (i64, [vp, vp, i64, vp, i64, ctypes.c_double, vp, i64, i64, vp, vp, vp, vp, vp]),
                "gnn_shard_alloc": (ctypes.c_int, [sz, ctypes.POINTER(vp), ctypes.c_char_p]),
                "gnn_shard_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(vp)]),
                "gnn_shard_close": (ctypes.c_int, [vp]),
                "gnn_shard_free": (ctypes.c_int, [vp]),
                "gnn_host_register": (ctypes.c_int, [vp, sz, ctypes.POINTER(vp)]),
                "gnn_host_unregister": (ctypes.c_int, [vp]),
            }
            for name, (res, args) in protos.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            lib._gnn_protos = protos
            _cabi = lib
        return _cabi


def check(rc: int, what: str = "gnn_b200"):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {cabi().gnn_error_string(rc).decode()} (code {rc})")
