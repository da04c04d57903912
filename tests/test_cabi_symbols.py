"""CPU-side checks of the boundary: the C-ABI library loads and exports every function
include/gnn_b200.h declares, the pybind layer keeps the reference's names, and argument
errors come back as codes (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from gnn_b200 import _native
    return _native


def _declared():
    text = open(os.path.join(REPO, "include", "gnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gnn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    names = _declared()
    assert len(names) >= 18
    lib = ctypes.CDLL(built.KERNELS_SO)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gnn_b200.h but not exported"
    # and the ctypes prototypes cover exactly the header
    assert sorted(built.cabi()._gnn_protos) == names


def test_abi_version_and_errors(built):
    lib = built.cabi()
    assert lib.gnn_abi_version() == 4
    assert b"workspace" in lib.gnn_error_string(-2)
    assert lib.gnn_error_string(0) == b"success"
    # argument errors are reported before any CUDA call
    assert lib.gnn_build_adj(None, None, None, 3, None, 4, 4, 4, None, None, None, None, None) == -1
    assert lib.gnn_csr_spmm_f32(None, None, None, -1, 1, 1, 1, None, 1, None, 1, None, 0, None) == -1
    assert lib.gnn_csr_spmm_f32_ex(None, None, None, None, 1, 1, 1, 1, None, 1, None, 1, None, None, 0, 2, None) == -1   # unknown flag
    assert lib.gnn_csr_spmm_t_f32(None, None, None, None, -1, 1, 1, 1, None, 1, None, 1, None) == -1
    assert (lib.gnn_csr_spmm_counter_bytes(100, 1000, 64) + lib.gnn_csr_spmm_partial_bytes(100, 1000, 64)
            == lib.gnn_csr_spmm_workspace_bytes(100, 1000, 64))
    # transpose bitmap budget: process-wide knob, returns the previous value; the workspace never exceeds budget + tables
    prev = lib.gnn_set_transpose_budget(1 << 20)
    assert prev == 512 << 20
    assert lib.gnn_csr_transpose_workspace_bytes(200000, 150000, 10) <= (1 << 20) + 150000 * 8 + 150000 * 8 + 4096
    assert lib.gnn_set_transpose_budget(0) == 1 << 20
    assert lib.gnn_set_transpose_budget(-5) == -1
    assert lib.gnn_csr_spmm_workspace_bytes(100, 1000, 64) >= 2 * (1000 // 64) * 64 * 4
    assert lib.gnn_csr_transpose_workspace_bytes(64, 10, 5) >= 10 * 2 * 4
    # planner hint: returns the previous value, rejects negatives, no CUDA call involved
    g = lib.gnn_host_gather_ctas()
    assert 1 <= g <= 148
    assert lib.gnn_set_corunner_ctas(g) == 0
    assert lib.gnn_set_corunner_ctas(-1) == -1
    assert lib.gnn_set_corunner_ctas(0) == g


def test_extension_keeps_reference_names(built):
    ext = built.extension()
    for name in ["spmm_naive", "spmm_load_balance", "create_coo_tensor"]:   # reference spmm.cpp:52-56
        assert callable(getattr(ext, name))
    import custom_sparse_ops as cso
    for name in ["spmm", "create_coo_tensor", "spmm_forward_time", "spmm_backward_time", "SparseDenseMM", "spmm_cpp"]:
        assert hasattr(cso, name)
    assert cso.spmm_forward_time == 0.0 and cso.spmm_backward_time == 0.0


def test_cpu_operands_raise(built):
    import torch
    import custom_sparse_ops as cso
    a = torch.sparse_coo_tensor(torch.tensor([[0, 1], [1, 0]]), torch.ones(2), (2, 2)).coalesce()
    with pytest.raises(RuntimeError, match="CUDA"):
        cso.spmm(a, torch.ones(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        cso.spmm_cpp.spmm_load_balance(a, torch.ones(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        cso.create_coo_tensor(torch.zeros(3, dtype=torch.int32), torch.zeros(3, dtype=torch.int32),
                              torch.zeros(0, dtype=torch.int16), torch.zeros(2), 2, 2)
