"""Drop-in ``custom_sparse_ops`` for HPC-Research-Lab/GNN (reference custom_sparse_ops.py).

The reference's main.py / models.py / sampler.py do ``import custom_sparse_ops`` from
the working directory; putting this repository root first on ``sys.path`` (or
copying this three-line shim next to them) switches them to the B200-native path.
See INTEGRATION.md."""
from gnn_b200.custom_sparse_ops import (  # noqa: F401
    Adjacency, SparseDenseMM, adjacency_of, create_coo_tensor, spmm, spmm_backward_time, spmm_cpp, spmm_forward_time,
)
