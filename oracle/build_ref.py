"""Compile the reference's own CUDA extension, unmodified, into oracle/_ref/.

TEST INFRASTRUCTURE.  Sources are compiled where they lie under /root/reference
(spmm_cpp/spmm.cpp, spmm_cpp/cuda_spmm.cu - the two files custom_sparse_ops.py:8
JIT-builds); nothing is copied into the repo.  The output (spmm_ref.so) is
git-ignored but travels to the GPU box with the gpurun snapshot, where
tests/golden/make_golden_gpu.py and bench.py's reference-kernel leg load it.
"""
import os
import sys


def build(ref_root="/root/reference", verbose=False):
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_ref")
    srcs = [os.path.join(ref_root, "spmm_cpp", "spmm.cpp"), os.path.join(ref_root, "spmm_cpp", "cuda_spmm.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return None
    so = os.path.join(out, "spmm_ref.so")
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(out, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="spmm_ref", sources=srcs, build_directory=out, verbose=verbose, is_python_module=False)
    return so


def load_ref():
    """Import oracle/_ref/spmm_ref.so as a Python module (GPU box or here); None if absent."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "spmm_ref.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("spmm_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", verbose=True))
