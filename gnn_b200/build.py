"""In-tree build of the native pieces (run by __graft_entry__.build()).

  gnn_b200/lib/libgnn_b200.so  <- csrc/gnn_kernels.cu   (nvcc, sm_100a, C ABI of include/gnn_b200.h)
  gnn_b200/lib/spmm.so         <- csrc/spmm_ext.cpp     (g++, pybind/torch layer named like the
                                                         reference's `spmm` extension)

Both land inside the tree so they travel with the gpurun snapshot; they are
git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib")
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(REPO, "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-Bsymbolic", "--use_fast_math=false"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_kernels(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libgnn_b200.so")
    src = os.path.join(CSRC, "gnn_kernels.cu")
    hdr = os.path.join(INC, "gnn_b200.h")
    deps = [src, hdr] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if force or _newer(out, deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
        _run([nvcc, *flags, "-I", INC, "-I", CSRC, "-o", out + ".tmp", src], verbose)
        os.replace(out + ".tmp", out)
    return out


def build_extension(verbose: bool = False, force: bool = False) -> str:
    from torch.utils import cpp_extension
    import torch
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "spmm.so")
    src = os.path.join(CSRC, "spmm_ext.cpp")
    hdr = os.path.join(INC, "gnn_b200.h")
    kern = build_kernels(verbose, force)
    if force or _newer(out, [src, hdr]):
        # The image's $CXX (/opt/gcc/bin/g++) is a wrapper whose -B tree only has a static libstdc++;
        # a second copy of libstdc++ inside the extension crashes on the first formatted TORCH_CHECK
        # (two sets of locale facets in one process).  Link against the shared libstdc++ torch uses.
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", shutil.which("g++") or "g++")
        incs = []
        for p in cpp_extension.include_paths(device_type="cuda") + [sysconfig.get_paths()["include"], INC]:
            incs += ["-isystem" if "site-packages" in p or "python" in p or "cuda" in p else "-I", p]
        torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
        abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=spmm",
               "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *incs, src,
               "-o", out + ".tmp", f"-L{LIB}", "-lgnn_b200", f"-L{torch_lib}",
               "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python",
               "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
        _run(cmd, verbose)
        os.replace(out + ".tmp", out)
    assert os.path.exists(kern)
    return out


def build_all(verbose: bool = False, force: bool = False):
    return build_kernels(verbose, force), build_extension(verbose, force)


if __name__ == "__main__":
    print(build_all(verbose=True, force="--force" in sys.argv))
