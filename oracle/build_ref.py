"""Build recipe for oracle/_ref (test infrastructure).

Compiles the reference's own CPU extension  THB/THB_extensions/compute_pts.cpp  (module THB_eval:
cpp_forward / cpp_backward) from where it lies under /root/reference with a direct g++ command -- the
reference's setup.py is not run, no reference source is copied into the repo.  Output goes to
oracle/_ref/THB_eval.so only (git-ignored; it travels to the GPU box with the snapshot).

    python oracle/build_ref.py            # build if the reference tree is present and the .so is stale
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/THB/THB_extensions/compute_pts.cpp"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "THB_eval.so")


def build(force=False, verbose=False):
    """Returns the path of the built module, or None when the reference tree is absent (GPU box)."""
    if not os.path.exists(REF_SRC):
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(REF_SRC):
        return OUT
    import torch
    from torch.utils import cpp_extension

    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in cpp_extension.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (
        ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", REF_SRC, "-o", OUT]
        + inc
        + [
            "-DTORCH_EXTENSION_NAME=THB_eval",
            "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
            f"-L{libdir}",
            f"-Wl,-rpath,{libdir}",
            "-ltorch",
            "-ltorch_cpu",
            "-ltorch_python",
            "-lc10",
        ]
    )
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


REF_CUDA_SRCS = ["/root/reference/THB/THB_extensions/compute_pts_cuda.cpp", "/root/reference/THB/THB_extensions/compute_pts_cuda_kernels.cu"]
OUT_CUDA_DIR = os.path.join(OUT_DIR, "cuda")
OUT_CUDA = os.path.join(OUT_CUDA_DIR, "THB_eval.so")


def build_cuda(force=False, verbose=False):
    """The reference's CUDA extension (compute_pts_cuda.cpp + compute_pts_cuda_kernels.cu), compiled unmodified
    for sm_100a -- the GPU kernels to beat for THB_eval.forward/backward.  Output: oracle/_ref/cuda/THB_eval.so"""
    if not all(os.path.exists(p) for p in REF_CUDA_SRCS):
        return OUT_CUDA if os.path.exists(OUT_CUDA) else None
    if os.path.exists(OUT_CUDA) and not force:
        return OUT_CUDA
    import torch
    from torch.utils import cpp_extension

    os.makedirs(OUT_CUDA_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in cpp_extension.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", "-I/usr/local/cuda/include"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    defs = ["-DTORCH_EXTENSION_NAME=THB_eval", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    o1, o2 = os.path.join(OUT_CUDA_DIR, "shim.o"), os.path.join(OUT_CUDA_DIR, "kernels.o")
    cmds = [
        ["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-c", REF_CUDA_SRCS[0], "-o", o1] + inc + defs,
        ["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-w",
         "--expt-relaxed-constexpr", "-c", REF_CUDA_SRCS[1], "-o", o2] + inc + defs,
        ["g++", "-shared", o1, o2, "-o", OUT_CUDA, f"-L{libdir}", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-ltorch", "-ltorch_cpu",
         "-ltorch_cuda", "-ltorch_python", "-lc10", "-lc10_cuda", "-lcudart"],
    ]
    for c in cmds:
        if verbose:
            print(" ".join(c))
        subprocess.check_call(c)
    return OUT_CUDA


def load_cuda():
    import importlib.util

    import torch  # noqa: F401

    if not os.path.exists(OUT_CUDA):
        return None
    spec = importlib.util.spec_from_file_location("THB_eval", OUT_CUDA)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import the compiled reference module (None if it was never built)."""
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    if not os.path.exists(OUT):
        return None
    spec = importlib.util.spec_from_file_location("THB_eval", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", p)
    print("oracle/_ref/cuda:", build_cuda(force="--force" in sys.argv, verbose=True))
