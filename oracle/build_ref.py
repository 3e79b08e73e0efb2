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
