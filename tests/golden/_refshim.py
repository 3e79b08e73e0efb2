"""Run the UNMODIFIED reference (/root/reference/THB) in this container to produce golden vectors.

Only usable where /root/reference exists (the build container); never imported by tests, bench or the
product.  The reference needs jax / matplotlib / pyvista, none of which are installed and there is no
network, so this module installs throw-away stand-ins into ``sys.modules``:

* ``jax`` / ``jax.numpy`` -> numpy (``jit`` = identity, ``vmap`` = python loop, ``lax.dynamic_slice`` =
  slicing).  ``jacfwd`` is NOT emulated (derivative goldens come from scipy instead).
* ``matplotlib``, ``pyvista`` -> empty modules (only plotting code uses them).

and loads the reference sources *from where they lie* with exactly one textual repair, applied in memory:
the three type annotations at THB/core.py:606,608,612 that raise TypeError at ``def`` time on
python >= 3.9.  No other line of the reference is altered; nothing is copied into the repo.

``fp16_tables=False`` maps ``jnp.float16`` to float64 so that the fp16 quantisation of the coefficient
tables at THB/core.py:111 is disabled; ``True`` keeps it.
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"


def _make_fake_jax(fp16_tables):
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            try:
                setattr(jnp, name, getattr(np, name))
            except Exception:
                pass

    def _einsum(sub, *ops, **kw):
        # JAX accepts "i... -> i" (sum over trailing axes); numpy does not.
        if sub.replace(" ", "") == "i...->i":
            a = np.asarray(ops[0])
            return a.reshape(a.shape[0], -1).sum(axis=1)
        kw.pop("optimize", None)
        return np.einsum(sub, *ops)

    jnp.einsum = _einsum
    jnp.array = lambda x, dtype=None: np.array(x, dtype=dtype)
    jnp.float16 = np.float16 if fp16_tables else np.float64
    jnp.ndarray = np.ndarray

    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)

    def vmap(f, in_axes=0):
        def run(*args):
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
            n = next(len(a) for a, ax in zip(args, axes) if ax is not None)
            out = []
            for i in range(n):
                out.append(f(*[a[i] if ax is not None else a for a, ax in zip(args, axes)]))
            return np.stack(out)
        return run

    def jacfwd(f, argnums=0):
        raise NotImplementedError("jacfwd is not emulated by the golden shim")

    lax = types.ModuleType("jax.lax")

    def dynamic_slice(x, starts, sizes):
        return x[tuple(slice(int(s), int(s) + int(n)) for s, n in zip(starts, sizes))]

    lax.dynamic_slice = dynamic_slice
    jax.vmap, jax.jacfwd, jax.lax = vmap, jacfwd, lax
    return jax, jnp, lax


def load_reference(fp16_tables=False):
    """Returns a namespace with the reference modules: .bspline_funcs, .core, .space"""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present; goldens can only be regenerated in the build container")
    for k in [k for k in sys.modules if k == "THB" or k.startswith("THB.") or k == "jax" or k.startswith("jax.")]:
        del sys.modules[k]
    jax, jnp, lax = _make_fake_jax(fp16_tables)
    sys.modules["jax"], sys.modules["jax.numpy"], sys.modules["jax.lax"] = jax, jnp, lax
    for name in ("matplotlib", "matplotlib.pyplot", "pyvista"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    pkg = types.ModuleType("THB")
    pkg.__path__ = [os.path.join(REF_ROOT, "THB")]
    sys.modules["THB"] = pkg

    def load(modname, patch=None):
        path = os.path.join(REF_ROOT, "THB", modname + ".py")
        src = open(path).read()
        if patch:
            src = patch(src)
        mod = types.ModuleType("THB." + modname)
        mod.__file__ = path
        sys.modules["THB." + modname] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        setattr(pkg, modname, mod)
        return mod

    def fix_annotations(src):
        # core.py:606,608,612 -- invalid typing subscripts
        src2 = src.replace("List[List[int, Tuple[int]]]", "list")
        src2 = src2.replace("Dict[int : np.ndarray]", "dict")
        src2 = src2.replace("List[bool, bool]", "list")
        assert src2 != src
        return src2

    ns = types.SimpleNamespace()
    ns.utils = load("utils")
    ns.bspline_funcs = load("bspline_funcs")
    ns.core = load("core", fix_annotations)
    ns.space = load("multilevel_spline_space")
    return ns
