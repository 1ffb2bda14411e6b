"""TEST INFRASTRUCTURE ONLY -- loads the real reference scripts from /root/reference.

Used by oracle/make_golden.py and by the `not gpu` tests that pin the oracle restatements against the
reference itself (they skip when the reference tree is absent, e.g. on the GPU box).  Nothing in the product
package imports this module.

The reference scripts are flat research files that import plotting / dataset libraries which are not
installed here and overwrite CUDA_VISIBLE_DEVICES at import time (try_with_torch.py:19-21); both are
neutralised below.  Their `main()` is guarded by `if __name__ == '__main__'`.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HG_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "pycocotools", "pycocotools.coco",
    "apex", "apex.amp", "tensorboardX", "torchstat", "skimage", "skimage.feature", "graphviz", "torchviz",
    "pydensecrf", "pydensecrf.densecrf", "pydensecrf.utils",
]


def available():
    return os.path.isdir(REFERENCE_ROOT) and os.path.isfile(os.path.join(REFERENCE_ROOT, "try_with_torch.py"))


class _Anything:
    """Attribute sink for stubbed libraries (SummaryWriter, COCO, amp, ...)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _install_stubs():
    import numpy as np
    import numpy.matlib  # noqa: F401  (the reference calls np.matlib.repmat)

    if not hasattr(np, "int"):
        np.int = int  # removed in numpy >= 1.24; reference uses astype(np.int)
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:  # noqa: BLE001
            pass
        m = types.ModuleType(name)

        def _getattr(attr, _n=name):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Anything

        m.__getattr__ = _getattr  # type: ignore[attr-defined]
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)


_CACHE = {}


def load(script, fresh=False):
    """Import /root/reference/<script>.py as a module object (cached unless fresh=True)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if script in _CACHE and not fresh:
        return _CACHE[script]
    _install_stubs()
    saved = {k: os.environ.get(k) for k in ("CUDA_VISIBLE_DEVICES", "CUDA_DEVICE_ORDER")}
    path = os.path.join(REFERENCE_ROOT, script + ".py")
    spec = importlib.util.spec_from_file_location("hgref_" + script, path)
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    _CACHE[script] = mod
    return mod
