"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the UNMODIFIED reference (ViktorLiang/GW-Depth, mounted read-only at
/root/reference) inside this container so that `oracle/make_golden.py` can
generate golden vectors and `tests/` can pin `oracle/gwdepth_oracle.py` against
the real thing.  /root/reference does not exist on the GPU box, so nothing on
the `-m gpu` / smoke / bench path may import this module.

The reference touches a handful of GUI / plotting modules at import time that
are not installed here; they are replaced by inert stubs.
"""
import argparse
import os
import sys
import types
from unittest.mock import MagicMock

import torch.nn as nn

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
# the build container mounts the reference read-only at /root/reference; oracle/stage_ref.sh stages an unmodified copy under
# baseline/_ref (git-ignored, travels with gpurun) for the GPU box
REFERENCE_ROOT = os.environ.get("GWD_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/src/models") else _STAGED)

# the flag set of the only working full configuration
DEFAULT_FLAGS = ["--device", "cpu", "--num_queries", "100", "--with_line",
                 "--with_center", "--with_dense"]


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _DropPath(nn.Module):
    """Stochastic depth; the reference only ever builds it with p == 0."""

    def __init__(self, p=0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        assert self.p == 0.0 or not self.training
        return x


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "tkinter" not in sys.modules:
        tk = _stub("tkinter")
        tk.messagebox = _stub("tkinter.messagebox", NO="no")
    _stub("turtle", forward=lambda *a, **k: None, color=lambda *a, **k: None)
    for n in ("matplotlib", "matplotlib.path", "matplotlib.transforms",
              "matplotlib.image", "matplotlib.pyplot", "matplotlib.colors",
              "matplotlib.cm"):
        if n not in sys.modules:
            sys.modules[n] = MagicMock(name=n)
    if "imp" not in sys.modules:
        _stub("imp")
    if "docopt" not in sys.modules:
        _stub("docopt", docopt=lambda *a, **k: {})
    if "timm" not in sys.modules:
        timm = _stub("timm")
        timm.models = _stub("timm.models")
        timm.models.layers = _stub(
            "timm.models.layers", DropPath=_DropPath,
            to_2tuple=lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x),
            trunc_normal_=nn.init.trunc_normal_)
    for p in (os.path.join(REFERENCE_ROOT, "src"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import models.backbone as bb  # noqa: E402  (reference module)
    bb.is_main_process = lambda: False  # no ImageNet download (no network)
    _installed = True


def reference_args(extra_flags=()):
    install()
    from args import get_args_parser  # reference module
    parser = argparse.ArgumentParser(parents=[get_args_parser()])
    return parser.parse_args(list(DEFAULT_FLAGS) + list(extra_flags))


def build_reference(extra_flags=()):
    """-> (model, criterions, postprocessors, args) built by the reference's own build_model."""
    install()
    from models import build_model  # reference module
    args = reference_args(extra_flags)
    model, criterions, post = build_model(args)
    return model, criterions, post, args
