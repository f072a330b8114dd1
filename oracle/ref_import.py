"""Container-only: import the reference's own module files from /root/reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  /root/reference does not exist
on the GPU box, so nothing that runs there may call into this file; it is used
by ``make_golden.py`` and by the CPU-side pinning tests (skipped when the
reference tree is absent).

The reference's ``models/__init__.py`` imports every model family (several of
which are unimportable), so the two top-level packages
``models`` and ``utils`` are registered as bare namespace shells and only the
hot-path submodules are executed.
"""
import argparse
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _default_root():
    """/root/reference in the build container; on the GPU box the copy oracle/make_ref.py staged under oracle/_ref/."""
    for cand in (os.environ.get("MLG_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if cand and os.path.isdir(os.path.join(cand, "models", "gcn_lib")):
            return cand
    return "/root/reference"


REF_ROOT = _default_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "gcn_lib"))


def _shell(name, path):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def load():
    """Returns a namespace with the reference classes/functions on the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    from . import pyg_stub
    pyg_stub.install()
    _shell("models", os.path.join(REF_ROOT, "models"))
    _shell("models.gcn_lib", os.path.join(REF_ROOT, "models", "gcn_lib"))
    _shell("models.gcn_lib.sparse", os.path.join(REF_ROOT, "models", "gcn_lib", "sparse"))
    _shell("models.gcn_lib.dense", os.path.join(REF_ROOT, "models", "gcn_lib", "dense"))
    _shell("utils", os.path.join(REF_ROOT, "utils"))
    ns = types.SimpleNamespace()
    ns.torch_message = importlib.import_module("models.gcn_lib.sparse.torch_message")
    ns.torch_vertex = importlib.import_module("models.gcn_lib.sparse.torch_vertex")
    ns.torch_edge = importlib.import_module("models.gcn_lib.sparse.torch_edge")
    ns.torch_nn = importlib.import_module("models.gcn_lib.sparse.torch_nn")
    ns.dense_edge = importlib.import_module("models.gcn_lib.dense.torch_edge")
    ns.dense_nn = importlib.import_module("models.gcn_lib.dense.torch_nn")
    ns.multilevel_gnn = importlib.import_module("models.multilevel_gnn")
    ns.deepergcn = importlib.import_module("models.deepergcn")
    ns.diff_pooling = importlib.import_module("models.diff_pooling")
    ns.vae = importlib.import_module("models.vae")
    return ns


def default_args(config=None, **overrides):
    """opt.py's argparse defaults overlaid with a YAML config (opt.py:437-444), as a Namespace."""
    import yaml
    saved = sys.argv
    sys.argv = ["train.py"]
    try:
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        opt = importlib.import_module("opt")
        args = opt.parser.parse_args([])
    finally:
        sys.argv = saved
        if REF_ROOT in sys.path:
            sys.path.remove(REF_ROOT)
    if config is not None:
        with open(os.path.join(REF_ROOT, "config", config)) as f:
            for k, v in yaml.safe_load(f).items():
                setattr(args, k, v)
    for k, v in overrides.items():
        setattr(args, k, v)
    return argparse.Namespace(**vars(args))
