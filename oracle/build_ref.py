"""Byte-compile the UNMODIFIED reference's hot-path modules into ``oracle/_ref/`` (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is a python code drop: "compiling it from the
sources where they lie" means compiling ``realworld_benchmark/nets/**/*.py`` into SOURCELESS code objects
(``<module>.refbin`` = marshalled bytecode; ``*.pyc`` files are stripped from gpurun snapshots) under the git-ignored
``oracle/_ref/nets/`` (no reference source is copied into the repository; the directory is not gpurun-ignored, so the
compiled modules travel to the GPU box like the built ``.so``).  ``bench.py --impl
reference`` and the ``cpu_baseline`` leg import them from there - on top of the DGL-0.4.2 stand-in of
``oracle/standin`` - and time the reference's own python path (``kind: "reference"``); when ``oracle/_ref`` is absent
they fall back to the oracle port (``kind: "port"``).

    python -m oracle.build_ref            # needs /root/reference (or DGN_REFERENCE)
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import marshal
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DGN_REFERENCE", "/root/reference/realworld_benchmark")
OUT = os.path.join(REPO, "oracle", "_ref")
EXT = ".refbin"


def build_ref(verbose: bool = False) -> int:
    """Returns the number of compiled modules (0 when the reference is not mounted)."""
    src_root = os.path.join(REF, "nets")
    if not os.path.isdir(src_root):
        return 0
    n = 0
    for dirpath, _, files in os.walk(src_root):
        rel = os.path.relpath(dirpath, REF)
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            dst_dir = os.path.join(OUT, rel)
            os.makedirs(dst_dir, exist_ok=True)
            dst = os.path.join(dst_dir, f[:-3] + EXT)
            with open(os.path.join(dirpath, f), "rb") as fh:
                code = compile(fh.read(), os.path.join("<reference>", rel, f), "exec", dont_inherit=True)
            with open(dst, "wb") as fh:
                fh.write(marshal.dumps(code))
            n += 1
            if verbose:
                print("compiled", os.path.join(rel, f), "->", os.path.relpath(dst, REPO))
    with open(os.path.join(OUT, "PYTHON"), "w") as fh:           # sourceless .pyc files are interpreter specific
        fh.write("%d.%d\n" % sys.version_info[:2])
    return n


def ref_available() -> bool:
    tag = os.path.join(OUT, "PYTHON")
    try:
        return (open(tag).read().strip() == "%d.%d" % sys.version_info[:2] and
                os.path.exists(os.path.join(OUT, "nets", "dgn_layer" + EXT)))
    except OSError:
        return False


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports ``nets`` / ``nets.*`` from the marshalled code objects under ``oracle/_ref``."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "nets" and not fullname.startswith("nets."):
            return None
        base = os.path.join(OUT, *fullname.split("."))
        if os.path.isdir(base):
            spec = importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        if os.path.exists(base + EXT):
            return importlib.machinery.ModuleSpec(fullname, self, origin=base + EXT)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        origin = module.__spec__.origin
        if origin is None:                       # a (namespace-like) package directory
            return
        module.__file__ = origin
        with open(origin, "rb") as fh:
            exec(marshal.loads(fh.read()), module.__dict__)


def import_ref():
    """Puts the DGL stand-in and ``oracle/_ref`` on ``sys.path`` and returns the reference's modules
    ``(nets.aggregators, nets.scalers, nets.dgn_layer, ZINC DGNNet)``.  The reference's ``message_func`` moves the
    eigenvectors to 'cuda' whenever a GPU is visible (rb/nets/dgn_layer.py:82-84): CPU timing runs must hide the
    GPUs (``CUDA_VISIBLE_DEVICES=""``) before torch initialises CUDA."""
    from oracle import use_standin_dgl
    use_standin_dgl()
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    import nets.aggregators as ra
    import nets.scalers as rs
    import nets.dgn_layer as rl
    from nets.molecules_graph_regression.dgn_net import DGNNet
    assert os.path.abspath(rl.__file__).startswith(OUT), "reference modules shadowed by %s" % rl.__file__
    return ra, rs, rl, DGNNet


if __name__ == "__main__":
    print("%d reference modules compiled into %s" % (build_ref(verbose=True), OUT))
