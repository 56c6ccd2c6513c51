"""CPU oracle for the DGN directional-aggregation path.  TEST INFRASTRUCTURE - NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import anything from this package, and only as the checker or as the timed CPU
baseline.  ``dgn_b200`` never imports it; the product path fails loudly when its CUDA
library is missing instead of falling back to this code.

What is here
------------
* ``oracle.standin.dgl``   restatement of the DGL-0.4.2 calls the reference makes (the one
                           part of the path that lives in an absent third-party dependency).
* ``oracle.mailbox_ops``   the 24 aggregators + 3 scalers (realworld_benchmark/nets/aggregators.py,
                           realworld_benchmark/nets/scalers.py) restated on ``[n, D, F]`` mailboxes.
* ``oracle.mlp``           FCLayer / MLP (realworld_benchmark/nets/layers.py).
* ``oracle.directional_layers``  DGNLayer{Simple,Complex,Tower} (realworld_benchmark/nets/dgn_layer.py).
* ``oracle.task_nets``     the ZINC / SBM / superpixel DGNNet heads that call the layers.
* ``oracle.make_golden``   imports the UNMODIFIED reference from /root/reference (build
                           container only) and writes tests/golden/*.npz.

Pinning status
--------------
The reference ships no tests, golden vectors or fixtures, so the oracle is
pinned against outputs of the reference itself: ``oracle/make_golden.py`` runs the real
``realworld_benchmark/nets/*.py`` on the DGL stand-in and commits inputs + outputs + gradients
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against
them.  The DGL-0.4.2 degree-bucketing semantics themselves are restated from the published
behaviour of that release and cannot be executed here ("DGL semantics unpinned").
"""
import os
import sys

STANDIN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standin")


def use_standin_dgl():
    """Put the DGL stand-in first on ``sys.path`` unless a real ``dgl`` is already imported."""
    if "dgl" not in sys.modules and STANDIN_PATH not in sys.path:
        sys.path.insert(0, STANDIN_PATH)
    import dgl  # noqa: F401
    return sys.modules["dgl"]
