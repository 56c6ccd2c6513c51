"""dgn_b200 - B200-native directional graph network (DGN) aggregation engine.

Public surface (mirrors the reference's ``realworld_benchmark/nets`` package):

* ``dgn_b200.nets.dgn_layer.DGNLayer`` / ``dgn_b200.nets.aggregators.AGGREGATORS`` /
  ``dgn_b200.nets.scalers.SCALERS``  - the plugin registry and layer factory,
* ``dgn_b200.task_nets.<task>.DGNNet``  - task networks,
* ``dgn_b200.graph.BatchedGraph`` / ``collate``  - batched CSR graph with a DGL-like surface,
* ``dgn_b200.ops``  - autograd ops over the C ABI in ``include/dgn_b200.h``.

Importing the package loads ``libdgn_b200.so``; it raises if the library is missing.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)

__version__ = "0.1.0"
