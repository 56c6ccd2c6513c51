"""``ogb.graphproppred.mol_encoder`` surface for the overlay (the reference's HIV / PCBA nets import it,
realworld_benchmark/nets/HIV_graph_classification/dgn_net.py:6); only used when the real ``ogb`` is not installed."""
