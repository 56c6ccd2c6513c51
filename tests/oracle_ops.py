"""Oracle-side evaluation of one aggregation on a stand-in graph (test helper)."""
import torch

from oracle import mailbox_ops as mo
from oracle import use_standin_dgl


def oracle_aggregate(n_nodes, src, dst, eig, h_in, messages, agg_names, scaler_names, avg_log):
    """messages [E,F] in edge-id order -> [N, S*A*F] via degree-bucketed reduce (autograd-capable)."""
    dgl = use_standin_dgl()
    g = dgl.DGLGraph(n_nodes, src, dst)
    g.ndata["h"], g.ndata["eig"] = h_in, eig
    g.edata["e"] = messages
    aggs = [mo.AGGREGATORS[a] if a in mo.AGGREGATORS else mo.build_aggregator_registry(6)[a] for a in agg_names]
    scs = [mo.SCALERS[s] for s in scaler_names]
    avg = {"log": torch.tensor(avg_log, dtype=torch.float32)}

    def edge_udf(edges):
        return {"eig_s": edges.src["eig"], "eig_d": edges.dst["eig"]}

    def msg_udf(edges):
        return {"e": edges.data["e"], "eig_s": edges.data["eig_s"], "eig_d": edges.data["eig_d"]}

    def red_udf(nodes):
        b = nodes.mailbox
        return {"h": mo.reduce_bucket(b["e"], b["eig_s"], b["eig_d"], nodes.data["h"], aggs, scs, avg)}

    g.apply_edges(edge_udf)
    g.update_all(msg_udf, red_udf)
    return g.ndata["h"]
