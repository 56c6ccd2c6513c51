"""OGB molecule encoders (ogb==1.2.2 ``AtomEncoder`` / ``BondEncoder``): the sum of one embedding table per integer
feature column, with the reference's parameter names (``atom_embedding_list.{i}.weight`` / ``bond_embedding_list.{i}``)
so checkpoints interchange.  On CUDA the lookup + sum runs as ONE gather over the concatenated tables."""
import torch
import torch.nn as nn

ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_FEATURE_DIMS = [5, 6, 2]


class _SumEmbedding(nn.Module):
    _list_name = "embedding_list"

    def __init__(self, dims, emb_dim):
        super().__init__()
        tables = nn.ModuleList()
        for dim in dims:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            tables.append(emb)
        setattr(self, self._list_name, tables)
        offs = torch.tensor([0] + list(dims[:-1])).cumsum(0)
        self.register_buffer("_offsets", offs, persistent=False)

    def forward(self, x):
        tables = getattr(self, self._list_name)
        n_col = x.shape[1]
        # one embedding_bag(sum) over the stacked tables: a single gather + reduction instead of 9 lookups and 8 adds
        weight = torch.cat([t.weight for t in tables[:n_col]], dim=0)
        idx = x.long() + self._offsets[:n_col].to(x.device)
        return nn.functional.embedding_bag(idx, weight, mode="sum")


class AtomEncoder(_SumEmbedding):
    _list_name = "atom_embedding_list"

    def __init__(self, emb_dim):
        super().__init__(ATOM_FEATURE_DIMS, emb_dim)


class BondEncoder(_SumEmbedding):
    _list_name = "bond_embedding_list"

    def __init__(self, emb_dim):
        super().__init__(BOND_FEATURE_DIMS, emb_dim)
