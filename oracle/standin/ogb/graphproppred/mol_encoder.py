"""``AtomEncoder`` / ``BondEncoder`` of ogb 1.2.2 restated (TEST INFRASTRUCTURE, see oracle/__init__.py).

Call sites: realworld_benchmark/nets/HIV_graph_classification/dgn_net.py:6,45,48 and
realworld_benchmark/nets/PCBA_graph_classification/dgn_net.py:6,36,39.  Published algorithm: one ``nn.Embedding`` per
integer feature column (vocabulary sizes = ``get_atom_feature_dims()`` / ``get_bond_feature_dims()`` of
``ogb.utils.features``), Xavier-uniform initialised, summed over the columns.  "parity unpinned" for this file: the
third-party source is absent, only its call sites are pinned by the golden net fixtures.
"""
import torch

# ogb.utils.features.allowable_features (1.2.x): atomic number (118 + misc), chirality, degree (0..10 + misc), formal
# charge (-5..5 + misc), #H (0..8 + misc), radical electrons (0..4 + misc), hybridisation, aromatic, in ring
full_atom_feature_dims = [119, 4, 12, 12, 10, 6, 6, 2, 2]
# bond type (4 + misc), stereo, conjugated
full_bond_feature_dims = [5, 6, 2]


class AtomEncoder(torch.nn.Module):
    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = torch.nn.ModuleList()
        for dim in full_atom_feature_dims:
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)

    def forward(self, x):
        x_embedding = 0
        for i in range(x.shape[1]):
            x_embedding = x_embedding + self.atom_embedding_list[i](x[:, i])
        return x_embedding


class BondEncoder(torch.nn.Module):
    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = torch.nn.ModuleList()
        for dim in full_bond_feature_dims:
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)

    def forward(self, edge_attr):
        bond_embedding = 0
        for i in range(edge_attr.shape[1]):
            bond_embedding = bond_embedding + self.bond_embedding_list[i](edge_attr[:, i])
        return bond_embedding
