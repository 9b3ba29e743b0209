"""Stand-in base models for the downstream parity tests (written for this
repo; the parameter creation order matches the reference's CompatibleGCN,
src/gnn/model.py:24-41, so a shared seed gives identical initial weights)."""
import torch.nn.functional as F
from torch import nn


class FixedLogits(nn.Module):
    """Stub base model of SURVEY 8d: fixed logits, ignores x and adj."""

    def __init__(self, logits):
        super().__init__()
        self.register_buffer("logits", logits)

    def forward(self, x, adj):
        return self.logits


class DenseGCN(nn.Module):
    """Row-normalised two-layer dense GCN: relu(W1 (D^-1 A) x) -> dropout ->
    W2 (D^-1 A) h, the computation of src/gnn/model.py:43-53."""

    def __init__(self, nfeat, nclass, nhid=64, dropout=0.5):
        super().__init__()
        self.gc1 = nn.Linear(nfeat, nhid)
        self.gc2 = nn.Linear(nhid, nclass)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, adj):
        dev = self.gc1.weight.device
        x, adj = x.to(dev), adj.to(dev)
        deg = adj.sum(dim=1, keepdim=True)
        deg[deg == 0] = 1
        a = adj / deg
        h = self.dropout(F.relu(self.gc1(a @ x)))
        return self.gc2(a @ h)
