"""CPU oracle for the ML (GNN candidate-reduction) path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's ``src/models/modelML.py`` cannot be imported
anywhere in this project (it needs ``torch_geometric==1.7.0`` and
``torch_scatter==2.0.6``, ``requirements.txt:6-7``, neither vendored in
``/root/reference`` nor installable offline) and the reference ships no test,
fixture or golden vector for it.  This file is therefore a *restatement* of

* ``src/models/modelML.py:9-29``    NodeEncoder (only table 0 is reachable)
* ``src/models/modelML.py:55-115``  Net.__init__ layer layout / parameter names
* ``src/models/modelML.py:131-176`` Net.forward

plus the published algorithms of the third-party operators it calls (PyG 1.7.0):
``GINConv`` (sum aggregation at the target, ``out += (1+eps) * x``, then the
MLP), ``GCNConv`` + ``gcn_norm`` + ``add_remaining_self_loops(fill=1)`` (flow
source_to_target: ``edge_index[0]`` = source j, ``edge_index[1]`` = target i;
transform first, then aggregate; bias after aggregation; weight stored
``[in,out]``, glorot init), ``scatter(reduce='mean')`` (sum / max(count,1)) and
``Batch.from_data_list`` collation including its ``*index*`` increment rule
(SURVEY 8a-5').  Every report that cites ML parity says "restated oracle".

Aggregations are done with ``index_add_`` in edge order (deterministic,
sequential per destination on CPU), which is the order the CUDA CSR kernel
reproduces.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------
# sparse helpers (PyG 1.7.0 semantics)
# --------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor],
                             num_nodes: int, fill_value: float = 1.0):
    """Non-loop edges keep their order; one loop per node is appended, carrying
    the weight of an existing loop on that node if there was one else ``fill``."""
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loops = torch.arange(num_nodes, dtype=edge_index.dtype)
    new_index = torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)
    new_weight = None
    if edge_weight is not None:
        lw = torch.full((num_nodes,), fill_value, dtype=edge_weight.dtype)
        old = edge_weight[~keep]
        if old.numel() > 0:
            lw[src[~keep]] = old
        new_weight = torch.cat([edge_weight[keep], lw])
    return new_index, new_weight


def gcn_norm(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int):
    """``deg[i] = sum_{e: dst=i} w_e``;  ``norm_e = deg^-1/2[src] * w_e * deg^-1/2[dst]`` (inf -> 0)."""
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.shape[1], dtype=torch.float32)
    ei, ew = add_remaining_self_loops(edge_index, edge_weight, num_nodes, 1.0)
    src, dst = ei[0], ei[1]
    deg = torch.zeros(num_nodes, dtype=ew.dtype).index_add_(0, dst, ew)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    return ei, dis[src] * ew * dis[dst]


def aggregate_sum(x: torch.Tensor, edge_index: torch.Tensor, weight: Optional[torch.Tensor] = None):
    """``out[dst] += w * x[src]`` in edge order."""
    msg = x[edge_index[0]]
    if weight is not None:
        msg = weight.view(-1, 1) * msg
    return torch.zeros_like(x).index_add_(0, edge_index[1], msg)


def segment_mean(x: torch.Tensor, seg: torch.Tensor, num_segments: Optional[int] = None):
    """torch_scatter ``scatter(x, seg, dim=0, reduce='mean')``."""
    n = int(seg.max()) + 1 if num_segments is None else num_segments
    out = torch.zeros(n, x.shape[1], dtype=x.dtype).index_add_(0, seg, x)
    cnt = torch.zeros(n, dtype=x.dtype).index_add_(0, seg, torch.ones_like(seg, dtype=x.dtype))
    return out / cnt.clamp(min=1).view(-1, 1)


# --------------------------------------------------------------------------
# layers with the reference's parameter names
# --------------------------------------------------------------------------
class NodeEncoderO(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.embeddings = nn.ModuleList([nn.Embedding(100, channels) for _ in range(9)])

    def reset_parameters(self):
        for e in self.embeddings:
            e.reset_parameters()

    def forward(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(1)
        out = 0
        for i in range(x.size(1)):
            out = out + self.embeddings[i](x[:, i].long())
        return out


class GINConvO(nn.Module):
    def __init__(self, mlp: nn.Module, train_eps: bool = True):
        super().__init__()
        self.nn = mlp
        self.eps = nn.Parameter(torch.zeros(1)) if train_eps else None

    def reset_parameters(self):
        for m in self.nn:
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()
        if self.eps is not None:
            self.eps.data.fill_(0.0)

    def forward(self, x, edge_index):
        out = aggregate_sum(x, edge_index)
        out = out + (1 + (self.eps if self.eps is not None else 0.0)) * x
        return self.nn(out)


class GCNConvO(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cin, cout))
        self.bias = nn.Parameter(torch.empty(cout))
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))
        self.weight.data.uniform_(-a, a)
        self.bias.data.zero_()

    def forward(self, x, edge_index, edge_weight=None):
        ei, norm = gcn_norm(edge_index, edge_weight, x.size(0))
        x = x @ self.weight
        return aggregate_sum(x, ei, norm) + self.bias


class NetO(nn.Module):
    """Restated ``Net`` (modelML.py:55-176); same constructor, same state_dict keys."""

    def __init__(self, hiddenChannels, outChannels, embeddingChannels, numLayersGIN, numLayersGCN,
                 isServices=True, dropout=0.0):
        super().__init__()
        H, E = hiddenChannels, embeddingChannels
        self.numLayersGIN, self.numLayersGCN = numLayersGIN, numLayersGCN
        self.dropout, self.outChannels, self.isService = dropout, outChannels, isServices
        self.nodeEncoder, self.serviceEncoder = NodeEncoderO(E), NodeEncoderO(E)
        self.nodeConvs, self.nodeBatchNorms = nn.ModuleList(), nn.ModuleList()
        for layer in range(numLayersGIN):
            cin = E + 2 * 3 if layer == 0 else H
            mlp = nn.Sequential(nn.Linear(cin, 2 * H), nn.BatchNorm1d(2 * H), nn.ReLU(), nn.Linear(2 * H, H))
            self.nodeConvs.append(GINConvO(mlp, True))
            self.nodeBatchNorms.append(nn.BatchNorm1d(H))
        self.nodeLin = nn.Linear(H, H)
        self.serviceConvs, self.serviceBatchNorms = nn.ModuleList(), nn.ModuleList()
        for layer in range(numLayersGCN):
            self.serviceConvs.append(GCNConvO(E + 4 if layer == 0 else 2 * H, 2 * H))
            self.serviceBatchNorms.append(nn.BatchNorm1d(2 * H))
        self.serviceLin = nn.Linear(2 * H, H)
        self.noServicesLins = nn.ModuleList(
            [nn.Linear(E + 4 if layer == 0 else 2 * H, 2 * H) for layer in range(numLayersGCN)])

    def reset_parameters(self):
        self.nodeEncoder.reset_parameters()
        self.serviceEncoder.reset_parameters()
        for conv, bn in zip(self.nodeConvs, self.nodeBatchNorms):
            conv.reset_parameters()
            bn.reset_parameters()
        self.nodeLin.reset_parameters()
        for conv, bn in zip(self.serviceConvs, self.serviceBatchNorms):
            conv.reset_parameters()
            bn.reset_parameters()
        self.serviceLin.reset_parameters()

    def forward(self, data, return_internals: bool = False):
        x = data.x.squeeze()
        x = torch.cat((self.nodeEncoder(x[:, 0].view(-1, 1).long()), x[:, 1:]), -1)
        for conv, bn in zip(self.nodeConvs, self.nodeBatchNorms):
            x = F.dropout(F.relu(bn(conv(x, data.edge_index))), self.dropout, training=self.training)
        xs = data.x_service.squeeze()
        xs = torch.cat((self.serviceEncoder(xs[:, 0].view(-1, 1).long()), xs[:, 1:]), -1)
        for i in range(self.numLayersGCN):
            if self.isService:
                xs = self.serviceConvs[i](xs, data.edge_index_service, data.edge_attr_service)
            else:
                xs = self.noServicesLins[i](xs)
            xs = F.dropout(F.relu(self.serviceBatchNorms[i](xs)), self.dropout, training=self.training)
        xs_gcn = xs
        xs = self.serviceLin(xs)
        x = self.nodeLin(x)
        x = segment_mean(x, data.batch)
        B = x.size(0)
        service_batch = torch.arange(self.outChannels).repeat(B)      # modelML.py:167-171
        xs = segment_mean(xs, service_batch, self.outChannels)
        scores = torch.sigmoid(x @ xs.t())
        if return_internals:
            return scores, {"x_req": x, "x_service": xs, "x_service_gcn": xs_gcn}
        return scores


# --------------------------------------------------------------------------
# PyG 1.7.0 collation
# --------------------------------------------------------------------------
def collate(samples: Sequence, faithful_quirk: bool = True):
    """``Batch.from_data_list`` for the attributes TrainML creates (trainML.py:93-114).

    Attributes whose name contains ``index`` are concatenated on the last dim
    and incremented by the running sum of each sample's ``num_nodes`` = the
    REQUEST graph's node count.  With ``faithful_quirk`` that also applies to
    ``edge_index_service`` (what PyG 1.7.0 does); otherwise service edges are
    offset by S per sample (the "sane" collation).
    """
    xs, ys, eis, batch, xsv, eisv, easv = [], [], [], [], [], [], []
    node_off = 0
    svc_off = 0
    for g, s in enumerate(samples):
        n = s.x.shape[0]
        xs.append(s.x)
        ys.append(s.y)
        eis.append(s.edge_index + node_off)
        batch.append(torch.full((n,), g, dtype=torch.long))
        xsv.append(s.x_service)
        eisv.append(s.edge_index_service + (node_off if faithful_quirk else svc_off))
        easv.append(s.edge_attr_service)
        node_off += n
        svc_off += s.x_service.shape[0]
    return SimpleNamespace(
        x=torch.cat(xs), y=torch.cat(ys), edge_index=torch.cat(eis, dim=1), batch=torch.cat(batch),
        x_service=torch.cat(xsv), edge_index_service=torch.cat(eisv, dim=1),
        edge_attr_service=torch.cat(easv), num_graphs=len(samples))
