"""Drop-in for the reference's ``src/models/trainML.py``: ``TrainML(dataset1, numLayersGIN, numLayersGCN,
hiddenChannels, embeddingChannels, dropout, lr, epochs).start()`` with the same side effects
(``solutions/ML/<ds>/model-{e}.pkl`` and ``testServices-epoch{e}.txt``), minus torch_geometric:

* samples are kept as tensors in memory (the reference writes one ``.pt`` per instance, each embedding the
  whole service graph, and re-reads it every step -- trainML.py:91-107);
* collation restates PyG 1.7.0 ``Batch.from_data_list`` including its ``*index*`` increment rule, which
  offsets ``edge_index_service`` by the REQUEST graph's node count (SURVEY 8a-5'); ``faithful_quirk=False``
  gives the S-offset collation instead;
* P@1 / P@5 and the rankings are computed on the device for the whole batch (trainML.py:49-72 loops in Python
  with one device sync per comparison).
Kept quirks: ``ReduceLROnPlateau(mode='min')`` stepped on P@1 (trainML.py:134-141) and the training split being
re-ranked through the *shuffled* loader (trainML.py:146).
"""
from __future__ import annotations

import json
import os
import time
from types import SimpleNamespace
from typing import List, Sequence

import numpy as np
import torch
from torch.nn import BCELoss
from torch.optim import Adam
from torch.optim.lr_scheduler import ReduceLROnPlateau

from .loadData import loadData, ml_arrays
from .modelML import Net


def build_samples(arrays) -> List[SimpleNamespace]:
    """trainML.py:91-114: one Data(x, y, edge_index) per instance + the shared service map."""
    nodefeatures, services, edge_indices, ei_s, ea_s, labels, _ = arrays
    xs = torch.tensor(services, dtype=torch.float)
    eis = torch.tensor(ei_s, dtype=torch.long).view(2, -1)
    eas = torch.tensor(ea_s, dtype=torch.float)
    out = []
    for nf, ei, lab in zip(nodefeatures, edge_indices, labels):
        out.append(SimpleNamespace(x=torch.tensor(nf, dtype=torch.float), y=torch.tensor(lab, dtype=torch.float),
                                   edge_index=torch.tensor(ei, dtype=torch.long).view(2, -1), x_service=xs,
                                   edge_index_service=eis, edge_attr_service=eas))
    return out


def collate(samples: Sequence[SimpleNamespace], faithful_quirk: bool = True, device=None) -> SimpleNamespace:
    node_off, svc_off = 0, 0
    x, y, ei, batch, xs, eis, eas = [], [], [], [], [], [], []
    for g, s in enumerate(samples):
        n = s.x.shape[0]
        x.append(s.x); y.append(s.y); xs.append(s.x_service); eas.append(s.edge_attr_service)
        ei.append(s.edge_index + node_off)
        eis.append(s.edge_index_service + (node_off if faithful_quirk else svc_off))
        batch.append(torch.full((n,), g, dtype=torch.long))
        node_off += n
        svc_off += s.x_service.shape[0]
    b = SimpleNamespace(x=torch.cat(x), y=torch.cat(y), edge_index=torch.cat(ei, 1), batch=torch.cat(batch),
                        x_service=torch.cat(xs), edge_index_service=torch.cat(eis, 1),
                        edge_attr_service=torch.cat(eas), num_graphs=len(samples))
    if device is not None:
        for k, v in vars(b).items():
            if torch.is_tensor(v):
                setattr(b, k, v.to(device, non_blocking=True))
    return b


def collate_requests(samples: Sequence[SimpleNamespace], device=None, pin: bool = False) -> SimpleNamespace:
    """Request side only (x, edge_index, batch, num_graphs) of a collated batch -- what ``Net.score_requests`` reads.  The
    service graph is static and encoded once (``Net.service_encodings``), so it is not replicated per sample."""
    sizes = [s.x.shape[0] for s in samples]
    offs = torch.tensor([0] + sizes[:-1]).cumsum(0)
    b = SimpleNamespace(
        x=torch.cat([s.x for s in samples]),
        edge_index=torch.cat([s.edge_index + int(o) for s, o in zip(samples, offs)], 1),
        batch=torch.repeat_interleave(torch.arange(len(samples)), torch.tensor(sizes)),
        num_graphs=len(samples))
    for k, v in list(vars(b).items()):
        if torch.is_tensor(v):
            if pin:
                v = v.pin_memory()
            setattr(b, k, v.to(device, non_blocking=True) if device is not None else v)
    return b


class _Loader:
    def __init__(self, samples, batch_size, shuffle, faithful_quirk=True):
        self.samples, self.batch_size, self.shuffle, self.quirk = samples, batch_size, shuffle, faithful_quirk
        self.dataset = samples

    def __iter__(self):
        order = torch.randperm(len(self.samples)).tolist() if self.shuffle else range(len(self.samples))
        order = list(order)
        for i in range(0, len(order), self.batch_size):
            yield collate([self.samples[j] for j in order[i:i + self.batch_size]], self.quirk)


def precision_at(scores: torch.Tensor, y: torch.Tensor, ks=(1, 5)):
    """Per-row descending ranking + P@k, batched on the device (trainML.py:59-70)."""
    order = scores.argsort(dim=1, descending=True, stable=True)
    hits = y.gather(1, order[:, : max(ks)]) == 1
    return order, [hits[:, :k].float().mean(dim=1) for k in ks]


class TrainML:
    def __init__(self, dataset1, numLayersGIN, numLayersGCN, hiddenChannels, embeddingChannels, dropout, lr, epochs,
                 root=".", faithful_quirk=True):
        self.dataset1 = dataset1
        self.hiddenChannels, self.embeddingChannels = hiddenChannels, embeddingChannels
        self.numLayersGIN, self.numLayersGCN = numLayersGIN, numLayersGCN
        self.epochs, self.dropout, self.lr = epochs, dropout, lr
        self.root, self.faithful_quirk = root, faithful_quirk
        self.device = torch.device('cuda')
        self.criterion = BCELoss()
        self.train_loader = self.val_loader = self.model = self.optimizer = None

    def train(self):
        self.model.train()
        total = 0.0
        for data in self.train_loader:
            data = _to(data, self.device)
            self.optimizer.zero_grad()
            x = self.model(data).squeeze()
            x = x.view(data.num_graphs, -1)
            loss = self.criterion(x, data.y.view(x.size(0), x.size(1)))
            loss.backward()
            total += loss.item() * data.num_graphs
            self.optimizer.step()
        return total / len(self.train_loader.dataset)

    @torch.no_grad()
    def test(self, loader):
        self.model.eval()
        rankings, p1, p5 = [], [], []
        for data in loader:
            data = _to(data, self.device)
            x = self.model(data).view(data.num_graphs, -1)
            order, (a, b) = precision_at(x, data.y.view(x.size(0), x.size(1)))
            rankings.append(order)
            p1.append(a); p5.append(b)
        idx = torch.cat(rankings).cpu().numpy().tolist()
        return idx, [float(torch.cat(p1).mean()), float(torch.cat(p5).mean())]

    def start(self, arrays=None):
        arrays = arrays if arrays is not None else loadData(self.dataset1, self.root)
        samples = build_samples(arrays)
        n = len(samples)
        split = n // 4 * 3
        self.dataset1 += "/"
        self.train_loader = _Loader(samples[:split], 2, True, self.faithful_quirk)
        self.val_loader = _Loader(samples[split:], 2, False, self.faithful_quirk)
        t = time.time()
        self.model = Net(hiddenChannels=self.hiddenChannels, outChannels=len(arrays[5][0]),
                         embeddingChannels=self.embeddingChannels, numLayersGIN=self.numLayersGIN,
                         numLayersGCN=self.numLayersGCN, isServices=True, dropout=0.0).to(self.device)
        print(f"\nRun {0}:\n")
        self.model.reset_parameters()
        self.optimizer = Adam(self.model.parameters(), lr=self.lr)
        scheduler = ReduceLROnPlateau(self.optimizer, mode='min', factor=0.5, patience=3, min_lr=0.00001)
        out_dir = os.path.join(self.root, "solutions", "ML", self.dataset1)
        os.makedirs(out_dir, exist_ok=True)
        for epoch in range(self.epochs):
            lr = scheduler.optimizer.param_groups[0]['lr']
            loss = self.train()
            val_idx, val_p = self.test(self.val_loader)
            scheduler.step(val_p[0])
            print(f"Epoch: {epoch:03d}, LR: {lr:.5f}, Loss: {loss:.4f}, ValP@1: {val_p[0]:.4f}, ValP@5: {val_p[1]:.4f}")
            print(time.time() - t)
            test_idx, _ = self.test(self.train_loader)
            torch.save(self.model, os.path.join(out_dir, f"model-{epoch}.pkl"))
            with open(os.path.join(out_dir, f"testServices-epoch{epoch}.txt"), "w") as f:
                json.dump(test_idx + val_idx, f)


def _to(b: SimpleNamespace, device) -> SimpleNamespace:
    out = SimpleNamespace(**vars(b))
    for k, v in vars(out).items():
        if torch.is_tensor(v):
            setattr(out, k, v.to(device, non_blocking=True))
    return out
