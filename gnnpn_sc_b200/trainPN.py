"""Drop-ins for the reference's ``src/models/trainPNLow.py`` and ``trainPNHigh.py``: ``PNLow(...).start()`` and
``PNHigh(...).start()`` with the same constructor arguments, files and REINFORCE update
(trainPNLow.py:70-146, trainPNHigh.py:70-151):

    b <- mean(R) on the first batch, then beta*b + (1-beta)*mean(R);  advantage = R - b
    logp = sum_k log p_k(action_k),  logp[logp < -1000] = 0;  loss = mean(advantage * logp)
    clip_grad_norm_(actor, max_grad_norm);  Adam over model.actor.parameters()

Sampling / greedy decoding and the reward run in the CUDA kernels; the gradient comes from the windowed
replay of ``PointerNet.replay_action_probs``.  With ``torch.distributed`` initialised the batch is sharded over
ranks and gradients + the reward mean are all-reduced (``gnnpn_sc_b200.parallel``) so replicas stay identical.
Plotting (matplotlib/IPython in the reference) is skipped when those packages are absent.
Kept quirks: PNLow validates with sampling (trainPNLow.py:131), PNHigh greedily (trainPNHigh.py:139);
``allActions`` has K+2 slots for PNLow (trainPNLow.py:122); PNHigh saves PNLow's weights next to its own.
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.optim as optim
from torch.utils.data import DataLoader, Dataset

from . import parallel
from .loadData import loadDataPN
from .modelPN import CombinatorialRL, reward


class SCDataset(Dataset):
    """trainPNLow.py:15-42: drops the leading category column unless ``embeddingTag``."""

    def __init__(self, dataset, targets, embeddingTag=False):
        super().__init__()
        self.data_set = [torch.FloatTensor([row if embeddingTag else row[1:] for row in data]) for data in dataset]
        self.label = list(targets)
        self.serviceNumbers = [0] * len(self.data_set)
        self.size = len(self.data_set)

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return self.data_set[idx], self.label[idx]


class TrainModel:
    def __init__(self, model, train_dataset, val_dataset, epochDiv, beta, USE_CUDA, dataset, serCategory, lr=0.5e-4,
                 batch_size=128, threshold=None, max_grad_norm=2., low_model=None, root=".", level="PNLow"):
        self.model, self.low_model = model, low_model
        self.train_dataset, self.val_dataset = train_dataset, val_dataset
        self.batch_size, self.threshold, self.epochDiv, self.beta = batch_size, threshold, epochDiv, beta
        self.USE_CUDA, self.dataset, self.serCategory = USE_CUDA, dataset, serCategory
        # data parallel: replicas must start identical and shuffle identically (each rank then takes its slice of the
        # same batch, parallel.shard) -- weights broadcast from rank 0, loader generator seeded with a shared seed
        parallel.broadcast_parameters(model.actor.parameters())
        self._loader_gen = torch.Generator()
        self._loader_gen.manual_seed(parallel.shared_seed())
        self.train_loader = DataLoader(train_dataset, batch_size=batch_size, shuffle=True, num_workers=0,
                                       generator=self._loader_gen)
        self.val_loader = DataLoader(val_dataset, batch_size=batch_size, shuffle=False, num_workers=0)
        self.actor_optim = optim.Adam(model.actor.parameters(), lr=lr)
        self.max_grad_norm = max_grad_norm
        self.train_tour, self.val_tour = [], []
        self.epochs = 0
        self.out_dir = os.path.join(root, "solutions", level, dataset)
        self.level = level

    def reinforce_step(self, inputs, labs, baseline, first):
        """One REINFORCE update; returns (mean reward, new baseline).  With several ranks ``inputs`` is this rank's
        shard (possibly empty for a ragged last batch): the loss is normalised by the GLOBAL batch size and gradients
        are summed, so the update equals the single-process one on the whole batch."""
        params = list(self.model.actor.parameters())
        self.actor_optim.zero_grad()
        if inputs.shape[0]:
            latent = None
            if self.low_model is not None:
                # the reference back-propagates into PNLow's graph but never uses that gradient
                # (trainPNHigh.py:83, optimiser over model.actor only): decode it without autograd
                with torch.no_grad():
                    _, _, _, _, latent = self.low_model(inputs, labs, sample="greedy", training="SL")
            R, probs, actions, actions_idxs, _ = self.model(inputs, labs, latent)
        else:
            R, probs = torch.zeros(0, device=inputs.device), []
        r_mean, n_global = parallel.global_mean_count(R)
        baseline = r_mean if first else baseline * self.beta + (1. - self.beta) * r_mean
        if inputs.shape[0]:
            advantage = R - baseline
            logprobs = 0
            for prob in probs:
                logprobs = logprobs + torch.log(prob)
            logprobs = torch.where(logprobs < -1000, torch.zeros_like(logprobs), logprobs)
            actor_loss = (advantage * logprobs).sum() / n_global              # == .mean() on one process
            actor_loss.backward()
        parallel.allreduce_gradients(params, average=False)
        torch.nn.utils.clip_grad_norm_(params, float(self.max_grad_norm), norm_type=2)
        self.actor_optim.step()
        return r_mean, baseline.detach()

    def train_and_validate(self, n_epochs, epochDiv):
        baseline = torch.zeros(1, device="cuda")
        t = time.time()
        os.makedirs(self.out_dir, exist_ok=True)
        for epoch in range(1, n_epochs + 1):
            for batch_id, (sample_batch, labs) in enumerate(self.train_loader):
                self.model.train()
                inputs = parallel.shard(sample_batch).cuda(non_blocking=True)
                r_mean, baseline = self.reinforce_step(inputs, labs, baseline, batch_id == 0)
                self.train_tour.append(float(r_mean))
            if self.threshold and self.train_tour[-1] < self.threshold:
                print("EARLY STOPPAGE!")
                break
            if epoch % epochDiv == 0:
                self.checkpoint_and_validate(epoch, epochDiv, t)
            self.epochs += 1

    def checkpoint_and_validate(self, epoch, epochDiv, t0):
        n = self.epochs // epochDiv
        if parallel.rank() == 0:
            torch.save({"epoch": epoch, "model": self.model.state_dict(), "optimizer": self.actor_optim.state_dict()},
                       os.path.join(self.out_dir, f"epoch{n}.model"))
            if self.low_model is not None:
                torch.save({"epoch": epoch, "model": self.low_model.state_dict(),
                            "optimizer": self.actor_optim.state_dict()}, os.path.join(self.out_dir, f"epoch{n}_low.model"))
        self.model.eval()
        high = self.low_model is not None
        if high:
            self.low_model.eval()
        K = self.serCategory
        allActions = [[] for _ in range(K if high else K + 2)]
        allR = []
        with torch.no_grad():
            for val_batch, labs in self.val_loader:
                inputs = val_batch.cuda(non_blocking=True)
                if high:
                    _, _, _, _, latent = self.low_model(inputs, labs, sample="greedy", training="SL")
                    R, probs, actions, _, _ = self.model(inputs, labs, latent, sample="greedy")
                else:
                    R, probs, actions, _, _ = self.model(inputs, labs, sample="sample")
                acts = torch.stack(actions).cpu().numpy().tolist()          # one D2H for the batch
                for a in range(len(acts)):
                    allActions[a] += acts[a]
                allR += R.cpu().numpy().tolist()
                self.val_tour.append(R.mean().item())
        if parallel.rank() != 0:
            return
        with open(os.path.join(self.out_dir, f"allActions{n}.txt"), "w") as f:
            json.dump(allActions, f)
        if not high:
            with open(os.path.join(self.out_dir, f"allR{n}.txt"), "w") as f:
                if allR:                                                    # trainPNLow.py:123-141
                    json.dump({"quality": allR, "averageQ": sum(allR) / len(allR)}, f)
        print(time.time() - t0)
        with open(os.path.join(self.out_dir, f"val{n}.txt"), "w") as f:
            json.dump(self.val_tour, f)
        if high:
            with open(os.path.join(self.out_dir, f"time{n}.txt"), "w") as f:
                json.dump([time.time() - t0], f)
        self.plot(n)

    def plot(self, n):
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return
        plt.figure(figsize=(20, 5))
        plt.subplot(131); plt.plot(self.train_tour[-2000:]); plt.grid()
        plt.subplot(132); plt.plot(self.val_tour); plt.grid()
        plt.savefig(os.path.join(self.out_dir, f"epoch{n}.png"))
        plt.close()


def _make_model(level, embedding_size, cfg):
    return CombinatorialRL(embedding_size, cfg.hidden_size, cfg.serCategory * cfg.serNumber, cfg.n_glimpses,
                           cfg.tanh_exploration, cfg.use_tanh, reward, attention="Dot", level=level,
                           use_cuda=cfg.USE_CUDA, sNumber=cfg.serNumber, sCategory=cfg.serCategory)


class PNLow:
    """trainPNLow.py:169-223."""
    epochs = 50

    def __init__(self, dataset, embeddingTag, USE_CUDA, serCategory, epochDiv, serNumber, hidden_size, n_glimpses,
                 tanh_exploration, use_tanh, beta, max_grad_norm, lr, epochML, root="."):
        self.dataset = dataset + "/"
        self.embeddingTag, self.USE_CUDA, self.serCategory, self.epochDiv = embeddingTag, USE_CUDA, serCategory, epochDiv
        self.serNumber, self.hidden_size, self.n_glimpses = serNumber, hidden_size, n_glimpses
        self.tanh_exploration, self.use_tanh, self.beta = tanh_exploration, use_tanh, beta
        self.max_grad_norm, self.lr, self.epochML, self.root = max_grad_norm, lr, epochML, root

    def _datasets(self, data=None):
        feats, labels = data if data is not None else loadDataPN(
            epoch=self.epochML, dataset=self.dataset[:-1], serviceNumber=self.serNumber, root=self.root)
        split = len(feats) // 4 * 3
        return (SCDataset(feats[:split], labels[:split], self.embeddingTag),
                SCDataset(feats[split:], labels[split:], self.embeddingTag))

    def start(self, data=None, n_epochs=None):
        train_ds, val_ds = self._datasets(data)
        if self.embeddingTag:
            self.dataset += "20embeddings/"
        model = _make_model("Low", 20 if self.embeddingTag else 0, self).cuda()
        self.trainer = TrainModel(model, train_ds, val_ds, self.epochDiv, self.beta, self.USE_CUDA, self.dataset,
                                  self.serCategory, self.lr, 128, None, self.max_grad_norm, root=self.root, level="PNLow")
        self.trainer.train_and_validate(n_epochs or self.epochs, self.epochDiv)
        return self.trainer


class PNHigh(PNLow):
    """trainPNHigh.py:175-251."""
    epochs = 100

    def __init__(self, dataset, embeddingTag, USE_CUDA, serCategory, epochDiv, serNumber, hidden_size, n_glimpses,
                 tanh_exploration, use_tanh, beta, max_grad_norm, lr, epochML, epochPNLow, root="."):
        super().__init__(dataset, embeddingTag, USE_CUDA, serCategory, epochDiv, serNumber, hidden_size, n_glimpses,
                         tanh_exploration, use_tanh, beta, max_grad_norm, lr, epochML, root)
        self.epochPNLow = epochPNLow

    def start(self, data=None, n_epochs=None, low_state=None):
        train_ds, val_ds = self._datasets(data)
        if self.embeddingTag:
            self.dataset += "20embeddings/"
        emb = 20 if self.embeddingTag else 0
        model_low, model_high = _make_model("Low", emb, self), _make_model("High", emb, self)
        if low_state is None:
            if self.epochPNLow >= 0:
                path = os.path.join(self.root, "solutions", "PNLow", self.dataset, f"epoch{self.epochPNLow}.model")
            else:
                path = os.path.join(self.root, "solutions", "pretrained", f"{self.dataset[:-1]}-PNLow.model")
            low_state = torch.load(path, map_location="cpu")["model"]
        model_low.load_state_dict(low_state)
        self.trainer = TrainModel(model_high.cuda(), train_ds, val_ds, self.epochDiv, self.beta, self.USE_CUDA,
                                  self.dataset, self.serCategory, self.lr, 128, None, self.max_grad_norm,
                                  low_model=model_low.cuda(), root=self.root, level="PNHigh")
        self.trainer.train_and_validate(n_epochs or self.epochs, self.epochDiv)
        return self.trainer
