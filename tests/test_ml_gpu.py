"""GPU parity of the drop-in ``Net`` (gnnpn_sc_b200/modelML.py) against the restated oracle
(oracle/ml_oracle.py -- PARITY UNPINNED: PyG 1.7.0 / torch_scatter 2.0.6 are not available, see its header).
Scores within 1e-5 relative; ranking agreement on the top-k; both collations (faithful PyG quirk and S-offset)."""
import numpy as np
import pytest
import torch

from oracle import ml_oracle as mo

pytestmark = pytest.mark.gpu


def _setup(K, S, n_inst, gcn_layers, seed=0, is_services=True):
    from gnnpn_sc_b200 import synth, loadData, trainML, modelML
    ds = synth.ml_dataset(n_instances=n_inst, K=K, S=S, seed=seed, min_tasks=min(4, K))
    arrays = loadData.ml_arrays(ds)
    samples = trainML.build_samples(arrays)
    torch.manual_seed(seed)
    ref = mo.NetO(128, S, 20, 2, gcn_layers, isServices=is_services)
    ref.reset_parameters()
    for m in ref.modules():                     # non-trivial eval-mode BatchNorm
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    net = modelML.Net(128, S, 20, 2, gcn_layers, isServices=is_services)
    missing = net.load_state_dict(ref.state_dict(), strict=True)
    return samples, ref, net.cuda()


@pytest.mark.parametrize("quirk", [True, False])
@pytest.mark.parametrize("K,S,gcn", [(8, 120, 2), (47, 2507, 2), (12, 300, 4)])
def test_net_forward_eval_matches_oracle(K, S, gcn, quirk):
    from gnnpn_sc_b200 import trainML
    samples, ref, net = _setup(K, S, 6, gcn)
    ref.eval(); net.eval()
    for lo in (0, 2):
        batch = samples[lo:lo + 2]
        with torch.no_grad():
            want = ref(mo.collate(batch, faithful_quirk=quirk))
            got = net(trainML.collate(batch, faithful_quirk=quirk, device="cuda")).cpu()
        assert got.shape == want.shape == (2, S)
        err = (got - want).abs() / want.abs().clamp(min=1e-3)
        print(f"Net K={K} S={S} gcn={gcn} quirk={quirk}: max |dscore| {(got - want).abs().max():.2e}, max rel {err.max():.2e}")
        assert (got - want).abs().max() <= 1e-5
        top_w, top_g = want.topk(10, dim=1).indices, got.topk(10, dim=1).indices
        for r in range(2):                      # same top-10 set unless scores tie within tolerance
            diff = set(top_w[r].tolist()) ^ set(top_g[r].tolist())
            assert all(abs(float(want[r, i]) - float(want[r, top_w[r, -1]])) < 1e-5 for i in diff)


def test_net_without_service_graph_branch():
    """``isServices=False`` (modelML.py:157-162: plain Linear layers instead of GCNConv on the service side) -- never
    taken by TrainML, but part of the ``Net`` interface."""
    from gnnpn_sc_b200 import trainML
    samples, ref, net = _setup(12, 300, 4, 2, is_services=False)
    ref.eval(); net.eval()
    with torch.no_grad():
        want = ref(mo.collate(samples[:2]))
        got = net(trainML.collate(samples[:2], device="cuda")).cpu()
    assert got.shape == want.shape and (got - want).abs().max() <= 1e-5


def test_net_train_step_gradients_match_oracle():
    """fwd + BCE + bwd through the CUDA aggregation kernels (autograd Function) vs the oracle's autograd."""
    from gnnpn_sc_b200 import trainML
    samples, ref, net = _setup(8, 120, 4, 2, seed=3)
    ref.train(); net.train()
    batch = samples[:2]
    crit = torch.nn.BCELoss()
    out_r = ref(mo.collate(batch))
    loss_r = crit(out_r, torch.stack([s.y for s in batch]))
    loss_r.backward()
    data = trainML.collate(batch, device="cuda")
    out_g = net(data)
    loss_g = crit(out_g, data.y.view(2, -1))
    loss_g.backward()
    assert abs(loss_r.item() - loss_g.item()) < 1e-5
    gr = dict(ref.named_parameters())
    worst = 0.0
    for name, p in net.named_parameters():
        if gr[name].grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        # biases that feed a BatchNorm have an exactly-zero true gradient (1e-8 noise on both sides): compare on
        # an absolute floor instead of dividing noise by noise
        d = (p.grad.cpu() - gr[name].grad).abs().max() / gr[name].grad.abs().max().clamp(min=1e-3)
        worst = max(worst, float(d))
    print(f"train-mode gradient max relative deviation {worst:.2e}")
    assert worst < 1e-4      # achieved 3.5e-5 on a B200 (profiles/r02_parity.json)


def test_trainml_smoke(tmp_path):
    """Two epochs of the drop-in TrainML on a tiny synthetic dataset: files in the reference's layout."""
    import json, os
    from gnnpn_sc_b200 import synth, loadData, trainML
    ds = synth.ml_dataset(n_instances=16, K=6, S=60, seed=1, min_tasks=3)
    t = trainML.TrainML("tiny", 2, 2, 128, 20, 0.0, 0.001, 2, root=str(tmp_path))
    t.start(arrays=loadData.ml_arrays(ds))
    with open(os.path.join(tmp_path, "solutions", "ML", "tiny", "testServices-epoch1.txt")) as f:
        rank = json.load(f)
    assert len(rank) == 16 and sorted(rank[0]) == list(range(60))


def _reference_test_loop(x, y, log=(1, 5)):
    """trainML.py:49-72 restated on CPU tensors: per-row descending sort, P@k counting, list of index lists."""
    idx_list, total = [], [[] for _ in log]
    for _x, _y in zip(x, y):
        pat = [0] * len(log)
        _, indices = _x.sort(dim=0, descending=True, stable=True)      # the reference's CPU sort keeps ties in index order
        for k in range(len(log)):
            for idx in indices[:log[k]]:
                if _y[idx] == 1:
                    pat[k] += 1
        idx_list.append(indices.numpy().tolist())
        for k in range(len(log)):
            total[k].append(pat[k] / log[k])
    return idx_list, [float(np.average(t)) for t in total]


def test_ranking_and_precision_at_k_match_reference_loop_with_ties():
    """f3: the batched device ranking / P@1 / P@5 of TrainML.test against the reference's per-row python loop
    (trainML.py:49-72) on the SAME scores, with exact ties injected (scores quantised to 1/8 and constant rows)."""
    from gnnpn_sc_b200 import trainML
    g = torch.Generator().manual_seed(3)
    B, S = 48, 257
    x = torch.rand(B, S, generator=g)
    x[: B // 2] = (x[: B // 2] * 8).floor() / 8          # many exact ties
    x[0] = 0.5                                            # all equal: ranking = index order, P@k = label prefix
    y = (torch.rand(B, S, generator=g) < 0.05).float()
    y[0, :5] = torch.tensor([1, 0, 1, 0, 0.])
    order, (p1, p5) = trainML.precision_at(x.cuda(), y.cuda())
    want_idx, (w1, w5) = _reference_test_loop(x, y)
    assert order.cpu().numpy().tolist() == want_idx
    assert abs(float(p1.mean()) - w1) < 1e-7 and abs(float(p5.mean()) - w5) < 1e-7
    assert order[0].cpu().tolist() == list(range(S)) and float(p1[0]) == 1.0 and abs(float(p5[0]) - 0.4) < 1e-7


def test_trainml_test_rankings_and_file_layout(tmp_path):
    """TrainML.test end to end: rankings == descending stable sort of the model's own scores, P@k == the reference loop
    on them; testServices-epoch{e}.txt = train rankings (loader order) + validation rankings, S ints per instance
    (trainML.py:146-149)."""
    import json, os
    from gnnpn_sc_b200 import synth, loadData, trainML
    ds = synth.ml_dataset(n_instances=12, K=6, S=60, seed=2, min_tasks=3)
    t = trainML.TrainML("tiny", 2, 2, 128, 20, 0.0, 0.001, 1, root=str(tmp_path))
    t.start(arrays=loadData.ml_arrays(ds))
    idx, (p1, p5) = t.test(t.val_loader)
    scores, labels = [], []
    t.model.eval()
    with torch.no_grad():
        for data in t.val_loader:
            d = trainML._to(data, t.device)
            scores.append(t.model(d).view(d.num_graphs, -1).cpu())
            labels.append(d.y.view(d.num_graphs, -1).cpu())
    want_idx, (w1, w5) = _reference_test_loop(torch.cat(scores), torch.cat(labels))
    assert idx == want_idx and abs(p1 - w1) < 1e-6 and abs(p5 - w5) < 1e-6
    with open(os.path.join(tmp_path, "solutions", "ML", "tiny", "testServices-epoch0.txt")) as f:
        rank = json.load(f)
    assert len(rank) == 12 and all(sorted(r) == list(range(60)) for r in rank)
    assert rank[9:] == idx                               # validation quarter last, in loader order


@pytest.mark.parametrize("M,C,relu", [(97, 256, True), (5014, 256, True), (60, 128, False), (3, 40, True)])
def test_bn_train_kernels_match_torch(M, C, relu):
    """gnnpn_bn_train_forward/backward_f32 against torch.nn.BatchNorm1d (training mode) + ReLU in float64 on the CPU:
    output, running statistics, dx / dgamma / dbeta."""
    from gnnpn_sc_b200 import ops
    from conftest import record_parity
    g = torch.Generator().manual_seed(M + C)
    y = torch.randn(M, C, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    dout = torch.randn(M, C, generator=g)
    bn = torch.nn.BatchNorm1d(C).double().train()
    bn.weight.data, bn.bias.data = gamma.double().clone(), beta.double().clone()
    y64 = y.double().requires_grad_(True)
    ref = bn(y64)
    ref = torch.relu(ref) if relu else ref
    ref.backward(dout.double())
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    out, mean, rstd = ops.bn_train_forward(y.cuda(), gamma.cuda(), beta.cuda(), 1e-5, 0.1, relu, rm, rv)
    dx, dg, db = ops.bn_train_backward(y.cuda(), out, dout.cuda(), gamma.cuda(), mean, rstd, relu)
    rel = lambda a, b: float(((a.cpu().double() - b).abs() / b.abs().clamp(min=1)).max())
    errs = {"out": rel(out, ref.detach()), "running_mean": rel(rm, bn.running_mean), "running_var": rel(rv, bn.running_var),
            "dx": rel(dx, y64.grad), "dgamma": rel(dg, bn.weight.grad), "dbeta": rel(db, bn.bias.grad)}
    record_parity(f"bn_train_M{M}_C{C}_relu{int(relu)}", tolerance=2e-5, **errs)
    assert max(errs.values()) <= 2e-5, errs
