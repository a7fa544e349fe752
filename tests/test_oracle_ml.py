"""The ML oracle (oracle/ml_oracle.py) is a restatement -- no reference fixture can pin it (PyG / torch_scatter are not
available).  What CAN be checked without them: the sparse, edge-ordered restatement against an INDEPENDENT dense-matrix
evaluation of the published operators it restates --
  GCN (Kipf & Welling; PyG GCNConv):  out = D^-1/2 (A + I) D^-1/2 (X W) + b,  D_ii = sum_j (A + I)_ij over incoming edges,
                                      self loops of weight 1 added only where the graph has none;
  GIN (Xu et al.; PyG GINConv):       out = MLP((1 + eps) x_i + sum_{j -> i} x_j);
  scatter(reduce='mean'):             segment sums / max(count, 1).
Float64 dense algebra vs the float32 oracle: agreement to float32 round-off."""
import torch

from oracle import ml_oracle as mo


def _graph(n, e, seed, self_loops=False):
    g = torch.Generator().manual_seed(seed)
    src, dst = torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)
    if not self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    else:
        src[:5] = dst[:5]                                    # a few explicit self loops (must not be doubled)
    return torch.stack([src, dst]), torch.rand(src.numel(), generator=g) + 0.1


def _dense_adj(ei, w, n):
    A = torch.zeros(n, n, dtype=torch.float64)               # A[i, j] = total weight of edges j -> i
    A.index_put_((ei[1], ei[0]), w.double(), accumulate=True)
    return A


def test_gcn_conv_equals_dense_normalised_adjacency():
    for self_loops in (False, True):
        n, cin, cout = 40, 6, 9
        ei, w = _graph(n, 300, 3, self_loops)
        conv = mo.GCNConvO(cin, cout)
        conv.bias.data.normal_()
        x = torch.randn(n, cin)
        with torch.no_grad():
            got = conv(x, ei, w).double()
        A = _dense_adj(ei, w, n)
        has_loop = torch.zeros(n, dtype=torch.bool)
        has_loop[ei[0][ei[0] == ei[1]]] = True
        A = A + torch.diag((~has_loop).double())              # add_remaining_self_loops(fill_value=1)
        dinv = A.sum(dim=1).pow(-0.5)                          # in-degree (weights summed at the target)
        dinv[torch.isinf(dinv)] = 0
        ref = (dinv[:, None] * A * dinv[None, :]) @ (x.double() @ conv.weight.data.double()) + conv.bias.data.double()
        assert (got - ref).abs().max() <= 2e-6 * ref.abs().max()


def test_gin_conv_equals_dense_sum_aggregation():
    n, c = 35, 7
    ei, _ = _graph(n, 200, 5)
    mlp = torch.nn.Sequential(torch.nn.Linear(c, 11), torch.nn.ReLU(), torch.nn.Linear(11, 4))
    conv = mo.GINConvO(mlp, train_eps=True)
    conv.eps.data.fill_(0.3)
    x = torch.randn(n, c)
    with torch.no_grad():
        got = conv(x, ei).double()
        A = _dense_adj(ei, torch.ones(ei.shape[1]), n)
        ref = mlp.double()((1 + 0.3) * x.double() + A @ x.double())
    assert (got - ref).abs().max() <= 2e-6 * ref.abs().max()


def test_segment_mean_and_aggregate_sum():
    n, c = 50, 5
    x = torch.randn(n, c)
    seg = torch.sort(torch.randint(0, 6, (n,), generator=torch.Generator().manual_seed(1))).values
    seg[seg == 2] = 3                                          # an empty segment -> zeros (count clamped to 1)
    got = mo.segment_mean(x, seg, 6)
    onehot = torch.nn.functional.one_hot(seg, 6).double()
    ref = (onehot.T @ x.double()) / onehot.sum(0).clamp(min=1)[:, None]
    assert (got.double() - ref).abs().max() <= 1e-6
    ei, w = _graph(n, 400, 8)
    assert (mo.aggregate_sum(x, ei, w).double() - _dense_adj(ei, w, n) @ x.double()).abs().max() <= 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# hand-derived known-answer tests (tests/ml_kats.py): the restatement against arithmetic done on paper
# ---------------------------------------------------------------------------------------------------------------------
import ml_kats as kat


def _close_fr(got, want, tol=2e-7):
    return abs(float(got) - float(want)) <= tol * max(1.0, abs(float(want)))


def test_kat_gcn_norm_existing_duplicate_and_zero_weight_self_loops():
    ei = torch.tensor([[s for s, _, _ in kat.GCN_EDGES], [d for _, d, _ in kat.GCN_EDGES]])
    w = torch.tensor([x for _, _, x in kat.GCN_EDGES])
    ei2, norm = mo.gcn_norm(ei, w, kat.GCN_N)
    assert list(zip(ei2[0].tolist(), ei2[1].tolist())) == kat.GCN_EDGE_ORDER          # kept edges, then loops 0..n-1
    for (s, d), v in zip(kat.GCN_EDGE_ORDER, norm.tolist()):
        assert _close_fr(v, kat.GCN_NORM[(s, d)]), ((s, d), v)
    assert norm[-1].item() == 0.0                                                      # deg = 0 -> inf -> 0, not NaN


def test_kat_gcn_conv_layer():
    conv = mo.GCNConvO(2, 2)
    conv.weight.data = torch.eye(2)
    conv.bias.data = torch.tensor(kat.GCN_BIAS)
    ei = torch.tensor([[s for s, _, _ in kat.GCN_EDGES], [d for _, d, _ in kat.GCN_EDGES]])
    w = torch.tensor([x for _, _, x in kat.GCN_EDGES])
    with torch.no_grad():
        out = conv(torch.tensor(kat.GCN_X), ei, w)
    for i in range(kat.GCN_N):
        for c in range(2):
            assert _close_fr(out[i, c].item(), kat.GCN_OUT[i][c], 3e-7), (i, c, out[i, c].item())


def test_kat_gin_eps_and_multi_edges():
    conv = mo.GINConvO(torch.nn.Sequential(torch.nn.Identity()), True)
    conv.eps.data.fill_(kat.GIN_EPS)
    ei = torch.tensor([[s for s, _ in kat.GIN_EDGES], [d for _, d in kat.GIN_EDGES]])
    with torch.no_grad():
        pre = conv(torch.tensor(kat.GIN_X).view(-1, 1), ei).view(-1)
    assert pre.tolist() == [float(v) for v in kat.GIN_PRE]                             # small integers / halves: exact


def test_kat_segment_mean_with_empty_segments():
    out = mo.segment_mean(torch.tensor(kat.MEAN_X).view(-1, 1), torch.tensor(kat.MEAN_SEG), kat.MEAN_SEGMENTS)
    assert out.view(-1).tolist() == kat.MEAN_OUT


def test_kat_collation_inc_rule_oracle_and_product():
    from types import SimpleNamespace
    from gnnpn_sc_b200 import trainML
    svc = torch.tensor([[s for s, _ in kat.COLLATE_SVC_EDGES], [d for _, d in kat.COLLATE_SVC_EDGES]])
    samples = []
    for n, edges in zip(kat.COLLATE_REQ_NODES, kat.COLLATE_REQ_EDGES):
        samples.append(SimpleNamespace(
            x=torch.zeros(n, 7), y=torch.zeros(kat.COLLATE_S),
            edge_index=torch.tensor([[s for s, _ in edges], [d for _, d in edges]]),
            x_service=torch.zeros(kat.COLLATE_S, 5), edge_index_service=svc, edge_attr_service=torch.ones(2)))
    for collate in (mo.collate, trainML.collate):
        b = collate(samples, True)
        assert b.edge_index.tolist() == kat.COLLATE_EDGE_INDEX and b.batch.tolist() == kat.COLLATE_BATCH
        assert b.edge_index_service.tolist() == kat.COLLATE_SVC_FAITHFUL               # shifted by 3 request nodes
        assert collate(samples, False).edge_index_service.tolist() == kat.COLLATE_SVC_SANE
        assert b.x_service.shape[0] == 2 * kat.COLLATE_S
