"""GPU parity of the stage glue: candidate selection on the device against the numpy restatement of
``loadDataPN`` (gnnpn_sc_b200/loadData.py::pn_rows_from_ranking, ranking order, no shuffle), and the whole
ML -> candidates -> PNLow -> PNHigh pipeline against the stage-by-stage path."""
import numpy as np
import pytest
import torch

from oracle import pn_oracle as po

pytestmark = pytest.mark.gpu


def _dataset(K, S, n, seed=0):
    from gnnpn_sc_b200 import synth, loadData, trainML, modelML
    ds = synth.ml_dataset(n_instances=n, K=K, S=S, seed=seed, min_tasks=min(4, K))
    samples = trainML.build_samples(loadData.ml_arrays(ds))
    torch.manual_seed(seed)
    net = modelML.Net(128, S, 20, 2, 2, isServices=True)
    net.reset_parameters()
    return ds, samples, net.cuda().eval()


def _mirror_rows(ds, rankings, N):
    from gnnpn_sc_b200 import loadData
    sf = ds["serviceFeature"]
    ser2cat, ser2pos = [], []
    for key in sf.keys():
        ser2cat += [int(key) - 1] * len(sf[key])
        ser2pos += list(range(len(sf[key])))
    ser2cat, ser2pos = np.asarray(ser2cat), np.asarray(ser2pos)
    return np.asarray([loadData.pn_rows_from_ranking(nf, rk, sf, ser2cat, ser2pos, N, rng=False)
                       for nf, rk in zip(ds["nodefeatures"], rankings)], dtype=np.float32)


@pytest.mark.parametrize("K,S,N", [(12, 300, 5), (47, 2507, 5), (6, 40, 10), (3, 3000, 1000), (8, 400, 8), (5, 300, 3), (9, 500, 16), (7, 420, 17), (4, 240, 32)])
def test_select_candidates_matches_loaddatapn(K, S, N):
    from gnnpn_sc_b200 import ops
    from gnnpn_sc_b200.pipeline import constraint_arrays, service_arrays
    n = 9
    ds, _, _ = _dataset(K, S, n, seed=K)
    g = torch.Generator().manual_seed(S)
    scores = torch.rand(n, S, generator=g)
    m = (S // 7) * 7
    scores[:, 0:m:7] = scores[:, 1:m:7]                                       # exact ties: lower service id first
    rankings = torch.sort(scores, dim=1, descending=True, stable=True).indices.tolist()
    want = _mirror_rows(ds, rankings, N)                                      # [n, K*N, 9] with the category column
    local, used, glob = constraint_arrays(ds["nodefeatures"], K)
    qos, ptr = service_arrays(ds["serviceFeature"])
    rows, picked = ops.select_candidates(scores.cuda(), torch.from_numpy(qos).cuda(), torch.from_numpy(ptr).cuda(),
                                         torch.from_numpy(local).cuda(), torch.from_numpy(used).cuda(),
                                         torch.from_numpy(glob).cuda(), N, with_category=True, return_picked=True)
    assert np.array_equal(rows.cpu().numpy(), want)                           # bit-exact rows
    rows8 = ops.select_candidates(scores.cuda(), torch.from_numpy(qos).cuda(), torch.from_numpy(ptr).cuda(),
                                  torch.from_numpy(local).cuda(), torch.from_numpy(used).cuda(),
                                  torch.from_numpy(glob).cuda(), N)
    assert np.array_equal(rows8.cpu().numpy(), want[:, :, 1:])
    pk = picked.cpu().numpy()
    neutral = (want[:, :, 1:5] == np.array([0, 1, 1, 1], dtype=np.float32)).all(axis=2)
    assert np.array_equal(pk < 0, neutral) or (pk[neutral] < 0).all()          # every neutral row has no service id
    ok = pk >= 0
    assert np.array_equal(qos[pk[ok]], want[:, :, 1:5][ok])                   # ids point at the rows' QoS values


def test_ml2pn_pipeline_equals_stage_by_stage():
    from gnnpn_sc_b200 import modelPN as M, ops, trainML
    from gnnpn_sc_b200.pipeline import ML2PN, constraint_arrays
    K, S, N, n = 12, 300, 5, 16
    ds, samples, net = _dataset(K, S, n, seed=4)
    cfg = po.PNConfig(seq_len=K * N, s_number=N, s_category=K)
    nets = []
    for level, seed in (("Low", 1), ("High", 2)):
        m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level=level)
        m.load_state_dict(po.make_state_dict(cfg, seed))
        nets.append(m.cuda().eval())
    low, high = nets
    dev_sample = type(samples[0])(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(samples[0]).items()})
    pipe = ML2PN(net, low, high, dev_sample, ds["serviceFeature"], "cuda")
    local, used, glob = (torch.from_numpy(a).cuda() for a in constraint_arrays(ds["nodefeatures"], K))
    batch = trainML.collate(samples, faithful_quirk=False, device="cuda")
    out = pipe.compose(batch, local, used, glob)
    # stage by stage: Net.forward on S-offset collated pairs (the reference's batch size 2), host loadDataPN mirror, PNs
    with torch.no_grad():
        ref_scores = torch.cat([net(trainML.collate(samples[i:i + 2], faithful_quirk=False, device="cuda"))
                                for i in range(0, n, 2)])
    assert (out["scores"] - ref_scores).abs().max() <= 1e-5
    rankings = torch.sort(out["scores"].cpu(), dim=1, descending=True, stable=True).indices.tolist()
    want_rows = _mirror_rows(ds, rankings, N)[:, :, 1:]
    assert np.array_equal(out["rows"].cpu().numpy(), want_rows)
    x = torch.from_numpy(want_rows).cuda()
    with torch.no_grad():
        _, _, _, _, latent = low(x, None, sample="greedy", training="SL")
        R, _, _, idx, _ = high(x, None, latent, sample="greedy", training="RL")
    assert torch.equal(torch.stack(idx), out["idx_high"]) and torch.equal(R, out["reward"])
    # the chosen services respect the task windows and the requests' categories
    svc = out["services"].cpu().numpy()
    used_np = used.cpu().numpy().astype(bool)
    assert ((svc >= 0) == used_np).all() or ((svc >= 0) <= used_np).all()
    ptr = pipe.cat_ptr.cpu().numpy()
    for b in range(n):
        for k in range(K):
            if svc[b, k] >= 0:
                assert ptr[k] <= svc[b, k] < ptr[k + 1]


def test_ml2pn_run_from_host_batches_equals_compose():
    """ML2PN.run (pinned host batches, uploads on a copy stream, results pipelined by one batch) yields, in order, exactly
    what compose() gives on the same batches -- including a last batch of another size."""
    from gnnpn_sc_b200 import modelPN as M, trainML
    from gnnpn_sc_b200.pipeline import ML2PN, constraint_arrays
    K, S, N, n = 10, 200, 4, 12
    ds, samples, net = _dataset(K, S, n, seed=6)
    cfg = po.PNConfig(seq_len=K * N, s_number=N, s_category=K)
    nets = []
    for level, seed in (("Low", 3), ("High", 4)):
        m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level=level)
        m.load_state_dict(po.make_state_dict(cfg, seed))
        nets.append(m.cuda().eval())
    dev_sample = type(samples[0])(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(samples[0]).items()})
    pipe = ML2PN(net, nets[0], nets[1], dev_sample, ds["serviceFeature"], "cuda")
    cons = constraint_arrays(ds["nodefeatures"], K)
    parts = [(0, 5), (5, 10), (10, 12)]                      # three batches, the last one smaller
    host = [(trainML.collate_requests(samples[a:b], pin=True), *[torch.from_numpy(c[a:b]).pin_memory() for c in cons])
            for a, b in parts]
    got = list(pipe.run(host))
    assert len(got) == len(parts)
    for (a, b), (svc, obj) in zip(parts, got):
        rb = trainML.collate_requests(samples[a:b], device="cuda")
        want = pipe.compose(rb, *[torch.from_numpy(c[a:b]).cuda() for c in cons])
        assert torch.equal(svc, want["services"].to(torch.int32).cpu())
        assert torch.equal(obj, want["objective"].cpu(), ) or torch.allclose(obj, want["objective"].cpu(), equal_nan=True)
