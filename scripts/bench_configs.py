"""Secondary measurements for the BASELINE.json configs that are not the bench.py headline:

  ml        config 0: ML candidate-reduction GNN on synthetic QWS-shaped data (K=47, S=2507) -- GPU Net inference
            (reference batching of 2 graphs, and batched requests with the service side encoded once) and a
            training pass (batch 2, train-mode BatchNorm, Adam), with the restated reference (oracle NetO) timed on
            the host cores beside it on a bounded sample;
  pipeline  config 2: full ML+2PN inference on Normal-distributed QoS (K=50, N=10, L=500, GCN=4): Net scores ->
            device candidate selection -> PNLow -> PNHigh -> objective, per-stage CUDA-event times;
  scaleup   config 3: 100 tasks x 1000 candidates (L = 100,000): encoder scan + general decode per instance shard.

    python scripts/bench_configs.py ml|pipeline|scaleup [--out gpurun_out/x.jsonl]
One JSON line per measurement (CUDA events, >= 3 warm-up passes, inputs larger than L2 or L2 flushed between runs)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def ev_time(fn, iters, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def emit(out, **kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if out:
        with open(out, "a") as f:
            f.write(line + "\n")


def pn_pair(K, N, dev):
    from gnnpn_sc_b200 import modelPN as M
    from gnnpn_sc_b200.weights import reference_shaped_state_dict
    nets = []
    for level, seed in (("Low", 1), ("High", 2)):
        m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level=level)
        m.load_state_dict(reference_shaped_state_dict(256, 8, seed))
        nets.append(m.to(dev).eval())
    return nets


def run_ml(args):
    from gnnpn_sc_b200 import synth, loadData, trainML, modelML
    from oracle import ml_oracle as mo                       # CPU baseline leg only
    dev = torch.device("cuda")
    K, S, n_inst = 47, 2507, args.instances
    ds = synth.ml_dataset(n_instances=n_inst, K=K, S=S, seed=0)
    arrays = loadData.ml_arrays(ds)
    samples = trainML.build_samples(arrays)
    E_s = samples[0].edge_index_service.shape[1]
    torch.manual_seed(0)
    net = modelML.Net(128, S, 20, 2, 2, isServices=True).to(dev)
    net.reset_parameters()
    net.eval()
    flush = torch.empty(64 << 20, device=dev)                # 256 MB > 126 MB L2
    # (a) reference batching: 2 graphs per forward, the whole service graph re-encoded per sample (trainML.py:121)
    pairs = [trainML.collate(samples[i:i + 2], faithful_quirk=True, device=dev) for i in range(0, 64, 2)]
    def infer_pairs():
        with torch.no_grad():
            for b in pairs:
                net(b)
    med, best = ev_time(infer_pairs, 5, flush=flush)
    emit(args.out, config="ml_infer_batch2_faithful", K=K, S=S, service_edges=E_s, graphs=64, ms=med,
         graphs_per_s=64 / (med * 1e-3))
    # (b) batched requests, service side encoded once
    svc_sample = type(samples[0])(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(samples[0]).items()})
    big = trainML.collate(samples[:min(n_inst, 1024)], faithful_quirk=False, device=dev)
    xs = net.service_encodings(svc_sample)
    nb = big.num_graphs
    med, best = ev_time(lambda: net.score_requests(big, xs), 10, flush=flush)
    emit(args.out, config="ml_infer_batched_service_cached", K=K, S=S, graphs=nb, ms=med, graphs_per_s=nb / (med * 1e-3))
    med, best = ev_time(lambda: net.service_encodings(svc_sample), 10, flush=flush)
    emit(args.out, config="ml_service_encodings_once", S=S, service_edges=E_s, ms=med)
    # (c) training pass, batch 2 (train-mode BatchNorm), Adam -- GPU
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    crit = torch.nn.BCELoss()
    def train_steps():
        for b in pairs:
            opt.zero_grad()
            loss = crit(net(b), b.y.view(2, -1))
            loss.backward(); opt.step()
    med, best = ev_time(train_steps, 3, warm=1)
    emit(args.out, config="ml_train_batch2", iterations=len(pairs), ms=med, it_per_s=len(pairs) / (med * 1e-3))
    # (d) CPU: restated reference (oracle NetO, PARITY UNPINNED: PyG absent) on the host cores, bounded sample
    torch.set_num_threads(os.cpu_count() or 1)
    ref = mo.NetO(128, S, 20, 2, 2, isServices=True); ref.reset_parameters()
    cpu_pairs = [mo.collate(samples[i:i + 2]) for i in range(0, 16, 2)]
    ref.eval()
    t0 = time.perf_counter()
    with torch.no_grad():
        for b in cpu_pairs:
            ref(b)
    t_inf = time.perf_counter() - t0
    ref.train()
    opt_c = torch.optim.Adam(ref.parameters(), lr=1e-3)
    t0 = time.perf_counter()
    for b in cpu_pairs:
        opt_c.zero_grad()
        loss = crit(ref(b), torch.stack([s.y for s in samples[:2]]) if False else b.y.view(2, -1))
        loss.backward(); opt_c.step()
    t_tr = time.perf_counter() - t0
    emit(args.out, config="ml_cpu_reference_restatement", cores=os.cpu_count(), sample="8 batches of 2 graphs",
         infer_graphs_per_s=16 / t_inf, train_it_per_s=8 / t_tr, kind="port")


def run_pipeline(args):
    from gnnpn_sc_b200 import synth, loadData, trainML, modelML
    from gnnpn_sc_b200.pipeline import ML2PN, constraint_arrays
    dev = torch.device("cuda")
    K, N, S = 50, 10, 2500
    n = args.instances
    ds = synth.ml_dataset(n_instances=min(n, 512), K=K, S=S, seed=3, dist="normal")
    base = len(ds["nodefeatures"])
    samples = trainML.build_samples(loadData.ml_arrays(ds))
    reps = (n + base - 1) // base
    samples_n = (samples * reps)[:n]
    nodef = (ds["nodefeatures"] * reps)[:n]
    torch.manual_seed(0)
    net = modelML.Net(128, S, 20, 2, 4, isServices=True).to(dev)
    net.reset_parameters(); net.eval()
    low, high = pn_pair(K, N, dev)
    svc_sample = type(samples[0])(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(samples[0]).items()})
    pipe = ML2PN(net, low, high, svc_sample, ds["serviceFeature"], dev)
    batch = trainML.collate(samples_n, faithful_quirk=False, device=dev)
    local, used, glob = (torch.from_numpy(a).to(dev) for a in constraint_arrays(nodef, K))
    flush = torch.empty(64 << 20, device=dev)
    med, best = ev_time(lambda: pipe.compose(batch, local, used, glob), args.iters, flush=flush)
    out = pipe.compose(batch, local, used, glob)
    # per-stage
    from gnnpn_sc_b200 import ops
    t_ml, _ = ev_time(lambda: net.score_requests(batch, pipe.service_enc), args.iters, flush=flush)
    t_sel, _ = ev_time(lambda: ops.select_candidates(out["scores"], pipe.svc_qos, pipe.cat_ptr, local, used, glob, N),
                       args.iters, flush=flush)
    def pn_only():
        with torch.no_grad():
            _, _, _, _, lat = low(out["rows"], None, sample="greedy", training="SL")
            high(out["rows"], None, lat, sample="greedy", training="RL")
    t_pn, _ = ev_time(pn_only, args.iters, flush=flush)
    emit(args.out, config="ml2pn_pipeline_normal", K=K, N=N, L=K * N, S=S, gcn_layers=4, instances=n, ms=med,
         instances_per_s=n / (med * 1e-3), stage_ms={"net_scores": t_ml, "select_candidates": t_sel, "pnlow_pnhigh": t_pn},
         mean_violations=float(out["violations"].float().mean()), mean_objective=float(out["objective"].mean()))


def run_scaleup(args):
    from gnnpn_sc_b200.synth import pn_instances
    dev = torch.device("cuda")
    K, N = args.K, args.N
    n = args.instances
    low, high = pn_pair(K, N, dev)
    x = pn_instances(n, K, N, seed=11).to(dev)
    def both():
        with torch.no_grad():
            _, _, _, _, lat = low(x, None, sample="greedy", training="SL")
            return high(x, None, lat, sample="greedy", training="RL")
    med, best = ev_time(both, args.iters, warm=1)
    def enc_only():
        from gnnpn_sc_b200 import ops
        ew, _ = low.actor._packed_weights()
        ws = ops.pn_workspace(n, 256, dev, None)
        ops.lstm_encode(x, ew, 256, workspace=ws)
    t_enc, _ = ev_time(enc_only, args.iters, warm=1)
    emit(args.out, config="scaleup_pnlow_pnhigh", K=K, N=N, L=K * N, instances=n, ms=med, instances_per_s=n / (med * 1e-3),
         encoder_ms=t_enc, enc_out_gb=n * K * N * 1024 / 1e9,
         note="one GPU's shard; instances are independent, so G GPUs run G shards with no communication")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["ml", "pipeline", "scaleup"])
    ap.add_argument("--instances", type=int, default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--N", type=int, default=1000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.instances is None:
        a.instances = {"ml": 256, "pipeline": 4096, "scaleup": 256}[a.what]
    {"ml": run_ml, "pipeline": run_pipeline, "scaleup": run_scaleup}[a.what](a)
