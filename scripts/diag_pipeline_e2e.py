"""Per-batch wall times of ML2PN.run over repeated host batches (QWS shape, 18,944 instances): where does an e2e outlier sit?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gnnpn_sc_b200 import synth, loadData, trainML, modelML, modelPN as M
from gnnpn_sc_b200.pipeline import ML2PN, constraint_arrays
from gnnpn_sc_b200.weights import reference_shaped_state_dict
dev = torch.device("cuda")
cfg = bench.PIPE_SHAPES["qws"]; K, N, S = cfg["K"], cfg["N"], cfg["S"]; n = 18944
ds = synth.ml_dataset(n_instances=256, K=K, S=S, seed=3, dist=cfg["dist"])
samples = trainML.build_samples(loadData.ml_arrays(ds))
reps = (n + len(samples) - 1) // len(samples)
samples_n, nodef = (samples * reps)[:n], (ds["nodefeatures"] * reps)[:n]
torch.manual_seed(0)
net = modelML.Net(128, S, 20, 2, cfg["gcn"], isServices=True).to(dev); net.reset_parameters(); net.eval()
pn = []
for level, seed in (("Low", 1), ("High", 2)):
    m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level=level)
    m.load_state_dict(reference_shaped_state_dict(256, 8, seed)); pn.append(m.to(dev).eval())
svc = type(samples[0])(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(samples[0]).items()})
pipe = ML2PN(net, pn[0], pn[1], svc, ds["serviceFeature"], dev)
host = trainML.collate_requests(samples_n, pin=True)
cons_host = [torch.from_numpy(a).pin_memory() for a in constraint_arrays(nodef, K)]
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ts = []
    for res in pipe.run((host, *cons_host) for _ in range(8)):
        ts.append(time.perf_counter())
    torch.cuda.synchronize()
    gaps = [round(1e3 * (b - a), 2) for a, b in zip([t0] + ts[:-1], ts)]
    print(f"rep {rep}: total {1e3 * (time.perf_counter() - t0):.1f} ms, per-yield gaps {gaps}, "
          f"reserved {torch.cuda.memory_reserved() >> 20} MiB", flush=True)
