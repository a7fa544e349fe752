import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import ml_oracle as mo
from test_ml_gpu import _setup
from gnnpn_sc_b200 import trainML
samples, ref, net = _setup(8, 120, 4, 2, seed=3)
ref.train(); net.train()
batch = samples[:2]
crit = torch.nn.BCELoss()
out_r = ref(mo.collate(batch)); loss_r = crit(out_r, torch.stack([s.y for s in batch])); loss_r.backward()
data = trainML.collate(batch, device="cuda")
out_g = net(data); loss_g = crit(out_g, data.y.view(2, -1)); loss_g.backward()
print("loss", loss_r.item(), loss_g.item(), "out diff", (out_g.cpu()-out_r).abs().max().item())
gr = dict(ref.named_parameters())
for name, p in net.named_parameters():
    if gr[name].grad is None: continue
    a, b = p.grad.cpu(), gr[name].grad
    print(f"{name:45s} max|g_ref| {b.abs().max():.3e}  max|diff| {(a-b).abs().max():.3e}  rel {((a-b).abs().max()/b.abs().max().clamp(min=1e-12)):.2e}")
