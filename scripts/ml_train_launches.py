"""Two TrainML training iterations (batch of 2 request graphs + the service graph, train-mode BatchNorm, Adam) on a small
synthetic dataset -- run under `ncu --metrics gpu__time_duration.sum` to list the kernels of an ML training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import synth, loadData, trainML, modelML
ds = synth.ml_dataset(n_instances=8, K=47, S=2507, seed=1)
samples = trainML.build_samples(loadData.ml_arrays(ds))
dev = torch.device("cuda")
torch.manual_seed(0)
net = modelML.Net(128, 2507, 20, 2, 2, isServices=True).to(dev)
net.reset_parameters(); net.train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
crit = torch.nn.BCELoss()
for i in range(0, 4, 2):
    b = trainML.collate(samples[i:i + 2], faithful_quirk=True, device=dev)
    opt.zero_grad()
    loss = crit(net(b), b.y.view(2, -1))
    loss.backward(); opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
