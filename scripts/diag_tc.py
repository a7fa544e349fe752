"""GPU-box bring-up diagnostic for the tcgen05 3xTF32 kernels (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops

def gemm_case(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(a.double(), w.double(), b.double())
    y = ops.gemm_bias_act(a.cuda(), w.cuda(), bias=b.cuda(), impl="tc").cpu().double()
    yf = ops.gemm_bias_act(a.cuda(), w.cuda(), bias=b.cuda(), impl="ffma").cpu().double()
    err = (y - ref).abs().max().item(); errf = (yf - ref).abs().max().item()
    tf = torch.nn.functional.linear((a.view(torch.int32) & ~0x1fff).view(torch.float32).double(), (w.view(torch.int32) & ~0x1fff).view(torch.float32).double(), b.double())
    print(f"gemm M={M} N={N} K={K}: tc max err {err:.3e} (ffma {errf:.3e}; plain-tf32 would be {(tf-ref).abs().max():.3e}) ref max {ref.abs().max():.2f}", flush=True)
    if err > 1e-4:
        bad = (y - ref).abs() > 1e-4
        print("   bad fraction", bad.float().mean().item(), "bad rows", bad.any(1).nonzero().flatten()[:10].tolist(), "bad cols", bad.any(0).nonzero().flatten()[:10].tolist())
        print("   y[0,:8]", y[0,:8].tolist()); print("   r[0,:8]", ref[0,:8].tolist())

for shp in [(128, 256, 32), (128, 16, 8), (300, 128, 256), (1000, 256, 24), (5014, 256, 256), (4096, 1024, 288), (130, 2507, 128)]:
    gemm_case(*shp)
    torch.cuda.synchronize()
print("gemm done", flush=True)

# LSTM recurrence: tc vs ffma on the same packed weights
from gnnpn_sc_b200 import modelPN as M
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict
for n, K, N in [(64, 6, 4), (300, 47, 5)]:
    x = pn_instances(n, K, N, seed=5).cuda()
    m = M.CombinatorialRL(0, 256, K * N, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
    m.load_state_dict(reference_shaped_state_dict(256, 8, 77)); m = m.cuda().eval()
    outs = {}
    for impl in ("ffma", "tc"):
        m.actor.impl = impl
        with torch.no_grad():
            _, idx, lg = m.actor(x, None, sample="greedy")
        torch.cuda.synchronize()
        outs[impl] = (m.actor.last["enc_out"].clone(), torch.stack(idx), m.actor.last["win_logits"].clone(), m.actor.last["dec_h"].clone())
    e = (outs["tc"][0] - outs["ffma"][0]).abs()
    print(f"lstm n={n} K={K} N={N}: enc_out tc-vs-ffma max {e.max():.3e} (t=0 {e[:,0].max():.2e}, t=1 {e[:,1].max():.2e}, last {e[:,-1].max():.2e}); dec_h {(outs['tc'][3]-outs['ffma'][3]).abs().max():.3e}; "
          f"win_logits {(outs['tc'][2]-outs['ffma'][2]).abs().max():.3e}; picks differ {(outs['tc'][1]!=outs['ffma'][1]).sum().item()}", flush=True)
