"""Run-to-run determinism of the full-wave launches (n = 18,944, 148 CTAs): the blocked encoder and the fused decoder are
run several times on the same inputs and compared bitwise with each other and with the row-major encoder + decoder
(separate pointer phase), which share the canonical arithmetic.  Prints where any difference sits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnnpn_sc_b200 import ops, modelPN as M
from gnnpn_sc_b200.synth import pn_instances
from gnnpn_sc_b200.weights import reference_shaped_state_dict

dev = torch.device("cuda")
K, N = (50, 10) if "--normal" in sys.argv else (47, 5)
n = int(os.environ.get("DIAG_N", "18944"))
reps = int(os.environ.get("DIAG_REPS", "8"))
L = K * N
x = pn_instances(n, K, N, seed=1234, dist="normal" if "--normal" in sys.argv else "qws").to(dev)
mm = M.CombinatorialRL(0, 256, L, 0, 10, 1, M.reward, "Dot", N, K, level="Low")
mm.load_state_dict(reference_shaped_state_dict(256, 8, 1)); mm = mm.cuda().eval()
enc_w, dec_w = mm.actor._packed_weights()
ws = ops.pn_workspace(n, 256, dev, "tc")
ops.set_option("scan", 0)


def decode(enc, c0, lay, with_h):
    c = c0.clone()
    out = (torch.empty(n, K, 256, device=dev) if with_h else None, torch.empty(K, n, device=dev, dtype=torch.int32),
           torch.full((n, L), float("nan"), device=dev), torch.full((n, L), float("nan"), device=dev))
    ops.pn_decode_greedy(x, enc, c, dec_w, K, N, out=out, workspace=ws, enc_layout=lay)
    torch.cuda.synchronize()
    return out[1], out[2], out[3], c


def where(a, b, name):
    bad = (a != b) & ~(torch.isnan(a) & torch.isnan(b))
    if not bad.any():
        return f"{name}: identical"
    nz = bad.nonzero()
    inst = nz[:, 0] if a.shape[0] == n else nz[:, 1]
    col = nz[:, 1] if a.shape[0] == n else nz[:, 0]
    d = (a.double() - b.double()).abs()[bad]
    rows128 = torch.bincount(inst % 128, minlength=128)
    return (f"{name}: {int(bad.sum())} differ, max |d| {d.max().item():.3e}, instances {inst.unique().numel()}, "
            f"CTAs {(inst // 128).unique().tolist()[:12]}, steps {(col // N if a.shape[0] == n else col).unique().tolist()[:12]}, "
            f"rows%128 hist(nonzero) {[(i, int(v)) for i, v in enumerate(rows128.tolist()) if v][:12]}")


# row-major reference (separate pointer phase)
enc_r = ops.enc_out_empty(n, L, 256, ops.ENC_ROWMAJOR, dev)
c_r = torch.empty(n, 256, device=dev)
ops.lstm_encode(x, enc_w, 256, enc_r, c_r, workspace=ws, layout=ops.ENC_ROWMAJOR)
torch.cuda.synchronize()
idx_r, wl_r, wp_r, cf_r = decode(enc_r, c_r, ops.ENC_ROWMAJOR, True)
del enc_r
# blocked encoder, repeated
encs = []
for i in range(2):
    e = ops.enc_out_empty(n, L, 256, ops.ENC_BLOCKED128, dev)
    c = torch.empty(n, 256, device=dev)
    ops.lstm_encode(x, enc_w, 256, e, c, workspace=ws, layout=ops.ENC_BLOCKED128)
    torch.cuda.synchronize()
    encs.append((e, c))
print("encoder run0 vs run1:", "identical" if torch.equal(encs[0][0], encs[1][0]) and torch.equal(encs[0][1], encs[1][1]) else "DIFFER",
      "| final c vs row-major:", "identical" if torch.equal(encs[0][1], c_r) else "DIFFER", flush=True)
enc_b, c_b = encs[0]
del encs
for r in range(reps):
    idx, wl, wp, cf = decode(enc_b, c_b, ops.ENC_BLOCKED128, False)
    print(f"fused run {r}: ", where(wl, wl_r, "win_logits vs row-major"), "|", where(idx.float(), idx_r.float(), "idx"), "|",
          where(cf, cf_r, "final c"), flush=True)
ops.set_option("scan", -1)
