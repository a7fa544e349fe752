"""Generate ``tests/golden/pn_*.npz`` by running the REAL reference.

Run in the build container only (``/root/reference`` is mounted read-only there
and does not exist on the GPU box):

    python oracle/make_golden.py

The reference's ``src/models/modelPN.py`` is imported unmodified; its one
hard-coded ``.cuda()`` (modelPN.py:151) is neutralised by an identity patch of
``torch.Tensor.cuda`` for the duration of the import/run.  Weights come from
``oracle.pn_oracle.make_state_dict`` (numpy PCG64) and are loaded with
``load_state_dict`` so a fixture is fully described by (config, seeds) plus the
reference's outputs.  Nothing of the reference's source is written anywhere.
"""
from __future__ import annotations

import io
import contextlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pn_oracle import PNConfig, make_state_dict  # noqa: E402
from gnnpn_sc_b200.synth import pn_instances  # noqa: E402

REFERENCE = "/root/reference"

CASES = {
    # name: (cfg kwargs, batch, weight seeds (low, high), input seed, gain, dist)
    "qws_b4": (dict(s_category=47, s_number=5), 4, (101, 202), 1234, 1.0, "qws"),
    "normal_b2": (dict(s_category=50, s_number=10), 2, (103, 204), 4321, 1.0, "normal"),
    "small_b16": (dict(s_category=6, s_number=4), 16, (105, 206), 7, 1.0, "qws"),
    "sharp_b4": (dict(s_category=12, s_number=5), 4, (107, 208), 8, 3.0, "qws"),
    "bahdanau_b3": (dict(s_category=8, s_number=4, attention="Bahdanau"), 3, (109, 210), 9, 1.0, "qws"),
    "glimpse_b3": (dict(s_category=6, s_number=3, n_glimpses=1), 3, (111, 212), 10, 1.0, "qws"),
    "embed20_b3": (dict(s_category=6, s_number=3, embedding_size=20), 3, (113, 214), 11, 1.0, "qws"),
    "notanh_b3": (dict(s_category=6, s_number=3, use_tanh=False), 3, (115, 216), 12, 1.0, "qws"),
}


def build_inputs(cfg: PNConfig, B: int, seed: int, dist: str) -> torch.Tensor:
    x = pn_instances(B, cfg.s_category, cfg.s_number, seed=seed, dist=dist)
    if cfg.embedding_size:
        cat = torch.arange(cfg.s_category).repeat_interleave(cfg.s_number).float()
        x = torch.cat([cat.view(1, -1, 1).expand(B, -1, 1), x], dim=2)
    return x.contiguous()


def case_config(name: str) -> PNConfig:
    kw = dict(CASES[name][0])
    kw["seq_len"] = kw["s_category"] * kw["s_number"]
    return PNConfig(**kw)


def import_reference():
    if not os.path.isdir(REFERENCE):
        raise SystemExit("reference tree not present; fixtures can only be generated in the build container")
    torch.Tensor.cuda = lambda self, *a, **k: self      # modelPN.py:151 workaround, CPU run
    # loaded by file path under a private name: this repo ships its own ``src.models.modelPN`` shim
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_modelPN",
                                                  os.path.join(REFERENCE, "src", "models", "modelPN.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def run_reference(ref, name: str):
    kw, B, (s_lo, s_hi), s_in, gain, dist = CASES[name]
    cfg = case_config(name)
    x = build_inputs(cfg, B, s_in, dist)
    out = {}
    models = {}
    for level, seed in (("Low", s_lo), ("High", s_hi)):
        m = ref.CombinatorialRL(cfg.embedding_size, cfg.hidden_size, cfg.seq_len, cfg.n_glimpses,
                                cfg.tanh_exploration, int(cfg.use_tanh), ref.reward, cfg.attention,
                                cfg.s_number, cfg.s_category, use_cuda=False, level=level)
        missing = m.load_state_dict(make_state_dict(cfg, seed, gain), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        m.eval()
        models[level] = m
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        pr_lo, ap_lo, act_lo, idx_lo, latent = models["Low"](x, None, sample="greedy", training="SL")
        r_lo = ref.reward(act_lo, None, cfg.s_category, USE_CUDA=False, level="Low",
                          embedding_size=cfg.embedding_size)
        pr_hi, ap_hi, act_hi, idx_hi, lg_hi = models["High"](x, None, latent, sample="greedy", training="SL")
        r_hi = ref.reward(act_hi, None, cfg.s_category, USE_CUDA=False, level="High",
                          embedding_size=cfg.embedding_size)
        r_hi_as_low = ref.reward(act_hi, None, cfg.s_category, USE_CUDA=False, level="Low",
                                 embedding_size=cfg.embedding_size)
    # the reference itself evaluated in float64 (same modules, .double()): used to state how far the
    # reference's own fp32 run is from exact arithmetic when a stress case amplifies rounding noise
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for m in models.values():
            m.double()
            m.actor.alpha = m.actor.alpha.double()
        _, _, _, idx_lo64, latent64 = models["Low"](x.double(), None, sample="greedy", training="SL")
        _, _, _, idx_hi64, lg_hi64 = models["High"](x.double(), None, latent64, sample="greedy", training="SL")
    same64 = bool((torch.stack(idx_lo64) == torch.stack(idx_lo)).all() and (torch.stack(idx_hi64) == torch.stack(idx_hi)).all())
    out["picks_equal_in_f64"] = np.array(same64)
    if torch.stack(latent64).numel() <= 20000:          # keep the big-shape fixtures small
        out.update(logits_low_f64=torch.stack(latent64).numpy(), logits_high_f64=torch.stack(lg_hi64).numpy())
    out.update(
        objfunc_high=np.array([ref.calc([act_hi[k][b][0:4] for k in range(cfg.s_category)] if cfg.embedding_size == 0 else
                                        [act_hi[k][b][1:5] for k in range(cfg.s_category)],
                                        [[[0.0, 2.0]], [[0.0, 2.0]]], cfg.s_category)[1] for b in range(B)], dtype=np.float32),
        idx_low=torch.stack(idx_lo).numpy().astype(np.int32),
        logits_low=torch.stack(latent).numpy(),
        action_probs_low=torch.stack(ap_lo).numpy(),
        idx_high=torch.stack(idx_hi).numpy().astype(np.int32),
        logits_high=torch.stack(lg_hi).numpy(),
        action_probs_high=torch.stack(ap_hi).numpy(),
        actions_high=torch.stack(act_hi).numpy(),
        reward_low=r_lo.numpy(), reward_high=r_hi.numpy(), viol_high=r_hi_as_low.numpy(),
        probs_high_step1=pr_hi[min(1, cfg.s_category - 1)].numpy(),
    )
    return out


def main():
    ref = import_reference()
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    for name in CASES:
        res = run_reference(ref, name)
        path = os.path.join(gold, f"pn_{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
