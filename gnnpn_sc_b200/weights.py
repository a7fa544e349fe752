"""Random-initialised pointer-network weights with the reference's key names and init family
(modelPN.py:153-163: nn.Linear / nn.LSTM defaults, U(-1/sqrt(H), 1/sqrt(H)) start token), drawn from
numpy PCG64 so a (seed) names the same bytes on every box.  There are no pretrained weights to load:
the reference ships none (solutions/pretrained/.gitkeep)."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def reference_shaped_state_dict(hidden: int = 256, in_features: int = 8, seed: int = 0,
                                gain: float = 1.0) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    H, Fin = hidden, in_features
    b = 1.0 / math.sqrt(H)

    def u(*shape, bound=b, g=1.0):
        return torch.from_numpy((rng.uniform(-bound, bound, size=shape) * g).astype(np.float32))

    sd = {"actor.decoder_start_input": u(H)}
    sd["actor.embedding2.weight"] = u(H, Fin, bound=1.0 / math.sqrt(Fin))
    sd["actor.embedding2.bias"] = u(H, bound=1.0 / math.sqrt(Fin))
    for rnn in ("encoder", "decoder"):
        sd[f"actor.{rnn}.weight_ih_l0"] = u(4 * H, H, g=gain)
        sd[f"actor.{rnn}.weight_hh_l0"] = u(4 * H, H, g=gain)
        sd[f"actor.{rnn}.bias_ih_l0"] = u(4 * H)
        sd[f"actor.{rnn}.bias_hh_l0"] = u(4 * H)
    return sd
