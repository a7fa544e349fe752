"""gnnpn_sc_b200 -- the B200 (sm_100a) hot path of GNNPN-SC behind the reference's module interfaces.

Numerics policy: fp32 end to end.  The reference pins torch 1.8.1, whose CPU path is strict fp32; torch 2.x lets
cuDNN use TF32 for fp32 RNNs/convolutions by default, which would put ~5e-4 relative error into the library-backed
gradient replay (``PointerNet.replay_action_probs``), so it is switched off here for the whole process.
"""
import torch as _torch

_torch.backends.cudnn.allow_tf32 = False
_torch.backends.cuda.matmul.allow_tf32 = False
