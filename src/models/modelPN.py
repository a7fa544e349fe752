from gnnpn_sc_b200.modelPN import *  # noqa: F401,F403
from gnnpn_sc_b200 import modelPN as _impl
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
