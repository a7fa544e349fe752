from gnnpn_sc_b200.trainPNLow import *  # noqa: F401,F403
from gnnpn_sc_b200 import trainPNLow as _impl
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
