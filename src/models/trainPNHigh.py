from gnnpn_sc_b200.trainPNHigh import *  # noqa: F401,F403
from gnnpn_sc_b200 import trainPNHigh as _impl
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
