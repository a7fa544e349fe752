from gnnpn_sc_b200.WOA import *  # noqa: F401,F403
from gnnpn_sc_b200 import WOA as _impl
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith('__')})
