"""Reference import layout (``src.models.*``, ``src.ML2PN``, ``src.loadData``) mapped onto gnnpn_sc_b200."""
