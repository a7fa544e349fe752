"""Reference import path shim: ``src/models/trainPNLow.py`` -> gnnpn_sc_b200.trainPN."""
from .trainPN import SCDataset, TrainModel, PNLow, PNHigh  # noqa: F401
