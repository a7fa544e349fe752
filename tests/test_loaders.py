"""Host-side data layer (gnnpn_sc_b200/loadData.py) against the reference's own loaders executed live
(only where /root/reference is mounted) and against structural invariants everywhere."""
import json
import os
import sys

import numpy as np
import pytest

from gnnpn_sc_b200 import loadData as ld
from gnnpn_sc_b200 import synth


@pytest.fixture(scope="module")
def tiny(tmp_path_factory):
    root = tmp_path_factory.mktemp("ds")
    ds = synth.ml_dataset(n_instances=12, K=7, S=70, seed=5, min_tasks=3)
    synth.write_dataset(str(root), "tiny", ds)
    rng = np.random.default_rng(0)
    rankings = [rng.permutation(70).tolist() for _ in range(12)]
    os.makedirs(root / "solutions" / "ML" / "tiny", exist_ok=True)
    with open(root / "solutions" / "ML" / "tiny" / "testServices-epoch0.txt", "w") as f:
        json.dump(rankings, f)
    return root, ds


def test_pn_rows_layout(tiny):
    root, ds = tiny
    feats, labels = ld.loadDataPN(0, "tiny", 3, root=str(root), rng=False)
    assert len(feats) == 12 and labels == ds["minCostList"]
    a = np.asarray(feats[0])
    assert a.shape == (7 * 3, 9)
    assert (a[:, 0] == np.repeat(np.arange(7), 3)).all()
    assert (a[3:, 5:] == 0).all()                                   # global bounds only on category 0 rows
    neutral = (a[:, 1:5] == [0, 1, 1, 1]).all(axis=1)
    used = {int(np.argmax(np.asarray(n[:-6]) == 1)) for n in ds["nodefeatures"][0] if n[0] != 1}
    assert set(np.unique(a[~neutral, 0]).astype(int) + 1) <= used


def test_loaders_equal_live_reference(tiny, reference_available, monkeypatch):
    if not reference_available:
        pytest.skip("reference tree not mounted (GPU box)")
    root, ds = tiny
    sys.path.insert(0, "/root/reference")
    try:
        for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[m]
        import importlib
        ref = importlib.import_module("src.loadData")
    finally:
        sys.path.pop(0)
    monkeypatch.chdir(root)
    got = ld.loadData("tiny", root=str(root))
    want = ref.loadData("tiny")
    for g, w in zip(got[:6], want[:6]):
        assert g == w or np.allclose(np.asarray(g, dtype=float), np.asarray(w, dtype=float), rtol=0, atol=0)
    np.testing.assert_allclose(got[6], want[6])
    np.random.seed(123)
    f_ref, l_ref = ref.loadDataPN(epoch=0, dataset="tiny", serviceNumber=3)
    np.random.seed(123)
    f_got, l_got = ld.loadDataPN(0, "tiny", 3, root=str(root))
    assert l_got == l_ref
    assert np.array_equal(np.asarray(f_got, dtype=float), np.asarray(f_ref, dtype=float))
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[m]
