"""pytest configuration: markers, import paths, shared fixture loaders."""
import json
import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running (full BASELINE sizes)")


def _cuda_ready():
    """(ok, reason): a CUDA device is visible and the in-tree extension is built."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as exc:                                     # pragma: no cover
        return False, f"torch unavailable: {exc}"
    if not os.path.exists(os.path.join(ROOT, "simrank_b200", "libsimrank_b200.so")):
        return False, "simrank_b200/libsimrank_b200.so is not built"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """Tests marked ``gpu`` are SKIPPED (not errors) on a machine without a CUDA device or without
    the built extension, so a plain ``pytest tests`` is green on a CPU-only box."""
    ok, reason = _cuda_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_notebook():
    return json.load(open(os.path.join(GOLDEN, "notebook_outputs.json")))


def load_ref_index():
    idx = json.load(open(os.path.join(GOLDEN, "ref_index.json")))
    return {k: v for k, v in idx.items() if not k.startswith("_")}, idx


def load_ref_case(name):
    """-> (meta, edge DataFrame, arrays dict) for tests/golden/ref_<name>.npz."""
    cases, _ = load_ref_index()
    meta = cases[name]
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=False)
    df = pd.DataFrame({c: z[c] for c in meta["columns"]})
    arrays = {k: z[k] for k in z.files if k not in meta["columns"]}
    return meta, df, arrays


def notebook_directed_df():
    nb = load_notebook()["inputs_directed"]
    return pd.DataFrame({"ORIGIN_AIRPORT_ID": nb["from"], "DEST_AIRPORT_ID": nb["to"], "flights": nb["flights"]})


def notebook_bipartite_df():
    nb = load_notebook()["inputs_bipartite"]
    return pd.DataFrame({"userId": nb["userId"], "movieId": nb["movieId"], "rating": nb["rating"]})


@pytest.fixture(scope="session")
def notebook():
    return load_notebook()
