import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def load_val_graph():
    z = np.load(os.path.join(GOLDEN, "val_graph.npz"))
    return dict(edge_index=torch.from_numpy(z["edge_index"].astype(np.int64)),
                edge_type=torch.from_numpy(z["edge_type"].astype(np.int64)),
                num_nodes=int(z["num_nodes"]), num_relations=int(z["num_relations"]),
                rows=torch.from_numpy(z["rows"].astype(np.int64)), emb_rows=torch.from_numpy(z["emb_rows"]),
                emb_colsum=torch.from_numpy(z["emb_colsum"]), seed=int(z["seed"]), conv_scale=float(z["conv_scale"]))


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) the C-ABI library; nvcc cross-compiles without a GPU."""
    import __graft_entry__ as entry
    entry.build()
    import primekg_rgcn_linkprediction_b200 as pkg
    return pkg
