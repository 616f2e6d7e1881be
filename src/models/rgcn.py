"""Drop-in for the reference's ``src/models/rgcn.py``: same three classes, served by the B200 kernels.

``from models.rgcn import DrugDiseaseModel`` (reference src/train.py:26-28) and
``from src.models.rgcn import ...`` both resolve here when this repo's ``src/`` replaces the reference's.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from primekg_rgcn_linkprediction_b200 import (DrugDiseaseModel, DrugDiseaseRGCN, LinkPredictor,  # noqa: E402,F401
                                               RGCNConv)

__all__ = ["DrugDiseaseRGCN", "LinkPredictor", "DrugDiseaseModel", "RGCNConv"]
