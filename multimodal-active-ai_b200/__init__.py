"""B200-native NT-Xent hot path of Multimodal-Active-AI (SimCLR/Objective.py), see DESIGN.md.

The directory name carries a hyphen (repo convention); import it through the ``maai_b200`` shim at
the repo root:  ``import maai_b200; maai_b200.contrastive_loss(...)``.
"""
from . import Model_Util, _lib  # noqa: F401
from .Model_Util import top_k_accuracy  # noqa: F401
from .Objective import LARGE_NUM, GraphedNTXentLoss, NTXentLoss, contrastive_loss, padded_dim  # noqa: F401

__all__ = ["contrastive_loss", "NTXentLoss", "GraphedNTXentLoss", "padded_dim", "LARGE_NUM", "top_k_accuracy"]
