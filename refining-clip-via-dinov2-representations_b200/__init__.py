"""B200-native DINO-Soft loss path (drop-in for the reference's ``open_clip.loss.ClipLossWithDINOEnhancements``).

Import as ``dinosoft_b200`` (see the shim at the repo root).  Contents:

    loss.py      the nn.Module with the reference's ctor/forward signature, gather_features,
                 compute_student_tau, install_into_open_clip
    feature_store.py  device-resident DINO feature store: bf16 table in HBM, gather kernel with on-device range
                 check writing straight into the packed operand buffer
    cyclip.py    CyCLIPLoss drop-in (loss.py:813-905): CLIP term on the kernels, consistency terms through D x D
                 moment matrices
    graphed.py   CUDA-graph capture of the loss forward + backward (launch-bound small batches)
    pair_stats.py  CLIP-blind pair statistics on the Gram-tile kernel (helpers.py:221-285)
    _cabi.py     ctypes binding of libdsoft.so (include/dsoft.h)
    _build.py    nvcc recipe for csrc/ (sm_100a only)
    csrc/        hand-written tcgen05 / TMEM / TMA kernels + the C ABI
"""
from . import _build, _cabi
from ._build import build
from .feature_store import DinoFeatureStore, DinoRows, lookup as dino_lookup, to_device_table
from .cyclip import CyCLIPLoss
from .graphed import make_graphed
from .pair_stats import pair_stats
from .loss import (
    ClipLossWithDINOEnhancements,
    CudaBackend,
    compute_student_tau,
    gather_features,
    install_into_open_clip,
    uninstall_from_open_clip,
)

__version__ = "0.1.0"
__all__ = [
    "ClipLossWithDINOEnhancements",
    "CudaBackend",
    "compute_student_tau",
    "gather_features",
    "install_into_open_clip",
    "uninstall_from_open_clip",
    "build",
    "DinoFeatureStore",
    "DinoRows",
    "to_device_table",
    "dino_lookup",
    "pair_stats",
    "CyCLIPLoss",
    "make_graphed",
]
