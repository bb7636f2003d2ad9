"""frk_b200 -- B200 (sm_100a) implementation of FRK's Gaussian fixed-rank-kriging hot path.

Public API mirrors the reference's R API for the path (names, argument meaning, error behaviour):
``plane/sphere/real_line``, ``local_basis/auto_basis/TensorP/eval_basis``, ``auto_BAUs``, ``SRE``,
``SRE_fit`` (``SRE.fit`` in R), ``predict``, ``loglik`` (``logLik``).  Everything below the host
control flow is libfrkb200.so (include/frkb200.h); importing this package without the built library
raises -- there is no CPU fallback.
"""
from ._lib import FRKError, LIB_PATH, lib  # noqa: F401  (raises ImportError if the .so is missing)
from .basis import (Basis, BuildD, TensorP, TensorP_Basis, auto_basis, eval_basis, eval_basis_device,  # noqa: F401
                    local_basis, nbasis)
from .geometry import (GridBAUs, Manifold, PointBAUs, PolygonBAUs, STBAUs, SpatialPointsDataFrame, SpatialPolygonsDataFrame, auto_BAUs, map_data_to_BAUs,  # noqa: F401
                       STplane, STsphere, plane, real_line, sphere)
from .runtime import Device, get_device, partition_rows  # noqa: F401
from .sre import SRE, SRE_fit, SREModel, loglik, predict  # noqa: F401

__version__ = "0.1.0"
