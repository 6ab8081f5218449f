"""odevio_b200 -- B200-native (sm_100a) implementation of ODE-VIO's latent-dynamics integration
hot path: the PoseODERNN / PoseCDE regressors behind the reference's own module interfaces.

Public surface (mirrors reference src/models/{PoseODERNN,PoseCDE,ODEFunc,FusionModule}.py):
    PoseODERNN, PoseCDE, ODEFunc, CDEFunc, FusionModule; InertialEncoder (src/models/Encoder.py:39-74)
C ABI underneath: include/odevio.h  ->  odevio_b200/lib/libodevio_b200.so (build: odevio_b200.build)
"""

from .modules import PoseODERNN, PoseCDE, ODEFunc, CDEFunc, FusionModule, InertialEncoder  # noqa: F401
from ._lib import OdevioError, LIB_PATH  # noqa: F401
from .streaming import StreamingPoseODERNN  # noqa: F401

__all__ = ["PoseODERNN", "PoseCDE", "ODEFunc", "CDEFunc", "FusionModule", "InertialEncoder", "OdevioError", "LIB_PATH", "StreamingPoseODERNN"]
