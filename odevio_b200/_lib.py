"""ctypes binding of libodevio_b200.so (the C ABI declared in include/odevio.h).

PyTorch only supplies device memory and the current stream; no torch types cross the
boundary.  There is no CPU fallback: if the library cannot be loaded, or a tensor is not a
contiguous fp32 CUDA tensor, the call raises.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ODEVIO_LIB_PATH") or os.path.join(_HERE, "lib", "libodevio_b200.so")

MAX_ODE_LINEARS = 6
MAX_RNN_LAYERS = 4

ACT = {"tanh": 0, "relu": 1, "leaky_relu": 2, "softplus": 3}
RNN = {"rnn": 0, "gru": 1}
SOLVER = {"dopri5": 0, "tsit5": 1, "heun": 2, "euler": 3, "rk4": 4, "rk4_38": 5}
PRECISION = {"fp32": 0, "tf32x3": 1, "fp16x3": 2}     # ODEVIO_PRECISION_*
STATUS_OK, STATUS_MAX_STEPS, STATUS_INFINITE_NORM = 0, 1, 2


class OdeRnnCfg(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("S", C.c_int32), ("D", C.c_int32), ("H", C.c_int32),
        ("n_hidden", C.c_int32), ("L", C.c_int32), ("activation", C.c_int32),
        ("rnn_type", C.c_int32), ("solver", C.c_int32), ("substeps", C.c_int32),
        ("atol", C.c_float), ("rtol", C.c_float), ("dt0", C.c_float),
        ("safety", C.c_float), ("factor_min", C.c_float), ("factor_max", C.c_float),
        ("accept_strict", C.c_int32), ("floor_factor", C.c_int32), ("endpoint_dense", C.c_int32),
        ("max_steps", C.c_int32), ("precision", C.c_int32), ("save_checkpoints", C.c_int32),
        ("rows_per_tile", C.c_int32), ("exact_landing", C.c_int32), ("trace_steps", C.c_int32),
        ("ckpt_loops", C.c_int32), ("evolve_only", C.c_int32), ("weights_prepacked", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


_FP = C.c_void_p


class OdeRnnWeights(C.Structure):
    _fields_ = [
        ("ode_w", _FP * MAX_ODE_LINEARS), ("ode_b", _FP * MAX_ODE_LINEARS),
        ("rnn_w_ih", _FP * MAX_RNN_LAYERS), ("rnn_w_hh", _FP * MAX_RNN_LAYERS),
        ("rnn_b_ih", _FP * MAX_RNN_LAYERS), ("rnn_b_hh", _FP * MAX_RNN_LAYERS),
        ("reg_w0", _FP), ("reg_b0", _FP), ("reg_w1", _FP), ("reg_b1", _FP),
        ("fuse_w", _FP), ("fuse_b", _FP),
    ]


class CdeCfg(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("S", C.c_int32), ("So", C.c_int32), ("Hc", C.c_int32),
        ("n_layers", C.c_int32), ("activation", C.c_int32), ("solver", C.c_int32), ("interp", C.c_int32),
        ("atol", C.c_float), ("rtol", C.c_float), ("step_size", C.c_double),
        ("max_steps", C.c_int32), ("rows_per_tile", C.c_int32), ("precision", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class CdeWeights(C.Structure):
    _fields_ = [
        ("cde_w", _FP * MAX_ODE_LINEARS), ("cde_b", _FP * MAX_ODE_LINEARS),
        ("init_w", _FP), ("init_b", _FP),
        ("reg_w0", _FP), ("reg_b0", _FP), ("reg_w1", _FP), ("reg_b1", _FP),
    ]


class CdeGrads(C.Structure):
    _fields_ = CdeWeights._fields_


CDE_SOLVER = {"dopri5": 0, "rk4": 1}
CDE_INTERP = {"linear": 0, "cubic": 1}


class OdeRnnGrads(C.Structure):
    _fields_ = OdeRnnWeights._fields_


class ImuEncoderWeights(C.Structure):
    """odevio_imu_encoder_weights (include/odevio.h)."""
    _fields_ = [
        ("conv_w", C.c_void_p * 3), ("conv_b", C.c_void_p * 3),
        ("bn_weight", C.c_void_p * 3), ("bn_bias", C.c_void_p * 3), ("bn_mean", C.c_void_p * 3), ("bn_var", C.c_void_p * 3),
        ("bn_eps", C.c_float), ("proj_w", C.c_void_p), ("proj_b", C.c_void_p),
    ]


ABI_VERSION = 4
_lib = None


class OdevioError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OdevioError(
            f"{LIB_PATH} is missing: build it with `python -m odevio_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    lib.odevio_version.restype = C.c_int32
    lib.odevio_error_string.restype = C.c_char_p
    lib.odevio_error_string.argtypes = [C.c_int32]
    lib.odevio_odernn_default_cfg.restype = None
    lib.odevio_odernn_default_cfg.argtypes = [C.POINTER(OdeRnnCfg)]
    lib.odevio_odernn_workspace_bytes.restype = C.c_size_t
    lib.odevio_odernn_workspace_bytes.argtypes = [C.POINTER(OdeRnnCfg)]
    lib.odevio_odernn_forward.restype = C.c_int32
    lib.odevio_odernn_forward.argtypes = [
        C.POINTER(OdeRnnCfg), C.POINTER(OdeRnnWeights),
        _FP, _FP, C.c_int32, _FP, _FP, _FP, _FP, _FP, _FP, _FP, C.c_size_t, _FP, C.c_size_t, _FP]
    lib.odevio_odernn_geometry.restype = C.c_int32
    lib.odevio_odernn_geometry.argtypes = [C.POINTER(OdeRnnCfg), C.POINTER(C.c_int32)]
    lib.odevio_odernn_ckpt_bytes.restype = C.c_size_t
    lib.odevio_odernn_ckpt_bytes.argtypes = [C.POINTER(OdeRnnCfg)]
    lib.odevio_odernn_backward_workspace_bytes.restype = C.c_size_t
    lib.odevio_odernn_backward_workspace_bytes.argtypes = [C.POINTER(OdeRnnCfg), C.c_int64]
    lib.odevio_odernn_backward.restype = C.c_int32
    lib.odevio_odernn_backward.argtypes = [
        C.POINTER(OdeRnnCfg), C.POINTER(OdeRnnWeights), _FP, _FP, C.c_int32, _FP, C.c_size_t,
        _FP, C.c_int64, _FP, _FP, C.POINTER(OdeRnnGrads), _FP, _FP, _FP, C.c_size_t, _FP]
    lib.odevio_odernn_backward_range.restype = C.c_int32
    lib.odevio_odernn_backward_range.argtypes = [
        C.POINTER(OdeRnnCfg), C.POINTER(OdeRnnWeights), _FP, _FP, C.c_int32, _FP, C.c_size_t,
        _FP, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _FP, _FP, C.POINTER(OdeRnnGrads), _FP, _FP, _FP, C.c_size_t, _FP]
    lib.odevio_cde_default_cfg.restype = None
    lib.odevio_cde_default_cfg.argtypes = [C.POINTER(CdeCfg)]
    lib.odevio_cde_workspace_bytes.restype = C.c_size_t
    lib.odevio_cde_workspace_bytes.argtypes = [C.POINTER(CdeCfg)]
    lib.odevio_cde_forward.restype = C.c_int32
    lib.odevio_cde_forward.argtypes = [
        C.POINTER(CdeCfg), C.POINTER(CdeWeights), _FP, _FP, _FP, C.c_int32, _FP, _FP,
        _FP, _FP, _FP, _FP, _FP, C.c_size_t, _FP]
    lib.odevio_cde_ckpt_bytes.restype = C.c_size_t
    lib.odevio_cde_ckpt_bytes.argtypes = [C.POINTER(CdeCfg), C.c_int32]
    lib.odevio_cde_forward_ckpt.restype = C.c_int32
    lib.odevio_cde_forward_ckpt.argtypes = [
        C.POINTER(CdeCfg), C.POINTER(CdeWeights), _FP, _FP, _FP, C.c_int32, _FP, _FP,
        _FP, _FP, _FP, _FP, _FP, C.c_size_t, C.c_int32, _FP, C.c_size_t, _FP]
    lib.odevio_cde_backward_workspace_bytes.restype = C.c_size_t
    lib.odevio_cde_backward_workspace_bytes.argtypes = [C.POINTER(CdeCfg), C.c_int32]
    lib.odevio_cde_backward.restype = C.c_int32
    lib.odevio_cde_backward.argtypes = [
        C.POINTER(CdeCfg), C.POINTER(CdeWeights), _FP, _FP, _FP, C.c_int32, _FP, C.c_int32, _FP, _FP,
        _FP, C.c_size_t, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
        _FP, _FP, C.POINTER(CdeGrads), _FP, _FP, _FP, C.c_size_t, _FP]
    lib.odevio_odefunc_workspace_bytes.restype = C.c_size_t
    lib.odevio_odefunc_workspace_bytes.argtypes = [C.c_int32] * 4
    lib.odevio_odefunc_forward.restype = C.c_int32
    lib.odevio_odefunc_forward.argtypes = [C.c_int32] * 5 + [C.POINTER(_FP), C.POINTER(_FP), _FP, _FP, _FP, C.c_size_t, _FP]
    lib.odevio_mlp_forward.restype = C.c_int32
    lib.odevio_mlp_forward.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(_FP), C.POINTER(_FP), _FP, _FP, _FP]
    lib.odevio_imu_encoder_workspace_bytes.restype = C.c_size_t
    lib.odevio_imu_encoder_workspace_bytes.argtypes = [C.c_int32]
    lib.odevio_imu_encoder_forward.restype = C.c_int32
    lib.odevio_imu_encoder_forward.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(ImuEncoderWeights), _FP, _FP,
                                               _FP, C.c_size_t, _FP]
    lib.odevio_train_glue_workspace_bytes.restype = C.c_size_t
    lib.odevio_train_glue_workspace_bytes.argtypes = []
    lib.odevio_pose_loss.restype = C.c_int32
    lib.odevio_pose_loss.argtypes = [C.c_int64, _FP, _FP, C.c_float, C.c_float, _FP, _FP, _FP, C.c_size_t, _FP]
    lib.odevio_adam_step.restype = C.c_int32
    lib.odevio_adam_step.argtypes = [C.c_int64, _FP, _FP, _FP, _FP, C.c_int32] + [C.c_float] * 6 + [_FP, _FP, C.c_size_t, _FP]
    lib.odevio_adam_step_groups.restype = C.c_int32
    lib.odevio_adam_step_groups.argtypes = ([C.c_int64, C.c_int64, _FP, _FP, _FP, _FP, C.c_int32] + [C.c_float] * 7 +
                                            [_FP, _FP, C.c_size_t, _FP])
    lib.odevio_allreduce_adam_peer_workspace_bytes.restype = C.c_size_t
    lib.odevio_allreduce_adam_peer_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
    lib.odevio_allreduce_adam_peer.restype = C.c_int32
    lib.odevio_allreduce_adam_peer.argtypes = ([C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(_FP), C.POINTER(_FP),
                                                C.POINTER(_FP), _FP, _FP, C.c_int32, C.c_uint32] + [C.c_float] * 8 +
                                               [_FP, _FP, C.c_size_t, _FP])
    lib.odevio_debug_tc_geometry.restype = C.c_int32
    lib.odevio_debug_tc_geometry.argtypes = [C.POINTER(C.c_int32)]
    lib.odevio_debug_tc_timing.restype = C.c_int32
    lib.odevio_debug_tc_timing.argtypes = [C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.odevio_microbench_ffma.restype = C.c_int32
    lib.odevio_microbench_ffma.argtypes = [C.c_int32, C.c_int32, _FP, C.POINTER(C.c_double), _FP]
    if lib.odevio_version() != ABI_VERSION:
        raise OdevioError("libodevio_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().odevio_error_string(code).decode()
        raise OdevioError(f"odevio call failed ({code}): {msg}")


def dptr(t, name="tensor"):
    """Device pointer of a contiguous fp32 (or int32) CUDA tensor; None -> NULL."""
    import torch
    if t is None:
        return None
    if not t.is_cuda:
        raise OdevioError(f"{name} must be a CUDA tensor (odevio_b200 has no CPU path)")
    if not t.is_contiguous():
        raise OdevioError(f"{name} must be contiguous")
    if t.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64, torch.uint8):
        raise OdevioError(f"{name} must be float32/int32, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


def default_cde_cfg():
    cfg = CdeCfg()
    load().odevio_cde_default_cfg(C.byref(cfg))
    return cfg


def default_odernn_cfg():
    cfg = OdeRnnCfg()
    load().odevio_odernn_default_cfg(C.byref(cfg))
    return cfg
