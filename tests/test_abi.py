"""The C-ABI library loads on a machine without a GPU and exports every symbol include/odevio.h
declares; argument validation returns error codes without touching a device."""

import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "odevio.h")
DEBUG_HEADER = os.path.join(ROOT, "include", "odevio_debug.h")       # diagnostics: not part of the product ABI


@pytest.fixture(scope="module")
def lib():
    from odevio_b200.build import build_library
    build_library()
    from odevio_b200 import _lib
    return _lib.load()


def declared_symbols(header=HEADER):
    src = open(header).read()
    return sorted(set(re.findall(r"ODEVIO_API[^;(]*?\b(odevio_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("odevio_version", "odevio_odernn_forward", "odevio_odernn_workspace_bytes",
                 "odevio_odernn_default_cfg", "odevio_error_string"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/odevio.h but not exported"


def test_debug_hooks_live_in_their_own_header(lib):
    assert not [n for n in declared_symbols() if "debug" in n or "microbench" in n]
    dbg = declared_symbols(DEBUG_HEADER)
    assert dbg and all("debug" in n or "microbench" in n for n in dbg)
    for name in dbg:
        assert hasattr(lib, name), f"{name} declared in include/odevio_debug.h but not exported"


def test_struct_layout_matches_header(lib):
    from odevio_b200 import _lib
    cfg = _lib.default_odernn_cfg()
    assert C.sizeof(_lib.OdeRnnCfg) == 30 * 4       # 24 ints + 6 floats, see include/odevio.h
    assert (cfg.B, cfg.S, cfg.D, cfg.H, cfg.n_hidden, cfg.L) == (1, 10, 768, 512, 3, 2)
    assert abs(cfg.atol - 1e-6) < 1e-12 and abs(cfg.rtol - 1e-2) < 1e-9 and abs(cfg.dt0 - 1e-4) < 1e-11
    assert (cfg.accept_strict, cfg.floor_factor, cfg.endpoint_dense, cfg.exact_landing) == (1, 0, 0, 1)
    assert lib.odevio_version() == _lib.ABI_VERSION == 4


def test_workspace_and_validation_without_gpu(lib):
    from odevio_b200 import _lib
    cfg = _lib.default_odernn_cfg()
    cfg.B = 1024
    assert lib.odevio_odernn_workspace_bytes(C.byref(cfg)) > 0
    bad = _lib.default_odernn_cfg()
    bad.D = 770                                  # not a multiple of 8
    assert lib.odevio_odernn_workspace_bytes(C.byref(bad)) == 0
    bad = _lib.default_odernn_cfg()
    bad.solver = 17
    assert lib.odevio_odernn_workspace_bytes(C.byref(bad)) == 0
    w = _lib.OdeRnnWeights()
    rc = lib.odevio_odernn_forward(C.byref(cfg), C.byref(w), None, None, 768, None, None, None, None, None,
                                   None, None, 0, None, 0, None)
    assert rc == -1                              # ODEVIO_E_NULL, before any device work
    assert b"NULL" in lib.odevio_error_string(rc)


def test_training_geometry_and_sizes_without_gpu(lib):
    """Checkpoint / backward-workspace planning is pure host arithmetic."""
    from odevio_b200 import _lib
    cfg = _lib.default_odernn_cfg()
    cfg.B, cfg.save_checkpoints = 100, 1
    geo = (C.c_int32 * 8)()
    assert lib.odevio_odernn_geometry(C.byref(cfg), geo) == 0
    RT, R, ntiles, ns, CK = geo[0], geo[1], geo[2], geo[3], geo[4]
    # dopri5: 6 stages enter y1 (FSAL); 100 sequences: 4-sequence tiles (25 CTAs) rather than 13 CTAs of 8
    assert (RT, R, ntiles, ns, CK) == (4, 8, 25, 6, 16)
    per_iv = 2 * 768 * R + CK * (768 * R + 2 * R)
    assert lib.odevio_odernn_ckpt_bytes(C.byref(cfg)) >= ntiles * 10 * per_iv * 4
    small = lib.odevio_odernn_backward_workspace_bytes(C.byref(cfg), 0)
    big = lib.odevio_odernn_backward_workspace_bytes(C.byref(cfg), 10000)
    assert small > 0 and big - small >= 10000 * 4 * (768 + 512 + 512 + 512 + 512 + 768)
    cfg.rnn_type = 1                                               # GRU: wider G records for the jump
    assert lib.odevio_odernn_backward_workspace_bytes(C.byref(cfg), 0) > small
    cfg.endpoint_dense = 1                                         # literal dense end point: no backward
    assert lib.odevio_odernn_ckpt_bytes(C.byref(cfg)) == 0
    cfg.endpoint_dense = 0
    cfg.rnn_type, cfg.solver = 0, 4                                # rk4: 4 stages, substeps iterations
    cfg.substeps = 2
    assert lib.odevio_odernn_geometry(C.byref(cfg), geo) == 0 and (geo[3], geo[4]) == (4, 2)
    g = _lib.OdeRnnGrads()
    rc = lib.odevio_odernn_backward(C.byref(cfg), C.byref(_lib.OdeRnnWeights()), None, None, 768, None, 0,
                                    None, 0, None, None, C.byref(g), None, None, None, 0, None)
    assert rc == -1


def test_cde_training_sizes_without_gpu(lib):
    """CDE checkpoint / backward-workspace planning is pure host arithmetic."""
    from odevio_b200 import _lib
    cfg = _lib.default_cde_cfg()
    cfg.B, cfg.Hc, cfg.interp = 40, 32, 1
    R, ntiles = 8, 5
    one = lib.odevio_cde_ckpt_bytes(C.byref(cfg), 1)
    grow = lib.odevio_cde_ckpt_bytes(C.byref(cfg), 11) - one
    assert 0 <= grow - 10 * ntiles * 9 * 32 * R * 4 <= 10 * 48 + 256             # Z, Y1, K0..K6 per step + the log entries
    assert lib.odevio_cde_ckpt_bytes(C.byref(cfg), 0) == 0
    small = lib.odevio_cde_backward_workspace_bytes(C.byref(cfg), 8)
    big = lib.odevio_cde_backward_workspace_bytes(C.byref(cfg), 16)
    # per pullback and row: inputs of the 4 Linears, gradients of the 3 hidden ones, the final Linear's 33 (+pad) channels
    assert small > 0 and big - small >= 8 * ntiles * R * 4 * (4 * 32 + 3 * 32 + 33 * 32)
    assert lib.odevio_cde_backward_workspace_bytes(C.byref(cfg), 4) == 0                           # below one step's pullbacks
    rc = lib.odevio_cde_backward(C.byref(cfg), C.byref(_lib.CdeWeights()), None, None, None, 32, None, 0, None, None,
                                 None, 0, 0, None, 0, 8, None, None, C.byref(_lib.CdeGrads()), None, None, None, 0, None)
    assert rc == -1


def test_module_fails_loudly_without_cuda():
    import torch
    import odevio_b200
    from oracle.pose_odernn import default_opt
    m = odevio_b200.PoseODERNN(default_opt())
    fv, fi, ts = torch.zeros(2, 10, 512), torch.zeros(2, 10, 256), torch.zeros(2, 11)
    with torch.no_grad(), pytest.raises(odevio_b200.OdevioError):
        m(fv, fi, ts)


def test_reference_menu_errors():
    import odevio_b200
    from oracle.pose_odernn import default_opt
    with pytest.raises(ValueError):
        odevio_b200.PoseODERNN(default_opt(ode_solver="bosh3"))
    with pytest.raises(ValueError):
        odevio_b200.PoseODERNN(default_opt(ode_rnn_type="lstm"))
    with pytest.raises(ValueError):
        odevio_b200.ODEFunc(8, 8, 2, "gelu")
