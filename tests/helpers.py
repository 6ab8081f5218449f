"""Shared helpers for the parity tests: build an oracle regressor and the CUDA-backed module
with IDENTICAL weights (reference init rule, DeepVIO.py:77-87) and compare their outputs."""

import copy

import torch

from oracle.modules import deepvio_initialization
from oracle.pose_odernn import OraclePoseODERNN, default_opt
from odevio_b200 import synth

POSE_RTOL = 1e-5      # north_star: relative poses within 1e-5 relative error in fp32
STATE_RTOL = 5e-5     # hidden state (not a north_star quantity), reported alongside


def make_pair(device, seed=0, bias_std=0.0, **opt_over):
    """(oracle on CPU, odevio_b200 module on `device`) sharing one state_dict."""
    import odevio_b200
    opt = default_opt(**opt_over)
    torch.manual_seed(seed)
    ref = OraclePoseODERNN(opt)
    deepvio_initialization(ref)
    if bias_std > 0:                       # trained-like: non-zero biases so f(0) != 0
        g = torch.Generator().manual_seed(seed + 7)
        for n, p in ref.named_parameters():
            if n.endswith("bias") and not n.startswith("rnn"):
                p.data.normal_(0.0, bias_std, generator=g)
    ref.eval()
    mod = odevio_b200.PoseODERNN(copy.copy(opt))
    missing = mod.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    mod = mod.to(device).eval()
    return ref, mod


def inputs(B, S=10, irregular=False, seed=0, offset=0.0, D=(512, 256)):
    fv, fi = synth.features(B, S, D[0], D[1], seed=seed)
    ts = synth.timestamps(B, S, irregular=irregular, seed=seed, offset=offset)
    return fv, fi, ts


def rel_err(a, b):
    """max |a-b| / max |b| (max-norm relative error)."""
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def run_pair(ref, mod, fv, fi, ts, prev=None, device="cuda"):
    with torch.no_grad():
        p_ref, h_ref = ref(fv, fi, ts, prev=prev)
        p, h = mod(fv.to(device), fi.to(device), ts.to(device),
                   prev=None if prev is None else prev.to(device))
    torch.cuda.synchronize()
    out = dict(pose_ref=p_ref, h_ref=h_ref, pose=p.cpu(), h=h.cpu(),
               pose_err=rel_err(p.cpu(), p_ref), h_err=rel_err(h.cpu(), h_ref))
    if mod.last_stats is not None and ref.last_stats is not None:
        st = mod.last_stats.cpu().long()
        out["steps_equal"] = bool((st[..., 0] == ref.last_stats["n_steps"]).all())
        out["acc_equal"] = bool((st[..., 1] == ref.last_stats["n_accepted"]).all())
        out["n_mismatch_rows"] = int(((st[..., 0] != ref.last_stats["n_steps"]) |
                                      (st[..., 1] != ref.last_stats["n_accepted"])).any(0).any(0).sum())
        out["status_max"] = int(mod.last_status.max().item())
    return out
