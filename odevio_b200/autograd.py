"""Autograd bridge for the fused regressors (discretise-then-optimise backward).

``loss.backward()`` in the reference (scripts/train_model.py:78) is plain autograd through
torchode's solver loop (``to.AutoDiffAdjoint``, src/models/PoseODERNN.py:58-60).  Here the
forward kernel stores per-iteration checkpoints and ``odevio_odernn_backward`` replays them;
accepted step sizes are treated as constants.  PyTorch only carries the tensors: every gradient
is computed by the sm_100a kernels behind the C ABI (no eager fallback).
"""

import ctypes as C

import torch

from . import _lib


def _param_list(module):
    """Flat parameter order shared by forward() and backward()."""
    ps = []
    for lin in module.ode_func.linears():
        ps += [lin.weight, lin.bias]
    for l in range(module.rnn_num_layers):
        ps += [getattr(module.rnn, f"weight_ih_l{l}"), getattr(module.rnn, f"weight_hh_l{l}"),
               getattr(module.rnn, f"bias_ih_l{l}"), getattr(module.rnn, f"bias_hh_l{l}")]
    ps += [module.regressor[0].weight, module.regressor[0].bias,
           module.regressor[2].weight, module.regressor[2].bias]
    return ps


class _OdeRnnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, Dv, ts_in, fvc, fic, h0, *params):
        pose, hT, (cfg, ckpt, ckpt_bytes) = module._launch(fvc, fic, Dv, ts_in, h0, save_ckpt=True)
        ctx.module, ctx.cfg, ctx.Dv = module, cfg, Dv
        ctx.ckpt, ctx.ckpt_bytes = ckpt, ckpt_bytes
        ctx.status = module.last_status
        ctx.has_fi, ctx.has_h0 = fic is not None, h0 is not None
        ctx.save_for_backward(fvc, *([fic] if fic is not None else []), *params)
        ctx.mark_non_differentiable()
        return pose, hT

    @staticmethod
    def backward(ctx, gpose, ghT):
        lib = _lib.load()
        module, cfg = ctx.module, ctx.cfg
        saved = ctx.saved_tensors
        fvc = saved[0]
        fic = saved[1] if ctx.has_fi else None
        params = saved[2 if ctx.has_fi else 1:]
        dev = fvc.device
        B, S, D, L = cfg.B, cfg.S, cfg.D, cfg.L

        geo = (C.c_int32 * 8)()
        _lib.check(lib.odevio_odernn_geometry(C.byref(cfg), geo))
        R, ntiles, ns = geo[1], geo[2], geo[3]
        nloops = ctx.ckpt[: ntiles * S * 4].view(torch.int32).to(torch.int64)
        rows = nloops * (ns * R)
        rec_base = (torch.cumsum(rows, 0) - rows).contiguous()
        # the one host read of the training step: record-stream length (+ solver status)
        ode_rows = int(rows.sum().item())
        bad = int(ctx.status.max().item())
        if bad != 0:
            what = {1: "max_steps reached", 2: "non-finite error norm",
                    3: "more solver iterations per interval than ode_ckpt_loops"}.get(bad, str(bad))
            raise RuntimeError(f"odevio_b200: cannot back-propagate, forward solve failed: {what}")

        nbytes = lib.odevio_odernn_backward_workspace_bytes(C.byref(cfg), ode_rows)
        if nbytes == 0:
            raise _lib.OdevioError("odevio_odernn_backward: unsupported configuration")
        # The record streams dominate (tens of GB at B = 4096).  The buffer is kept on the module and reused
        # across steps: handed back to torch's caching allocator it gets split by the next forward's small
        # allocations and the following request of ~the same size no longer fits (measured: OOM / retry stalls).
        ws = getattr(module, "_bwd_workspace", None)
        if ws is None or ws.device != dev or ws.numel() < nbytes:
            module._bwd_workspace = ws = None
            gran = 1 << 30 if nbytes >= (8 << 30) else 1 << 26
            module._bwd_workspace = ws = torch.empty((nbytes + gran - 1) // gran * gran, dtype=torch.uint8, device=dev)
        w = _lib.OdeRnnWeights()
        g = _lib.OdeRnnGrads()
        grads = [torch.empty_like(p) for p in params]
        NL = cfg.n_hidden + 1
        k = 0
        for j in range(NL):
            w.ode_w[j], w.ode_b[j] = _lib.dptr(params[k]), _lib.dptr(params[k + 1])
            g.ode_w[j], g.ode_b[j] = _lib.dptr(grads[k]), _lib.dptr(grads[k + 1])
            k += 2
        for l in range(L):
            w.rnn_w_ih[l], w.rnn_w_hh[l] = _lib.dptr(params[k]), _lib.dptr(params[k + 1])
            w.rnn_b_ih[l], w.rnn_b_hh[l] = _lib.dptr(params[k + 2]), _lib.dptr(params[k + 3])
            g.rnn_w_ih[l], g.rnn_w_hh[l] = _lib.dptr(grads[k]), _lib.dptr(grads[k + 1])
            g.rnn_b_ih[l], g.rnn_b_hh[l] = _lib.dptr(grads[k + 2]), _lib.dptr(grads[k + 3])
            k += 4
        w.reg_w0, w.reg_b0, w.reg_w1, w.reg_b1 = (_lib.dptr(params[k + i]) for i in range(4))
        g.reg_w0, g.reg_b0, g.reg_w1, g.reg_b1 = (_lib.dptr(grads[k + i]) for i in range(4))

        gpose = gpose.contiguous().float()
        ghT_c = None if ghT is None else ghT.contiguous().float()
        need_in = ctx.needs_input_grad
        gfused = torch.empty(B, S, D, dtype=torch.float32, device=dev) if (need_in[3] or need_in[4]) else None
        gh0 = torch.empty(L, B, D, dtype=torch.float32, device=dev) if (ctx.has_h0 and need_in[5]) else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.odevio_odernn_backward(
                C.byref(cfg), C.byref(w), _lib.dptr(fvc), _lib.dptr(fic), ctx.Dv,
                _lib.dptr(ctx.ckpt), ctx.ckpt_bytes, _lib.dptr(rec_base), ode_rows,
                _lib.dptr(gpose), _lib.dptr(ghT_c), C.byref(g), _lib.dptr(gfused), _lib.dptr(gh0),
                _lib.dptr(ws), ws.numel(), C.c_void_p(stream))
        _lib.check(rc)
        gfv = gfi = None
        if gfused is not None:
            if ctx.has_fi:
                gfv, gfi = gfused[..., :ctx.Dv], gfused[..., ctx.Dv:]
            else:
                gfv = gfused
        return (None, None, None, gfv, gfi, gh0, *grads)


def odernn_apply(module, fvc, fic, Dv, ts_in, h0):
    if ts_in.requires_grad:
        raise _lib.OdevioError("gradients with respect to timestamps are not supported")
    return _OdeRnnFunction.apply(module, Dv, ts_in, fvc, fic, h0, *_param_list(module))
