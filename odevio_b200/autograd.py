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
        with torch.cuda.device(dev):              # the planner reads the SM count of the CURRENT device
            _lib.check(lib.odevio_odernn_geometry(C.byref(cfg), geo))
        R, ntiles, ns = geo[1], geo[2], geo[3]
        nloops = ctx.ckpt[: ntiles * S * 4].view(torch.int32).to(torch.int64)
        rows = (nloops * (ns * R)).view(ntiles, S)
        # the one host read of the training step: record rows per observation interval (+ solver status)
        host = torch.cat([rows.sum(0), ctx.status.max().to(torch.int64).reshape(1)]).cpu()      # ONE device -> host copy
        per_iv, bad = [int(v) for v in host[:-1]], int(host[-1])
        if bad != 0:
            what = {1: "max_steps reached", 2: "non-finite error norm",
                    3: "more solver iterations per interval than opt.ode_ckpt_loops (raise it; a solver without error "
                       "control such as euler takes interval / ode_dt0 steps per interval)"}.get(bad, str(bad))
            raise RuntimeError(f"odevio_b200: cannot back-propagate, forward solve failed: {what}")
        # Interval ranges, walked from the last interval to the first, each within the record budget
        # (opt.ode_bwd_record_gb): the record streams are ~20 MB per sequence at configs[3] (84 GB at B = 4096 in one piece)
        with torch.cuda.device(dev):
            b0 = lib.odevio_odernn_backward_workspace_bytes(C.byref(cfg), 0)
            b1 = lib.odevio_odernn_backward_workspace_bytes(C.byref(cfg), R * 1024)
        if b0 == 0 or b1 == 0:
            raise _lib.OdevioError("odevio_odernn_backward: unsupported configuration")
        per_row = (b1 - b0) / (R * 1024)
        cap_rows = max(int((module.bwd_record_gb * (1 << 30) - b0) / per_row), 0)
        ranges, hi, acc = [], S - 1, 0
        for i in range(S - 1, -1, -1):
            if i < hi and acc + per_iv[i] > cap_rows:
                ranges.append((i + 1, hi, acc))
                hi, acc = i, 0
            acc += per_iv[i]
        ranges.append((0, hi, acc))
        plan_rows = max(r[2] for r in ranges)
        with torch.cuda.device(dev):
            nbytes = lib.odevio_odernn_backward_workspace_bytes(C.byref(cfg), plan_rows)
        if nbytes == 0:
            raise _lib.OdevioError("odevio_odernn_backward: unsupported configuration")
        # The buffer is kept on the module and reused across steps: handed back to torch's caching allocator it gets split by
        # the next forward's small allocations and the following request of ~the same size no longer fits (measured: OOM /
        # retry stalls).
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
        for lo, hi, nrows in ranges:
            sub = rows[:, lo:hi + 1].reshape(-1)
            base = torch.zeros(ntiles, S, dtype=torch.int64, device=dev)
            base[:, lo:hi + 1] = (torch.cumsum(sub, 0) - sub).view(ntiles, hi - lo + 1)      # relative to the range
            base = base.contiguous()
            with torch.cuda.device(dev):
                rc = lib.odevio_odernn_backward_range(
                    C.byref(cfg), C.byref(w), _lib.dptr(fvc), _lib.dptr(fic), ctx.Dv,
                    _lib.dptr(ctx.ckpt), ctx.ckpt_bytes, _lib.dptr(base), nrows, plan_rows, lo, hi,
                    _lib.dptr(gpose), _lib.dptr(ghT_c), C.byref(g), _lib.dptr(gfused), _lib.dptr(gh0),
                    _lib.dptr(ws), ws.numel(), C.c_void_p(stream))
            _lib.check(rc)
        module.last_bwd_ranges = [(lo, hi) for lo, hi, _ in ranges]
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


class _CdeFunction(torch.autograd.Function):
    """PoseCDE under autograd: ``odevio_cde_forward_ckpt`` + ``odevio_cde_backward`` (reference
    src/models/PoseCDE.py:98-101 ``cdeint(adjoint=False)`` + scripts/train_model.py:78)."""

    @staticmethod
    def forward(ctx, module, Dv, tobs, tout, fvc, fic, z0_in, *params):
        pose, z0, hidden, stats, (cfg, ckpt, cap) = module._launch(tobs, fvc, fic, Dv, tout, z0_in, save_ckpt=True)
        module.last_stats = stats
        # the one host read of the training step: the step log's pullback prefix counts (+ the solver status)
        log = ckpt[: 48 * (1 + cap)].view(torch.int32).view(1 + cap, 12).cpu()
        n_acc, n_vjp, status = (int(v) for v in log[0, :3])
        if status != 0:
            what = {1: "max_steps reached", 2: "non-finite error norm",
                    3: f"more accepted solver steps than cde_ckpt_steps={cap} (raise opt.cde_ckpt_steps)"}.get(status, str(status))
            raise RuntimeError(f"odevio_b200: CDE solve failed: {what}")
        ctx.vjp_base = [int(v) for v in log[1:1 + n_acc, 10]] + [n_vjp]
        ctx.n_acc, ctx.cap = n_acc, cap
        ctx.module, ctx.cfg, ctx.Dv, ctx.ckpt = module, cfg, Dv, ckpt
        ctx.has_fi, ctx.has_prev = fic is not None, z0_in is not None
        ctx.save_for_backward(tobs, tout, fvc, hidden, z0, *([fic] if fic is not None else []), *params)
        ctx.mark_non_differentiable(hidden)
        ctx.set_materialize_grads(False)
        return pose, z0, hidden

    @staticmethod
    def backward(ctx, gpose, gz0, _ghidden):
        lib = _lib.load()
        module, cfg = ctx.module, ctx.cfg
        saved = ctx.saved_tensors
        tobs, tout, fvc, hidden, z0 = saved[:5]
        fic = saved[5] if ctx.has_fi else None
        params = saved[6 if ctx.has_fi else 5:]
        dev = fvc.device
        B, S, So, Hc = cfg.B, cfg.S, cfg.So, cfg.Hc
        n_acc, n_vjp = ctx.n_acc, ctx.vjp_base[-1]
        # bound the record streams: halve the pullbacks per launch until the workspace fits the budget
        per_step = max([b - a for a, b in zip(ctx.vjp_base[:-1], ctx.vjp_base[1:])] + [8])
        chunk = max(n_vjp, per_step)
        budget = int(module.bwd_record_gb * (1 << 30))
        with torch.cuda.device(dev):
            nbytes = lib.odevio_cde_backward_workspace_bytes(C.byref(cfg), chunk)
            while nbytes > budget and chunk > per_step:
                chunk = max(per_step, chunk // 2)
                nbytes = lib.odevio_cde_backward_workspace_bytes(C.byref(cfg), chunk)
        if nbytes == 0:
            raise _lib.OdevioError("odevio_cde_backward: unsupported configuration")
        ws = getattr(module, "_bwd_workspace", None)
        if ws is None or ws.device != dev or ws.numel() < nbytes:
            module._bwd_workspace = ws = None
            module._bwd_workspace = ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        w, keep = module._weight_struct(params)
        g = _lib.CdeGrads()
        grads = [torch.zeros_like(p) for p in params]
        n = cfg.n_layers + 1
        for j in range(n):
            g.cde_w[j], g.cde_b[j] = _lib.dptr(grads[2 * j]), _lib.dptr(grads[2 * j + 1])
        k = 2 * n
        g.init_w, g.init_b, g.reg_w0, g.reg_b0, g.reg_w1, g.reg_b1 = (_lib.dptr(grads[k + i]) for i in range(6))
        gpose = (torch.zeros(B, S, 6, dtype=torch.float32, device=dev) if gpose is None else gpose.contiguous().float())
        gz0_c = None if gz0 is None else gz0.contiguous().float()
        need_in = ctx.needs_input_grad
        gx = torch.zeros(B, So, Hc + 1, dtype=torch.float32, device=dev) if (need_in[4] or need_in[5]) else None
        gprev = torch.zeros(B, Hc, dtype=torch.float32, device=dev) if ctx.has_prev else None
        base = (C.c_int32 * len(ctx.vjp_base))(*ctx.vjp_base)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.odevio_cde_backward(
                C.byref(cfg), C.byref(w), _lib.dptr(tobs), _lib.dptr(fvc), _lib.dptr(fic), ctx.Dv, _lib.dptr(tout),
                1 if ctx.has_prev else 0, _lib.dptr(hidden), _lib.dptr(z0), _lib.dptr(ctx.ckpt), ctx.ckpt.numel(), ctx.cap,
                base, n_acc, chunk, _lib.dptr(gpose), _lib.dptr(gz0_c), C.byref(g), _lib.dptr(gx), _lib.dptr(gprev),
                _lib.dptr(ws), ws.numel(), C.c_void_p(stream))
        _lib.check(rc)
        del keep
        gfv = gfi = None
        if gx is not None:
            feats = gx[..., 1:]                     # channel 0 is the timestamp
            if ctx.has_fi:
                gfv, gfi = feats[..., :ctx.Dv].contiguous(), feats[..., ctx.Dv:].contiguous()
            else:
                gfv = feats.contiguous()
        if ctx.has_prev:                            # initial() was not on the path
            grads[k] = grads[k + 1] = None
        return (None, None, None, None, gfv, gfi, gprev if need_in[6] else None, *grads)


def cde_apply(module, Dv, tobs, tout, fvc, fic, z0_in):
    if tobs.requires_grad:
        raise _lib.OdevioError("gradients with respect to timestamps are not supported")
    return _CdeFunction.apply(module, Dv, tobs, tout, fvc, fic, z0_in, *module._param_list())
