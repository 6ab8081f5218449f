"""Host-side mirror of the reference's regressor interfaces for the latent-dynamics path.

Same constructor namespace (``opt``), ``forward`` signatures, attribute names and state_dict
keys as the reference, so ``DeepVIO.forward`` (reference src/models/DeepVIO.py:61-68) is a
drop-in and reference checkpoints load:

  * :class:`ODEFunc`      -- reference src/models/ODEFunc.py:5-39
  * :class:`CDEFunc`      -- reference src/models/ODEFunc.py:44-84
  * :class:`FusionModule` -- reference src/models/FusionModule.py:8-29
  * :class:`PoseODERNN`   -- reference src/models/PoseODERNN.py:39-154
  * :class:`InertialEncoder` -- reference src/models/Encoder.py:39-74 (the step upstream of the regressors)

The modules only *hold* parameters (ordinary ``nn.Parameter``) and pack pointers; all of the
path's arithmetic runs in the sm_100a kernels behind the C ABI (``include/odevio.h``).  There
is no eager/PyTorch/CPU fallback: without the CUDA library or on CPU tensors, ``forward`` raises.

Optional ``opt`` attributes beyond the reference's argparse namespace (reference values are the
defaults; ``getattr(opt, name, default)``):
  ode_atol=1e-6, ode_rtol=1e-2, ode_dt0=1e-4   (hard-coded at PoseODERNN.py:57,72)
  ode_substeps=1        steps per interval for the fixed-step solvers {"rk4", "rk4_38"}
  ode_max_steps=100000  per-interval guard (rows still running get status MAX_STEPS)
  ode_accept_strict=True, ode_floor_factor=False                         (SURVEY.md A.1 switches)
  ode_endpoint="y1" | "dense", ode_exact_landing=True   ("dense"/False = literal fp32 torchode
                        arithmetic, which is ill-conditioned; see oracle/torchode_like.py).  "dense" runs in
                        the FMA kernel ("fp32") and in the one-launch tcgen05 kernel ("fp16x3"); training takes
                        "dense" with exact landing (x = 1: the dense output is y1), not the literal variant
  ode_rows_per_tile=0   (auto) | 4 | 8 | 16
  ode_ckpt_loops=0      training: stored solver iterations per interval and tile (0 = 16)
  ode_bwd_record_gb=24  training: bound on the record streams of the deferred weight-gradient GEMMs; the backward
                        walks the observation intervals in as many ranges as that takes (B = 4096: 84 GB in one piece)

Training: with grad enabled, ``forward`` goes through ``odevio_b200.autograd`` -- the fused
forward with checkpoints, then ``odevio_odernn_backward`` (discretise-then-optimise, step sizes
constant; reference: autograd through torchode's AutoDiffAdjoint, scripts/train_model.py:78).
"""

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

_ACTS = {"tanh": nn.Tanh, "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "softplus": nn.Softplus}


def _activation(name):
    if name not in _ACTS:
        raise ValueError(f"Activation function {name} not supported")     # ODEFunc.py:33-34
    return _ACTS[name]()


def _vector_field_net(sizes, activation):
    layers = []
    last = len(sizes) - 2
    for i in range(len(sizes) - 1):
        lin = nn.Linear(sizes[i], sizes[i + 1])
        nn.init.normal_(lin.weight, mean=0.0, std=0.1)                      # ODEFunc.py:18-21
        nn.init.zeros_(lin.bias)
        layers += [lin, nn.Tanh() if i == last else _activation(activation)]
    return nn.Sequential(*layers)


def _mlp_forward_kernel(net, x2, last_act):
    """Evaluate an nn.Sequential of Linear/activation pairs with odevio_mlp_forward (CUDA cores)."""
    lib = _lib.load()
    lins = [m for m in net if isinstance(m, nn.Linear)]
    acts = [m for m in net if not isinstance(m, nn.Linear)]
    names = {nn.Tanh: 0, nn.ReLU: 1, nn.LeakyReLU: 2, nn.Softplus: 3}
    act_ids = [names[type(a)] for a in acts]
    assert len(act_ids) == len(lins) and act_ids[-1] == last_act
    dims = [lins[0].in_features] + [l.out_features for l in lins]
    keep = [(_f32c(l.weight.detach(), "weight"), _f32c(l.bias.detach(), "bias")) for l in lins]
    out = torch.empty(x2.shape[0], dims[-1], dtype=torch.float32, device=x2.device)
    n = len(lins)
    with torch.cuda.device(x2.device):
        rc = lib.odevio_mlp_forward(x2.shape[0], n, (C.c_int32 * (n + 1))(*dims), (C.c_int32 * n)(*act_ids),
                                    (C.c_void_p * n)(*[t[0].data_ptr() for t in keep]),
                                    (C.c_void_p * n)(*[t[1].data_ptr() for t in keep]),
                                    _lib.dptr(x2), _lib.dptr(out),
                                    C.c_void_p(torch.cuda.current_stream(x2.device).cuda_stream))
    _lib.check(rc)
    del keep
    return out


class ODEFunc(nn.Module):
    """Parameter container for the autonomous vector field
    f(t, x) = tanh(W_n a(... a(W_0 x + b_0) ...) + b_n)  (keys ``net.{0,2,..}.{weight,bias}``)."""

    def __init__(self, feature_dim, hidden_dim, num_hidden_layers=3, activation="tanh"):
        super().__init__()
        self.feature_dim, self.hidden_dim = feature_dim, hidden_dim
        self.num_hidden_layers, self.activation = num_hidden_layers, activation
        self.net = _vector_field_net([feature_dim] + [hidden_dim] * num_hidden_layers + [feature_dim],
                                     activation)

    def linears(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]

    def forward(self, t, x):
        """f(t, x) for callers that evaluate the field directly (ODEFunc.py:38-39; `t` is ignored).
        Runs the tcgen05 3xTF32 kernel behind ``odevio_odefunc_forward``; the fused solver kernels
        read the weights themselves and never come through here.  No CPU / autograd path."""
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.OdevioError("ODEFunc.forward needs a CUDA tensor: odevio_b200 has no CPU path")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise _lib.OdevioError("ODEFunc.forward is inference-only; gradients flow through PoseODERNN's fused backward")
        shape = x.shape
        x2 = _f32c(x.reshape(-1, self.feature_dim), "x")
        M = x2.shape[0]
        nbytes = lib.odevio_odefunc_workspace_bytes(M, self.feature_dim, self.hidden_dim, self.num_hidden_layers)
        if nbytes == 0:
            # shapes outside the tensor-core kernel's tiling (D, H multiples of 256 up to 1024): CUDA-core kernel
            return _mlp_forward_kernel(self.net, x2, 0).reshape(shape)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        out = torch.empty_like(x2)
        lins = self.linears()
        keep = [(_f32c(l.weight.detach(), "weight"), _f32c(l.bias.detach(), "bias")) for l in lins]
        wp = (C.c_void_p * len(lins))(*[t[0].data_ptr() for t in keep])
        bp = (C.c_void_p * len(lins))(*[t[1].data_ptr() for t in keep])
        with torch.cuda.device(x.device):
            rc = lib.odevio_odefunc_forward(M, self.feature_dim, self.hidden_dim, self.num_hidden_layers,
                                            _lib.ACT[self.activation], wp, bp, _lib.dptr(x2), _lib.dptr(out),
                                            _lib.dptr(ws), nbytes,
                                            C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        _lib.check(rc)
        del keep
        return out.reshape(shape)


class CDEFunc(nn.Module):
    """Parameter container for g(t, z) = tanh(MLP(z)).view(B, hidden, channels)."""

    def __init__(self, feature_dim, hidden_dim, num_hidden_layers=3, activation="tanh"):
        super().__init__()
        self.hidden_dim, self.feature_dim = hidden_dim, feature_dim
        self.num_hidden_layers, self.activation = num_hidden_layers, activation
        self.net = _vector_field_net([hidden_dim] * (num_hidden_layers + 1) + [hidden_dim * feature_dim],
                                     activation)

    def linears(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]

    def forward(self, t, z):
        """g(t, z) -> [B, hidden, channels] for callers that evaluate the field directly
        (ODEFunc.py:81-84); the fused CDE kernel never materialises this tensor.  Runs
        ``odevio_mlp_forward``; no CPU / autograd path."""
        _lib.load()
        if not z.is_cuda:
            raise _lib.OdevioError("CDEFunc.forward needs a CUDA tensor: odevio_b200 has no CPU path")
        if torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise _lib.OdevioError("CDEFunc.forward is inference-only (the fused CDE backward is not built)")
        out = _mlp_forward_kernel(self.net, _f32c(z.reshape(-1, self.hidden_dim), "z"), 0)
        return out.view(z.size(0), self.hidden_dim, self.feature_dim)


class FusionModule(nn.Module):
    """cat / soft / hard fusion (FusionModule.py:17-29).  Inside PoseODERNN, "cat" is folded into
    the kernel's feature load (no concatenated tensor is materialised) and "soft" runs in the
    kernel prologue of every jump for inference; while training "soft" stays this torch op so that
    autograd carries its gradient, and "hard" (Gumbel noise from torch's RNG) always does."""

    def __init__(self, feature_dim, fuse_method):
        super().__init__()
        self.fuse_method, self.f_len = fuse_method, feature_dim
        if fuse_method == "soft":
            self.net = nn.Sequential(nn.Linear(feature_dim, feature_dim))
        elif fuse_method == "hard":
            self.net = nn.Sequential(nn.Linear(feature_dim, 2 * feature_dim))

    def forward(self, v, i):
        cat = torch.cat((v, i), -1)
        if self.fuse_method == "cat":
            return cat
        if self.fuse_method == "soft":
            return cat * self.net(cat)
        if self.fuse_method == "hard":
            w = self.net(cat).view(v.shape[0], v.shape[1], self.f_len, 2)
            return cat * F.gumbel_softmax(w, tau=1, hard=True, dim=-1)[:, :, :, 0]
        raise ValueError(f"fuse method {self.fuse_method} not supported")


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise _lib.OdevioError(f"{name} must be float32 (the path computes in fp32), got {t.dtype}")
    return t.contiguous()


class InertialEncoder(nn.Module):
    """IMU encoder (reference src/models/Encoder.py:39-74), the step immediately upstream of the regressors: raw IMU rows
    ``[B, 10*S + 1, 6]`` -> ``fi [B, S, i_f_len]``.  Same construction order, module types and state_dict keys as the
    reference (``encoder_conv.{0..11}``, ``proj``), so its checkpoints load; ``forward`` launches ``odevio_imu_encoder_forward``
    (one kernel: windowing, 3 x Conv1d+BatchNorm+LeakyReLU, projection).  Inference only: in training mode BatchNorm
    uses batch statistics and Dropout is active, for which there is no kernel -- ``forward`` raises."""

    def __init__(self, opt):
        super().__init__()
        self.seq_len = opt.seq_len
        drop = getattr(opt, "imu_dropout", 0.0)
        self.encoder_conv = nn.Sequential(
            nn.Conv1d(6, 64, kernel_size=3, padding=1), nn.BatchNorm1d(64), nn.LeakyReLU(0.1, inplace=True), nn.Dropout(drop),
            nn.Conv1d(64, 128, kernel_size=3, padding=1), nn.BatchNorm1d(128), nn.LeakyReLU(0.1, inplace=True), nn.Dropout(drop),
            nn.Conv1d(128, 256, kernel_size=3, padding=1), nn.BatchNorm1d(256), nn.LeakyReLU(0.1, inplace=True), nn.Dropout(drop),
        )
        self.proj = nn.Linear(256 * 1 * 11, opt.i_f_len)
        self.i_f_len = opt.i_f_len

    def forward(self, x):
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.OdevioError("InertialEncoder.forward needs a CUDA tensor: odevio_b200 has no CPU path")
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
                             and x.requires_grad):
            raise _lib.OdevioError("InertialEncoder.forward is inference-only (eval mode: running BatchNorm statistics, "
                                   "no Dropout); the reference trains it with the encoders, which are out of scope")
        if x.dim() != 3 or x.shape[2] != 6 or x.shape[1] < 11:
            raise _lib.OdevioError(f"imu must be [B, 10*S + 1, 6], got {tuple(x.shape)}")
        B, S = x.shape[0], (x.shape[1] - 1) // 10                                   # Encoder.py:61
        xs = _f32c(x[:, :10 * S + 1].detach(), "imu")
        keep = []

        def ptr(t):
            t = _f32c(t.detach(), "weight")
            keep.append(t)
            return t.data_ptr()

        w = _lib.ImuEncoderWeights()
        for k in range(3):
            conv, bn = self.encoder_conv[4 * k], self.encoder_conv[4 * k + 1]
            w.conv_w[k], w.conv_b[k] = ptr(conv.weight), ptr(conv.bias)
            w.bn_weight[k], w.bn_bias[k] = ptr(bn.weight), ptr(bn.bias)
            w.bn_mean[k], w.bn_var[k] = ptr(bn.running_mean), ptr(bn.running_var)
        w.bn_eps = self.encoder_conv[1].eps
        w.proj_w, w.proj_b = ptr(self.proj.weight), ptr(self.proj.bias)
        nbytes = lib.odevio_imu_encoder_workspace_bytes(self.i_f_len)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        out = torch.empty(B, S, self.i_f_len, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.odevio_imu_encoder_forward(B, S, self.i_f_len, C.byref(w), _lib.dptr(xs), _lib.dptr(out),
                                                _lib.dptr(ws), nbytes,
                                                C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        _lib.check(rc)
        del keep
        return out


class PoseODERNN(nn.Module):
    """ODE-RNN pose regressor; ``forward`` launches the fused sm_100a kernel
    (``odevio_odernn_forward``)."""

    SOLVERS = ("dopri5", "heun", "tsit5", "euler", "rk4", "rk4_38")

    def __init__(self, opt):
        super().__init__()
        self.f_len = opt.v_f_len + opt.i_f_len
        self.rnn_hidden_dim = getattr(opt, "rnn_hidden_dim", self.f_len)    # unused, as in the reference
        self.rnn_num_layers = opt.rnn_num_layers
        self.fuse_method = opt.fuse_method
        self.ode_func = ODEFunc(feature_dim=self.f_len, hidden_dim=opt.ode_hidden_dim,
                                num_hidden_layers=opt.ode_fn_num_layers,
                                activation=opt.ode_activation_fn)
        self.ode_solver = self._set_solver(opt.ode_solver)
        self.rnn_type = opt.ode_rnn_type
        self.rnn = self._set_rnn(opt.ode_rnn_type)
        self.rnn_drop_out = nn.Dropout(getattr(opt, "rnn_dropout_out", 0.0))   # constructed, never applied
        self.fuse = FusionModule(feature_dim=self.f_len, fuse_method=self.fuse_method)
        self.regressor = nn.Sequential(nn.Linear(self.f_len, 128), nn.LeakyReLU(0.1, inplace=True),
                                       nn.Linear(128, 6))
        # solver knobs (reference values hard-coded at PoseODERNN.py:57,72)
        self.atol = float(getattr(opt, "ode_atol", 1e-6))
        self.rtol = float(getattr(opt, "ode_rtol", 1e-2))
        self.dt0 = float(getattr(opt, "ode_dt0", 1e-4))
        self.substeps = int(getattr(opt, "ode_substeps", 1))
        self.max_steps = int(getattr(opt, "ode_max_steps", 100000))
        self.accept_strict = bool(getattr(opt, "ode_accept_strict", True))
        self.floor_factor = bool(getattr(opt, "ode_floor_factor", False))
        self.endpoint = getattr(opt, "ode_endpoint", "y1")
        self.exact_landing = bool(getattr(opt, "ode_exact_landing", True))
        self.rows_per_tile = int(getattr(opt, "ode_rows_per_tile", 0))
        self.collect_stats = bool(getattr(opt, "ode_collect_stats", True))
        self.trace_steps = int(getattr(opt, "ode_trace_steps", 0))   # diagnostic: (dt, ratio) of first T steps
        self.ckpt_loops = int(getattr(opt, "ode_ckpt_loops", 0))     # training: stored solver iterations per interval (0 = 16)
        self.bwd_record_gb = float(getattr(opt, "ode_bwd_record_gb", 24.0))   # training: bound on the backward's record streams
        self.last_bwd_ranges = None                                  # interval ranges the last backward was walked in
        # "fp32": CUDA-core FFMA kernel; "tf32x3": ODEFunc GEMMs on tcgen05 (3xTF32, fp32-accurate), inference only
        self.precision = getattr(opt, "ode_precision", "fp32")
        if self.precision not in _lib.PRECISION:
            raise ValueError(f"Precision {self.precision} not supported")
        self.last_stats = None      # int32 [S, L, B, 2] = (n_steps, n_accepted) of the last forward
        self.last_trace = None      # float32 [S, L, B, T, 2] = (dt, error ratio) when trace_steps = T > 0
        self.last_status = None     # int32 [B]
        self.last_prepacked = False # the last forward reused the packed weight images of an earlier one
        self._ws_cache = None

    # -- reference menu (PoseODERNN.py:125-148) ------------------------------------------
    def _set_solver(self, ode_solver):
        if ode_solver not in self.SOLVERS:
            raise ValueError(f"Solver {ode_solver} not supported")
        return ode_solver

    def _set_rnn(self, rnn_type):
        if rnn_type == "rnn":
            return nn.RNN(input_size=self.f_len, hidden_size=self.f_len, num_layers=self.rnn_num_layers,
                          batch_first=True)
        if rnn_type == "gru":
            return nn.GRU(input_size=self.f_len, hidden_size=self.f_len, num_layers=self.rnn_num_layers,
                          batch_first=True)
        raise ValueError(f"RNN type {rnn_type} not supported")

    def get_regressor_params(self):
        return self.regressor.parameters()

    def get_other_params(self):
        return [p for n, p in self.named_parameters() if not n.startswith("regressor")]

    # -- C-ABI plumbing -------------------------------------------------------------------
    def _cfg(self, B, S):
        cfg = _lib.default_odernn_cfg()
        cfg.B, cfg.S, cfg.D, cfg.H = B, S, self.f_len, self.ode_func.hidden_dim
        cfg.n_hidden, cfg.L = self.ode_func.num_hidden_layers, self.rnn_num_layers
        cfg.activation = _lib.ACT[self.ode_func.activation]
        cfg.rnn_type = _lib.RNN[self.rnn_type]
        cfg.solver = _lib.SOLVER[self.ode_solver]
        cfg.substeps = self.substeps
        cfg.atol, cfg.rtol, cfg.dt0 = self.atol, self.rtol, self.dt0
        cfg.accept_strict = int(self.accept_strict)
        cfg.floor_factor = int(self.floor_factor)
        cfg.endpoint_dense = int(self.endpoint == "dense")
        cfg.exact_landing = int(self.exact_landing)
        cfg.max_steps = self.max_steps
        cfg.rows_per_tile = self.rows_per_tile
        cfg.trace_steps = self.trace_steps
        cfg.precision = _lib.PRECISION[self.precision]
        return cfg

    def _weights(self, fuse_in_kernel=False):
        w = _lib.OdeRnnWeights()
        keep = []

        def ptr(p, name):
            t = _f32c(p.detach(), name)
            keep.append(t)
            return _lib.dptr(t, name)

        for j, lin in enumerate(self.ode_func.linears()):
            w.ode_w[j] = ptr(lin.weight, f"ode_func.net.{2 * j}.weight")
            w.ode_b[j] = ptr(lin.bias, f"ode_func.net.{2 * j}.bias")
        for l in range(self.rnn_num_layers):
            w.rnn_w_ih[l] = ptr(getattr(self.rnn, f"weight_ih_l{l}"), f"rnn.weight_ih_l{l}")
            w.rnn_w_hh[l] = ptr(getattr(self.rnn, f"weight_hh_l{l}"), f"rnn.weight_hh_l{l}")
            w.rnn_b_ih[l] = ptr(getattr(self.rnn, f"bias_ih_l{l}"), f"rnn.bias_ih_l{l}")
            w.rnn_b_hh[l] = ptr(getattr(self.rnn, f"bias_hh_l{l}"), f"rnn.bias_hh_l{l}")
        w.reg_w0 = ptr(self.regressor[0].weight, "regressor.0.weight")
        w.reg_b0 = ptr(self.regressor[0].bias, "regressor.0.bias")
        w.reg_w1 = ptr(self.regressor[2].weight, "regressor.2.weight")
        w.reg_b1 = ptr(self.regressor[2].bias, "regressor.2.bias")
        if fuse_in_kernel:
            w.fuse_w = ptr(self.fuse.net[0].weight, "fuse.net.0.weight")
            w.fuse_b = ptr(self.fuse.net[0].bias, "fuse.net.0.bias")
        return w, keep

    def _prepare_inputs(self, fv, fi, ts, prev):
        """Reference pre-processing (PoseODERNN.py:93-100): fusion, ts - ts[:, :1] without prev."""
        if not fv.is_cuda:
            raise _lib.OdevioError("PoseODERNN.forward needs CUDA tensors: odevio_b200 has no CPU path")
        B = fv.shape[0]
        if self.fuse_method == "cat" or (self.fuse_method == "soft" and not self._fuse_in_torch()):
            # "cat" is folded into the kernel's feature load; "soft" (cat * Linear(cat)) runs in the
            # kernel prologue of every jump when no gradient is needed (the raw features go in)
            fvc, fic, Dv = _f32c(fv, "fv"), _f32c(fi, "fi"), fv.shape[2]
        else:
            fvc, fic, Dv = _f32c(self.fuse(fv, fi), "fused"), None, self.f_len
        ts = _f32c(ts, "ts")
        ts_in = ((ts - ts[:, :1]) if prev is None else ts).contiguous()        # PoseODERNN.py:100
        h0 = None if prev is None else _f32c(prev, "prev")
        if h0 is not None and tuple(h0.shape) != (self.rnn_num_layers, B, self.f_len):
            raise _lib.OdevioError(f"prev must be [L,B,D]={self.rnn_num_layers, B, self.f_len}, got {tuple(h0.shape)}")
        return fvc, fic, Dv, ts_in, h0

    def _fuse_in_torch(self):
        """'soft' fusion stays a torch op (autograd) while training; 'hard' (Gumbel noise) always."""
        return self.fuse_method == "hard" or (
            torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()))

    def forward(self, fv, fi, ts, prev=None, do_profile=False):
        _lib.load()
        fvc, fic, Dv, ts_in, h0 = self._prepare_inputs(fv, fi, ts, prev)
        needs_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self.parameters()) or fvc.requires_grad
            or (fic is not None and fic.requires_grad) or (h0 is not None and h0.requires_grad))
        if needs_grad:
            from .autograd import odernn_apply                    # fused discretise-then-optimise backward
            return odernn_apply(self, fvc, fic, Dv, ts_in, h0)
        with torch.no_grad():
            pose, hT, _ = self._launch(fvc, fic, Dv, ts_in, h0, save_ckpt=False, do_profile=do_profile)
        return pose, hT

    def _launch(self, fvc, fic, Dv, ts_in, h0, save_ckpt=False, do_profile=False):
        """One odevio_odernn_forward call.  Returns (pose, hT, ctx) where ctx carries what the
        backward needs (cfg, checkpoint buffer) when save_ckpt is set."""
        lib = _lib.load()
        B, S = fvc.shape[0], fvc.shape[1]
        dev = fvc.device
        cfg = self._cfg(B, S)
        ckpt, ckpt_bytes = None, 0
        if save_ckpt:
            # training: "fp16x3" keeps the checkpointed forward on tcgen05 (one launch; the library falls back to the FMA
            # forward for GRU / L > 2), the fused backward is the FMA kernel either way; "tf32x3" has no training forward
            if cfg.precision == _lib.PRECISION["tf32x3"]:
                cfg.precision = _lib.PRECISION["fp32"]
            if self.endpoint == "dense" and self.exact_landing:
                # with exact landing the step that reaches t_end has x = (t_end - t) / dt = 1, where the quartic dense
                # output IS y1 (its coefficients sum to y1 - y0 identically): the training forward / backward use the y1
                # rule -- the same function and gradient, the forward differs from the dense evaluation by rounding only
                cfg.endpoint_dense = 0
            cfg.save_checkpoints = 1
            cfg.ckpt_loops = self.ckpt_loops
            if cfg.rows_per_tile == 16:
                cfg.rows_per_tile = 8
            with torch.cuda.device(dev):          # the planner reads the SM count of the CURRENT device
                ckpt_bytes = lib.odevio_odernn_ckpt_bytes(C.byref(cfg))
            if ckpt_bytes == 0:
                raise _lib.OdevioError("training through the fused path needs ode_endpoint='y1' (or 'dense' with "
                                       "ode_exact_landing=True, which is the same function), rows_per_tile in "
                                       "{0, 4, 8} and D, H multiples of 128 "
                                       f"(endpoint={self.endpoint}, D={cfg.D}, H={cfg.H})")
            ckpt = torch.empty(ckpt_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            nbytes = lib.odevio_odernn_workspace_bytes(C.byref(cfg))
        if nbytes == 0:
            raise _lib.OdevioError("unsupported PoseODERNN configuration for the fused kernel "
                                   f"(D={cfg.D}, H={cfg.H}, L={cfg.L}, n={cfg.n_hidden})")
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws = self._cached_workspace(cfg, nbytes, dev, stream) if not save_ckpt else None
        if ws is None:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        pose = torch.empty(B, S, 6, dtype=torch.float32, device=dev)
        hT = torch.empty(self.rnn_num_layers, B, self.f_len, dtype=torch.float32, device=dev)
        T = self.trace_steps
        stats = (torch.zeros(S, self.rnn_num_layers, B, 2 + 2 * T, dtype=torch.int32, device=dev)
                 if (self.collect_stats or T) else None)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        w, keep = self._weights(fuse_in_kernel=(self.fuse_method == "soft" and fic is not None))
        if do_profile:
            torch.cuda.nvtx.range_push("odeint")                               # PoseODERNN.py:103-104
        with torch.cuda.device(dev):
            rc = lib.odevio_odernn_forward(
                C.byref(cfg), C.byref(w), _lib.dptr(fvc, "fv"), _lib.dptr(fic, "fi"), Dv,
                _lib.dptr(ts_in, "ts"), _lib.dptr(h0, "prev"), _lib.dptr(pose), _lib.dptr(hT),
                _lib.dptr(stats), _lib.dptr(status), _lib.dptr(ckpt), ckpt_bytes,
                _lib.dptr(ws), nbytes, C.c_void_p(stream))
        if do_profile:
            torch.cuda.nvtx.range_pop()
        _lib.check(rc)
        del keep
        self.last_status = status
        self.last_stats = None if stats is None else stats[..., :2]
        self.last_trace = (stats[..., 2:].contiguous().view(torch.float32).view(S, self.rnn_num_layers, B, T, 2)
                           if T else None)
        return pose, hT, (cfg, ckpt, ckpt_bytes)

    def _cached_workspace(self, cfg, nbytes, dev, stream):
        """"Prepare the weights once" (fp16x3 inference): the library packs the weights into the workspace on every call
        unless ``cfg.weights_prepacked`` says the images of an earlier call are still there.  The module keeps that
        workspace and hands it back while nothing the images depend on has changed: same cfg, device, stream (forwards of
        one module on one stream are ordered, so the scratch part of the workspace is never shared), and every parameter
        at the same address and version counter (writes through ``param.data`` bypass the counter: call
        ``invalidate_packed_weights()`` after such an update).  Returns None when the precision has no packed-weight reuse."""
        self.last_prepacked = False
        if cfg.precision != _lib.PRECISION["fp16x3"]:
            return None
        fields = tuple(getattr(cfg, n) for n, _ in cfg._fields_ if n not in ("weights_prepacked", "reserved"))
        params = tuple((p.data_ptr(), p._version) for p in self.parameters())
        key = (fields, str(dev), stream, params, nbytes)
        cached = getattr(self, "_ws_cache", None)
        if cached is not None and cached[0] == key:
            cfg.weights_prepacked = 1
            self.last_prepacked = True
            return cached[1]
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self._ws_cache = (key, ws)
        return ws

    def invalidate_packed_weights(self):
        self._ws_cache = None

    def evolve_state(self, state, ts):
        """``PoseODERNN.evolve_state`` (reference PoseODERNN.py:70-75): the IVP ``y' = ODEFunc(y)`` from
        ``ts[:, 0]`` to ``ts[:, -1]`` for one layer's hidden state ``[B, D]`` (intermediate columns of
        ``ts`` are solved as consecutive intervals; the reference passes two).  Runs the fused kernel
        with ``evolve_only = 1``: solver loop only, no jump, no pose head.  Inference only."""
        lib = _lib.load()
        if not state.is_cuda:
            raise _lib.OdevioError("PoseODERNN.evolve_state needs CUDA tensors: odevio_b200 has no CPU path")
        y0 = _f32c(state.detach(), "state")
        tsc = _f32c(ts.detach(), "ts")
        if y0.dim() != 2 or y0.shape[1] != self.f_len or tsc.dim() != 2 or tsc.shape[0] != y0.shape[0] or tsc.shape[1] < 2:
            raise _lib.OdevioError(f"evolve_state expects state [B,{self.f_len}] and ts [B,>=2], got "
                                   f"{tuple(state.shape)}, {tuple(ts.shape)}")
        B, S = y0.shape[0], tsc.shape[1] - 1
        dev = y0.device
        cfg = self._cfg(B, S)
        cfg.L = 1
        cfg.evolve_only = 1
        nbytes = lib.odevio_odernn_workspace_bytes(C.byref(cfg))
        if nbytes == 0:
            raise _lib.OdevioError("unsupported configuration for the fused kernel")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(1, B, self.f_len, dtype=torch.float32, device=dev)
        T = self.trace_steps
        stats = torch.zeros(S, 1, B, 2 + 2 * T, dtype=torch.int32, device=dev) if (self.collect_stats or T) else None
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        w, keep = self._weights()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.odevio_odernn_forward(
                C.byref(cfg), C.byref(w), None, None, 0, _lib.dptr(tsc, "ts"), _lib.dptr(y0.unsqueeze(0), "state"),
                None, _lib.dptr(out), _lib.dptr(stats), _lib.dptr(status), None, 0,
                _lib.dptr(ws), nbytes, C.c_void_p(stream))
        _lib.check(rc)
        del keep
        self.last_status = status
        self.last_stats = None if stats is None else stats[..., :2]
        return out[0]

    def update_method(self):
        """Reference PoseODERNN.update_method (PoseODERNN.py:77-86): switch the solver to Euler.  Euler has no embedded
        error estimate, so the controller accepts every step and never changes the step size: each interval is walked in
        steps of ``ode_dt0`` (1e-4 s: ~1000 steps per 0.1 s interval), as restated in oracle/torchode_like.py [recalled
        torchode semantics].  Training through it therefore needs ``opt.ode_ckpt_loops`` >= interval / ode_dt0
        checkpoint slots per interval (default 16): the forward reports ODEVIO_STATUS_CKPT_OVERFLOW otherwise and
        ``loss.backward()`` raises, naming the knob."""
        self.ode_solver = self._set_solver("euler")

    def check_status(self):
        """Synchronising check of the last forward's per-row solver status."""
        if self.last_status is None:
            return
        bad = int(self.last_status.max().item())
        if bad != 0:
            what = {1: "max_steps reached", 2: "non-finite error norm",
                    3: "more solver iterations per interval than ode_ckpt_loops (training checkpoints; a solver without "
                       "error control such as euler takes interval / ode_dt0 steps)"}.get(bad, str(bad))
            raise RuntimeError(f"ODE solve failed for some rows: {what}")


class PoseCDE(nn.Module):
    """Neural-CDE pose regressor (reference src/models/PoseCDE.py:41-112); ``forward`` launches the
    fused cooperative sm_100a kernel (``odevio_cde_forward``).

    Same constructor namespace, attributes and state_dict keys as the reference: ``fuse``,
    ``reduction_net`` (constructed, never used -- PoseCDE.py:53-57 -- so the fused feature width must
    equal ``cde_hidden_dim``), ``initial``, ``cde_func``, ``regressor``; eval-mode ``history`` growth
    and absolute eval timestamps (PoseCDE.py:81,88-92) are reproduced.  Returns ``(poses, z0)``.

    Optional ``opt`` attributes (reference values are the defaults): ``cde_interp`` "linear"
    (rectilinear, integrated over batch row 0's times in seconds) | "cubic" (north_star: Hermite
    cubics with backward differences, integrated over the knot grid), ``cde_atol`` 1e-6,
    ``cde_rtol`` 1e-4 (PoseCDE.py:101), ``cde_step_size`` (fixed-grid rk4), ``cde_max_steps``, ``cde_history_limit``
    (cubic mode: eval-mode history bounded to that many observations, same poses; default: the reference's unbounded growth),
    ``cde_rows_per_tile``, ``cde_precision`` ("auto" | "fp32" | "fp16x3": CUDA-core kernel or the tcgen05 kernel with the
    final Linear's weights resident in shared memory, both fp32-accurate), ``cde_ckpt_steps`` (training: accepted solver steps the checkpoints hold, default 256),
    ``cde_bwd_record_gb`` (training: bound on the record streams of the deferred weight-gradient GEMMs, default 4).
    Training: under autograd ``forward`` runs ``odevio_cde_forward_ckpt`` and ``loss.backward()`` runs the fused
    ``odevio_cde_backward`` (odevio_b200/autograd.py) -- discretise-then-optimise with the accepted step sizes as
    constants, which is what the reference's ``cdeint(adjoint=False)`` computes up to torchdiffeq's step-size terms; the
    reference's ``adjoint`` flag only changes how (not which) gradients are computed and is accepted and ignored.
    """

    SOLVERS = ("dopri5", "rk4")

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.adjoint = getattr(opt, "adjoint", False)
        self.f_len = opt.v_f_len + opt.i_f_len
        self.input_dim = opt.cde_hidden_dim + 1
        self.cde_hidden_dim = opt.cde_hidden_dim
        self.cde_num_layers = getattr(opt, "cde_num_layers", 3)        # unused, as in the reference
        self.cde_fn_num_layers = opt.cde_fn_num_layers
        self.fuse_method = opt.fuse_method
        self.fuse = FusionModule(self.f_len, opt.fuse_method)
        self.reduction_net = nn.Sequential(nn.Linear(self.f_len, self.f_len // 2), nn.LeakyReLU(0.1, inplace=True),
                                           nn.Linear(self.f_len // 2, opt.cde_hidden_dim))
        self.initial = nn.Sequential(nn.Linear(opt.cde_hidden_dim + 1, opt.cde_hidden_dim), nn.Tanh())
        self.cde_func = CDEFunc(feature_dim=self.input_dim, hidden_dim=opt.cde_hidden_dim,
                                num_hidden_layers=opt.cde_fn_num_layers, activation=opt.cde_activation_fn)
        self.regressor = nn.Sequential(nn.Linear(self.cde_hidden_dim, 128), nn.LeakyReLU(0.1, inplace=True),
                                       nn.Linear(128, 6))
        if opt.cde_solver not in self.SOLVERS:
            raise ValueError(f"Solver {opt.cde_solver} not supported")
        self.solver = opt.cde_solver
        self.interp = getattr(opt, "cde_interp", "linear")
        if self.interp not in _lib.CDE_INTERP:
            raise ValueError(f"control path {self.interp} not supported")
        self.atol = float(getattr(opt, "cde_atol", 1e-6))
        self.rtol = float(getattr(opt, "cde_rtol", 1e-4))
        self.step_size = getattr(opt, "cde_step_size", None)
        self.max_steps = int(getattr(opt, "cde_max_steps", 100000))
        self.rows_per_tile = int(getattr(opt, "cde_rows_per_tile", 0))
        self.history_limit = getattr(opt, "cde_history_limit", None)     # cubic mode: observations kept across windows
        self.precision = getattr(opt, "cde_precision", "auto")           # "auto" | "fp32" | "fp16x3"
        if self.precision not in ("auto", "fp32", "fp16x3"):
            raise ValueError(f"cde_precision {self.precision} not supported")
        self.last_precision = None
        self.ckpt_steps = int(getattr(opt, "cde_ckpt_steps", 256))       # training: accepted solver steps the checkpoints hold
        self.bwd_record_gb = float(getattr(opt, "cde_bwd_record_gb", 4.0))   # training: bound on the backward's record streams
        self.history = None          # (tobs [B,n], fv [B,n,.], fi [B,n,.] | None) in eval mode
        self.last_stats = None       # int32 [4]: n_steps, n_accepted, n_f_evals, status
        self.last_hidden = None      # [B,S,Hc]

    def get_reduction_net_params(self):
        return self.reduction_net.parameters()

    def get_regressor_params(self):
        return self.regressor.parameters()

    def get_other_params(self):
        return [p for n, p in self.named_parameters() if not n.startswith("regressor")]

    def forward(self, fv, fi, ts, prev=None, do_profile=False):
        lib = _lib.load()
        if not fv.is_cuda:
            raise _lib.OdevioError("PoseCDE.forward needs CUDA tensors: odevio_b200 has no CPU path")
        if self.f_len != self.cde_hidden_dim:
            raise _lib.OdevioError(f"PoseCDE needs v_f_len + i_f_len == cde_hidden_dim (reduction_net is unused in "
                                   f"the reference, PoseCDE.py:53-57,62): {self.f_len} != {self.cde_hidden_dim}")
        B, S = fv.shape[0], fv.shape[1]
        dev = fv.device
        if self.fuse_method == "cat":
            fvc, fic, Dv = _f32c(fv, "fv"), _f32c(fi, "fi"), fv.shape[2]
        else:
            fvc, fic, Dv = _f32c(self.fuse(fv, fi), "fused"), None, self.f_len
        ts = _f32c(ts, "ts")
        ts_diff = (ts - ts[:, :1]) if self.training else ts                       # PoseCDE.py:81
        tobs = ts_diff[:, 1:].contiguous()
        if not self.training:                                                     # PoseCDE.py:88-92
            if prev is not None and self.history is not None:
                h_t, h_v, h_i = self.history
                tobs = torch.cat([h_t, tobs], 1).contiguous()
                fvc = torch.cat([h_v, fvc], 1).contiguous()
                fic = None if fic is None else torch.cat([h_i, fic], 1).contiguous()
            lim = self.history_limit
            if lim and self.interp == "cubic" and tobs.shape[1] > max(lim, S + 1):
                # Bounded history.  The reference lets `history` grow without bound (PoseCDE.py:88-92)
                # although the window's solve only reads the path on [knot n - S, knot n - 1]: the Hermite cubic with backward
                # differences needs one observation before the window and nothing older, so keeping the last
                # max(limit, S + 1) observations leaves the poses unchanged (tests/test_cde_gpu.py).  Not offered for the
                # reference's rectilinear path: it is integrated over ABSOLUTE seconds on an integer knot grid, so its
                # result depends on how many knots precede the window.
                keep_n = max(lim, S + 1)
                tobs, fvc = tobs[:, -keep_n:].contiguous(), fvc[:, -keep_n:].contiguous()
                fic = None if fic is None else fic[:, -keep_n:].contiguous()
            self.history = (tobs.detach(), fvc.detach(), None if fic is None else fic.detach())
        else:
            self.history = None
        So = tobs.shape[1]
        if self.interp == "linear":
            tout = ts_diff[0, 1:].double().contiguous()       # batch row 0's times, seconds (PoseCDE.py:101)
        else:
            tout = torch.arange(So - S, So, dtype=torch.float64, device=dev)
        z0_in = None if prev is None else _f32c(prev, "prev")
        if z0_in is not None and tuple(z0_in.shape) != (B, self.cde_hidden_dim):
            raise _lib.OdevioError(f"prev must be [B,Hc]={B, self.cde_hidden_dim}, got {tuple(z0_in.shape)}")

        needs_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self._param_list()) or fvc.requires_grad or
            (fic is not None and fic.requires_grad) or (z0_in is not None and z0_in.requires_grad))
        if do_profile:
            torch.cuda.nvtx.range_push("cdeint")
        if needs_grad:
            from .autograd import cde_apply
            pose, z0, hidden = cde_apply(self, Dv, tobs, tout, fvc, fic, z0_in)
        else:
            pose, z0, hidden, stats, _ = self._launch(tobs, fvc, fic, Dv, tout, z0_in)
            self.last_stats = stats
        if do_profile:
            torch.cuda.nvtx.range_pop()
        self.last_hidden = hidden
        return pose, z0                                                          # PoseCDE.py:103 returns z0

    def _param_list(self):
        """Flat parameter order shared by the autograd bridge's forward and backward."""
        ps = []
        for lin in self.cde_func.linears():
            ps += [lin.weight, lin.bias]
        ps += [self.initial[0].weight, self.initial[0].bias, self.regressor[0].weight, self.regressor[0].bias,
               self.regressor[2].weight, self.regressor[2].bias]
        return ps

    def _cfg(self, B, S, So, train=False):
        """``precision``: "fp32" = the CUDA-core kernel, "fp16x3" = the tensor-core kernel (loud failure for shapes it does
        not take), "auto" (default) = the tensor-core kernel whenever it takes the shape.  Training forwards (checkpoints for
        the fused backward) follow the same rule: both kernels write them, the backward reads either layout."""
        cfg = self._cfg_base(B, S, So)
        if self.precision == "fp32":
            return cfg
        cfg.precision = _lib.PRECISION["fp16x3"]
        if self.precision == "auto" and _lib.load().odevio_cde_workspace_bytes(C.byref(cfg)) == 0:
            cfg.precision = _lib.PRECISION["fp32"]
        return cfg

    def _cfg_base(self, B, S, So):
        cfg = _lib.default_cde_cfg()
        cfg.B, cfg.S, cfg.So, cfg.Hc = B, S, So, self.cde_hidden_dim
        cfg.n_layers = self.cde_fn_num_layers
        cfg.activation = _lib.ACT[self.cde_func.activation]
        cfg.solver, cfg.interp = _lib.CDE_SOLVER[self.solver], _lib.CDE_INTERP[self.interp]
        cfg.atol, cfg.rtol = self.atol, self.rtol
        step = self.step_size
        if step is None and self.solver == "rk4" and self.interp == "linear":
            step = 1.0            # torchcde injects min(diff(grid_points)) for fixed-grid solvers
        cfg.step_size = float(step or 0.0)
        cfg.max_steps, cfg.rows_per_tile = self.max_steps, self.rows_per_tile
        return cfg

    def _weight_struct(self, params):
        """odevio_cde_weights over the flat parameter list (``_param_list`` order)."""
        w = _lib.CdeWeights()
        keep = [_f32c(p.detach(), "parameter") for p in params]
        n = self.cde_fn_num_layers + 1
        for j in range(n):
            w.cde_w[j], w.cde_b[j] = _lib.dptr(keep[2 * j]), _lib.dptr(keep[2 * j + 1])
        k = 2 * n
        w.init_w, w.init_b, w.reg_w0, w.reg_b0, w.reg_w1, w.reg_b1 = (_lib.dptr(keep[k + i]) for i in range(6))
        return w, keep

    def _launch(self, tobs, fvc, fic, Dv, tout, z0_in, save_ckpt=False):
        """One ``odevio_cde_forward[_ckpt]`` launch.  Returns pose, z0, hidden, stats and (cfg, ckpt, ckpt_steps)."""
        lib = _lib.load()
        B, So = tobs.shape
        S = tout.shape[0]
        dev = tobs.device
        with torch.cuda.device(dev):
            cfg = self._cfg(B, S, So, train=save_ckpt)
            nbytes = lib.odevio_cde_workspace_bytes(C.byref(cfg))
        if nbytes == 0:
            raise _lib.OdevioError(f"unsupported PoseCDE configuration for the fused kernel (Hc={cfg.Hc}, "
                                   f"n={cfg.n_layers}, S={S}, So={So})")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        pose = torch.empty(B, S, 6, dtype=torch.float32, device=dev)
        z0 = torch.empty(B, self.cde_hidden_dim, dtype=torch.float32, device=dev)
        hidden = torch.empty(B, S, self.cde_hidden_dim, dtype=torch.float32, device=dev)
        stats = torch.zeros(4, dtype=torch.int32, device=dev)
        w, keep = self._weight_struct(self._param_list())
        stream = torch.cuda.current_stream(dev).cuda_stream
        ckpt, cap = None, 0
        with torch.cuda.device(dev):
            if save_ckpt:
                cap = self.ckpt_steps
                cbytes = lib.odevio_cde_ckpt_bytes(C.byref(cfg), cap)
                ckpt = torch.empty(cbytes, dtype=torch.uint8, device=dev)
                rc = lib.odevio_cde_forward_ckpt(
                    C.byref(cfg), C.byref(w), _lib.dptr(tobs, "tobs"), _lib.dptr(fvc, "fv"), _lib.dptr(fic, "fi"), Dv,
                    _lib.dptr(tout, "tout"), _lib.dptr(z0_in, "prev"), _lib.dptr(pose), _lib.dptr(z0),
                    _lib.dptr(hidden), _lib.dptr(stats), _lib.dptr(ckpt), cbytes, cap, _lib.dptr(ws), nbytes,
                    C.c_void_p(stream))
            else:
                rc = lib.odevio_cde_forward(
                    C.byref(cfg), C.byref(w), _lib.dptr(tobs, "tobs"), _lib.dptr(fvc, "fv"), _lib.dptr(fic, "fi"), Dv,
                    _lib.dptr(tout, "tout"), _lib.dptr(z0_in, "prev"), _lib.dptr(pose), _lib.dptr(z0),
                    _lib.dptr(hidden), _lib.dptr(stats), _lib.dptr(ws), nbytes, C.c_void_p(stream))
        _lib.check(rc)
        del keep
        self.last_precision = "fp16x3" if cfg.precision == _lib.PRECISION["fp16x3"] else "fp32"
        return pose, z0, hidden, stats, (cfg, ckpt, cap)

    def check_status(self):
        """Synchronising check of the last forward's solver status."""
        if self.last_stats is None:
            return
        bad = int(self.last_stats[3].item())
        if bad != 0:
            what = {1: "max_steps reached", 2: "non-finite error norm"}.get(bad, str(bad))
            raise RuntimeError(f"CDE solve failed: {what}")
