"""Training-loop glue of the regressor path on the device.

Reference: scripts/train_model.py:69-86 (loss = 100 * MSE(angles) + MSE(translations), backward, global-norm clip 5,
optimizer.step) and src/utils/utils.py:143-157 (Adam over the two ``Pose_net`` parameter groups, lr, betas (0.9, 0.999),
eps 1e-8, weight_decay 5e-5).  Here:

* :func:`fused_pose_loss` -- loss value(s) and d loss / d poses in ONE kernel (``odevio_pose_loss``), as an autograd
  function so it drops into the reference's ``loss.backward()``;
* :class:`FusedPoseNetAdam` -- the ``Pose_net`` parameters live as views of one flat fp32 buffer (state_dict keys and
  shapes unchanged, reference checkpoints still load with ``load_state_dict``); the gradients are gathered into the flat
  bucket that the NCCL all-reduce needs anyway, and ``odevio_adam_step`` runs clip + Adam on that bucket in three
  launches with the clip coefficient kept on the device -- the optimiser step is the all-reduce's epilogue;
* :func:`fused_train_step` -- forward (fused kernels), loss, backward (fused kernels), all-reduce, clip + Adam.

No CPU path: CUDA tensors only (the reference-equivalent torch ops remain in ``odevio_b200.distributed`` and are what the
tests compare against -- torch.optim.Adam / clip_grad_norm_ / mse_loss are the reference's own dependencies).
"""

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .distributed import pose_net_params


def _workspace(dev):
    lib = _lib.load()
    n = lib.odevio_train_glue_workspace_bytes()
    return torch.empty(n, dtype=torch.uint8, device=dev), n


class _PoseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, poses, gts, w_angle):
        lib = _lib.load()
        if not poses.is_cuda:
            raise _lib.OdevioError("fused_pose_loss needs CUDA tensors: odevio_b200 has no CPU path")
        p = poses.detach().to(torch.float32).contiguous()
        g = gts.detach().to(torch.float32).contiguous()
        if p.shape != g.shape or p.shape[-1] != 6:
            raise _lib.OdevioError(f"poses / gts must both be [..., 6], got {tuple(p.shape)} / {tuple(g.shape)}")
        n_rows = p.numel() // 6
        loss3 = torch.empty(3, dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p) if poses.requires_grad else None
        ws, nbytes = _workspace(p.device)
        with torch.cuda.device(p.device):
            rc = lib.odevio_pose_loss(n_rows, _lib.dptr(p), _lib.dptr(g), float(w_angle), 1.0, _lib.dptr(loss3),
                                      _lib.dptr(grad), _lib.dptr(ws), nbytes,
                                      C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream))
        _lib.check(rc)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3
    
    @staticmethod
    def backward(ctx, gout, _g3):
        (grad,) = ctx.saved_tensors
        return (None if grad is None else grad * gout), None, None


def fused_pose_loss(poses, gts, w_angle=100.0, with_parts=False):
    """scripts/train_model.py:72-76.  Returns the scalar loss (differentiable w.r.t. ``poses``); with ``with_parts`` also the
    device tensor [loss, angle_mse, translation_mse] the reference logs (:91)."""
    loss, loss3 = _PoseLoss.apply(poses, gts, w_angle)
    return (loss, loss3) if with_parts else loss


class FusedPoseNetAdam:
    """Adam (+ global-norm clip) over the reference's ``Pose_net`` parameter set on one flat bucket.

    The reference builds TWO parameter groups, ``[other parameters, regressor]`` in that order (utils/utils.py:116-119),
    both starting at ``lr_warmup`` = 1e-4; its epoch loop then re-schedules only ``param_groups[0]['lr']``
    (scripts/train_model.py:215-216: the group-1 line is commented out), so the regressor stays at 1e-4 while the rest
    drops to 1e-5 / 1e-6.  ``param_groups`` here is the same two-entry list in the same order and ``step`` reads each
    group's ``lr`` from it, so the reference loop drives this optimiser unchanged.  The bucket itself is laid out
    regressor first (``pose_net_params``): elements ``[0, split)`` are the regressor segment."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-5, max_norm=5.0):
        reg = [p for p in model.get_regressor_params() if p.requires_grad]
        other = [p for p in model.get_other_params() if p.requires_grad]
        self.params = reg + other
        if not self.params or not self.params[0].is_cuda:
            raise _lib.OdevioError("FusedPoseNetAdam needs the model on a CUDA device: odevio_b200 has no CPU path")
        dev = self.params[0].device
        # every parameter starts on a 256-byte boundary of the bucket (the kernels' bulk-TMA / 128-bit accesses assume the
        # alignment of a fresh allocation); the padding stays exactly zero under clip + Adam (g = 0, wd * 0 = 0)
        self.offsets, off = [], 0
        for i, p in enumerate(self.params):
            if i == len(reg):
                self.split = off                   # first element of the "other" segment (a multiple of 64)
            self.offsets.append(off)
            off += (p.numel() + 63) // 64 * 64
        if not other:
            self.split = off
        self.numel = sum(p.numel() for p in self.params)
        self.flat = self._bucket(off, dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):      # parameters become views of the flat buffer (keys / shapes unchanged)
                n = p.numel()
                self.flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + n].view_as(p.data)
        self.grads = self._bucket(off, dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.norm_coef = torch.zeros(2, dtype=torch.float32, device=dev)      # [total grad norm, clip coefficient]
        self.step_count = 0
        self.betas, self.eps, self.weight_decay, self.max_norm = betas, eps, weight_decay, max_norm
        self.param_groups = [{"params": other, "lr": lr}, {"params": reg, "lr": lr}]      # utils/utils.py:116-119 order
        self._ws, self._ws_bytes = _workspace(dev)

    def _bucket(self, n, dev):
        return torch.zeros(n, dtype=torch.float32, device=dev)

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):                            # one lr for both groups (the reference's constructor state)
        for g in self.param_groups:
            g["lr"] = value

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def _check_aliasing(self):
        """model.to() / .double() or `p.data = ...` after construction would silently detach a parameter from the bucket."""
        base = self.flat.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() != base + 4 * o:
                raise _lib.OdevioError("a Pose_net parameter no longer aliases the optimiser's flat bucket (model.to() / "
                                       "re-assigned .data after FusedPoseNetAdam was built): rebuild the optimiser")

    def gather_grads(self, weight=1.0):
        """p.grad of every parameter -> the flat bucket (scaled); missing gradients count as zero."""
        for p, o in zip(self.params, self.offsets):
            dst = self.grads[o:o + p.numel()]
            if p.grad is None:
                dst.zero_()
            else:
                dst.copy_(p.grad.reshape(-1))
        if weight != 1.0:
            self.grads.mul_(weight)
        return self.grads

    def step(self):
        """clip_grad_norm_(max_norm) + Adam on the bucket (three launches, no host synchronisation); the regressor segment
        steps with ``param_groups[1]['lr']``, the rest with ``param_groups[0]['lr']``."""
        lib = _lib.load()
        self._check_aliasing()
        self.step_count += 1
        dev = self.flat.device
        with torch.cuda.device(dev):
            rc = lib.odevio_adam_step_groups(self.flat.numel(), self.split, _lib.dptr(self.flat), _lib.dptr(self.grads),
                                             _lib.dptr(self.exp_avg), _lib.dptr(self.exp_avg_sq), self.step_count,
                                             float(self.param_groups[1]["lr"]), float(self.param_groups[0]["lr"]),
                                             self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                             float(self.max_norm or 0.0), _lib.dptr(self.norm_coef),
                                             _lib.dptr(self._ws), self._ws_bytes,
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc)

    def state_dict(self):
        """Moments, step count and group learning rates (resume: scripts/train_model.py saves optimizer.state_dict())."""
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": [g["lr"] for g in self.param_groups], "betas": self.betas, "eps": self.eps,
                "weight_decay": self.weight_decay, "max_norm": self.max_norm}

    def load_state_dict(self, sd):
        if sd["exp_avg"].numel() != self.exp_avg.numel():
            raise _lib.OdevioError("optimizer state does not match this model's Pose_net bucket")
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, lr in zip(self.param_groups, sd["lr"]):
            g["lr"] = lr
        self.betas, self.eps = tuple(sd["betas"]), sd["eps"]
        self.weight_decay, self.max_norm = sd["weight_decay"], sd["max_norm"]


class PeerFusedPoseNetAdam(FusedPoseNetAdam):
    """FusedPoseNetAdam whose buckets live in symmetric memory: ``step_allreduce`` runs gradient all-reduce + clip + Adam as
    ONE kernel over NVLink peer memory (``odevio_allreduce_adam_peer``: reduce-scatter by pull, norm exchange through flag
    pads, Adam on this rank's slice, all-gather by push) instead of NCCL all-reduce + three launches.  The optimiser moments
    are sharded: each rank only ever touches its slice (``state_dict`` gathers them).  All ranks must hold identical
    parameters at construction (as under DDP) and call ``step_allreduce`` collectively.  ``world_size == 1`` (tests) uses
    plain device tensors."""

    def __init__(self, model, group=None, **kw):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._handles = []
        super().__init__(model, **kw)
        dev = self.flat.device
        self.pad = self._bucket(64, dev)                    # 256 bytes of flags / partial norms, viewed as uint32 by the kernel
        lib = _lib.load()
        n = self.flat.numel()
        self._peer_ws_bytes = lib.odevio_allreduce_adam_peer_workspace_bytes(n, self.world)
        if self._peer_ws_bytes == 0:
            raise _lib.OdevioError(f"odevio_allreduce_adam_peer: unsupported bucket / world size ({n}, {self.world})")
        self._peer_ws = torch.empty(self._peer_ws_bytes, dtype=torch.uint8, device=dev)
        self.epoch = 0

        def ptrs(t):
            if self.world == 1:
                return (C.c_void_p * 1)(t.data_ptr())
            import torch.distributed._symmetric_memory as symm_mem
            h = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            self._handles.append(h)
            return (C.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])

        self._p_params, self._p_grads, self._p_pad = ptrs(self.flat), ptrs(self.grads), ptrs(self.pad)
        if self.world > 1:
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.group)                  # every rank's buckets are zeroed / filled before any peer access

    def _bucket(self, n, dev):
        if self.world == 1:
            return torch.zeros(n, dtype=torch.float32, device=dev)
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        t.zero_()
        return t

    def step_allreduce(self):
        """All-reduce (mean) of the gathered gradient bucket + clip + Adam, one launch, collectively on every rank."""
        lib = _lib.load()
        self._check_aliasing()
        self.step_count += 1
        self.epoch += 1
        dev = self.flat.device
        with torch.cuda.device(dev):
            rc = lib.odevio_allreduce_adam_peer(
                self.flat.numel(), self.split, self.rank, self.world, self._p_params, self._p_grads, self._p_pad,
                _lib.dptr(self.exp_avg), _lib.dptr(self.exp_avg_sq), self.step_count, self.epoch, 1.0 / self.world,
                float(self.param_groups[1]["lr"]), float(self.param_groups[0]["lr"]), self.betas[0], self.betas[1], self.eps,
                self.weight_decay, float(self.max_norm or 0.0), _lib.dptr(self.norm_coef), _lib.dptr(self._peer_ws),
                self._peer_ws_bytes, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc)

    def state_dict(self):
        sd = super().state_dict()
        if self.world > 1:                                  # the moments of slice r live on rank r only
            for k in ("exp_avg", "exp_avg_sq"):
                dist.all_reduce(sd[k], op=dist.ReduceOp.SUM, group=self.group)     # the other slices are exactly zero
        return sd

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        if self.world > 1:                                  # keep only this rank's slice (the kernel never reads the rest)
            n = self.flat.numel()
            per = ((n // 4 + self.world - 1) // self.world) * 4
            lo, hi = min(per * self.rank, n), min(per * (self.rank + 1), n)
            for t in (self.exp_avg, self.exp_avg_sq):
                t[:lo].zero_(); t[hi:].zero_()


def fused_train_step(model, opt, fv, fi, ts, gts, world_size=1, group=None, events=None):
    """One optimisation step on this rank's shard (scripts/train_model.py:69-86) with the glue on the device: fused forward
    (checkpoints), fused loss + d loss / d poses, fused backward, ONE flat-bucket all-reduce, clip + Adam as its epilogue.
    ``events``: optional list that receives (name, start, end) CUDA-event triples of the phases (bench.py)."""
    def mark():
        if events is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    opt.zero_grad()
    t0 = mark()
    poses, _ = model(fv, fi, ts, prev=None)
    t1 = mark()
    loss = fused_pose_loss(poses, gts)
    loss.backward()
    peer = isinstance(opt, PeerFusedPoseNetAdam)
    flat = opt.gather_grads(1.0 / world_size if (world_size > 1 and not peer) else 1.0)
    t2 = mark()
    if world_size > 1 and not peer:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    t3 = mark()
    if peer:
        opt.step_allreduce()          # all-reduce + clip + Adam in one kernel over NVLink peer memory
    else:
        opt.step()
    t4 = mark()
    if events is not None:
        events += [("forward", t0, t1), ("loss_backward", t1, t2), ("allreduce", t2, t3), ("clip_adam", t3, t4)]
    return loss.detach()
