"""Synthetic KITTI-shaped inputs for the regressors (tests, smoke, bench).

Shapes follow the reference: fv [B,S,512], fi [B,S,256] (scripts/config.py:50-51), S = seq_len-1
= 10 (config.py:55), timestamps [B,S+1] in seconds at KITTI's 10 Hz (src/data/KITTI_eval.py:256).
Irregular sampling restates the reference's frame-drop loop (src/data/KITTI_dataset.py:63-74):
every interior frame is dropped with probability p, which merges the gaps, so the kept frames
are 0.1*m seconds apart with m >= 1 geometric; p ~ U[0, p_max] per sequence.
"""

import torch


def features(B, S=10, v_f_len=512, i_f_len=256, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    fv = torch.randn(B, S, v_f_len, generator=g)
    fi = torch.randn(B, S, i_f_len, generator=g)
    return fv.to(device), fi.to(device)


def timestamps(B, S=10, irregular=False, p_max=0.5, jitter=0.01, offset=0.0, seed=0, device="cpu",
               frame_dt=0.1):
    """[B, S+1] float32 timestamps.  offset > 0 gives the absolute-time variant used when a
    previous hidden state is carried (reference src/models/PoseODERNN.py:100)."""
    g = torch.Generator().manual_seed(seed + 1)
    if not irregular:
        gaps = torch.full((B, S), frame_dt, dtype=torch.float64)
    else:
        p = torch.rand(B, 1, generator=g, dtype=torch.float64) * p_max
        u = torch.rand(B, S, generator=g, dtype=torch.float64)
        # m ~ Geometric(1-p) on {1,2,...}: number of frames until one is kept
        m = torch.floor(torch.log1p(-u) / torch.log(p.clamp_min(1e-12))).clamp_min(0) + 1
        m = torch.where(p > 0, m, torch.ones_like(m)).clamp_max(12)
        jit = 1.0 + jitter * (2 * torch.rand(B, S, generator=g, dtype=torch.float64) - 1)
        gaps = frame_dt * m * jit
    ts = torch.cat([torch.zeros(B, 1, dtype=torch.float64), gaps.cumsum(1)], 1) + offset
    return ts.to(torch.float32).to(device)
