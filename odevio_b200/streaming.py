"""Windowed / streaming inference driver.

The reference evaluates a trajectory as a chain of 11-frame windows stepping by 10, carrying the
regressor state across windows and using ABSOLUTE timestamps once a state is carried
(src/data/KITTI_eval.py:124-160, src/models/PoseODERNN.py:97-100) -- at batch 1.  This helper does
the same for many trajectories in lock-step (one row per trajectory), so long sequences run as
repeated fused-kernel launches with `prev`; nothing here computes on the path.
"""

import torch


class StreamingPoseODERNN:
    """Carry the hidden state of a :class:`odevio_b200.PoseODERNN` across consecutive windows.

        stream = StreamingPoseODERNN(model)
        for fv, fi, ts in windows:              # fv [B,S,512], fi [B,S,256], ts [B,S+1] absolute seconds
            poses = stream.step(fv, fi, ts)     # [B,S,6]

    The first window is run exactly like the reference's first call (prev=None: zero state, times
    relative to the window start); every later window continues from the carried state with absolute
    timestamps.  `reset(rows)` zeroes the state of trajectories that restart."""

    def __init__(self, model):
        self.model = model
        self.state = None

    def reset(self, rows=None):
        if rows is None or self.state is None:
            self.state = None
        else:
            self.state[:, rows] = 0.0

    @torch.no_grad()
    def step(self, fv, fi, ts):
        poses, self.state = self.model(fv, fi, ts, prev=self.state)
        return poses

    @torch.no_grad()
    def run(self, fv, fi, ts, window=10):
        """Whole trajectories: fv [B,T,.], fi [B,T,.], ts [B,T+1] -> poses [B,T,6], windows of `window` steps."""
        out = []
        T = fv.shape[1]
        for a in range(0, T, window):
            b = min(a + window, T)
            out.append(self.step(fv[:, a:b].contiguous(), fi[:, a:b].contiguous(), ts[:, a:b + 1].contiguous()))
        return torch.cat(out, 1)
