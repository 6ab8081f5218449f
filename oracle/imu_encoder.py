"""CPU restatement of the reference's InertialEncoder (src/models/Encoder.py:39-74).  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and the golden generator may import it.

PINNED: `oracle/make_imu_golden.py` (run in the authoring container, where /root/reference is mounted) imports the
reference's own `src.models.Encoder.InertialEncoder`, checks this restatement against it bit for bit on CPU and freezes
tests/golden/imu_encoder.pt (seeded weights incl. non-trivial BatchNorm statistics, input, reference output)."""

import torch
import torch.nn as nn
import torch.nn.functional as F


class OracleInertialEncoder(nn.Module):
    """Same construction order, module types and state_dict keys as the reference (Encoder.py:40-58)."""

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
        """x [B, 10*S + 1, 6] -> [B, S, i_f_len] (Encoder.py:60-74), written out functionally for eval mode."""
        B = x.shape[0]
        S = (x.shape[1] - 1) // 10                                                   # Encoder.py:61
        win = torch.stack([x[:, i * 10:i * 10 + 11, :] for i in range(S)], dim=1)    # [B, S, 11, 6]   :62-65
        h = win.reshape(B * S, 11, 6).permute(0, 2, 1)                               # [B*S, 6, 11]    :69-72
        if self.training:
            h = self.encoder_conv(h)
        else:
            for k in range(3):
                conv, bn = self.encoder_conv[4 * k], self.encoder_conv[4 * k + 1]
                h = F.conv1d(h, conv.weight, conv.bias, padding=1)
                h = F.batch_norm(h, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
                h = F.leaky_relu(h, 0.1)
        out = F.linear(h.reshape(B * S, -1), self.proj.weight, self.proj.bias)       # channel-major flatten  :73
        return out.view(B, S, self.i_f_len)


def randomize_batchnorm(model, seed=0):
    """Non-trivial affine parameters and running statistics (a fresh BatchNorm is the identity)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.weight.data = 0.5 + torch.rand(m.num_features, generator=g)
            m.bias.data = 0.2 * torch.randn(m.num_features, generator=g)
            m.running_mean.data = 0.3 * torch.randn(m.num_features, generator=g)
            m.running_var.data = 0.5 + torch.rand(m.num_features, generator=g)


def imu_like(B, S, seed=0):
    """Synthetic IMU rows with the per-channel scale of KITTI's accelerometer / gyroscope (src/data/transforms.py:24-26
    normalises them; here: unit-variance accelerations around gravity on z, small angular rates)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 10 * S + 1, 6, generator=g)
    x[..., :3] *= 1.0
    x[..., 2] += 9.8
    x[..., 3:] *= 0.1
    return x
