"""Oracle restatement of the torchcde 0.2.5 pieces ``PoseCDE.forward`` uses (test
infrastructure; parity unpinned -- see oracle/__init__.py).  Call sites: reference
src/models/PoseCDE.py:94-96,101.  Semantics follow SURVEY.md A.2 / Appendix B.

  * :func:`linear_interpolation_coeffs` -- ``cde.linear_interpolation_coeffs(x, rectilinear=0)``
  * :class:`LinearInterpolation`       -- ``cde.LinearInterpolation(coeffs)`` (knots on the integer grid)
  * :class:`HermiteCubicBackward`      -- north_star "cubic-spline control path": Hermite cubics
    with backward differences on the integer knot grid (torchcde's
    ``hermite_cubic_coefficients_with_backward_differences`` + ``CubicSpline``), NOT in the reference.
"""

import torch


def linear_interpolation_coeffs(x, rectilinear=None):
    """x [B,S,C] -> coeffs.  rectilinear=0 lags channel 0 (time): length 2S-1,
    (t1,x1),(t2,x1),(t2,x2),... -- time moves first, then the values."""
    if rectilinear is None:
        return x
    r = x.repeat_interleave(2, dim=-2)
    r = r.clone()
    r[..., :-1, rectilinear] = r[..., 1:, rectilinear].clone()
    return r[..., :-1, :]


def segment_index(t, grid):
    """torchcde ``_interpret_t``: idx = clamp(bucketize(t, grid) - 1, 0, len(grid) - 2), bucketize
    with right=False (a t exactly on knot k > 0 belongs to segment k - 1)."""
    maxlen = grid.numel() - 2
    idx = (torch.bucketize(t.detach(), grid) - 1).clamp(0, maxlen)
    return int(idx)


class LinearInterpolation:
    def __init__(self, coeffs):
        self.coeffs = coeffs                                   # [B, n, C]
        n = coeffs.shape[-2]
        self.grid = torch.linspace(0, n - 1, n, dtype=coeffs.dtype, device=coeffs.device)

    @property
    def interval(self):
        return torch.stack([self.grid[0], self.grid[-1]])

    @property
    def grid_points(self):
        return self.grid

    def evaluate(self, t):
        i = segment_index(t, self.grid)
        frac = t - self.grid[i]
        prev, nxt = self.coeffs[..., i, :], self.coeffs[..., i + 1, :]
        return prev + frac * (nxt - prev) / (self.grid[i + 1] - self.grid[i])

    def derivative(self, t):
        i = segment_index(t, self.grid)
        return (self.coeffs[..., i + 1, :] - self.coeffs[..., i, :]) / (self.grid[i + 1] - self.grid[i])


class HermiteCubicBackward:
    """X on knots 0..S-1 through x [B,S,C]; on segment i (s = t - i in [0,1]):
    X'(s) = m_i + (d_i - m_i) * (4 - 3 s) * s,  d_i = x_{i+1} - x_i,  m_i = d_{i-1} (m_0 = d_0)
    (SURVEY.md Appendix B with h = 1: b = m_i, 2c = 4 (d_i - m_i), 3d = 3 (m_i - d_i))."""

    def __init__(self, x):
        self.x = x
        n = x.shape[-2]
        self.grid = torch.linspace(0, n - 1, n, dtype=x.dtype, device=x.device)

    @property
    def interval(self):
        return torch.stack([self.grid[0], self.grid[-1]])

    @property
    def grid_points(self):
        return self.grid

    def evaluate(self, t):
        i = segment_index(t, self.grid)
        s = t - self.grid[i]
        d = self.x[..., i + 1, :] - self.x[..., i, :]
        m = d if i == 0 else self.x[..., i, :] - self.x[..., i - 1, :]
        # a + b s + c s^2 + d3 s^3 with c = 2 (d - m), d3 = (m - d)
        return self.x[..., i, :] + (m + ((d - m) * 2.0 + (m - d) * s) * s) * s

    def derivative(self, t):
        i = segment_index(t, self.grid)
        s = t - self.grid[i]
        d = self.x[..., i + 1, :] - self.x[..., i, :]
        m = d if i == 0 else self.x[..., i, :] - self.x[..., i - 1, :]
        return m + (d - m) * ((4.0 - 3.0 * s) * s)
