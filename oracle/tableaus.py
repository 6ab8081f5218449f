"""Butcher tableaus used by the oracle and (as the same numbers) by the CUDA path.

All coefficients are held in float64 here; solvers cast them to the state dtype.
``a`` is a list of rows (row i has i entries), ``b`` the solution weights,
``e`` the error weights (b - b_low, sign irrelevant under abs), ``b_mid`` the
weights giving y(t0 + dt/2) for the 4th-order dense output.

Sources: Dormand-Prince 5(4) [Dormand & Prince 1980] — identical to
``scipy.integrate._ivp.rk.RK45.{A,B,C,E,P}`` (checked in tests/test_tableaus.py);
Tsitouras 5(4) [Tsitouras 2011], verified in the tests through the order
conditions; Heun/Euler/RK4 are textbook.  torchode's menu is the reference's
``PoseODERNN._set_solver`` (src/models/PoseODERNN.py:125-137).
"""

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass(frozen=True)
class Tableau:
    name: str
    c: List[float]
    a: List[List[float]]          # a[i] has i entries (a[0] == [])
    b: List[float]                # len == n_stages
    e: Optional[List[float]]      # error weights or None (no embedded method)
    order: int                    # convergence order used by the step controller
    fsal: bool                    # last stage's vf equals next step's first
    ssal: bool                    # y1 equals the last stage's argument
    b_mid: Optional[List[float]] = None   # dense-output midpoint weights
    extra: dict = field(default_factory=dict)

    @property
    def n_stages(self) -> int:
        return len(self.b)


# ---------------------------------------------------------------- dopri5 ----
_DP_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_B = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
_DP_BLOW = [5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40]
_DP_E = [b - bl for b, bl in zip(_DP_B, _DP_BLOW)]
# torchdiffeq DPS_C_MID == RK45.P @ [1/2, 1/4, 1/8, 1/16]
_DP_BMID = [
    6025192743 / 30085553152 / 2,
    0.0,
    51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2,
    11237099 / 235043384 / 2,
]

DOPRI5 = Tableau(
    name="dopri5",
    c=[0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0],
    a=_DP_A, b=_DP_B, e=_DP_E, order=5, fsal=True, ssal=True, b_mid=_DP_BMID,
)

# ----------------------------------------------------------------- tsit5 ----
_TS_A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
     -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
     -3.290069515436081, 2.324710524099774],
]
_TS_B = _TS_A[6] + [0.0]
_TS_E = [
    -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
    -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 1 / 66,
]


def tsit5_dense_weights(theta: float) -> List[float]:
    """Tsitouras' free 4th-order interpolant b_i(theta); b_i(1) == b_i."""
    t = theta
    b1 = -1.0530884977290216 * t * (t - 1.3299890189751412) * (
        t * t - 1.4364028541716351 * t + 0.7139816917074209)
    b2 = 0.1017 * t * t * (t * t - 2.1966568338249754 * t + 1.2949852507374631)
    b3 = 2.490627285651252793 * t * t * (
        t * t - 2.38535645472061657 * t + 1.57803468208092486)
    b4 = -16.54810288924490272 * (t - 1.21712927295533244) * (
        t - 0.61620406037800089) * t * t
    b5 = 47.37952196281928122 * (t - 1.203071208372362603) * (
        t - 0.658047292653547382) * t * t
    b6 = -34.87065786149660974 * (t - 1.2) * (t - 0.666666666666666667) * t * t
    b7 = 2.5 * (t - 1.0) * (t - 0.6) * t * t
    return [b1, b2, b3, b4, b5, b6, b7]


TSIT5 = Tableau(
    name="tsit5",
    c=[0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0],
    a=_TS_A, b=_TS_B, e=_TS_E, order=5, fsal=True, ssal=True,
    b_mid=tsit5_dense_weights(0.5),
)

# ------------------------------------------------------------ heun / euler --
# torchode Heun: explicit trapezoid with embedded Euler, adaptive, order 2.
HEUN = Tableau(
    name="heun", c=[0.0, 1.0], a=[[], [1.0]], b=[0.5, 0.5], e=[-0.5, 0.5],
    order=2, fsal=False, ssal=False, b_mid=None,
)
# torchode Euler: no embedded error estimate -> every step accepted, dt constant.
EULER = Tableau(
    name="euler", c=[0.0], a=[[]], b=[1.0], e=None, order=1, fsal=False, ssal=False,
)

# ------------------------------------------------------- fixed-grid RK4s ----
# classic RK4 (north_star "fixed-step rk4")
RK4 = Tableau(
    name="rk4", c=[0.0, 0.5, 0.5, 1.0],
    a=[[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]],
    b=[1 / 6, 1 / 3, 1 / 3, 1 / 6], e=None, order=4, fsal=False, ssal=False,
)
# torchdiffeq's "rk4" is the 3/8 rule
RK4_38 = Tableau(
    name="rk4_38", c=[0.0, 1 / 3, 2 / 3, 1.0],
    a=[[], [1 / 3], [-1 / 3, 1.0], [1.0, -1.0, 1.0]],
    b=[1 / 8, 3 / 8, 3 / 8, 1 / 8], e=None, order=4, fsal=False, ssal=False,
)

BY_NAME = {t.name: t for t in (DOPRI5, TSIT5, HEUN, EULER, RK4, RK4_38)}
