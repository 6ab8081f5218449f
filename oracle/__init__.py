"""CPU oracle for the ODE-VIO latent-dynamics integration path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the
reported CPU baseline.  The product path (``odevio_b200``) never imports this
package and fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The arithmetic of the reference's hot path lives in three
third-party packages that are neither vendored under /root/reference nor
installable offline: ``torchode==0.2.0``, ``torchcde==0.2.5`` and
``torchdiffeq==0.2.3`` (reference ``requirements.txt:9-11``).  The reference
ships no tests, golden vectors or saved outputs for this path.  This package is
therefore a plain-PyTorch *restatement* of those libraries' published
algorithms, anchored on

  * the reference's own call sites (``src/models/PoseODERNN.py:55-60,70-75,88-123``,
    ``src/models/PoseCDE.py:76-103``, ``src/models/ODEFunc.py:5-39,44-84``,
    ``src/models/FusionModule.py:17-29``),
  * the reference's importable classes (``ODEFunc``, ``CDEFunc``,
    ``FusionModule``): ``oracle/make_golden.py`` checks our module restatement
    against them bit-for-bit and freezes golden vectors under ``tests/golden/``,
  * SciPy's Dormand-Prince tableau / dense output (``scipy.integrate._ivp.rk.RK45``)
    and ``solve_ivp`` converged solutions, and ``torch.nn.RNN`` / ``nn.GRU``.

Every behavioural choice that could not be verified against the real libraries
is a named switch in :class:`oracle.torchode_like.ControllerOptions` /
:mod:`oracle.torchdiffeq_like` with the believed-reference default.
"""

from . import tableaus  # noqa: F401
