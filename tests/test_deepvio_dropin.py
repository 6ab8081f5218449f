"""The drop-in claim of INTEGRATION.md section 1, executed.

CPU part (runs where /root/reference is mounted, i.e. in the authoring container; skipped elsewhere): the reference's
UNMODIFIED ``src.models.DeepVIO`` is imported with its un-installable pip dependencies stubbed (torchode / torchcde /
ncps / fvcore are only touched at construction time by the reference's own regressors), the two regressor names are
re-bound the way INTEGRATION.md tells a maintainer to (``from odevio_b200 import PoseODERNN, PoseCDE``) and then
  * ``DeepVIO(opt)`` constructs with the reference's own config defaults (scripts/config.py) and runs its own
    ``initialization`` over our modules (DeepVIO.py:43, 77-87),
  * a state_dict written by the reference-built DeepVIO -- including the duplicate ``Pose_net.solver._orig_mod.*`` keys of
    the compiled torchode module -- loads with ``strict=False``: nothing missing, nothing unexpected but those duplicates,
  * the reference's ``get_optimizer`` (utils/utils.py:115-130) builds its two parameter groups from our accessors,
  * ``DeepVIO.forward`` reaches our module with the reference's call signature (DeepVIO.py:61-68) and fails LOUDLY on
    CPU tensors (no CPU path).
GPU part (no access to the reference on the GPU box): ``DeepVIO.forward``'s two lines (DeepVIO.py:64-67) with
``odevio_b200.InertialEncoder`` and ``odevio_b200.PoseODERNN`` swapped in, chunked over two windows with the carried state
``hc``, against the oracle restatements.
"""

import copy
import importlib
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

REFERENCE = "/root/reference"


def _stub_modules():
    """Minimal stand-ins for the reference's un-installable dependencies: enough for import + construction.  The
    torchode stand-ins are nn.Modules that hold their arguments, so that the reference-built model has the same
    duplicate `solver._orig_mod.*` parameter names as with the real library."""
    class Holder(nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__()
            k = 0
            for a in list(args) + list(kwargs.values()):
                if isinstance(a, nn.Module):
                    setattr(self, ("term", "step_method", "step_size_controller")[k] if type(self).__name__ == "AutoDiffAdjoint"
                            else ("f" if type(self).__name__ == "ODETerm" else f"m{k}"), a)
                    k += 1

    to = types.ModuleType("torchode")
    for name in ("ODETerm", "Dopri5", "Tsit5", "Heun", "Euler", "IntegralController", "FixedStepController", "AutoDiffAdjoint"):
        setattr(to, name, type(name, (Holder,), {}))
    to.InitialValueProblem = lambda **kw: kw
    mods = {"torchode": to, "torchcde": types.ModuleType("torchcde"), "ncps": types.ModuleType("ncps"),
            "ncps.torch": types.ModuleType("ncps.torch"), "ncps.wirings": types.ModuleType("ncps.wirings"),
            "fvcore": types.ModuleType("fvcore"), "fvcore.nn": types.ModuleType("fvcore.nn")}
    mods["ncps.torch"].CfC = mods["ncps.torch"].LTC = object
    mods["ncps.wirings"].AutoNCP = mods["ncps.wirings"].FullyConnected = object
    mods["fvcore.nn"].FlopCountAnalysis = object
    return mods


@pytest.fixture
def reference_deepvio(monkeypatch):
    if not os.path.isdir(os.path.join(REFERENCE, "src", "models")):
        pytest.skip("the reference tree is only mounted in the authoring container")
    for name, mod in _stub_modules().items():
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.syspath_prepend(REFERENCE)
    for name in [n for n in sys.modules if n == "src" or n.startswith("src.") or n == "scripts" or n.startswith("scripts.")]:
        monkeypatch.delitem(sys.modules, name)
    # torch.compile is lazy, but keep construction independent of the inductor tool chain
    monkeypatch.setattr(torch, "compile", lambda m, *a, **k: types.SimpleNamespace(_orig_mod=m) if not isinstance(m, nn.Module)
                        else _Compiled(m))
    D = importlib.import_module("src.models.DeepVIO")
    yield D
    for name in [n for n in sys.modules if n == "src" or n.startswith("src.")]:
        sys.modules.pop(name, None)


class _Compiled(nn.Module):
    """Stand-in for torch._dynamo.OptimizedModule: registers the wrapped module as `_orig_mod` (the source of the
    reference checkpoints' duplicate `solver._orig_mod.*` keys)."""

    def __init__(self, m):
        super().__init__()
        self._orig_mod = m


def _reference_opt(**over):
    """scripts/config.py defaults (reference get_args()), small images so the (out-of-scope) image encoder stays cheap."""
    sys.path.insert(0, REFERENCE)
    try:
        cfg = importlib.import_module("scripts.config")
    finally:
        sys.path.remove(REFERENCE)
    argv, sys.argv = sys.argv, ["x"]
    try:
        opt = cfg.get_args()
    finally:
        sys.argv = argv
    opt.img_w, opt.img_h = 128, 128
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


@pytest.mark.parametrize("model_type", ["ode-rnn", "cde"])
def test_reference_deepvio_constructs_and_loads_with_the_dropins(reference_deepvio, model_type):
    import odevio_b200
    from odevio_b200 import _lib
    D = reference_deepvio
    over = dict(model_type=model_type)
    if model_type == "cde":
        over.update(v_f_len=64, i_f_len=64, cde_hidden_dim=128)       # PoseCDE.py:53-57: fused width == cde_hidden_dim
    opt = _reference_opt(**over)
    torch.manual_seed(0)
    ref_model = D.DeepVIO(copy.copy(opt))                              # the reference's own regressor classes (stubbed solver libs)
    ref_sd = ref_model.state_dict()

    # the maintainer's change is the import line of DeepVIO.py
    D.PoseODERNN, D.PoseCDE = odevio_b200.PoseODERNN, odevio_b200.PoseCDE
    model = D.DeepVIO(copy.copy(opt))
    assert type(model.Pose_net).__module__.startswith("odevio_b200")
    res = model.load_state_dict(ref_sd, strict=False)
    assert res.missing_keys == []
    assert all(k.startswith("Pose_net.solver.") for k in res.unexpected_keys), res.unexpected_keys
    if model_type == "ode-rnn":
        assert any("solver._orig_mod" in k for k in res.unexpected_keys)      # the duplicates do exist in reference checkpoints
    ours = model.state_dict()
    assert set(ours) == {k for k in ref_sd if not k.startswith("Pose_net.solver.")}
    for k, v in ours.items():
        assert torch.equal(v, ref_sd[k]), k

    # utils/utils.py:115-130 get_optimizer: [other params, regressor params], model wrapped in DataParallel (.module)
    sys.path.insert(0, REFERENCE)
    try:
        spec = importlib.util.spec_from_file_location("ref_utils_utils", os.path.join(REFERENCE, "utils", "utils.py"))
        utils = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(utils)
        except ImportError as exc:                                       # its plotting / metric imports are out of scope
            utils = None
            reason = str(exc)
    finally:
        sys.path.remove(REFERENCE)
    if utils is not None:
        optim = utils.get_optimizer(types.SimpleNamespace(module=model), opt)
        n_opt = sum(p.numel() for g in optim.param_groups for p in g["params"])
        assert n_opt == sum(p.numel() for p in model.Pose_net.parameters())
        assert len(optim.param_groups) == 2 and optim.param_groups[0]["lr"] == opt.lr_warmup
    else:
        n_reg = sum(p.numel() for p in model.Pose_net.get_regressor_params())
        n_other = sum(p.numel() for p in model.Pose_net.get_other_params())
        assert n_reg + n_other == sum(p.numel() for p in model.Pose_net.parameters()), reason

    # DeepVIO.forward (DeepVIO.py:61-68) reaches the drop-in with the reference's signature; CPU tensors fail loudly
    model.eval()
    B, S = 2, 3
    img = torch.zeros(B, S + 1, 3, opt.img_h, opt.img_w)
    imu = torch.zeros(B, 10 * S + 1, 6)
    ts = torch.arange(S + 1, dtype=torch.float32).repeat(B, 1) * 0.1
    with torch.no_grad(), pytest.raises(_lib.OdevioError, match="no CPU path"):
        model(img, imu, ts, hc=None)


class _DeepVIOForward(nn.Module):
    """DeepVIO.forward (reference src/models/DeepVIO.py:61-68) around given encoders / regressor."""

    def __init__(self, image_net, inertial_net, pose_net):
        super().__init__()
        self.Image_net, self.Inertial_net, self.Pose_net = image_net, inertial_net, pose_net

    def forward(self, img, imu, timestamps, hc=None):
        fv, fi = self.Image_net(img), self.Inertial_net(imu)
        poses, h_T = self.Pose_net(fv, fi, timestamps, prev=hc)
        return poses, h_T


class _StubImageNet(nn.Module):
    """Stand-in for the (out-of-scope) FlowNet image encoder: [B, S + 1, F] 'images' -> fv [B, S, v_f_len]."""

    def __init__(self, v_f_len):
        super().__init__()
        self.visual_head = nn.Linear(32, v_f_len)

    def forward(self, img):
        return self.visual_head(torch.cat((img[:, :-1], img[:, 1:]), dim=2))


@pytest.mark.gpu
def test_deepvio_forward_with_dropins_two_windows_on_gpu(cuda_device):
    """DeepVIO.forward with InertialEncoder + PoseODERNN from odevio_b200, two consecutive windows with the carried state
    (the reference's evaluation loop, src/data/KITTI_eval.py: `hc` of one window seeds the next), vs the oracle."""
    import odevio_b200
    from helpers import POSE_RTOL, rel_err
    from oracle.imu_encoder import OracleInertialEncoder
    from oracle.modules import deepvio_initialization
    from oracle.pose_odernn import OraclePoseODERNN, default_opt
    opt = default_opt(ode_solver="dopri5", ode_rtol=1e-3, ode_precision="fp16x3", imu_dropout=0.0)
    torch.manual_seed(0)
    ref = _DeepVIOForward(_StubImageNet(opt.v_f_len), OracleInertialEncoder(opt), OraclePoseODERNN(opt))
    deepvio_initialization(ref)
    ref.eval()
    ours = _DeepVIOForward(_StubImageNet(opt.v_f_len), odevio_b200.InertialEncoder(copy.copy(opt)), odevio_b200.PoseODERNN(copy.copy(opt)))
    res = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    ours = ours.to(cuda_device).eval()
    B, S = 24, 5
    g = torch.Generator().manual_seed(3)
    hc_ref = hc = None
    t0 = 0.0
    for window in range(2):
        img = torch.randn(B, S + 1, 16, generator=g)
        imu = torch.randn(B, 10 * S + 1, 6, generator=g)
        ts = t0 + torch.cumsum(0.1 + 0.1 * torch.rand(B, S + 1, generator=g), 1)
        t0 = float(ts.max())
        with torch.no_grad():
            p_ref, hc_ref = ref(img, imu, ts, hc=hc_ref)
            p, hc = ours(img.to(cuda_device), imu.to(cuda_device), ts.to(cuda_device), hc=hc)
        ours.Pose_net.check_status()
        assert rel_err(p.cpu(), p_ref) <= 4 * POSE_RTOL, (window, rel_err(p.cpu(), p_ref))
        assert rel_err(hc.cpu(), hc_ref) <= 2e-4
