"""Pin the oracle's module restatement on the reference's own classes and freeze golden vectors.

Run HERE (the authoring container), where /root/reference is mounted:

    python -m oracle.make_golden

1. Imports the reference's importable classes -- ``src.models.ODEFunc.{ODEFunc, CDEFunc}`` and
   ``src.models.FusionModule.FusionModule`` (the regressors themselves need torchode / torchcde,
   which are not installable offline) -- loads one state_dict into reference and oracle modules
   and checks that their outputs are bit-identical on CPU.
2. Writes small fixtures under tests/golden/:
     vector_fields.pt   reference-class outputs (the oracle, and on the GPU box the kernels, must
                        reproduce them)
     odernn_<case>.pt   oracle PoseODERNN forward on seeded weights / inputs / timestamps:
                        pose, h, n_steps, n_accepted
     cde_<case>.pt      oracle PoseCDE forward: pose, z0, n_steps, n_accepted, n_f_evals
/root/reference does not exist on the GPU box; nothing at test time reads it.
"""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

ODERNN_CASES = {
    # name: (opt overrides, B, S, irregular, bias_std)
    "rk4_regular": (dict(ode_solver="rk4"), 6, 4, False, 0.0),
    "rk4_38_sub2": (dict(ode_solver="rk4_38", ode_substeps=2), 6, 4, True, 0.05),
    "dopri5_ref_defaults": (dict(ode_solver="dopri5"), 6, 4, False, 0.0),
    "dopri5_irregular_rtol1e-3": (dict(ode_solver="dopri5", ode_rtol=1e-3), 6, 4, True, 0.05),
    "tsit5_gru": (dict(ode_solver="tsit5", ode_rnn_type="gru"), 5, 3, True, 0.05),
    "heun_softplus_L3": (dict(ode_solver="heun", ode_activation_fn="softplus", rnn_num_layers=3), 5, 3, True, 0.05),
}
SMALL = dict(v_f_len=24, i_f_len=8, ode_hidden_dim=16)


def _same(a, b):
    return a.shape == b.shape and torch.equal(a, b)


def check_against_reference():
    sys.path.insert(0, REF)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):           # the reference prints its activation
        from src.models.ODEFunc import ODEFunc as RefODEFunc, CDEFunc as RefCDEFunc
        from src.models.FusionModule import FusionModule as RefFusion
    from oracle.modules import OracleCDEFunc, OracleFusion, OracleODEFunc, ACTIVATIONS
    g = torch.Generator().manual_seed(0)
    fixtures = {}
    for act in ACTIVATIONS:
        with contextlib.redirect_stdout(io.StringIO()):
            ref = RefODEFunc(32, 16, 3, act)
            cref = RefCDEFunc(9, 8, 2, act)
        for p in list(ref.parameters()) + list(cref.parameters()):
            p.data.normal_(0, 0.3, generator=g)
        ora = OracleODEFunc(32, 16, 3, act)
        ora.load_state_dict(ref.state_dict())
        x = torch.randn(7, 32, generator=g)
        y_ref = ref(torch.zeros(()), x)
        assert _same(y_ref, ora(None, x)), f"ODEFunc restatement differs from the reference ({act})"
        cora = OracleCDEFunc(9, 8, 2, act)
        cora.load_state_dict(cref.state_dict())
        z = torch.randn(5, 8, generator=g)
        g_ref = cref(None, z)
        assert _same(g_ref, cora(None, z)), f"CDEFunc restatement differs from the reference ({act})"
        fixtures[act] = dict(ode_state={k: v.clone() for k, v in ref.state_dict().items()}, x=x, y=y_ref.detach(),
                             cde_state={k: v.clone() for k, v in cref.state_dict().items()}, z=z, g=g_ref.detach())
    for method in ("cat", "soft"):
        rf = RefFusion(32, method)
        of = OracleFusion(32, method)
        of.load_state_dict(rf.state_dict())
        v, i = torch.randn(3, 4, 20, generator=g), torch.randn(3, 4, 12, generator=g)
        out = rf(v, i)
        assert _same(out, of(v, i)), f"FusionModule restatement differs from the reference ({method})"
        fixtures["fuse_" + method] = dict(state={k: t.clone() for k, t in rf.state_dict().items()}, v=v, i=i,
                                          out=out.detach())
    torch.save(fixtures, os.path.join(OUT, "vector_fields.pt"))
    print("oracle modules == reference classes (bit-identical); wrote vector_fields.pt")


def make_odernn_goldens():
    from oracle.modules import deepvio_initialization
    from oracle.pose_odernn import OraclePoseODERNN, default_opt
    from odevio_b200 import synth
    for name, (over, B, S, irregular, bias_std) in ODERNN_CASES.items():
        opt_kw = dict(SMALL, **over)
        torch.manual_seed(1234)
        m = OraclePoseODERNN(default_opt(**opt_kw))
        deepvio_initialization(m)
        if bias_std:
            g = torch.Generator().manual_seed(99)
            for n, p in m.named_parameters():
                if n.endswith("bias") and not n.startswith("rnn"):
                    p.data.normal_(0, bias_std, generator=g)
        m.eval()
        fv, fi = synth.features(B, S, SMALL["v_f_len"], SMALL["i_f_len"], seed=3)
        ts = synth.timestamps(B, S, irregular=irregular, seed=3)
        with torch.no_grad():
            pose, h = m(fv, fi, ts)
            n_steps = m.last_stats["n_steps"].to(torch.int32)
            n_acc = m.last_stats["n_accepted"].to(torch.int32)
            prev = 0.3 * torch.randn(m.rnn_num_layers, B, m.f_len, generator=torch.Generator().manual_seed(5))
            pose_c, h_c = m(fv, fi, ts + 40.0, prev=prev)
        torch.save(dict(opt=opt_kw, state={k: v.clone() for k, v in m.state_dict().items()}, fv=fv, fi=fi, ts=ts,
                        pose=pose, h=h, n_steps=n_steps, n_accepted=n_acc,
                        prev=prev, ts_abs=ts + 40.0, pose_carry=pose_c, h_carry=h_c),
                   os.path.join(OUT, f"odernn_{name}.pt"))
        print("wrote", f"odernn_{name}.pt")


CDE_CASES = {
    # name: (opt overrides, B, S, irregular, feature scale)
    "linear_dopri5_ref": (dict(cde_solver="dopri5", cde_interp="linear"), 6, 6, False, 0.2),
    "linear_dopri5_knots": (dict(cde_solver="dopri5", cde_interp="linear"), 5, 6, True, 0.2),
    "cubic_dopri5": (dict(cde_solver="dopri5", cde_interp="cubic", cde_rtol=1e-3), 6, 5, True, 0.2),
    "cubic_rk4_step": (dict(cde_solver="rk4", cde_interp="cubic", cde_step_size=0.25), 6, 5, True, 0.2),
    "linear_rk4": (dict(cde_solver="rk4", cde_interp="linear"), 6, 6, True, 0.2),
}
CDE_SMALL = dict(v_f_len=16, i_f_len=16, cde_hidden_dim=32, cde_fn_num_layers=2)


def make_cde_goldens():
    from oracle.modules import deepvio_initialization
    from oracle.pose_cde import OraclePoseCDE
    from oracle.pose_odernn import default_opt
    from odevio_b200 import synth
    for name, (over, B, S, irregular, scale) in CDE_CASES.items():
        opt_kw = dict(CDE_SMALL, **over)
        torch.manual_seed(4321)
        m = OraclePoseCDE(default_opt(**opt_kw))
        deepvio_initialization(m)
        g = torch.Generator().manual_seed(98)
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.data.normal_(0, 0.05, generator=g)
        m.train()                                             # relative timestamps (PoseCDE.py:81)
        fv, fi = synth.features(B, S, 16, 16, seed=4)
        fv, fi = fv * scale, fi * scale
        ts = synth.timestamps(B, S, irregular=irregular, seed=4)
        if name == "linear_dopri5_knots":
            ts = ts * 4.0                                     # row 0 crosses several integer knots
        with torch.no_grad():
            pose, z0 = m(fv, fi, ts)
        st = m.last_stats
        torch.save(dict(opt=opt_kw, state={k: v.clone() for k, v in m.state_dict().items()}, fv=fv, fi=fi, ts=ts,
                        pose=pose, z0=z0, n_steps=st["n_steps"], n_accepted=st["n_accepted"], n_f_evals=st["n_f_evals"]),
                   os.path.join(OUT, f"cde_{name}.pt"))
        print("wrote", f"cde_{name}.pt", st["n_steps"], st["n_accepted"], st["n_f_evals"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if not os.path.isdir(REF):
        raise SystemExit("make_golden needs the reference tree at /root/reference")
    sys.path.insert(0, ROOT)
    check_against_reference()
    make_odernn_goldens()
    make_cde_goldens()
