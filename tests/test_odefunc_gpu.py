"""ODEFunc.forward on the tensor cores (odevio_odefunc_forward: tcgen05 kind::tf32, 3xTF32 split,
cluster of 8 CTAs per 128-row tile) vs an fp64 torch evaluation of the same MLP.
Tolerance: max-norm relative error <= 1e-5 (fp32-level; a single-pass TF32 GEMM would be ~1e-3)."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(dev, D, H, n, act, seed=0):
    import odevio_b200
    torch.manual_seed(seed)
    f = odevio_b200.ODEFunc(D, H, n, act)
    for m in f.net:
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.kaiming_normal_(m.weight.data)
            m.bias.data.normal_(0, 0.05)
    return f, f.to(dev)


@pytest.mark.parametrize("D,H,n,act,M", [
    (768, 512, 3, "tanh", 2048), (768, 512, 3, "softplus", 300), (768, 1024, 2, "relu", 129),
    (768, 256, 1, "leaky_relu", 128), (256, 256, 4, "tanh", 17), (1024, 768, 2, "tanh", 4000)])
def test_odefunc_tensor_core_matches_fp64(cuda_device, D, H, n, act, M):
    import copy
    f_cpu, _ = _pair("cpu", D, H, n, act)
    f64 = copy.deepcopy(f_cpu).double()
    f_gpu = copy.deepcopy(f_cpu).to(cuda_device)
    x = torch.randn(M, D, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = f64.net(x.double())
        out = f_gpu(None, x.to(cuda_device))
    torch.cuda.synchronize()
    err = ((out.cpu().double() - ref).abs().max() / ref.abs().max()).item()
    assert err <= 1e-5, err


def test_vector_fields_match_reference_class_outputs(cuda_device):
    """ODEFunc / CDEFunc called directly (odevio_mlp_forward for these small shapes) reproduce the outputs
    recorded from the reference's own classes (tests/golden/vector_fields.pt, src/models/ODEFunc.py)."""
    import os
    import odevio_b200
    vf = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vector_fields.pt"))
    for act in ("tanh", "relu", "leaky_relu", "softplus"):
        fx = vf[act]
        f = odevio_b200.ODEFunc(32, 16, 3, act)
        f.load_state_dict(fx["ode_state"])
        g = odevio_b200.CDEFunc(9, 8, 2, act)
        g.load_state_dict(fx["cde_state"])
        with torch.no_grad():
            y = f.to(cuda_device)(None, fx["x"].to(cuda_device)).cpu()
            out = g.to(cuda_device)(None, fx["z"].to(cuda_device)).cpu()
        assert out.shape == (5, 8, 9)
        assert torch.allclose(y, fx["y"], rtol=2e-6, atol=2e-7)
        assert torch.allclose(out, fx["g"], rtol=2e-6, atol=2e-7)


def test_vector_fields_fail_loudly_on_cpu():
    import odevio_b200
    with pytest.raises(odevio_b200.OdevioError):          # CPU tensor: no CPU path
        odevio_b200.ODEFunc(32, 16, 2, "tanh")(None, torch.zeros(4, 32))
    with pytest.raises(odevio_b200.OdevioError):
        odevio_b200.CDEFunc(9, 8, 2, "tanh")(None, torch.zeros(4, 8))
