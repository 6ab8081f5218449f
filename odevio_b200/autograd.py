"""Autograd bridge for the fused regressors (discretise-then-optimise backward).

Placeholder until the backward kernels land: training through the fused path raises rather
than silently falling back to an eager implementation.
"""


def odernn_apply(module, fv, fi, ts, prev):
    raise NotImplementedError(
        "odevio_b200: the fused backward is not built yet; call the regressor under "
        "torch.no_grad() / module.requires_grad_(False) for inference")
