"""Oracle restatement of the small torch.nn pieces on the path (test infrastructure).

Parameter names/shapes match the reference so state_dicts are interchangeable:
  * vector field MLP  -- reference src/models/ODEFunc.py:5-39 (ODEFunc), :44-84 (CDEFunc)
  * feature fusion    -- reference src/models/FusionModule.py:8-29
  * pose head         -- reference src/models/PoseODERNN.py:64-68, src/models/PoseCDE.py:67-71
  * init rule         -- reference src/models/DeepVIO.py:77-87 (kaiming-normal Linear, zero bias)
``oracle/make_golden.py`` checks these against the reference's own classes.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

ACTIVATIONS = ("tanh", "relu", "leaky_relu", "softplus")


def make_activation(name: str) -> nn.Module:
    table = {"tanh": nn.Tanh, "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "softplus": nn.Softplus}
    if name not in table:
        raise ValueError(f"Activation function {name} not supported")
    return table[name]()


def _mlp(sizes, activation):
    """Linear/act pairs; the last Linear is followed by Tanh (ODEFunc.py:13-14, :56-57)."""
    mods = []
    for idx, (fan_in, fan_out) in enumerate(zip(sizes[:-1], sizes[1:])):
        mods.append(nn.Linear(fan_in, fan_out))
        mods.append(nn.Tanh() if idx == len(sizes) - 2 else make_activation(activation))
    net = nn.Sequential(*mods)
    for m in net:
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, mean=0.0, std=0.1)
            nn.init.zeros_(m.bias)
    return net


class OracleODEFunc(nn.Module):
    """f(t, x) = tanh(W_n a(... a(W_0 x + b_0) ...) + b_n); autonomous."""

    def __init__(self, feature_dim, hidden_dim, num_hidden_layers=3, activation="tanh"):
        super().__init__()
        self.net = _mlp([feature_dim] + [hidden_dim] * num_hidden_layers + [feature_dim], activation)

    def forward(self, t, x):
        return self.net(x)


class OracleCDEFunc(nn.Module):
    """g(t, z) = tanh(MLP(z)) viewed as [B, hidden, channels]."""

    def __init__(self, feature_dim, hidden_dim, num_hidden_layers=3, activation="tanh"):
        super().__init__()
        self.hidden_dim, self.feature_dim = hidden_dim, feature_dim
        self.net = _mlp([hidden_dim] * (num_hidden_layers + 1) + [hidden_dim * feature_dim], activation)

    def forward(self, t, z):
        return self.net(z).view(z.shape[0], self.hidden_dim, self.feature_dim)


class OracleFusion(nn.Module):
    def __init__(self, feature_dim, fuse_method):
        super().__init__()
        self.fuse_method, self.f_len = fuse_method, feature_dim
        if fuse_method == "soft":
            self.net = nn.Sequential(nn.Linear(feature_dim, feature_dim))
        elif fuse_method == "hard":
            self.net = nn.Sequential(nn.Linear(feature_dim, 2 * feature_dim))

    def forward(self, v, i):
        cat = torch.cat((v, i), -1)
        if self.fuse_method == "cat":
            return cat
        if self.fuse_method == "soft":
            return cat * self.net(cat)          # no sigmoid (FusionModule.py:20-23)
        if self.fuse_method == "hard":
            logits = self.net(cat).view(v.shape[0], v.shape[1], self.f_len, 2)
            return cat * F.gumbel_softmax(logits, tau=1, hard=True, dim=-1)[..., 0]
        raise ValueError(f"fuse method {self.fuse_method} not supported")


def make_regressor(in_dim):
    return nn.Sequential(nn.Linear(in_dim, 128), nn.LeakyReLU(0.1, inplace=True), nn.Linear(128, 6))


def deepvio_initialization(net: nn.Module):
    """The part of DeepVIO.initialization that touches the regressors
    (DeepVIO.py:77-87): Linear -> kaiming_normal_, zero bias.  nn.RNN / nn.GRU
    are matched by no branch there and keep PyTorch's default init."""
    for m in net.modules():
        if isinstance(m, nn.Linear):
            nn.init.kaiming_normal_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()
