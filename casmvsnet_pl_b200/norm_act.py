"""Activated-batch-norm parameter holders.

The reference takes its norm-act class from the un-vendored, unpinned
``inplace_abn`` package (README.md:28; models/modules.py:5, eval.py:13,201,
train.py:10,41).  The engine only needs the *parameters* (weight, bias,
running_mean, running_var, eps, activation_param) — the 3D stack folds them
into the conv epilogue — and a plain forward for the 2D FeatureNet, which
stays PyTorch/cuDNN.  If the real ``inplace_abn`` is installed its classes can
be passed as ``norm_act`` instead: the same attributes are read.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ABN(nn.Module):
    """BatchNorm (eps 1e-5, momentum 0.1) + LeakyReLU(0.01), inplace_abn defaults."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, activation="leaky_relu",
                 activation_param=0.01):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.activation = activation
        self.activation_param = activation_param
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def forward(self, x):
        x = F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                         self.training, self.momentum, self.eps)
        return F.leaky_relu(x, self.activation_param)


class InPlaceABN(ABN):
    """Same math as ABN; the in-place memory trick (README.md:108-113) is a
    training-time optimisation that is out of scope for the inference engine."""


def folded_scale_shift(bn):
    """eval-mode ABN as y = x*alpha + beta' (what ATen's batch_norm computes)."""
    invstd = 1.0 / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = bn.weight.detach().float() if getattr(bn, "weight", None) is not None \
        else torch.ones_like(invstd)
    b = bn.bias.detach().float() if getattr(bn, "bias", None) is not None \
        else torch.zeros_like(invstd)
    alpha = invstd * w
    beta = b - bn.running_mean.detach().float() * alpha
    return alpha.contiguous(), beta.contiguous()


def activation_slope(bn):
    act = getattr(bn, "activation", "leaky_relu")
    if act == "leaky_relu":
        return float(getattr(bn, "activation_param", 0.01))
    if act == "relu":
        return 0.0
    if act in ("identity", "none"):
        return 1.0
    raise ValueError(f"unsupported norm_act activation {act!r}")
