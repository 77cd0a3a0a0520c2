"""Generate tests/golden/train_step_*.npz from the REAL reference (/root/reference): one
training step = CascadeMVSNet.forward in train mode (batch-statistics ABN) + the reference's
own SL1Loss (losses.py:10-17) + backward, on CPU.  TEST INFRASTRUCTURE.

    python oracle/make_golden_grad.py

Stored: inputs that are not regenerable from a seed (targets, masks), the three depth maps, the
loss and the gradients of a spread of parameters (first/last 2D convs, every kind of 3D layer,
ABN affine parameters, prob bias).  Weights come from make_golden.seeded_state_dict (checksum
guards RNG drift).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casmvsnet_pl_b200 import synth                                        # noqa: E402
from oracle import ref_loader                                               # noqa: E402
from oracle.make_golden import sd_checksum, seeded_state_dict              # noqa: E402

GRAD_KEYS = ["feature.conv0.0.conv.weight", "feature.conv2.2.bn.weight", "feature.smooth0.weight",
             "feature.lat1.bias", "cost_reg_2.conv0.conv.weight", "cost_reg_2.conv1.conv.weight",
             "cost_reg_2.conv6.bn.bias", "cost_reg_2.conv7.0.weight", "cost_reg_1.conv4.conv.weight",
             "cost_reg_1.conv9.0.weight", "cost_reg_0.conv0.bn.weight", "cost_reg_0.conv11.0.weight",
             "cost_reg_0.prob.weight", "cost_reg_0.prob.bias"]


def training_state_dict(G):
    """seeded_state_dict with the x50 scaling of prob.weight undone: that scaling makes the softmax
    nearly one-hot (a sharp parity target for INFERENCE), which makes the GRADIENTS exponentially
    sensitive to the forward rounding (a 2.7e-3 forward difference moved them by 20-80 %)."""
    sd = seeded_state_dict((8, 32, 48), (1, 2, 4), G, seed=0)
    for l in range(3):
        sd[f"cost_reg_{l}.prob.weight"] = sd[f"cost_reg_{l}.prob.weight"] / 50.0
    return sd


def case(tag, G, W, H, V, seed):
    mvsnet, modules, abn = ref_loader.load_reference_models()
    spec = importlib.util.spec_from_file_location("ref_losses", os.path.join(ref_loader.REFERENCE_ROOT, "losses.py"))
    losses = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(losses)
    sd = training_state_dict(G)
    model = mvsnet.CascadeMVSNet(num_groups=G, norm_act=abn)
    model.load_state_dict(sd)
    model.train()
    imgs, pm, dmin, dint = synth.make_inputs(B=1, V=V, W=W, H=H, seed=seed)
    g = torch.Generator().manual_seed(100 + seed)
    targets, masks = {}, {}
    for l in range(3):
        h, w = H >> l, W >> l
        targets[f"level_{l}"] = 425.0 + 2.65 * 190 * torch.rand(1, h, w, generator=g)
        masks[f"level_{l}"] = torch.rand(1, h, w, generator=g) > 0.2
    res = model(imgs, pm, dmin, dint)
    loss = losses.SL1Loss()(res, targets, masks)
    loss.backward()
    params = dict(model.named_parameters())
    out = {"sd_checksum": sd_checksum(sd), "loss": loss.item(), "G": G, "W": W, "H": H, "V": V,
           "seed": seed}
    for l in range(3):
        out[f"depth_{l}"] = res[f"depth_{l}"].detach()
        out[f"target_{l}"] = targets[f"level_{l}"]
        out[f"mask_{l}"] = masks[f"level_{l}"]
    for k in GRAD_KEYS:
        out["grad/" + k] = params[k].grad
    # the same step in float64: what exact arithmetic gives, to judge which fp32 digits mean anything
    model64 = mvsnet.CascadeMVSNet(num_groups=G, norm_act=abn)
    model64.load_state_dict(sd)
    model64 = model64.double().train()
    # the reference builds its pixel grid in float32 (kornia create_meshgrid): for the float64
    # run only, hand it the same grid in float64 (integers: exact in both)
    orig_grid = modules.create_meshgrid
    modules.create_meshgrid = lambda *a, **k: orig_grid(*a, **k).double()
    try:
        res64 = model64(imgs.double(), pm.double(), dmin, dint)
    finally:
        modules.create_meshgrid = orig_grid
    loss64 = losses.SL1Loss()(res64, {k: v.double() for k, v in targets.items()}, masks)
    loss64.backward()
    p64 = dict(model64.named_parameters())
    for k in GRAD_KEYS:
        out["grad64/" + k] = p64[k].grad.float()
    path = os.path.join(ROOT, "tests", "golden", f"train_step_{tag}.npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in out.items()})
    print(tag, "loss", loss.item(), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    case("var_96x64", 1, 96, 64, 3, 0)
    case("gwc8_96x64", 8, 96, 64, 3, 1)
