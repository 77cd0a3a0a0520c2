"""Synthetic DTU-shaped inputs — no dataset or checkpoint is
available offline, so tests and bench.py use this generator.

Camera model: at level 2 (quarter resolution of 640x512) K2 = [[361.54,0,82.9],
[0,360.4,66.4],[0,0,1]], doubled per finer level (reference datasets/dtu.py:68-72)
and scaled with the image size; reference pose [I|0]; source view i is rotated
about y by (-1)^i * 4deg * ceil(i/2) with its optical axis converging at 680 mm;
proj = (K[R|t])_4x4 @ inv((K[I|0])_4x4), rows 0..2 (reference datasets/dtu.py:181-186).
"""
import math

import numpy as np
import torch

DEPTH_MIN = 425.0       # DTU range 425..935 mm (reference README.md:99)
DEPTH_INTERVAL = 2.65   # reference opt.py:16
CONVERGE_MM = 680.0


def _rot_y(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rot_x(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_z(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def intrinsics(level, W=640, H=512):
    K = np.array([[361.54, 0, 82.9], [0, 360.4, 66.4], [0, 0, 1]], dtype=np.float64)
    K[0] *= W / 640.0
    K[1] *= H / 512.0
    K[:2] *= 2 ** (2 - level)
    return K


def projection_matrices(n_views, W=640, H=512, levels=3, stress=False, behind_view=None):
    """-> float32 tensor (V-1, levels, 3, 4), level 0 = finest.

    stress=True adds 2deg roll + 3deg pitch on odd source views (epipolar lines
    no longer axis-aligned); behind_view=i pushes source view i far enough along
    -z that part of the sweep hits the q_z <= 1e-7 branch (modules.py:76-79)."""
    out = np.zeros((n_views - 1, levels, 3, 4), dtype=np.float32)
    for i in range(1, n_views):
        theta = math.radians((-1) ** i * 4.0 * math.ceil(i / 2))
        R = _rot_y(theta)
        t = np.array([-CONVERGE_MM * math.sin(theta), 0.0,
                      CONVERGE_MM * (1 - math.cos(theta))])
        if stress and i % 2 == 1:
            R = _rot_z(math.radians(2.0)) @ _rot_x(math.radians(3.0)) @ R
        if behind_view is not None and i == behind_view:
            t = t + np.array([0.0, 0.0, -600.0])
        for l in range(levels):
            K = intrinsics(l, W, H)
            src = np.eye(4)
            src[:3, :4] = K @ np.concatenate([R, t[:, None]], 1)
            ref = np.eye(4)
            ref[:3, :3] = K
            out[i - 1, l] = (src @ np.linalg.inv(ref))[:3, :4].astype(np.float32)
    return torch.from_numpy(out)


def make_inputs(B=1, V=3, W=640, H=512, seed=0, stress=False, behind_view=None,
                device="cpu"):
    """imgs (B,V,3,H,W) ~ N(0,1) (ImageNet-normalised range, datasets/dtu.py:134-137),
    proj_mats (B,V-1,3,3,4), init_depth_min, depth_interval (python floats)."""
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randn(B, V, 3, H, W, generator=g)
    pm = projection_matrices(V, W, H, 3, stress, behind_view)
    pm = pm.unsqueeze(0).expand(B, -1, -1, -1, -1).contiguous()
    return imgs.to(device), pm.to(device), DEPTH_MIN, DEPTH_INTERVAL


def make_level_feats(B, V, level, W=640, H=512, seed=0, device="cpu", smooth=False):
    """Random feature pyramid level (B,V,C_l,h_l,w_l) ~ N(0,1) for kernel-only
    tests/benches.  smooth=True low-pass filters it (band-limited features make
    tolerance tests insensitive to sub-ulp sampling-position differences)."""
    C = 8 * 2 ** level
    h, w = H >> level, W >> level
    g = torch.Generator().manual_seed(seed + 17 * level)
    f = torch.randn(B, V, C, h, w, generator=g)
    if smooth:
        k = torch.tensor([1., 4., 6., 4., 1.])
        k = (k[:, None] * k[None, :]) / 256.0
        ff = f.reshape(B * V * C, 1, h, w)
        ff = torch.nn.functional.conv2d(ff, k.reshape(1, 1, 5, 5), padding=2)
        f = (ff / ff.std()).reshape(B, V, C, h, w)
    return f.to(device)


def randomize_model_(model, seed=0, prob_scale=50.0):
    """Make a randomly initialised CascadeMVSNet a meaningful parity target:
    BN statistics/affine randomised (gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1),
    var~U(.5,1.5)) so the fused epilogue is exercised, and prob.weight scaled so
    the softmax is peaked (SURVEY.md §7 'Precision': flat softmax hides errors)."""
    g = torch.Generator().manual_seed(seed + 1000)
    with torch.no_grad():
        for name, m in model.named_modules():
            if hasattr(m, "running_mean") and hasattr(m, "running_var"):
                n = m.running_mean.numel()
                m.weight.copy_(torch.rand(n, generator=g) + 0.5)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
        for l in range(3):
            getattr(model, f"cost_reg_{l}").prob.weight.mul_(prob_scale)
    return model
