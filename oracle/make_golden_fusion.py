"""Generate tests/golden/fusion_64x48.npz from the REAL reference's fusion functions
(/root/reference/eval.py:113-182), extracted from the source with their numba decorators and run
in the build container.  TEST INFRASTRUCTURE.

    python oracle/make_golden_fusion.py
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CASMVS_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)


def reference_functions():
    """eval.py cannot be imported (plyfile / inplace_abn are not installed): compile only its
    three geometry functions, unmodified, in a namespace with the modules they use."""
    import cv2
    from numba import jit
    src = open(os.path.join(REF, "eval.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and
            n.name in ("xy_ref2src", "xy_src2ref", "check_geo_consistency")]
    ns = {"np": np, "cv2": cv2, "jit": jit}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "reference_eval.py", "exec"), ns)
    return ns


def scene(W=64, H=48, S=3, seed=0):
    """A smooth synthetic surface seen by S+1 cameras; source depth maps are rendered from the
    same surface (+ noise, + an occluder in one view) so that masks have both outcomes."""
    import math
    rng = np.random.default_rng(seed)
    K = np.array([[80.0, 0, W / 2], [0, 80.0, H / 2], [0, 0, 1]], np.float64)

    def cam(theta, tx):
        R = np.array([[math.cos(theta), 0, math.sin(theta)], [0, 1, 0],
                      [-math.sin(theta), 0, math.cos(theta)]])
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = [tx, 0.0, 600.0 * (1 - math.cos(theta))]
        P = np.eye(4)
        P[:3, :4] = K @ E[:3, :4]
        return P.astype(np.float32)
    Ps = [cam(0.0, 0.0)] + [cam(math.radians(a), -600.0 * math.sin(math.radians(a)))
                            for a in (3.0, -3.0, 5.0)[:S]]
    ys, xs = np.mgrid[:H, :W].astype(np.float64)

    def depth_for(P):
        # plane z_world = 600 + 0.2*x_world + bumps, solved per pixel by 3 fixed-point steps
        Pinv = np.linalg.inv(P.astype(np.float64))
        d = np.full((H, W), 600.0)
        for _ in range(8):
            pts = Pinv @ np.stack([xs * d, ys * d, d, np.ones_like(d)]).reshape(4, -1)
            zw = 600.0 + 0.2 * pts[0] + 10 * np.sin(pts[0] / 40.0) * np.cos(pts[1] / 30.0)
            cam_pts = P.astype(np.float64) @ np.stack([pts[0], pts[1], zw, np.ones_like(zw)])
            d = cam_pts[2].reshape(H, W)
        return d
    depths = [depth_for(P).astype(np.float32) for P in Ps]
    depths = [d + rng.normal(0, 0.3, d.shape).astype(np.float32) for d in depths]
    depths[2][10:25, 20:40] -= 60.0                               # occluder in one source view
    images = [rng.uniform(0, 255, (H, W, 3)).astype(np.float32) for _ in Ps]
    proba = rng.uniform(0.99, 1.0, (H // 4, W // 4)).astype(np.float32)
    return Ps, depths, images, proba


def main():
    ns = reference_functions()
    Ps, depths, images, proba = scene()
    H, W = depths[0].shape
    out = {"P": np.stack(Ps), "depths": np.stack(depths), "images": np.stack(images), "proba": proba}
    rs, ms, is_ = [], [], []
    for s in range(1, len(Ps)):
        r, m, i2 = ns["check_geo_consistency"](depths[0], Ps[0], depths[s], Ps[s], images[0],
                                               images[s], (W, H))
        rs.append(r); ms.append(m); is_.append(i2)
    out.update(reproj=np.stack(rs), mask=np.stack(ms), img2ref=np.stack(is_))
    import cv2
    out["proba_up"] = cv2.resize(proba, None, fx=4, fy=4, interpolation=cv2.INTER_LINEAR)
    p = os.path.join(ROOT, "tests", "golden", "fusion_64x48.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p) // 1024, "KiB; mask fractions", [float(m.mean()) for m in ms])


if __name__ == "__main__":
    main()
