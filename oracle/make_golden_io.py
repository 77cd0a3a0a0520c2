"""Generate tests/golden/io/* from the REAL reference's dataset code (/root/reference/datasets),
run in the build container.  TEST INFRASTRUCTURE.

    python oracle/make_golden_io.py

Pins for casmvsnet_pl_b200/io.py: PFM files written by the reference's own
save_pfm, a cam.txt / pair.txt in the MVSNet format with the values the reference's
read_cam_file / build_metas / build_proj_mats / __getitem__ derive from them, and the
ToTensor + Normalize output for a small uint8 image.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CASMVS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "io")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cam_text(K, E, depth_min, interval):
    rows = ["extrinsic"] + [" ".join(f"{v:.6f}" for v in r) for r in E] + ["", "intrinsic"]
    rows += [" ".join(f"{v:.6f}" for v in r) for r in K] + ["", f"{depth_min} {interval}"]
    return "\n".join(rows) + "\n"


def main():
    os.makedirs(OUT, exist_ok=True)
    utils = _load(os.path.join(REF, "datasets", "utils.py"), "ref_datasets_utils")
    rng = np.random.default_rng(0)
    gray = rng.standard_normal((5, 7)).astype(np.float32) * 100
    color = rng.standard_normal((4, 6, 3)).astype(np.float32)
    utils.save_pfm(os.path.join(OUT, "gray_7x5.pfm"), gray)
    utils.save_pfm(os.path.join(OUT, "color_6x4.pfm"), color, scale=2)
    # big-endian file (the reference reader handles both)
    with open(os.path.join(OUT, "gray_be_7x5.pfm"), "wb") as f:
        f.write(b"Pf\n7 5\n1.000000\n")
        np.flipud(gray).astype(">f4").tofile(f)

    # cameras: 49 views on an arc (DTU has 49), DTU-like intrinsics at quarter resolution; the
    # reference's OWN build_metas / build_proj_mats / read_cam_file run on a temp tree
    import tempfile
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "root", "Cameras", "train"))
    os.makedirs(os.path.join(tmp, "cwd", "datasets", "lists", "dtu"))
    for vid in range(49):
        t = np.deg2rad(2.0 * vid)
        R = np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]])
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = [-15.0 * vid, 1.5 * vid, 2.0]
        K = np.array([[361.54125, 0.0, 82.900625], [0.0, 360.3975, 66.383875], [0.0, 0.0, 1.0]])
        txt = cam_text(K, E, 425.0 + vid, 2.5)
        for d in ("Cameras", os.path.join("Cameras", "train")):
            open(os.path.join(tmp, "root", d, f"{vid:08d}_cam.txt"), "w").write(txt)
        if vid < 4:
            open(os.path.join(OUT, f"{vid:08d}_cam.txt"), "w").write(txt)
    pair = ("4\n0\n3 1 2037.2 2 1500.1 3 900.5\n1\n3 0 2037.2 2 1800.0 3 700.0\n"
            "2\n2 1 1800.0 3 1700.25\n3\n1 2 1700.25\n")
    open(os.path.join(OUT, "pair.txt"), "w").write(pair)
    open(os.path.join(tmp, "root", "Cameras", "pair.txt"), "w").write(pair)
    open(os.path.join(tmp, "cwd", "datasets", "lists", "dtu", "val.txt"), "w").write("scan1\n")

    sys.path.insert(0, REF)
    from datasets.dtu import DTUDataset          # noqa: E402
    out = {}
    cwd = os.getcwd()
    os.chdir(os.path.join(tmp, "cwd"))
    for mode, img_wh in (("train", None), ("test", (1152, 864))):
        ds = object.__new__(DTUDataset)
        ds.root_dir, ds.split, ds.img_wh, ds.levels = os.path.join(tmp, "root"), "val", img_wh, 3
        ds.build_metas()                          # datasets/dtu.py:31-50
        ds.build_proj_mats()                      # datasets/dtu.py:52-75
        out[f"metas_{mode}"] = np.array([[m[1], m[2]] + m[3] + [-1] * (3 - len(m[3]))
                                         for m in ds.metas])
        mats = [ds.proj_mats[v][0] for v in range(4)]
        dmins = [ds.proj_mats[v][1] for v in range(4)]
        out[f"proj_{mode}"] = torch.stack(mats).numpy()
        # datasets/dtu.py:176-186: src_proj @ inv(ref_proj), rows 0..2
        ref_inv = torch.inverse(mats[0])
        out[f"rel_{mode}"] = torch.stack([mats[v] @ ref_inv for v in (1, 2)])[:, :, :3].numpy()
    os.chdir(cwd)
    K0, E0, _ = ds.read_cam_file(os.path.join(OUT, "00000000_cam.txt"))
    out["intrinsics0"], out["extrinsics0"] = K0, E0
    out["depth_min"] = np.array(dmins)
    # ToTensor + Normalize (datasets/dtu.py:130-137) on a small uint8 image batch
    from torchvision import transforms as T
    img = rng.integers(0, 256, size=(2, 6, 8, 3), dtype=np.uint8)
    tr = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    from PIL import Image
    out["img_u8"] = img
    out["img_norm"] = torch.stack([tr(Image.fromarray(i)) for i in img]).numpy()
    out["pfm_gray"], out["pfm_color"] = gray, color
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **out)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
