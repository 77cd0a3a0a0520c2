"""Input pipeline / on-disk formats against fixtures the REAL reference's
dataset code produced (oracle/make_golden_io.py): PFM both ways, cam / pair parsing, per-level
projection assembly, relative projections, image normalisation."""
import filecmp
import os

import numpy as np
import pytest
import torch

from casmvsnet_pl_b200 import io as cio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "io")


@pytest.fixture(scope="module")
def exp():
    with np.load(os.path.join(G, "expected.npz")) as z:
        return {k: z[k] for k in z.files}


def test_read_pfm_matches_reference_files(exp):
    d, s = cio.read_pfm(os.path.join(G, "gray_7x5.pfm"))
    assert d.dtype == np.float32 and s == 1.0 and np.array_equal(d, exp["pfm_gray"])
    d, s = cio.read_pfm(os.path.join(G, "color_6x4.pfm"))
    assert s == 2.0 and np.array_equal(d, exp["pfm_color"])
    d, s = cio.read_pfm(os.path.join(G, "gray_be_7x5.pfm"))          # big-endian payload
    assert s == 1.0 and np.array_equal(d.astype(np.float32), exp["pfm_gray"])


def test_save_pfm_is_byte_identical_to_reference(tmp_path, exp):
    p = tmp_path / "g.pfm"
    cio.save_pfm(p, exp["pfm_gray"])
    assert filecmp.cmp(p, os.path.join(G, "gray_7x5.pfm"), shallow=False)
    p = tmp_path / "c.pfm"
    cio.save_pfm(p, exp["pfm_color"], scale=2)
    assert filecmp.cmp(p, os.path.join(G, "color_6x4.pfm"), shallow=False)
    # H x W x 1 is accepted as greyscale; wrong dtype / shape are refused like the reference
    cio.save_pfm(tmp_path / "g1.pfm", exp["pfm_gray"][..., None])
    assert np.array_equal(cio.read_pfm(tmp_path / "g1.pfm")[0], exp["pfm_gray"])
    with pytest.raises(TypeError):
        cio.save_pfm(tmp_path / "x.pfm", exp["pfm_gray"].astype(np.float64))
    with pytest.raises(ValueError):
        cio.save_pfm(tmp_path / "x.pfm", np.zeros((2, 3, 4), np.float32))
    (tmp_path / "bad.pfm").write_bytes(b"P6\n1 1\n-1.0\n\0\0\0\0")
    with pytest.raises(ValueError):
        cio.read_pfm(tmp_path / "bad.pfm")


def test_pfm_roundtrip_edge_shapes(tmp_path):
    for shape in ((1, 1), (1, 9), (9, 1), (3, 5, 3)):
        a = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
        a.flat[0] = np.inf
        cio.save_pfm(tmp_path / "r.pfm", a)
        b, _ = cio.read_pfm(tmp_path / "r.pfm")
        assert np.array_equal(a, b)


def test_cam_and_pair_parsing(exp):
    K, E, dmin = cio.read_cam_file(os.path.join(G, "00000000_cam.txt"))
    assert K.dtype == np.float32 and np.array_equal(K, exp["intrinsics0"])
    assert np.array_equal(E, exp["extrinsics0"]) and dmin == exp["depth_min"][0]
    pairs = cio.read_pair_file(os.path.join(G, "pair.txt"))
    assert pairs == [(0, [1, 2, 3]), (1, [0, 2, 3]), (2, [1, 3]), (3, [2])]
    # the reference's metas (test mode: light 3 only) carry the same (ref, srcs) tuples
    got = [[3, r] + s + [-1] * (3 - len(s)) for r, s in pairs]
    assert np.array_equal(np.array(got), exp["metas_test"])


@pytest.mark.parametrize("mode,img_wh", [("train", None), ("test", (1152, 864))])
def test_projection_pyramid_bit_exact(exp, mode, img_wh):
    mats = {}
    for vid in range(4):
        K, E, _ = cio.read_cam_file(os.path.join(G, f"{vid:08d}_cam.txt"))
        mats[vid] = cio.pyramid_proj_mats(K, E, 3, img_wh)
        assert mats[vid].dtype == torch.float32
        assert np.array_equal(mats[vid].numpy(), exp[f"proj_{mode}"][vid])
    rel = cio.relative_proj_mats(mats, [0, 1, 2])
    assert rel.shape == (2, 3, 3, 4) and np.array_equal(rel.numpy(), exp[f"rel_{mode}"])


@pytest.mark.gpu
def test_normalize_u8_bit_exact_vs_torchvision(exp):
    img = torch.from_numpy(exp["img_u8"])
    out = cio.normalize_images(img.pin_memory(), device="cuda:0")
    assert torch.equal(out.cpu(), torch.from_numpy(exp["img_norm"]))
    # full-size batch, ragged H*W % 4 tail for N = 1, and the every-byte-value sweep
    big = torch.randint(0, 256, (3, 512, 640, 3), dtype=torch.uint8)
    ref = (big.permute(0, 3, 1, 2).float().div(255)
           .sub(torch.tensor(cio.IMAGENET_MEAN).view(1, 3, 1, 1))
           .div(torch.tensor(cio.IMAGENET_STD).view(1, 3, 1, 1)))
    assert torch.equal(cio.normalize_images(big.cuda()).cpu(), ref)
    odd = big[:1, :7, :9].contiguous()
    assert torch.equal(cio.normalize_images(odd.cuda()).cpu(), ref[:1, :, :7, :9])


@pytest.mark.gpu
def test_infer_views_writes_reference_layout(tmp_path):
    """eval.py:213-229 replacement: streamed views -> depth_XXXX.pfm / proba_XXXX.pfm equal to
    the eager forward of each view."""
    from casmvsnet_pl_b200 import ABN, synth
    from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet
    torch.manual_seed(0)
    model = CascadeMVSNet(norm_act=ABN, precision="tf32")
    synth.randomize_model_(model, 0)
    model = model.eval().cuda()
    views = []
    for vid in range(5):
        imgs, pm, dmin, dint = synth.make_inputs(B=1, V=3, W=160, H=128, seed=vid)
        views.append((("scan9", vid), imgs[0], pm[0]))
    writer = cio.DepthWriter(str(tmp_path / "depth"))
    n = cio.infer_views(model, iter(views), dmin, dint,
                        lambda key, d, c: writer(key[0], key[1], d, c))
    assert n == 5
    for (scan, vid), imgs, pm in views:
        res = model(imgs.unsqueeze(0).cuda(), pm.unsqueeze(0).cuda(), dmin, dint)
        d, _ = cio.read_pfm(tmp_path / "depth" / scan / f"depth_{vid:04d}.pfm")
        c, _ = cio.read_pfm(tmp_path / "depth" / scan / f"proba_{vid:04d}.pfm")
        assert np.array_equal(d, res["depth_0"][0].cpu().numpy())
        assert np.array_equal(c, res["confidence_2"][0].cpu().numpy())


@pytest.mark.gpu
def test_eval_pipeline_on_a_synthetic_scan(tmp_path):
    """eval.py end to end (depth inference -> filter -> fusion -> PLY / PFM) on a small scan written
    in the DTU test layout: images as PNG, MVSNet cam files, pair.txt."""
    cv2 = pytest.importorskip("cv2")
    from casmvsnet_pl_b200 import ABN, eval_pipeline, synth
    from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet
    root = tmp_path / "dtu"
    (root / "Cameras").mkdir(parents=True)
    (root / "Rectified" / "scan7").mkdir(parents=True)
    rng = np.random.default_rng(0)
    W, H, n = 160, 128, 4
    K4 = synth.intrinsics(0, W, H)      # cam files hold full-resolution intrinsics (dtu.py:61-63)
    for vid in range(n):
        t = np.deg2rad(3.0 * vid)
        E = np.eye(4)
        E[:3, :3] = [[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]]
        E[:3, 3] = [-680 * np.sin(t), 0, 680 * (1 - np.cos(t))]
        rows = ["extrinsic"] + [" ".join(f"{v:.6f}" for v in r) for r in E] + ["", "intrinsic"]
        rows += [" ".join(f"{v:.6f}" for v in r) for r in K4] + ["", "425.0 2.65"]
        (root / "Cameras" / f"{vid:08d}_cam.txt").write_text("\n".join(rows) + "\n")
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        cv2.imwrite(str(root / "Rectified" / "scan7" / f"rect_{vid + 1:03d}_3_r5000.png"), img)
    (root / "Cameras" / "pair.txt").write_text(
        "4\n0\n3 1 9.0 2 8.0 3 7.0\n1\n3 0 9.0 2 8.0 3 7.0\n2\n3 1 9.0 3 8.0 0 7.0\n3\n3 2 9.0 1 8.0 0 7.0\n")
    torch.manual_seed(0)
    model = CascadeMVSNet(norm_act=ABN, precision="tf32")
    synth.randomize_model_(model, 0)
    model = model.eval().cuda()
    scan = eval_pipeline.DTUTestScan(str(root), "scan7", (W, H), n_views=3, full_wh=(W, H))
    xyz, rgb = eval_pipeline.run_scan(model, scan, conf=0.0, min_geo_consistent=0,
                                      depth_dir=str(tmp_path / "depth"), ply_path=str(tmp_path / "s.ply"))
    assert xyz.shape[1] == 3 and len(xyz) == len(rgb) == n * H * W      # nothing filtered out
    d, _ = cio.read_pfm(tmp_path / "depth" / "scan7" / "depth_0002.pfm")
    assert d.shape == (H, W) and np.isfinite(d).all() and d.min() > 300
    assert os.path.getsize(tmp_path / "s.ply") > 15 * len(xyz)
    # with the reference's thresholds random images are (almost) never consistent in 5 views
    xyz2, _ = eval_pipeline.run_scan(model, scan, conf=0.999, min_geo_consistent=5)
    assert len(xyz2) < len(xyz)
