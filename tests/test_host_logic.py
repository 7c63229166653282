"""CPU: host-side logic of the boundary — tile-order RNG replay, sharding arithmetic, pose/focal math."""
import math

import numpy as np
import torch

import nerf_b200
from nerf_b200 import procedures, sharding


def test_get_patch_size_and_64x64():
    assert procedures.get_patch_size((400, 400)) == (50, (8, 8))
    assert procedures.get_patch_size((800, 800)) == (50, (16, 16))
    assert procedures.get_patch_size((120, 120)) == (40, (3, 3))
    # the reference raises UnboundLocalError here (nerf/procedures.py:24-31); the engine renders one tile
    assert procedures.get_patch_size((64, 64)) == (None, (1, 1))
    # width-only rule, like the reference: the height is truncated to whole tiles (rows past it are never rendered)
    assert procedures.get_patch_size((70, 100)) == (50, (1, 2)) and procedures.rendered_rows((70, 100)) == 50
    assert procedures.get_patch_size((420, 400)) == (50, (8, 8)) and procedures.rendered_rows((420, 400)) == 400
    assert procedures.get_patch_size((30, 100)) == (50, (0, 2)) and procedures.rendered_rows((30, 100)) == 0
    assert procedures.get_patch_size((100, 80)) == (40, (2, 2))
    assert procedures.rendered_rows((64, 64)) == 64


def test_reference_rng_replay_truncated_rows():
    """70x100: the reference draws for the two 50x50 tiles of the first tile row only."""
    torch.manual_seed(9)
    jit, u = procedures._reference_rng_draws((70, 100), 64, 129)
    assert jit.shape == (50 * 100, 64) and u.shape == (50 * 100, 129)
    torch.manual_seed(9)
    for j in range(2):
        a = torch.rand((50, 50, 64))
        b = torch.rand([2500, 129])
        assert torch.equal(jit.view(50, 100, 64)[:, 50 * j:50 * (j + 1)], a)
        assert torch.equal(u.view(50, 100, 129)[:, 50 * j:50 * (j + 1)], b.view(50, 50, 129))


def test_reference_rng_replay_order():
    """'reference' RNG mode consumes torch's CPU generator exactly like the reference's tile loop."""
    H = W = 100
    torch.manual_seed(7)
    jit, u = procedures._reference_rng_draws((H, W), 64, 129)
    torch.manual_seed(7)
    sz = 50
    for k in range(2):
        for j in range(2):
            a = torch.rand((sz, sz, 64))          # procedures.py:65
            b = torch.rand([sz * sz, 129])        # utils.py:115
            assert torch.equal(jit.view(H, W, 64)[sz * k:sz * (k + 1), sz * j:sz * (j + 1)], a)
            assert torch.equal(u.view(H, W, 129)[sz * k:sz * (k + 1), sz * j:sz * (j + 1)], b.view(sz, sz, 129))


def test_shard_range_covers_and_aligns():
    for n in (0, 1, 7, 160000, 640000, 640001):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world, align=2) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for s, c in spans:
                assert s == min(pos, n) or c == 0
                pos += c
            assert all(c % 2 == 0 for _, c in spans[:-1] if c and _ + c < n)


def test_fov2focal_and_pose():
    f = nerf_b200.fov2Focal(0.6911112070083618, (400, 400))
    assert abs(f[0] - 400 / math.tan(0.5 * 0.6911112070083618)) < 1e-9 and f[0] == f[1]
    fx = nerf_b200.fov2Focal((0.7, 0.5), (300, 400))
    assert abs(fx[0] - 0.5 * 300 / math.tan(0.25)) < 1e-9 and abs(fx[1] - 0.5 * 400 / math.tan(0.35)) < 1e-9
    p = nerf_b200.pose_spherical(30.0, -30.0, 4.0)
    assert p.shape == (4, 4) and abs(float(p[:3, 3].norm()) - 4.0) < 1e-5
    R = p[:3, :3]
    assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-6)


def test_lr_schedule_matches_formula():
    s = nerf_b200.DecayLrScheduler(0.1, 0.5, 100, 1e-3, warmup_step=10)
    assert abs(s.update_opt_lr(5)[1] - 1e-3 * (0.1 * 0.5 + 0.5)) < 1e-12
    assert abs(s.update_opt_lr(110)[1] - 1e-3 * 0.5) < 1e-12
    assert abs(s.update_opt_lr(100000)[1] - 1e-4) < 1e-12


def test_get_parser_matches_the_reference_defaults():
    """tests/golden/reference_parser_defaults.json = vars(get_parser().parse_args([])) of the UNMODIFIED reference
    (nerf/procedures.py:166-213)."""
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_parser_defaults.json")))
    got = vars(nerf_b200.get_parser().parse_args([]))
    assert got == ref
    a = nerf_b200.get_parser().parse_args(["-w", "-t", "--render_depth", "--sample_ray_num", "512", "--lr", "1e-3", "-e"])
    assert a.white_bkg and a.ref_nerf and a.render_depth and a.eval_poses and a.sample_ray_num == 512 and a.lr == 1e-3


def test_blender_dataset_loader(tmp_path):
    """nerf/dataset.py:22-114 behaviour on a tiny synthetic Blender-format scene: natural file order, fov / transforms from
    the JSON, white-background alpha compositing, scene_scale on the translation, AdaptiveResize."""
    import json
    from PIL import Image
    from torchvision import transforms
    from nerf_b200.dataset import AdaptiveResize, CustomDataSet
    root = tmp_path / "toy"
    (root / "test").mkdir(parents=True)
    frames = []
    for i in (0, 1, 2, 10):
        arr = np.zeros((8, 8, 4), dtype=np.uint8)
        arr[..., 0] = 10 * i + 5
        arr[..., 3] = 128 if i == 2 else 255
        Image.fromarray(arr, "RGBA").save(root / "test" / f"r_{i}.png")
        tf = np.eye(4)
        tf[:3, 3] = [i, 2 * i, 3.0]
        frames.append({"transform_matrix": tf.tolist()})
    Image.fromarray(np.zeros((8, 8, 4), dtype=np.uint8), "RGBA").save(root / "test" / "r_0_normal.png")
    json.dump({"camera_angle_x": 0.69, "frames": frames}, open(root / "transforms_test.json", "w"))
    tfm = transforms.Compose([AdaptiveResize(0.5), transforms.ToTensor()])
    ds = CustomDataSet(str(root) + "/", tfm, scene_scale=2.0, is_train=False, white_bkg=True)
    assert ds.total_imgs == ["r_0.png", "r_1.png", "r_2.png", "r_10.png"] and len(ds) == 4       # natural order, no *normal*
    fov, tfs = ds.getCameraParam()
    assert fov == 0.69 and tfs.shape == (4, 3, 4) and ds.r_c() == (4, 4)
    img, tf = ds[2]
    assert img.shape == (3, 4, 4)
    a = 128 / 255
    assert abs(float(img[0, 0, 0]) - ((25 / 255) * a + (1 - a))) < 2e-2 and abs(float(img[1, 0, 0]) - (1 - a)) < 2e-2
    assert torch.allclose(tf[:, 3], torch.tensor([2.0, 4.0, 3.0]) * 2.0)
    assert torch.allclose(ds.tfs[2][:, 3], torch.tensor([2.0, 4.0, 3.0]))                      # the stored transforms stay unscaled


def test_launch_plan_records_batches_and_rebases_dynamic_pointers():
    """linear.Program (host logic only: recording never calls the library): consecutive GEMMs form one batch, a non-GEMM
    call splits batches, reductions are deferred, pointers into bound tensors are re-based on every run."""
    import ctypes
    from nerf_b200 import _lib, linear
    assert ctypes.sizeof(_lib.ReduceDesc) == 56
    bf = torch.bfloat16
    x, w = torch.zeros((16, 8), dtype=bf), torch.zeros((4, 8), dtype=bf)
    y = torch.zeros((16, 8), dtype=bf)
    out = torch.zeros((16, 4), dtype=torch.float32)
    ws = torch.zeros(2 * 128 * 32, dtype=torch.float32)
    grads = torch.zeros(64, dtype=torch.float32)
    calls = []
    with linear.Program("cpu") as prog:
        prog.bind(out=out, grads=grads)
        linear.gemm(16, 8, [(x, False, w, False, 8)], out_hi=y)
        linear.gemm(16, 1, [(y, False, w, False, 8)], out_f32=out[:, 3:])
        prog.call(lambda: calls.append(1))
        linear.gemm(16, 3, [(y, False, w, False, 8)], out_f32=out[:, :3])
        linear.reduce_splits(ws, 2, 128 * 32, 4, 8, 32, grads[32:].view(4, 8))
    assert linear.recording() is None
    assert [n for _, n in prog._final] == [2, -1, 1] and len(prog._red) == 1 and prog.launches == 5
    # re-base by hand what run() does (run() itself needs the CUDA library)
    out2, grads2 = torch.zeros_like(out), torch.zeros_like(grads)
    base = {"out": out2.data_ptr(), "grads": grads2.data_ptr()}
    for elem, field, key, off in prog._patches:
        setattr(elem, field, base[key] + off)
    assert prog._final[0][0][1].out_f32 == out2.data_ptr() + 12
    assert prog._final[2][0][0].out_f32 == out2.data_ptr()
    assert prog._red[0].out == grads2.data_ptr() + 128 and prog._red[0].split_stride == 128 * 32
    assert prog._final[0][0][0].out_f32 is None          # not dynamic, not an fp32 output
