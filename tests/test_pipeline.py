"""Input pipeline: the host-side PIL coefficient restatement is pinned against PIL on the CPU; the GPU
kernels are then bit-exact against the PIL-based oracle of get_batches_fn (FCN.py:242-305)."""
import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as PO


def _numpy_resize(img, oh, ow):
    from semanticsegmentation_tensorflow_b200.pipeline import PRECISION_BITS, pil_bilinear_coeffs
    h, w, c = img.shape
    kx, bx, _ = pil_bilinear_coeffs(w, ow)
    ky, by, _ = pil_bilinear_coeffs(h, oh)
    tmp = np.zeros((h, ow, c), np.uint8)
    for xx in range(ow):
        x0, n = bx[xx]
        ss = (1 << (PRECISION_BITS - 1)) + np.tensordot(img[:, x0:x0 + n].astype(np.int64), kx[xx, :n].astype(np.int64), axes=([1], [0]))
        tmp[:, xx] = np.clip(ss >> PRECISION_BITS, 0, 255)
    out = np.zeros((oh, ow, c), np.uint8)
    for yy in range(oh):
        y0, n = by[yy]
        ss = (1 << (PRECISION_BITS - 1)) + np.tensordot(tmp[y0:y0 + n].astype(np.int64), ky[yy, :n].astype(np.int64), axes=([0], [0]))
        out[yy] = np.clip(ss >> PRECISION_BITS, 0, 255)
    return out


@pytest.mark.parametrize("shape", [(375, 1242, 160, 576), (100, 333, 160, 576), (348, 1150, 160, 576), (37, 50, 37, 50),
                                   (375, 1242, 384, 1248)])
def test_coefficient_tables_reproduce_pil_bilinear(shape):
    h, w, oh, ow = shape
    img = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(_numpy_resize(img, oh, ow), PO.imresize(img, (oh, ow)))


def _kitti_like(rng, h=375, w=1242):
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    gt = np.zeros((h, w, 3), np.uint8)
    gt[..., 0] = 255                                         # background (255, 0, 0)
    yy, xx = np.mgrid[0:h, 0:w]
    road = (yy > h // 2) & (np.abs(xx - w // 2) < (yy - h // 2) * 2)
    gt[road] = (255, 0, 255)                                 # KITTI road colour
    return img, gt


def test_rgba_resize_restatement_matches_pil():
    """PIL resizes RGBA through premultiplied alpha (RGBA -> RGBa, filter, RGBa -> RGBA); the integer formulas the
    kernels use (csrc/pipeline.cu) restated in NumPy and pinned against PIL itself."""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (37, 51, 4), dtype=np.uint8)
    a[:5, :5, 3] = 0
    a[5:10, :, 3] = 255
    al = a[..., 3:4].astype(np.int64)
    t = a[..., :3].astype(np.int64) * al + 128
    pm = np.concatenate([((t >> 8) + t) >> 8, al], -1).astype(np.uint8)
    from PIL import Image
    assert np.array_equal(pm, np.array(Image.fromarray(a, "RGBA").convert("RGBa")))
    r = _numpy_resize(pm, 16, 24).astype(np.int64)
    ra = r[..., 3:4]
    rgb = np.where((ra == 0) | (ra == 255), r[..., :3], np.minimum(255, (255 * r[..., :3]) // np.where(ra == 0, 1, ra)))
    got = np.concatenate([rgb, ra], -1).astype(np.uint8)
    assert np.array_equal(got, PO.imresize(a, (16, 24)))


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [3, 4])
def test_gpu_batcher_bit_exact_vs_reference_pipeline(cuda_device, channels):
    from semanticsegmentation_tensorflow_b200.pipeline import GpuBatcher
    rng = np.random.default_rng(1)
    imgs, gts, params = [], [], []
    for i in range(2):
        img, gt = _kitti_like(rng)
        if channels == 4:      # the reference's "merge" PNGs: RGB + a fourth (LiDAR) channel that PIL treats as alpha
            alpha = rng.integers(0, 256, img.shape[:2] + (1,), dtype=np.uint8)
            alpha[:40] = 255
            alpha[40:60] = 0
            img = np.concatenate([img, alpha], axis=2)
        imgs.append(img); gts.append(gt)
        params.append({"crop": (17 + i, 9, 1180 + i, int((1180 + i) / 3.3)), "contrast": 0.93 + 0.1 * i, "brightness": -20 + 35 * i})
    gb = GpuBatcher((160, 576), cuda_device)
    x, y = gb.batch(imgs, gts, params)
    torch.cuda.synchronize()
    x, y = x.cpu().numpy(), y.cpu().numpy()
    assert x.shape == (6, 160, 576, channels)
    for i in range(2):
        views, labels = PO.three_views(imgs[i], gts[i], (160, 576), params[i]["crop"], params[i]["contrast"], params[i]["brightness"])
        for v in range(3):
            assert np.array_equal(x[3 * i + v], views[v]), (i, v)
            assert np.array_equal(y[3 * i + v], labels[v]), (i, v)
    assert set(np.unique(y)) <= {0, 1} and 0 < y.mean() < 1
    # random draws stay inside the reference's ranges
    gb2 = GpuBatcher((160, 576), cuda_device, seed=3)
    for _ in range(20):
        x1, y1, nw, nh = gb2.crop_box(375, 1242)
        assert 1150 <= nw <= 1237 and nh == int(nw / 3.3) and 0 <= x1 <= 1242 - nw and 0 <= y1 <= 375 - nh
