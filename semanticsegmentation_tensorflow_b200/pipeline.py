"""GPU input pipeline: the per-image work of the reference's `get_batches_fn`
(`FCN.py:242-305`) after PNG decode, on the device — three views per decoded image (`:276-304`):
   (1) whole image resized + brightness/contrast (`bc_img`, `:186-192,287-289`)
   (2) random 3.3:1 crop resized (`crop_image`, `:176-182`)
   (3) horizontal flip resized (`flip_image`, `:184-185`)
and for each the ground-truth image resized the same way and colour-matched to class ids
(`process_gt_image`, `:194-201`; background (255,0,0) -> 0, road -> 1).

`scipy.misc.imresize` is PIL's `Image.resize(BILINEAR)`; its coefficient tables are restated here
(host, double precision, as PIL's precompute_coeffs / normalize_coeffs_8bpc) and the kernels apply
them in integer arithmetic, so results are bit-exact with PIL (tests pin this against PIL itself), including
PIL's premultiplied-alpha handling of 4-channel (RGBA) images.  PNG inflate stays on the host."""
from __future__ import annotations

import math
import random
from functools import lru_cache

import numpy as np
import torch

from .ops import Ops, _p, _stream

PRECISION_BITS = 32 - 8 - 2


@lru_cache(maxsize=256)
def pil_bilinear_coeffs(in_size: int, out_size: int):
    """(coeffs int32 [out][ksize], bounds int32 [out][2], ksize) exactly as PIL computes them."""
    scale = float(in_size) / out_size
    fscale = scale if scale > 1.0 else 1.0
    support = 1.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.float64)
    bounds = np.zeros((out_size, 2), np.int32)
    ss = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    scaled = kk * float(1 << PRECISION_BITS)
    ik = np.where(kk < 0, (scaled - 0.5).astype(np.int64), (scaled + 0.5).astype(np.int64)).astype(np.int32)
    return ik, bounds, ksize


class GpuBatcher:
    """Device-side `get_batches_fn`: feed decoded u8 images + ground-truth images, get the reference's
    3-view batch as (images u8 [3B,H,W,3], labels u8 [3B,H,W]) CUDA tensors."""

    def __init__(self, image_shape=(160, 576), device=None, seed=None):
        self.oh, self.ow = image_shape
        self.ops = Ops(device)
        self.device = torch.device("cuda", self.ops.ctx.device)
        self.rng = random.Random(seed)
        self._tables = {}

    def _table(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self._tables:
            ik, b, ks = pil_bilinear_coeffs(in_size, out_size)
            self._tables[key] = (torch.as_tensor(ik).to(self.device), torch.as_tensor(b).to(self.device), ks)
        return self._tables[key]

    def resize(self, src, out, crop=None, flip=False, mode=0, contrast=1.0, brightness=0.0):
        """src u8 [H,W,C] (device) -> out; crop = (x0, y0, w, h) of src or None."""
        H, W, C = src.shape
        x0, y0, cw, ch = crop if crop is not None else (0, 0, W, H)
        kx, bx, ksx = self._table(cw, self.ow)
        ky, by, ksy = self._table(ch, self.oh)
        tmp = torch.empty((ch, self.ow, C), dtype=torch.uint8, device=self.device)
        call = self.ops.call
        call("segk_resize_h_u8", _p(src), _p(tmp), _p(kx), _p(bx), ksx, W, C, x0, y0, cw, ch, self.ow, int(flip), _stream())
        call("segk_resize_v_u8", _p(tmp), _p(out), _p(ky), _p(by), ksy, C, self.ow, self.oh, mode, float(contrast),
             float(brightness), _stream())
        return out

    def crop_box(self, h, w):
        """crop_image (FCN.py:176-182): random width in [1150, w-5], height = int(width / 3.3)."""
        nw = self.rng.randint(min(1150, w - 5), w - 5)
        nh = int(nw / 3.3)
        x1 = self.rng.randint(0, w - nw)
        y1 = self.rng.randint(0, h - nh)
        return x1, y1, nw, nh

    def batch(self, images, gt_images, params=None):
        """images: decoded u8 [H,W,3] or [H,W,4] arrays / tensors (4 = the RGBA "merge" PNGs the reference trains on,
        FCN.py:225,257,312: resized with PIL's premultiplied-alpha semantics); gt_images: u8 [H,W,3].  `params`
        (optional) pins the random draws per image: dicts with crop=(x0,y0,w,h), contrast, brightness."""
        n = len(images)
        chans = {int(torch.as_tensor(im).shape[2]) for im in images}
        if len(chans) != 1 or next(iter(chans)) not in (3, 4):
            raise ValueError(f"images must all have 3 or all have 4 channels (got {sorted(chans)})")
        if any(torch.as_tensor(g).shape[2] != 3 for g in gt_images):
            raise ValueError("ground-truth images must be RGB (FCN.py:196: compared against [255, 0, 0])")
        out_x = torch.empty((3 * n, self.oh, self.ow, next(iter(chans))), dtype=torch.uint8, device=self.device)
        out_y = torch.empty((3 * n, self.oh, self.ow), dtype=torch.uint8, device=self.device)
        for i, (im, gt) in enumerate(zip(images, gt_images)):
            im = torch.as_tensor(im).to(self.device, non_blocking=True).contiguous()
            gt = torch.as_tensor(gt).to(self.device, non_blocking=True).contiguous()
            h, w = im.shape[:2]
            prm = params[i] if params is not None else {}
            crop = prm.get("crop") or self.crop_box(h, w)
            contrast = prm.get("contrast", self.rng.uniform(0.85, 1.15))        # FCN.py:287
            bright = prm.get("brightness", self.rng.randint(-45, 30))           # FCN.py:288
            # view 1: whole image + brightness/contrast; view 2: crop; view 3: flip (FCN.py:276-304)
            self.resize(im, out_x[3 * i], mode=1, contrast=contrast, brightness=bright)
            self.resize(gt, out_y[3 * i], mode=2)
            self.resize(im, out_x[3 * i + 1], crop=crop)
            self.resize(gt, out_y[3 * i + 1], crop=crop, mode=2)
            self.resize(im, out_x[3 * i + 2], flip=True)
            self.resize(gt, out_y[3 * i + 2], flip=True, mode=2)
        return out_x, out_y
