"""Oracle for the input pipeline (`get_batches_fn`, FCN.py:242-305, after PNG decode).  TEST
INFRASTRUCTURE ONLY.  PINNED: `scipy.misc.imresize(arr, shape)` (removed from SciPy >= 1.3) was a thin
wrapper `toimage(arr).resize((w, h), resample=BILINEAR)` over PIL, which IS installed — so the resize is
checked against PIL itself; crop / flip / bc_img / process_gt_image are the reference's NumPy lines."""
import numpy as np
from PIL import Image


def imresize(arr: np.ndarray, shape) -> np.ndarray:
    """scipy.misc.imresize(arr, (h, w)) with the default interp='bilinear' (FCN.py:273-281)."""
    h, w = shape
    return np.array(Image.fromarray(arr).resize((w, h), resample=Image.BILINEAR))


def bc_img(img, s=1.0, m=0.0):
    """FCN.py:186-192 (np.int was an alias of the builtin int)."""
    img = img.astype(np.int64)
    img = img * s + m
    img[img > 255] = 255
    img[img < 0] = 0
    return img.astype(np.uint8)


def process_gt_image(gt_image):
    """FCN.py:194-201 -> class ids (channel 1 of the reference's one-hot = not background)."""
    gt_bg = np.all(gt_image == np.array([255, 0, 0]), axis=2)
    return np.invert(gt_bg).astype(np.uint8)


def three_views(image, gt_image, image_shape, crop, contrast, bright):
    """One loop body of get_batches_fn (FCN.py:270-304) with the random draws injected.
    crop = (x1, y1, nw, nh) as crop_image draws them (FCN.py:176-182)."""
    x1, y1, nw, nh = crop
    image2, gt2 = image[y1:y1 + nh, x1:x1 + nw, :], gt_image[y1:y1 + nh, x1:x1 + nw, :]
    image3, gt3 = np.flip(image, axis=1), np.flip(gt_image, axis=1)
    im1 = bc_img(imresize(image, image_shape), contrast, bright)
    views = [im1, imresize(np.ascontiguousarray(image2), image_shape), imresize(np.ascontiguousarray(image3), image_shape)]
    gts = [process_gt_image(imresize(g, image_shape)) for g in
           (gt_image, np.ascontiguousarray(gt2), np.ascontiguousarray(gt3))]
    return views, gts
