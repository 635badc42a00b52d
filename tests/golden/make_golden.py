"""Generates the fixtures in tests/golden/ (run from the repo root: `python tests/golden/make_golden.py`).

WHAT THESE ARE -- AND ARE NOT.  The reference's arithmetic lives in TensorFlow 1.x, which cannot be
installed in the build container, and the reference ships no tests or vectors, so nothing
here was produced by the reference itself: parity stays UNPINNED.  The fixtures freeze

  * fcn_tiny.npz      outputs of the CPU oracle (`oracle/`, fp32, TF-1.x op semantics) for a small FCN-8s
                      (fc = 64, 32x64 image, He init): logits, loss, per-variable gradient norms and a few
                      full gradients, three TF-Adam steps.  Pins the oracle against accidental change and
                      gives the GPU tests a second, file-based target.
  * pil_resize.npz    outputs of PIL itself (the library `scipy.misc.imresize` in `get_batches_fn` calls,
                      FCN.py:270-304) for the three augmentation views + label resize, and of PIL's
                      alpha-composite for `paste_mask` (FCN.py:203-211).  This one IS pinned to the library
                      the reference uses (Pillow version recorded in the file).
  * pool_adam.npz     hand-checkable vectors: 2x2 max-pool with ties (first-max routing) and one TF-form
                      Adam step (closed form).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def fcn_tiny():
    import torch
    from oracle.fcn_oracle import FCN8sOracle, synthetic_batch
    from semanticsegmentation_tensorflow_b200 import plan as P
    from semanticsegmentation_tensorflow_b200.fcn import reference_init
    torch.set_num_threads(1)
    fc, n, h, w = 64, 2, 32, 64
    variables = reference_init(P.variable_shapes(3, 2, fc), 1234, "he")
    x, lab = synthetic_batch(n, h, w, seed=3, road_shaped=True)
    x = (x // 32).astype(np.uint8)
    orc = FCN8sOracle(variables)
    loss, logits, grads = orc.loss_and_grads(x, lab)
    out = {"x": x, "labels": lab, "fc": np.int64(fc), "loss": np.float64(loss), "logits": logits.numpy(),
           "grad_names": np.array(list(grads.keys())),
           "grad_norms": np.array([float(g.norm()) for g in grads.values()], dtype=np.float64)}
    for k in ("conv1_1/weights", "conv8/weights", "conv_t3/weights", "conv_t3/bias"):
        out["grad:" + k] = grads[k].numpy()
    losses = []
    orc2 = FCN8sOracle(variables)
    for _ in range(3):
        losses.append(orc2.train_step(x, lab)[0])
    out["adam_losses"] = np.array(losses, dtype=np.float64)
    out["conv8_weights_after_3_steps"] = orc2.vars["conv8/weights"].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "fcn_tiny.npz"), **out)


def pil_resize():
    import PIL
    from PIL import Image
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (37, 91, 3), dtype=np.uint8)
    lab = (rng.integers(0, 2, (37, 91)) * 255).astype(np.uint8)
    out = {"image": img, "label": lab, "pillow_version": np.array(PIL.__version__)}
    for name, (hh, ww) in {"up": (48, 160), "down": (16, 40), "same": (37, 91)}.items():
        out["resize_" + name] = np.asarray(Image.fromarray(img).resize((ww, hh), Image.BILINEAR))
        out["label_" + name] = np.asarray(Image.fromarray(lab).resize((ww, hh), Image.BILINEAR))
    # paste_mask: street image + [0,255,0,127] overlay where mask (FCN.py:203-211)
    street = Image.fromarray(img)
    seg = rng.integers(0, 2, (37, 91)).astype(bool)
    mask = np.dot(seg.reshape(37, 91, 1), np.array([[0, 255, 0, 127]])).astype(np.uint8)
    street.paste(Image.fromarray(mask, mode="RGBA"), box=None, mask=Image.fromarray(mask, mode="RGBA"))
    out["paste_seg"] = seg
    out["paste_result"] = np.asarray(street)
    np.savez_compressed(os.path.join(HERE, "pil_resize.npz"), **out)


def pool_adam():
    # 2x2 s2 max-pool on a 4x4x1 map with ties: strict '>' scan in (dy,dx) order keeps the first maximum
    x = np.array([[1, 1, 0, 2], [1, 0, 2, 2], [3, 3, 5, 4], [3, 3, 4, 5]], dtype=np.float32).reshape(1, 4, 4, 1)
    y = np.array([[1, 2], [3, 5]], dtype=np.float32).reshape(1, 2, 2, 1)
    idx = np.array([[0, 1], [0, 0]], dtype=np.uint8).reshape(1, 2, 2, 1)          # window index dy*2+dx
    dy = np.array([[10, 20], [30, 40]], dtype=np.float32).reshape(1, 2, 2, 1)
    dx = np.zeros((1, 4, 4, 1), np.float32)
    dx[0, 0, 0, 0], dx[0, 0, 3, 0], dx[0, 2, 0, 0], dx[0, 2, 2, 0] = 10, 20, 30, 40
    # TF ApplyAdam, step 1, from m = v = 0: lr_t = lr*sqrt(1-b2)/(1-b1); m = (1-b1) g; v = (1-b2) g^2
    g = np.array([0.5, -2.0, 1e-3, 0.0], dtype=np.float32)
    p0 = np.array([1.0, -1.0, 0.25, 3.0], dtype=np.float32)
    lr, b1, b2, eps = 1e-4, 0.9, 0.999, 1e-8
    lr_t = lr * np.sqrt(1 - b2) / (1 - b1)
    m = (1 - b1) * g.astype(np.float64)
    v = (1 - b2) * g.astype(np.float64) ** 2
    p1 = p0 - lr_t * m / (np.sqrt(v) + eps)
    np.savez_compressed(os.path.join(HERE, "pool_adam.npz"), pool_x=x, pool_y=y, pool_idx=idx, pool_dy=dy, pool_dx=dx,
                        adam_g=g, adam_p0=p0, adam_p1=p1.astype(np.float64), adam_m1=m, adam_v1=v)


if __name__ == "__main__":
    fcn_tiny()
    pil_resize()
    pool_adam()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
