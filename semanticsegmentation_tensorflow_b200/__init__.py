"""B200-native (sm_100a) segmentation hot path behind the reference's model-builder API.

Drop-in surface (names follow `Network/model/FCN.py` of the reference):
    FCN(x, keep_prob, num_classess).create() -> (pred, logits)       FCN.py:31-114
    conv_layer / deconv_layer / max_pool / dropout / fuse             FCN.py:117-171
    AdamOptimizer(lr).minimize(net) -> train_step(feed_dict)          FCN.py:338-340,398

All arithmetic runs in hand-written CUDA kernels reached through the C ABI of
`libsegk.so` (`include/segk.h`) via ctypes.  torch is used for device memory, streams and
`torch.distributed` only.  There is no CPU fallback: importing the ops without the built
library raises.
"""
from .build import build_library, library_path  # noqa: F401

__all__ = ["build_library", "library_path"]
