"""B200-native (sm_100a) segmentation hot path behind the reference's model-builder API.

Drop-in surface (names follow `Network/model/FCN.py` and `Network/utils/utils.py` of the reference):
    FCN(x, keep_prob, num_classess).create() -> (pred, logits)        FCN.py:31-114      (fcn.py)
    conv_layer / deconv_layer / max_pool / dropout / fuse              FCN.py:117-171     (layers.py)
    AdamOptimizer(lr).minimize(net) -> train_step(feed_dict)           FCN.py:338-340,398 (fcn.py)
    SegNet(x, num_classes) / UNet(x, num_classes) (graph.py); FCDenseNet(x, keep_prob, num_classes) (densenet.py)

All arithmetic runs in hand-written CUDA kernels reached through the C ABI of
`libsegk.so` (`include/segk.h`) via ctypes.  torch is used for device memory, streams and
`torch.distributed` only.  There is no CPU fallback: importing the ops without the built
library raises.
"""
from .build import build_library, library_path  # noqa: F401

_LAZY = {
    "FCN": "fcn", "AdamOptimizer": "fcn", "MomentumOptimizer": "fcn", "gen_test_output": "fcn",
    "conv_layer": "layers", "deconv_layer": "layers", "max_pool": "layers", "dropout": "layers", "fuse": "layers",
    "VariableStore": "layers",
    "SegNet": "graph", "UNet": "graph", "FCDenseNet": "densenet",
}

__all__ = ["build_library", "library_path"] + sorted(_LAZY)


def __getattr__(name):
    # torch-dependent modules load on first use, so `import semanticsegmentation_tensorflow_b200` (and the
    # CPU-only build check) stay light
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f".{_LAZY[name]}", __name__), name)
    raise AttributeError(name)
