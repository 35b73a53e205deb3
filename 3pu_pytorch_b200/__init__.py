"""3pu_pytorch_b200 -- B200-native (sm_100a) implementation of the 3PU patch-upsampling hot path.

The directory name starts with a digit, so import it with
    import importlib; pu3 = importlib.import_module("3pu_pytorch_b200")
or put 3pu_pytorch_b200/shim on sys.path to get the reference's own module names
(`sampling`, `losses`, `network.operations`, ...).

Layout: csrc/ CUDA kernels + C ABI (include/pu3_b200.h) -> lib/libpu3_b200.so, loaded by _lib.py;
sampling.py / losses.py mirror the reference's two pybind modules; operations.py, layers.py,
upsampler.py, model_loss.py mirror network/*.py; patches.py (data.py's patch extraction/augmentation on the GPU) and
formats.py (checkpoint dictionary, .xyz / PLY files) are the callers and formats either side of the path.
There is no CPU path for the kernels.
"""
from . import _lib  # noqa: F401
from . import sampling, losses, operations, model_loss, fused, layers, level_train, upsampler, dist, model, pipeline, patches, formats  # noqa: F401
from .model import Model  # noqa: F401
from .upsampler import Net, Level  # noqa: F401
from .model_loss import ChamferLoss  # noqa: F401

__all__ = ["sampling", "losses", "operations", "model_loss", "fused", "layers", "upsampler", "dist", "model", "pipeline", "patches", "formats", "Net", "Level", "ChamferLoss", "Model"]
