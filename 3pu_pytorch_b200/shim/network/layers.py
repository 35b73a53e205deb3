"""Drop-in for the reference's network/layers.py -> 3pu_pytorch_b200.layers"""
from importlib import import_module as _im

_impl = _im("3pu_pytorch_b200.layers")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
