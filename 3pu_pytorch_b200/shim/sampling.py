"""Drop-in for the reference's `sampling` extension module -> 3pu_pytorch_b200.sampling."""
from importlib import import_module as _im

_impl = _im("3pu_pytorch_b200.sampling")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("_")})
