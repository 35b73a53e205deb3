"""Drop-in for the reference's network/operations.py -> 3pu_pytorch_b200.operations"""
from importlib import import_module as _im

_impl = _im("3pu_pytorch_b200.operations")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
