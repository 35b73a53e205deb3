def get_cmap(name=None, lut=None):
    raise ImportError("matplotlib is not installed in this environment (3pu_pytorch_b200/shim/matplotlib is an import stand-in)")
