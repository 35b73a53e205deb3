"""Stand-in for h5py (not in this image).  The reference imports it at module level (data.py:3, reached through main.py:17) but
only opens a file when training starts (data.py:84): importing works, opening a file raises."""


class File:
    def __init__(self, *args, **kwargs):
        raise ImportError("h5py is not installed in this environment (3pu_pytorch_b200/shim/h5py is an import stand-in); "
                          "the training set loader needs the real package")
