"""Drop-in for the reference's `network` package -> 3pu_pytorch_b200."""
