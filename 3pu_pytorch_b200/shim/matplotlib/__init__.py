"""Stand-in for matplotlib (not in this image).  The reference imports `from matplotlib import cm` at module level
(utils/pc_utils.py:7) but only uses it in the colour-by-property PLY writer and the --phase vis GUI, both outside the hot path:
importing works, using raises."""
from . import cm  # noqa: F401
