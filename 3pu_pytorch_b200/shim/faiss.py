"""Stand-in for `faiss`, which the reference imports unconditionally (network/operations.py:2) but only uses on a
dead code path (search_index_pytorch / KNN, operations.py:33-106).  Nothing here is ever called."""


class StandardGpuResources:  # faiss_setup.py:4 constructs one at import time
    def __init__(self, *a, **k):
        pass
