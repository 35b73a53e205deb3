"""Stand-in for visdom (not in this image; imported by the reference only in --phase train, main.py:160,423): every plotting
call is accepted and ignored, so a training run does not need a visdom server."""


class Visdom:
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        def _noop(*args, **kwargs):
            return None
        return _noop
