"""IterativeCollisionCheckLink -- placeholder import target until icc.cu lands (same commit series)."""


class IterativeCollisionCheckLink:  # replaced below in this round
    def __init__(self, *a, **k):
        raise NotImplementedError("ICC kernel not built yet")
