class FancyArrowPatch:  # shim
    def __init__(self, *a, **k): pass
