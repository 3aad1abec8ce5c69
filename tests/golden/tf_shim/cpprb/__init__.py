class ReplayBuffer:  # shim: the update path never touches it
    def __init__(self, *a, **k): pass
