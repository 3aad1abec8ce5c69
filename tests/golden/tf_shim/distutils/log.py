def error(*a, **k): pass
