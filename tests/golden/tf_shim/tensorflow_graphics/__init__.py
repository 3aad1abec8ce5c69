"""Stub: elipse_cost.py imports tensorflow_graphics at module level; only ElipseCost3D (AUV, out of scope) uses it."""
