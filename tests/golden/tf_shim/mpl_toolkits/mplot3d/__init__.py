class Axes3D: pass
from . import proj3d
