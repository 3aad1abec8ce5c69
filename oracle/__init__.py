"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes front-end of the CPU oracle (oracle/mppi_oracle.c) plus a graph-faithful
op-for-op restatement of the reference's TensorFlow graph (oracle/graph_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this package; the product package mppi_tf_b200 never does.
"""
from .pyoracle import Oracle, load, build  # noqa: F401
