"""CPU oracle for the range-analysis hot path (TEST INFRASTRUCTURE ONLY -- never imported by the product).

A float32 NumPy restatement of the reference (nmwsharp/neural-implicit-queries, src/*.py); each function
cites the reference file:line it follows.  See net.py for the pinning statement.
"""
from . import mc, mc_tables, net, rays, tree  # noqa: F401
