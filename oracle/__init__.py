"""TEST INFRASTRUCTURE: CPU checkers for the WEmbed step (see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package `wembed_b200` never does.
"""
from .oracle import CpuEmbedder, OrcOptions, build, have, ref_read_edge_list  # noqa: F401
