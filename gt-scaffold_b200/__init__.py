"""gt-scaffold_b200 -- B200-native hot path of gt Scaffolder.

Scaffold-graph construction from distance records, repeat marking and the
polymorphic/inconsistent filter (reference gt_scaffolder_parser.c:295-394,
gt_scaffolder_algorithms.c:90-343) as hand-written sm_100a CUDA kernels behind
a C ABI (include/gtscaffold_b200.h).  The directory name carries a hyphen, so
import it with importlib.import_module("gt-scaffold_b200").
"""
from . import api, synth  # noqa: F401
from .api import ScaffoldGraphB200, load_library  # noqa: F401

__all__ = ["api", "synth", "ScaffoldGraphB200", "load_library"]
