"""montecarlolocalisation_b200 — B200-native particle-filter hot path of Bright8787/MonteCarloLocalisation.

The product is libmcl_b200.so (hand-written sm_100a CUDA behind the C-ABI in include/mcl.h). This package is the thin
Python host mirror used by the tests and bench.py; it never falls back to a CPU implementation.
"""
from ._lib import MODE_NS, MODE_REF, TRIG_CORRECTLY_ROUNDED, TRIG_LIBM, MclError, build, load  # noqa: F401
from .particle_filter import (NsShard, ParticleFilter, default_config, ns_first_slot, ns_shard_range,  # noqa: F401
                              ns_step_in_process, rasterise_map_txt, pose_to_cell, exact_pose)
