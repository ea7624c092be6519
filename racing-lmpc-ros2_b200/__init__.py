"""racing-lmpc-ros2_b200: B200-native batched LMPC solve behind the RacingMPC::solve surface.

Host-side Python mirror of the C-ABI in include/lmpc_b200.h (the product is the CUDA library in
csrc/; this package only marshals buffers).  No CPU fallback exists: every compute entry point
raises if the CUDA library or a GPU is missing.
"""
from . import configs, workload  # noqa: F401
