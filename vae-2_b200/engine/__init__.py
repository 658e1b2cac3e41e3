"""Execution engine of the B200-native VAE^2 path (plans, autograd glue, ctypes binding)."""
from .module import (EngineModule, set_precision, get_precision, use_cuda_graphs, launch_count,  # noqa: F401
                     activation_phase, ActArena)
from .elbo import elbo_terms, check_finite, finite_check_mode  # noqa: F401
from . import native  # noqa: F401
from . import peer  # noqa: F401
from .profiler import KernelProfile  # noqa: F401
