"""ampis_b200 -- B200-native (sm_100a) implementation of AMPIS's mask evaluation and
measurement hot path, behind AMPIS's own Python signatures.

    from ampis_b200 import analyze, structures, data_utils
    from ampis_b200.applications import powder

mirrors ``from ampis import analyze, structures, data_utils`` /
``from ampis.applications import powder`` for every function on that path (SURVEY.md
section 8a).  All mask arithmetic runs in hand-written CUDA kernels reached through the C ABI
of ``libampis_b200.so`` (include/ampis_b200.h); there is no CPU fallback.
"""
from . import containers
from . import engine
from . import structures
from . import analyze
from . import data_utils
from . import applications

__version__ = '0.1.0'
__all__ = ['analyze', 'data_utils', 'structures', 'applications', 'containers', 'engine', '__version__']
