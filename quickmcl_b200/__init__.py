"""quickmcl_b200 — B200-native Monte Carlo localisation core.

The particle-filter hot path of VorpalBlade/quickmcl as hand-written sm_100a
kernels behind the reference's own Map / IParticleFilter interface.  The CUDA
library (libqmcl_b200.so, C ABI in include/qmcl.h) is loaded lazily on first
use and there is no CPU fallback.
"""
from .api import (LikelihoodMapParameters, MapLikelihood, MapState, OccupancyGrid,  # noqa: F401
                  ParticleFilter, ParticleFilterParameters, ResampleType,
                  create_likelihood_map, create_particle_filter, kernel_launch_count,
                  make_particles, particles_array, transfer_bytes)

__version__ = "0.1.0"
