"""br2 -- B200-native drop-in for the rod-integration hot path of BR2-simulator.

Same package / class / method names as the reference (``br2.environment.Environment``,
``br2.free_simulator.FreeAssembly`` / ``BR2Simulator`` / ``FreeCallback``,
``br2.free_custom_systems``, ``br2.custom_activation``, ``br2.custom_constraint``); the PyElastica
system-collection hooks the reference imports from ``elastica`` live in ``br2.hooks``.
Integration happens only in the sm_100a CUDA kernels of ``lib/libbr2cuda.so``.
"""
from .version import VERSION  # noqa: F401
