"""hoig_b200 -- B200-native (sm_100a) implementation of HOGAN's generator inference hot path.

Public surface:
  * ``hoig_b200.generator.GeneratorB200`` / ``create``   drop-in for the reference ``Generator`` (boundary B1)
  * ``hoig_b200.compat``                                  call-compatible stand-ins for the reference's pybind ops (B2)
  * ``hoig_b200.renderer``                                batched rasterizer / condition-map API (B3)
  * ``hoig_b200.ops``                                     tensor wrappers over the C ABI in ``include/hoig_b200.h``
The compute path is hand-written CUDA in ``hoig_b200/csrc`` behind ``libhoig_b200.so``; it has no CPU fallback.
"""
__version__ = "0.1.0"
