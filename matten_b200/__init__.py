"""
matten_b200 -- B200-native (sm_100a) implementation of the equivariant message-passing
hot path of MatTen (wengroup/matten): hand-written CUDA kernels behind a C ABI
(include/matten_b200.h), bound with ctypes and exposed through modules that mirror the
reference's ``matten.nn`` / ``matten.model_factory`` API.
"""
ABI_VERSION = 2
__version__ = "0.1.0"
