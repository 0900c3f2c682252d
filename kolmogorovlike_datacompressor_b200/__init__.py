"""B200-native (sm_100a) per-block transform encode/decode hot path of KolmogorovLike-DataCompressor.

Drop-in modules (same names / signatures / error behaviour as the reference):
    kolm_final                     -> compress / decompress                ('KOLM' container)
    kolm_final_researched_v2_2     -> compress_blocks_fixed / _cdc / decompress ('KOLR' container)
Stage operators on torch CUDA tensors: `stages`.  Everything routes through the C-ABI library
libkolm_b200.so (include/kolm_abi.h) via ctypes; there is no CPU fallback.
"""
__version__ = "0.1.0"
