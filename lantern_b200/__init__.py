"""lantern_b200 — B200-native verification step for LANTERN relaxed speculative decoding.

Host side mirrors the reference's Python call surface; the hot path is
hand-written sm_100a CUDA behind the C ABI declared in ``include/lantern_b200.h``.
"""
__version__ = "0.1.0"
