"""varigraph_b200 -- varigraph's read k-mer counting hot path, B200-native (sm_100a).

The product is `libvgb200.so` (C ABI in include/vgb200.h; CUDA in csrc/) plus the C++ host
classes in host/ that keep the reference's FastqKmerKernel / BloomFilterKernel surface.
`varigraph_b200.capi` is the ctypes binding used by tests and bench.py; importing it fails
loudly when the shared library has not been built -- there is no CPU or PyTorch fallback.
`varigraph_b200.synth` generates the seeded synthetic workloads.
"""
__all__ = ["capi", "synth", "build"]
__version__ = "0.1.0"
