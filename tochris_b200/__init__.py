"""ref_cuda: B200-native batched replacement for the ref_soft rasteriser of viciious/tochris.
Product = tochris_b200/libref_cuda.so (C-ABI in include/ref_cuda.h); this package holds its
sources (csrc/), the ctypes binding (api.py) and the procedural scene generators (scenes/)."""
