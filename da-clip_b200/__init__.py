"""daclip_b200 - B200-native (sm_100a) implementation of DA-CLIP's universal-restoration inference path:
IR-SDE reverse sampling -> ConditionalUNet denoiser -> DaCLIP.encode_image(control=True) conditioning.
Python keeps the reference's model / sampler API; all arithmetic runs in hand-written CUDA kernels behind the
C ABI in include/dac_b200.h (libdac_b200.so).  No CPU fallback."""
__version__ = "0.1.0"
