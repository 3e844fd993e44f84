"""causaldiffae_b200 — B200-native implementation of the CausalDiffAE denoising hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every hot op is a hand-written
sm_100a CUDA kernel in libcdae.so reached through the C ABI of include/cdae.h.  There is no CPU fallback:
importing is allowed anywhere (so tooling can inspect the package), but any compute call raises without
libcdae.so and an sm_100 device.
"""
__version__ = "0.1.0"
