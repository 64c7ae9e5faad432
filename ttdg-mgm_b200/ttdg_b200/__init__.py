"""ttdg_b200 - B200-native runtime for the TTDG-MGM test-time-adaptation hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed); every device op is a
hand-written sm_100a kernel reached through the C-ABI library ``libttdg_sm100.so``
(``include/ttdg_b200.h``).  There is no CPU fallback: importing an op without the built library
raises.
"""
__version__ = "0.1.0"
