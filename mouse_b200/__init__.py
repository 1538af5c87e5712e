"""mouse_b200 - B200-native local orientational-ordering hot path of mglagolev/mouse.

Host code is Python (the reference's own language); the arithmetic is hand-written sm_100a CUDA
reached through the C-ABI of ``libmouse_b200.so`` (``include/mouse_b200.h``).  No CPU fallback."""
__version__ = "0.1.0"
