"""quokka_b200 -- B200-native (sm_100a) implementation of Quokka's hydro / radiation sweep hot path.

The product is csrc/libquokka_b200.so (hand-written CUDA behind the C ABI of include/quokka_b200.h).
This package is the thin host side: ctypes bindings (capi), problem set-ups (problems) and the
time-step driver that mirrors QuokkaSimulation's level advance (driver).  No CPU fallback exists.
"""
__version__ = "0.1.0"
