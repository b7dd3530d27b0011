"""simpledet_b200 — B200-native (sm_100a) detection hot path behind SimpleDet's operator names.

Layout:  csrc/ (CUDA kernels + C ABI) -> libsimpledet_b200.so;  _lib.py (ctypes binding);
ops.py (host-side mirror of the reference operator interface);  synth.py (seeded synthetic
workloads shared by tests and bench.py).
"""
__version__ = "0.1.0"
