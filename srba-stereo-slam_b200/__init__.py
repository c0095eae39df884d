"""B200-native stereo front-end for SRBA-Stereo-SLAM: Python view of the C-ABI library.

The product is `libsrba_b200.so` (CUDA, sm_100a) + the C++ adapter classes under host/.  This
package only binds the C ABI with ctypes for the test-suite and the benchmark; it contains no
compute and no CPU fallback: importing `frontend` fails loudly when the library is missing.
"""
__version__ = "0.1.0"
