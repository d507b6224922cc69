"""precondition_b200: B200-native (sm_100a) preconditioner hot path of
google-research/precondition's ``distributed_shampoo``.

Layout (only what the hot path needs):
  csrc/                    CUDA kernels + the C ABI (include/precond_b200.h)
  _lib.py                  ctypes binding of libprecond_b200.so
  ops.py                   operator-level wrappers (torch tensors = device memory)
  quantization_utils.py    QuantizedValue mirror (QU:25-113)
  distributed_shampoo.py   host-side mirror of the reference's optax-style API
"""
__version__ = "0.1.0"
