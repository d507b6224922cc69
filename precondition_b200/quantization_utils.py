"""QuantizedValue mirror (reference: precondition/quantization_utils.py:25-113).

Same fields and methods as the reference dataclass; arrays are CUDA torch tensors
and (de)quantisation runs in the library's kernels (``pc_quantize_batched`` /
``pc_dequantize_batched``).  int16 / int8 use per-column symmetric buckets with
optional diagonal extraction; bfloat16 is a plain cast; float32 is a no-op.
"""
from __future__ import annotations

import dataclasses
from typing import Any

import torch

from precondition_b200 import ops


@dataclasses.dataclass
class QuantizedValue:
  """State associated with quantized value (QU:25-35)."""
  quantized: Any
  diagonal: Any      # diagonal (if extract_diagonal is set)
  bucket_size: Any
  quantized_dtype: torch.dtype
  extract_diagonal: bool
  shape: Any

  @classmethod
  def from_float_value(cls, fvalue, quantized_dtype, extract_diagonal=False):
    """QU:37-45."""
    if isinstance(fvalue, list) and not fvalue:
      return cls([], [], [], quantized_dtype, extract_diagonal, [])
    quantized, diagonal, bucket = cls.quantize(fvalue, quantized_dtype, extract_diagonal)
    return cls(quantized, diagonal, bucket, quantized_dtype, extract_diagonal,
               list(quantized.shape))

  @classmethod
  def quantize(cls, fvalue, quantized_dtype, extract_diagonal=False):
    """Returns quantized value and the bucket (QU:49-95)."""
    if quantized_dtype == torch.float32:
      return fvalue, [], []
    if quantized_dtype == torch.bfloat16:
      q, _, _ = ops.quantize(_as_matrix(fvalue), torch.bfloat16)
      return q.reshape(fvalue.shape), [], []
    if quantized_dtype not in (torch.int8, torch.int16):
      raise ValueError(f"Quantized dtype {quantized_dtype} not supported.")
    if extract_diagonal and fvalue.dim() != 2:
      raise ValueError(
          f"Input array {fvalue} must be 2D to work with extract_diagonal.")
    if fvalue.dim() < 1:
      raise ValueError(
          f"Input array {fvalue} must have a strictly positive number of dimensions.")
    mat = _as_matrix(fvalue)  # max over axis 0 == per-column max of [d0, rest]
    q, diag, bucket = ops.quantize(mat, quantized_dtype, extract_diagonal)
    return (q.reshape(fvalue.shape), diag if extract_diagonal else [],
            bucket.reshape(fvalue.shape[1:]))

  def to_float(self):
    """Returns the float value (QU:97-113)."""
    if isinstance(self.quantized, list) and not self.quantized:
      return self.quantized
    if self.quantized_dtype == torch.float32:
      return self.quantized
    q = _as_matrix(self.quantized)
    if self.quantized_dtype == torch.bfloat16:
      return ops.dequantize(q, None, None).reshape(self.quantized.shape)
    bucket = self.bucket_size.reshape(-1)
    diag = self.diagonal if self.extract_diagonal else None
    return ops.dequantize(q, diag, bucket, self.extract_diagonal).reshape(self.quantized.shape)


def _as_matrix(x: torch.Tensor) -> torch.Tensor:
  if x.dim() == 1:
    return x.reshape(x.shape[0], 1).contiguous()
  return x.reshape(x.shape[0], -1).contiguous()
