"""numpy restatement of precondition/sm3.py (TEST INFRASTRUCTURE -- never imported by the
product).  SM3 (Anil, Gupta, Koren, Singer 2019): per-axis diagonal accumulators, the
second-moment estimate of an entry is the minimum of its axes' accumulators; int8 momentum
(reference: /root/reference/precondition/sm3.py:40-168, "SM3" below)."""
from __future__ import annotations

import functools
from typing import NamedTuple

import numpy as np

from oracle import numerics as N


class ParameterStats(NamedTuple):  # SM3:33-37
  diagonal_statistics: list
  diagonal_momentum: N.QuantizedValue


class SM3State(NamedTuple):
  count: int
  stats: list


class GradientTransformation(NamedTuple):
  init: callable
  update: callable


def sm3(learning_rate, beta1=0.9, beta2=0.999, diagonal_epsilon=1e-10, weight_decay=0.0,
        normalize_grads=False):
  f = np.float32

  def init_fn(params):  # SM3:71-80
    stats = []
    for p in params:
      acc = [np.zeros([s], np.float32) for s in p.shape]
      mom = N.QuantizedValue.from_float_value(np.zeros_like(p, dtype=np.float32), np.int8)
      stats.append(ParameterStats(acc, mom))
    return SM3State(0, stats)

  def update_fn(updates, state, params):
    new_updates, new_stats = [], []
    lr = learning_rate(state.count) if callable(learning_rate) else learning_rate
    for g, st, p in zip(updates, state.stats, params):
      g = np.asarray(g, np.float32)
      if normalize_grads:  # SM3:113-115
        g = g / (np.linalg.norm(g).astype(np.float32) + f(1e-16))
      w2 = f(1.0 - beta2) if beta2 != 1.0 else f(1.0)
      if g.ndim < 2:  # SM3:88-94
        nu = f(beta2) * st.diagonal_statistics[0] + w2 * g**2
      else:
        expanded = [st.diagonal_statistics[i].reshape(
            [1] * i + [g.shape[i]] + [1] * (g.ndim - i - 1)) for i in range(g.ndim)]
        nu = f(beta2) * functools.reduce(np.minimum, expanded) + w2 * g**2
      pre = f(1.0) / np.sqrt(nu + f(diagonal_epsilon))  # SM3:133-134
      pg = g * pre
      w1 = f(1.0 - beta1) if beta1 != 1.0 else f(1.0)
      mom = f(beta1) * st.diagonal_momentum.to_float() + w1 * pg  # SM3:96-98
      acc = []  # SM3:100-109
      for i in range(g.ndim):
        axes = tuple(a for a in range(g.ndim) if a != i)
        acc.append(np.max(nu, axis=axes) if axes else nu)
      if g.ndim == 1:
        acc[0] = nu
      new_stats.append(ParameterStats(acc, N.QuantizedValue.from_float_value(mom, np.int8)))
      out = mom + f(weight_decay) * p if weight_decay > 0.0 else mom  # SM3:154-158
      new_updates.append((-f(lr) * out).astype(np.float32))
    return new_updates, SM3State(state.count + 1, new_stats)

  return GradientTransformation(init_fn, update_fn)
