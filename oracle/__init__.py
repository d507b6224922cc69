"""CPU oracle for the preconditioner hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline that is timed beside the GPU path.  ``precondition_b200`` never imports
it and has no CPU fallback.

Contents
--------
``numerics.py``   numpy restatement (fp32 by default, fp64 twin via ``dtype=``) of
                  the reference's numerical kernels: ``power_iteration``,
                  ``mat_power``, ``matrix_inverse_pth_root``, ``QuantizedValue``,
                  Gram / frequent-directions statistics, FD sketch root,
                  low-rank root.  Every function cites the reference lines
                  (``DS`` = precondition/distributed_shampoo.py, ``QU`` =
                  precondition/quantization_utils.py) that it follows.
``optimizer.py``  numpy restatement of the optax-style transformation
                  (blocking, statistics, preconditioner scheduling, grafting,
                  momentum) -- ``distributed_shampoo(...).init/update``.
``jax_shim/``     a numpy-backed stand-in for the tiny part of jax / flax / chex /
                  optax the reference imports, used ONLY by ``gen_golden.py`` in the
                  build container to execute the UNMODIFIED reference sources
                  from ``/root/reference`` and record golden vectors into
                  ``tests/golden/``.  It never travels into product code.

Parity status: PINNED -- (1) against the known answers of the reference's own
unit tests (DST = precondition/distributed_shampoo_test.py, re-expressed in
``tests/test_oracle_*.py``), and (2) against golden vectors produced by running
the unmodified reference sources on the numpy shim (``oracle/gen_golden.py``;
array backend is numpy instead of XLA because JAX is not installable here, the
control flow and arithmetic order are the reference's own).
"""
