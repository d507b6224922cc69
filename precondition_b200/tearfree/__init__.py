"""tearfree front-end (precondition/tearfree in the reference) over the CUDA kernels of this
package: ``optimizer.tearfree`` and its parts (``grafting``, ``momentum``, ``second_order``,
``shampoo``, ``sketchy``, ``reshaper``, ``praxis_shim``) with the reference's option dataclasses and the
optax-style ``init`` / ``update`` protocol on pytrees of CUDA tensors."""
from precondition_b200.tearfree import grafting, momentum, optimizer, praxis_shim, reshaper
from precondition_b200.tearfree import second_order, shampoo, sketchy

__all__ = ["grafting", "momentum", "optimizer", "praxis_shim", "reshaper", "second_order",
           "shampoo", "sketchy"]
