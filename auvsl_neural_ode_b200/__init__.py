"""auvsl_neural_ode_b200 — B200-native batched differentiable Jackal rollouts (fp64).

The product is ``libavsim.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/avsim.h``).
This package is the thin Python host layer over it: a ctypes binding (``capi``), the synthetic
workload generator used by the benchmarks (``synth``) and a data-parallel training step that
mirrors the reference ``Trainer`` semantics over ``torch.distributed`` (``trainer``).
There is no CPU fallback: importing ``capi`` raises if the shared library has not been built.
"""
from . import capi  # noqa: F401
from .capi import Context, AvsError  # noqa: F401

__all__ = ["capi", "Context", "AvsError"]
