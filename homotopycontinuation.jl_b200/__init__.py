"""B200-native batched path tracker behind HomotopyContinuation.jl's track/solve interface.

The directory name carries a dot, so import it through the root-level ``hcb200`` shim."""
from . import modelkit, systems, capi, start_systems, polyhedral, flops, workloads, sharding  # noqa: F401
