"""B200-native Groth16 (BLS12-381) proving backend: hand-written sm_100a CUDA
behind a C ABI (include/b200zk.h) and a thin host mirror of the arkworks surface
the reference's prove routes use.  Import never touches the GPU; the first
Context() does, and raises if the library or a device is missing."""
from . import _ffi, codec, witness                          # noqa: F401
from .groth16 import (ConstraintMatrices, Context, DistributedProver, FixedBase, Groth16, LibsnarkReduction, PolynomialDegreeTooLarge,  # noqa: F401
                      ProvingKey, Radix2EvaluationDomain, SynthesisError, VariableBaseMSM, VerifyingKey)
