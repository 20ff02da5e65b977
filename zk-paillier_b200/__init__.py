"""b200-zk-paillier: batched Paillier zero-knowledge proofs on B200 (sm_100a).

Layout
  csrc/            CUDA kernels + the C ABI (include/zkp_b200.h) -> libzkp_b200.so
  native.py        ctypes binding of the C ABI (numpy limb arrays in / out)
  zkproofs.py      host-side mirror of the reference's `zkproofs::*` call surface
                   (RangeProofNi, NiCorrectKeyProof, ZeroProof, ...), batched
  serialize.py     serde wire codec (decimal-string BigInts) of src/serialize.rs

The CUDA library is the only compute path: importing `native` without the built
library, or creating a context without a GPU, raises.
"""
from . import native  # noqa: F401

__all__ = ["native"]
