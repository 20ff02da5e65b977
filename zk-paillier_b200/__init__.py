"""b200-zk-paillier: batched Paillier zero-knowledge proofs on B200 (sm_100a).

Layout
  csrc/            CUDA kernels + the C ABI (include/zkp_b200.h) -> libzkp_b200.so
  host/            C++ mirror of the reference's `zkproofs::*` call surface and serde wire format
                   (RangeProofNi, NiCorrectKeyProof, ZeroProof, ...; batched) -> libzkp_host.so
  native.py        ctypes binding of the C ABI (numpy limb arrays in / out); what the tests and bench.py drive
  workload.py      seeded synthetic statements
  sharding.py      one process per GPU: key broadcast, contiguous shards, gather of proof bytes

The CUDA library is the only compute path: importing `native` without the built
library, or creating a context without a GPU, raises.
"""
from . import native  # noqa: F401

__all__ = ["native"]
