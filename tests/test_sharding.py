"""CPU suite: the N>1 plumbing (key broadcast, shard bounds, ragged gather of per-proof records) on gloo with
world_size 2.  The per-shard compute stand-in here is the CPU oracle (tests only); on the GPU box the same
functions run over NCCL around the CUDA path (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT, c_oracle, keys, po
from zk_paillier_b200 import sharding, workload
from zk_paillier_b200.native import to_limbs


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    p, qq = keys(1024)[0]
    n = p * qq
    n_arr = sharding.broadcast_key(to_limbs(n, 32) if rank == 0 else np.zeros(1, np.uint32), dev)
    assert int.from_bytes(n_arr.tobytes(), "little") == n
    # every rank derives the same global workload and works on its own contiguous shard
    work = workload.rangeproof_batch(n, total, ef=4, seed=9, reject_every=3)
    lo, hi = sharding.shard_bounds(total, world, rank)
    sl = slice(lo, hi)
    pr = c_oracle.rangeproof_ni_prove(n_arr, 4, work["range"][sl], work["x"][sl], work["r"][sl], work["w1"][sl], work["swap"][sl],
                                      work["r1"][sl], work["r2"][sl], 1)
    cx = c_oracle.paillier_enc(n_arr, work["x_n"][sl], work["r"][sl], 1)
    acc, fault, dig, _ = c_oracle.rangeproof_ni_verify(n_arr, 4, work["range"][sl], cx, pr["c1"], pr["c2"], pr["kind"], pr["resp_w"],
                                                       pr["resp_r"], 1)
    rec = sharding.gather_records(np.concatenate([acc[:, None], dig], axis=1), dev)
    if rank == 0:
        q.put(rec)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    total, world = 7, 2      # ragged: 4 + 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    rec = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    p_, q_ = keys(1024)[0]
    n = p_ * q_
    n_arr = to_limbs(n, 32)
    work = workload.rangeproof_batch(n, total, ef=4, seed=9, reject_every=3)
    pr = c_oracle.rangeproof_ni_prove(n_arr, 4, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"], 2)
    cx = c_oracle.paillier_enc(n_arr, work["x_n"], work["r"], 2)
    acc, fault, dig, _ = c_oracle.rangeproof_ni_verify(n_arr, 4, work["range"], cx, pr["c1"], pr["c2"], pr["kind"], pr["resp_w"], pr["resp_r"], 2)
    assert rec.shape == (total, 33)
    assert np.array_equal(rec[:, 0], acc) and np.array_equal(rec[:, 1:], dig)
    assert acc.tolist() == [1, 1, 0, 1, 1, 0, 1]


def _worker_proofs(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    local = 3                                           # equal shards (weak scaling)
    c1 = rng.integers(0, 2**32, (local, 4, 8), dtype=np.uint32)
    kind = rng.integers(0, 3, (local, 4), dtype=np.uint8)
    resp = rng.integers(0, 2**32, (local, 4, 2, 3), dtype=np.uint32)
    out = sharding.gather_proof_bytes([c1, kind, resp], torch.device("cpu"))
    if rank == 0:
        q.put(out.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_gather_of_proof_bytes_two_ranks():
    """The final gather of the proofs themselves (bench.py, N > 1): rank-major, one byte record per proof."""
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_proofs, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == (2, 3, 4 * 8 * 4 + 4 + 4 * 2 * 3 * 4)
    for r in range(world):
        rng = np.random.default_rng(100 + r)
        c1 = rng.integers(0, 2**32, (3, 4, 8), dtype=np.uint32)
        kind = rng.integers(0, 3, (3, 4), dtype=np.uint8)
        resp = rng.integers(0, 2**32, (3, 4, 2, 3), dtype=np.uint32)
        want = np.concatenate([c1.view(np.uint8).reshape(3, -1), kind.reshape(3, -1), resp.view(np.uint8).reshape(3, -1)], axis=1)
        assert np.array_equal(got[r], want)
