"""Multi-GPU plumbing: one process per GPU, independent proofs sharded by index.

The path has NO exchange step (proof i depends only on key, statement_i, randomness_i), so the only
collectives are one broadcast of the public key before the first launch and one gather of fixed-size
per-proof records (verdict byte + 32-byte challenge hash, or whole proof rows) after the last kernel
(SURVEY.md section 8e).  Works on any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in
the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous block partition: ranks get ceil/floor(total/world) items, earlier ranks the larger blocks."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_key(n_limbs_arr, device, src: int = 0):
    """Rank `src` owns the Paillier modulus n (uint32 limbs); every rank returns it after one broadcast."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return np.ascontiguousarray(n_limbs_arr, dtype=np.uint32)
    meta = torch.zeros(1, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        meta[0] = int(np.asarray(n_limbs_arr).shape[-1])
    dist.broadcast(meta, src)
    t = torch.zeros(int(meta.item()), dtype=torch.int32, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(n_limbs_arr, dtype=np.uint32).view(np.int32)))
    dist.broadcast(t, src)
    return t.cpu().numpy().view(np.uint32)


def gather_records(records: np.ndarray, device, counts=None):
    """all_gather of per-proof byte records [local, width] (uint8) -> [total, width] in rank order.
    Shards may be ragged: every rank pads to the largest shard, `counts` (per-rank sizes) trims."""
    records = np.ascontiguousarray(records, dtype=np.uint8)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return records
    if counts is None:
        c = torch.tensor([records.shape[0]], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(allc, c)
        counts = [int(x.item()) for x in allc]
    width = records.shape[1]
    mx = max(counts)
    pad = torch.zeros((mx, width), dtype=torch.uint8, device=device)
    pad[: records.shape[0]] = torch.from_numpy(records).to(device)
    out = torch.empty((world, mx, width), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, pad) if device.type == "cuda" else dist.all_gather(list(out.unbind(0)), pad)
    out = out.cpu().numpy()
    return np.concatenate([out[r, : counts[r]] for r in range(world)], axis=0)


def gather_proof_bytes(arrays, device):
    """Final gather of the proofs themselves over NVLink (BASELINE.json north_star): every rank's proof rows -- the host
    arrays of one shard, each [local, ...] -- are staged into one device record per proof and all-gathered on the device
    (NCCL has no gather; the root is whoever reads rank-major slice [r]).  Equal shards (weak scaling).  Returns the device
    tensor [world, local, bytes_per_proof]; nothing is copied back to the host."""
    local = arrays[0].shape[0]
    flats = [np.ascontiguousarray(a).view(np.uint8).reshape(local, -1) for a in arrays]
    width = sum(f.shape[1] for f in flats)
    rec = torch.empty((local, width), dtype=torch.uint8, device=device)
    off = 0
    for f in flats:
        rec[:, off:off + f.shape[1]].copy_(torch.from_numpy(f), non_blocking=True)
        off += f.shape[1]
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return rec.unsqueeze(0)
    out = torch.empty((world, local, width), dtype=torch.uint8, device=device)
    if device.type == "cuda":
        dist.all_gather_into_tensor(out, rec)
    else:
        dist.all_gather(list(out.unbind(0)), rec)
    return out

