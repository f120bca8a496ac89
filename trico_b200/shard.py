"""Chunk sharding of one stream across the GPUs of a box.

A v1 stream is a list of independent chunks, so a large stream shards by contiguous, chunk-aligned
element ranges: rank r encodes elements [lo_r, hi_r) with its own kernels and there is NO
data-path collective.  Only two exchanges exist:
  1. the compressed byte count of every rank (one small all-gather) - every rank then knows its
     base offset inside the assembled stream;
  2. optionally the assembly itself: size tables and payloads are gathered to one rank.
Concatenating the ranks' size tables and payloads in rank order yields byte-for-byte the stream a
single GPU would have produced, because chunk boundaries and chunk contents do not depend on the
split.

Everything here is host-side plumbing over torch.distributed (NCCL between GPUs, gloo in the CPU
tests); the encode itself is passed in as a callable.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

V1_FIXED = 15
# A rank's piece is encoded as a stream of its own.  Index streams count triangles (3 indices per
# count), and a chunk-aligned range need not hold a whole number of triangles, so their pieces are
# encoded as plain integer lists of the same width: same codec, same chunks, only the header differs
# (and the header of a piece is discarded by the assembly).
PIECE_TYPE = {3: 19, 4: 20}


def plan_ranges(n_elements: int, log2_chunk: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous element ranges, one per rank, every boundary a multiple of the chunk size."""
    chunk = 1 << log2_chunk
    nchunks = (n_elements + chunk - 1) // chunk
    base, extra = divmod(nchunks, world)
    out, c = [], 0
    for r in range(world):
        take = base + (1 if r < extra else 0)
        lo, hi = min(c * chunk, n_elements), min((c + take) * chunk, n_elements)
        out.append((lo, hi))
        c += take
    return out


def stream_header(stream_type: int, count: int, codec_info: int, log2_chunk: int, payload_bytes: int) -> bytes:
    return (bytes([stream_type]) + int(count).to_bytes(4, "little") + bytes([codec_info, log2_chunk]) +
            int(payload_bytes).to_bytes(8, "little"))


def split_piece(piece: bytes, nchunks: int) -> Tuple[bytes, bytes]:
    """A rank's own v1 stream (header + sizes + payload) -> (size table, payload)."""
    total = int.from_bytes(piece[7:15], "little")
    sizes = piece[V1_FIXED:V1_FIXED + 2 * nchunks]
    payload = piece[V1_FIXED + 2 * nchunks:V1_FIXED + 2 * nchunks + total]
    assert len(payload) == total
    return sizes, payload


def exchange_sizes(dist, local_payload_bytes: int, device=None):
    """Exchange 1: every rank's payload byte count -> (list of counts, this rank's base offset)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([local_payload_bytes], dtype=torch.int64, device=device)
    allc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine)
    counts = [int(x) for x in allc.cpu()]
    return counts, sum(counts[:rank])


def assemble_stream(dist, stream_type: int, total_count: int, codec_info: int, log2_chunk: int,
                    sizes: bytes, payload: bytes, dst: int = 0, device=None):
    """Exchange 2: gather every rank's size table and payload on `dst`; returns the assembled v1
    stream there (None elsewhere).  Variable-size pieces travel as padded uint8 tensors."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = torch.tensor([len(sizes), len(payload)], dtype=torch.int64, device=device)
    metas = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.cpu().reshape(world, 2)
    cap = int((metas[:, 0] + metas[:, 1]).max())
    buf = torch.zeros(max(cap, 1), dtype=torch.uint8, device=device)
    blob = np.frombuffer(sizes + payload, np.uint8)
    if blob.size:
        buf[:blob.size] = torch.from_numpy(blob.copy()).to(buf.device)
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gathered, dst=dst)
    if rank != dst:
        return None
    tabs, pays = [], []
    for r in range(world):
        ns, npay = int(metas[r, 0]), int(metas[r, 1])
        raw = gathered[r].cpu().numpy().tobytes()
        tabs.append(raw[:ns])
        pays.append(raw[ns:ns + npay])
    total = sum(len(p) for p in pays)
    return stream_header(stream_type, total_count, codec_info, log2_chunk, total) + b"".join(tabs) + b"".join(pays)


def encode_sharded(dist, stream_type: int, total_count: int, per_count: int, log2_chunk: int,
                   encode_range: Callable[[int, int], bytes], nsub: int, dst: int = 0, device=None, assemble: bool = True):
    """Encode one stream across all ranks.

    encode_range(lo, hi) encodes elements [lo, hi) (in units of the stream's counted elements times
    per_count, i.e. scalars per component / plane) as this rank's own v1 stream and returns its bytes.
    nsub = components (FPC) or byte planes (LZ4) per chunk range.
    Returns (assembled stream on dst or None, counts of all ranks, this rank's base offset)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = total_count * per_count
    lo, hi = plan_ranges(n, log2_chunk, world)[rank]
    chunk = 1 << log2_chunk
    nch = ((hi - lo + chunk - 1) // chunk) * nsub
    piece = encode_range(lo, hi) if hi > lo else stream_header(stream_type, 0, 0, log2_chunk, 0)
    codec_info = piece[5]
    sizes, payload = split_piece(piece, nch)
    counts, base = exchange_sizes(dist, len(payload), device)
    # the codec_info byte is the same on every rank that encoded something; agree on it
    import torch
    ci = torch.tensor([codec_info if hi > lo else 0], dtype=torch.int64, device=device)
    dist.all_reduce(ci, op=dist.ReduceOp.MAX)
    stream = assemble_stream(dist, stream_type, total_count, int(ci), log2_chunk, sizes, payload, dst, device) if assemble else None
    return stream, counts, base
