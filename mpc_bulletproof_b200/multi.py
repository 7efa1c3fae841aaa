"""Multi-GPU forms of the hot path, one process per GPU over `torch.distributed`.

* A large MSM (reference call sites of `StarkPoint::msm_iter`, e.g. src/r1cs/verifier.rs:516-547)
  shards by STRIDE: term i lives on rank i mod world.  Every rank runs the whole Pippenger
  pipeline on its slice and owns one extended point per output; the only exchange is one
  all-gather of n_sets x 128 bytes per rank, after which every rank adds the partials and encodes.
* Batch verification of many proofs (BASELINE.json config 4) shards WHOLE PROOFS by stride; there
  is no data-path collective, only the gather of one accept/reject byte per proof.
* The 2-party MPC prover (reference src/r1cs_mpc/mpc_prover.rs:621-657) maps party p to rank p:
  a party's additive share of the scalars gives its additive share of each commitment, so
  "open" is the same exchange-and-add as the sharded MSM.  The party-to-party link of the
  reference is QUIC and stays untouched; here the process group stands in for it.

The arithmetic is behind an `engine` with three methods, so that the same sharding and exchange
logic runs under `nccl` with the CUDA engine and, in the CPU test-suite (`gloo`, world_size 2),
with a checker engine supplied by the tests:

    engine.partial(scalars: bytes, n_sets: int) -> torch.Tensor  # int32[n_sets*32], this rank's sums
    engine.combine(parts: torch.Tensor, n_parts: int, n_sets: int) -> list[bytes]
    engine.device                                                  # where the tensors live
"""
from __future__ import annotations

import ctypes
from typing import Callable, Sequence


def shard_indices(n: int, rank: int, world: int) -> range:
    """Indices owned by `rank` under the stride partition (i mod world == rank)."""
    return range(rank, n, world)


def shard_bytes(buf: bytes, rank: int, world: int, item: int = 32) -> bytes:
    """The items of `buf` (32-byte scalars or points) that `rank` owns."""
    n = len(buf) // item
    return b"".join(buf[item * i : item * i + item] for i in shard_indices(n, rank, world))


class CudaEngine:
    """This rank's slice of the points as a resident table on its GPU."""

    def __init__(self, ctx, table):
        import torch

        self.ctx, self.table = ctx, table
        self.device = torch.device("cuda", ctx.device)
        self._torch = torch

    def partial(self, scalars: bytes, n_sets: int = 1):
        torch = self._torch
        raw = self.table.msm_partial(scalars, n_sets)
        return torch.frombuffer(bytearray(raw), dtype=torch.int32).to(self.device)

    def partial_dev(self, d_scalars: int, n_sets: int, out):
        """Device-resident scalars (int32[n*8] per set); `out` int32[n_sets*32] on this GPU."""
        self.table.dev_msm(d_scalars, n_sets, out.data_ptr())
        return out

    def combine(self, parts, n_parts: int, n_sets: int = 1):
        from .api import dev_sum_encode

        torch = self._torch
        out = torch.empty(32 * n_sets, dtype=torch.uint8, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        # the context's stream must see the collective's result: order it after the current stream
        self.ctx.sync()
        cur.synchronize()
        dev_sum_encode(self.ctx, parts.data_ptr(), n_parts, n_sets, out.data_ptr())
        self.ctx.sync()
        raw = bytes(out.cpu().numpy().tobytes())
        return [raw[32 * i : 32 * i + 32] for i in range(n_sets)]


class PeerExchange:
    """The exchange fused with the combine: ONE kernel per rank stores its partial sums into every
    rank's peer-mapped buffer over NVLink, waits for all flags, adds and encodes
    (`bpg_dev_exchange_sum_encode`), in place of an NCCL all-gather followed by a combine launch.
    One process per GPU on one node; the 64-byte IPC handles travel once over the process group."""

    def __init__(self, ctx, max_sets: int = 4, group=None):
        import torch
        import torch.distributed as dist

        from ._lib import check, lib

        self.ctx = ctx
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._h = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        check(lib().bpg_peer_create(ctx._h, self.world, self.rank, max_sets, ctypes.byref(self._h), handle))
        dev = torch.device("cuda", ctx.device)
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(dev)
        if self.world > 1:
            allh = torch.empty(64 * self.world, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
        else:
            allh = mine
        check(lib().bpg_peer_connect(self._h, bytes(allh.cpu().numpy().tobytes())))
        if self.world > 1:
            dist.barrier(group=group)  # every rank has mapped every buffer before the first exchange

    def exchange_sum_encode(self, d_part: int, n_sets: int, d_out_bytes: int | None, d_out_ext: int | None = None):
        """Enqueue on the context's stream; d_part: this rank's n_sets x 128-byte partial sums."""
        from ._lib import check, lib

        check(
            lib().bpg_dev_exchange_sum_encode(
                self.ctx._h, self._h, ctypes.c_void_p(d_part), n_sets, ctypes.c_void_p(d_out_bytes or 0), ctypes.c_void_p(d_out_ext or 0)
            )
        )

    def ok(self) -> bool:
        from ._lib import check, lib

        st = ctypes.c_int()
        check(lib().bpg_peer_status(self._h, ctypes.byref(st)))
        return st.value == 0

    def close(self):
        from ._lib import lib

        if self._h:
            lib().bpg_peer_free(self._h)
            self._h = ctypes.c_void_p()


def allgather_combine(engine, part, n_sets: int = 1, group=None) -> list[bytes]:
    """Exchange the ranks' partial sums (n_sets x 128 bytes each) and add them: every rank
    returns the same n_sets encodings.  The one collective of a sharded MSM / an MPC open."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return engine.combine(part, 1, n_sets)
    parts = torch.empty(world * part.numel(), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(parts, part.contiguous(), group=group)
    return engine.combine(parts, world, n_sets)


def sharded_msm(engine, local_scalars: bytes, n_sets: int = 1, group=None) -> list[bytes]:
    """sum_i k_i P_i over all ranks' slices; `local_scalars` are this rank's (stride-sharded,
    set-major) scalars for the points its engine holds."""
    return allgather_combine(engine, engine.partial(local_scalars, n_sets), n_sets, group)


def open_shares(engine, share_scalars: bytes, n_sets: int = 1, group=None) -> list[bytes]:
    """MPC open of commitments: every party holds ALL points and an additive share of every
    scalar (or of its MAC); the opened commitments are the sum of the parties' partial sums."""
    return allgather_combine(engine, engine.partial(share_scalars, n_sets), n_sets, group)


def batch_verify_sharded(n_proofs: int, verify_one: Callable[[int], bool], group=None, device=None) -> list[bool]:
    """Per-proof results for proofs 0..n_proofs-1 with proof i verified by rank i mod world
    (`verify_one(i)` -> accept?).  No data-path collective: only the result bytes are gathered."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return [bool(verify_one(i)) for i in range(n_proofs)]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_proofs + world - 1) // world
    mine = torch.zeros(per, dtype=torch.uint8)
    for slot, i in enumerate(shard_indices(n_proofs, rank, world)):
        mine[slot] = 1 if verify_one(i) else 0
    if device is not None:
        mine = mine.to(device)
    allr = torch.empty(world * per, dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(allr, mine, group=group)
    allr = allr.cpu().view(world, per)
    return [bool(allr[i % world, i // world]) for i in range(n_proofs)]


def gather_results(local: Sequence[bool], n_total: int, group=None, device=None) -> list[bool]:
    """Gather stride-sharded per-item booleans (rank r holds items r, r+world, ...)."""
    it = iter(local)
    return batch_verify_sharded(n_total, lambda _i: next(it), group, device)
