"""Two (or more) B200s, one process per GPU over NCCL: the multi-GPU forms of the hot path
(mpc_bulletproof_b200.multi) checked against this rank's own single-GPU result, which the GPU
test-suite pins to the oracle.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/multi_gpu_check.py
Prints one JSON object on rank 0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mpc_bulletproof_b200 import Comb, Context, Table, multi  # noqa: E402
from mpc_bulletproof_b200 import protocol as P  # noqa: E402

BASE = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
L = P.L


def scalars(n, seed):
    import numpy as np

    g = np.random.Generator(np.random.PCG64(seed))
    a = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x0F
    return a.tobytes()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    comb = Comb(ctx, BASE)
    out = {"world": world}

    # 1. stride-sharded MSM, 2^16 points, 2 scalar sets
    n, sets = 1 << 16, 2
    pts = comb.mul(scalars(n, 1))  # every rank can regenerate all points; it keeps only its stride
    ks = [scalars(n, 10 + s) for s in range(sets)]
    t_loc = Table(ctx, multi.shard_bytes(pts, rank, world)).set_windows(0)
    eng = multi.CudaEngine(ctx, t_loc)
    loc_sc = b"".join(multi.shard_bytes(k, rank, world) for k in ks)
    got = multi.sharded_msm(eng, loc_sc, n_sets=sets)
    t_all = Table(ctx, pts).set_windows(0)
    want = t_all.msm(b"".join(ks), n_sets=sets)
    out["sharded_msm_equal"] = got == want
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        multi.sharded_msm(eng, loc_sc, n_sets=sets)
    out["sharded_msm_ms"] = (time.perf_counter() - t0) / 5 * 1e3

    # 1b. the same sum with the exchange fused into the combine kernel (peer-mapped buffers, no NCCL)
    px = multi.PeerExchange(ctx, max_sets=sets)
    part = eng.partial(loc_sc, sets)
    outb = torch.zeros(32 * sets, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    fused_ok = True
    for _ in range(4):  # both buffer parities, twice
        px.exchange_sum_encode(part.data_ptr(), sets, outb.data_ptr())
        ctx.sync()
        raw = bytes(outb.cpu().numpy().tobytes())
        fused_ok = fused_ok and [raw[32 * i : 32 * i + 32] for i in range(sets)] == want
    out["fused_exchange_equal"] = fused_ok and px.ok()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        px.exchange_sum_encode(part.data_ptr(), sets, outb.data_ptr())
    ctx.sync()
    out["fused_exchange_us"] = (time.perf_counter() - t0) / 50 * 1e6
    parts = torch.empty(world * part.numel(), dtype=part.dtype, device=dev)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        dist.all_gather_into_tensor(parts, part)
        torch.cuda.synchronize()
        from mpc_bulletproof_b200.api import dev_sum_encode
        dev_sum_encode(ctx, parts.data_ptr(), world, sets, outb.data_ptr())
        ctx.sync()
    out["nccl_allgather_plus_combine_us"] = (time.perf_counter() - t0) / 50 * 1e6
    dist.barrier()
    px.close()

    # 2. MPC open: party p = rank p holds additive shares of the scalars over ALL points
    x = ks[0]
    xs = [int.from_bytes(x[32 * i : 32 * i + 32], "little") for i in range(1024)]
    r0 = scalars(1024, 77)
    r0s = [int.from_bytes(r0[32 * i : 32 * i + 32], "little") for i in range(1024)]
    if rank == 0:
        share = r0s
    elif rank == 1:
        share = [(a - b) % L for a, b in zip(xs, r0s)]
    else:
        share = [0] * 1024
    t_small = Table(ctx, pts[: 32 * 1024]).set_windows(0)
    opened = multi.open_shares(multi.CudaEngine(ctx, t_small), b"".join(P.sc_bytes(v) for v in share))
    out["mpc_open_equal"] = opened == t_small.msm(x[: 32 * 1024])

    # 3. batch verification: 16 proofs, whole proofs per rank
    m = 1 << 10
    gens = P.Gens(ctx, pts[: 32 * m], comb.mul(scalars(m, 2)), BASE, comb.mul(scalars(1, 3)))
    nproofs = 16
    proofs = []
    for i in range(nproofs):  # every rank needs every proof's bytes; proving is cheap at this size
        p = P.Prover(gens, P.Transcript(b"multi"))
        p.square_chain(p.commit_public(100 + i), m)
        proofs.append(p.prove(500 + i))
    bad = bytearray(proofs[5])
    bad[33] ^= 2
    proofs[5] = bytes(bad)
    truth = [i != 5 for i in range(nproofs)]

    def verify_one(i):
        v = P.Verifier(gens, P.Transcript(b"multi"))
        v.square_chain(v.commit_public(100 + i), m)
        return P.batch_verify([(v, proofs[i])])[0]

    dist.barrier()
    t0 = time.perf_counter()
    res = multi.batch_verify_sharded(nproofs, verify_one, device=dev)
    out["batch_verify_ms"] = (time.perf_counter() - t0) * 1e3
    out["batch_verify_equal"] = res == truth
    ok = torch.tensor([int(all(v for k, v in out.items() if k.endswith("_equal")))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(ok.item())
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if out["all_ranks_ok"] else 1)


if __name__ == "__main__":
    main()
