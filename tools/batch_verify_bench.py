"""BASELINE.json config 4, second half: batch verification of many R1CS proofs, whole proofs per GPU
(no data-path collective), per-proof accept/reject as in the reference (`Verifier::verify`,
src/r1cs/verifier.rs:393).  One process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29512 tools/batch_verify_bench.py [n_proofs] [lg_multipliers]
(or bare for one GPU).  Proofs are of the reference's bench circuit with different public inputs;
one of them is tampered.  Prints one JSON object on rank 0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mpc_bulletproof_b200 import Comb, Context, multi  # noqa: E402
from mpc_bulletproof_b200 import protocol as P  # noqa: E402

BASE = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")


def scalars(n, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    a = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x0F
    return a.tobytes()


def main():
    n_proofs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    n_ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 1  # contexts (host threads) per GPU
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    comb = Comb(ctx, BASE)
    n = 1 << lg
    gens = P.Gens(ctx, comb.mul(scalars(n, 1)), comb.mul(scalars(n, 2)), BASE, comb.mul(scalars(1, 3)))
    # every rank needs the bytes of the proofs it verifies: it proves them itself here (not timed)
    mine = list(multi.shard_indices(n_proofs, rank, world))
    proofs = {}
    t0 = time.perf_counter()
    for i in mine:
        p = P.Prover(gens, P.Transcript(b"batch"))
        p.square_chain(p.commit_public(1000 + i), n)
        proofs[i] = p.prove(7000 + i)
    t_prove = time.perf_counter() - t0
    bad = n_proofs // 2
    if bad in proofs:
        b = bytearray(proofs[bad])
        b[45] ^= 4
        proofs[bad] = bytes(b)
    truth = [i != bad for i in range(n_proofs)]

    def build(i):
        v = P.Verifier(gens, P.Transcript(b"batch"))
        v.square_chain(v.commit_public(1000 + i), n)
        return v

    # one context per host thread: a context is single-owner, several may share a GPU, and the library
    # calls release the GIL, so the host part of one proof overlaps the device part of another
    ctxs, gens_k = [ctx], [gens]
    for k in range(1, n_ctx):
        c = Context(local)
        cb = Comb(c, BASE)
        ctxs.append(c)
        gens_k.append(P.Gens(c, cb.mul(scalars(n, 1)), cb.mul(scalars(n, 2)), BASE, cb.mul(scalars(1, 3))))

    def build_k(i, k):
        v = P.Verifier(gens_k[k], P.Transcript(b"batch"))
        v.square_chain(v.commit_public(1000 + i), n)
        return v

    lanes = [[(pos, build_k(i, k), proofs[i]) for pos, i in enumerate(mine) if pos % n_ctx == k] for k in range(n_ctx)]
    local_ok = [False] * len(mine)

    def run_lane(k):
        res = P.batch_verify([(v, pr) for _, v, pr in lanes[k]])
        for (pos, _, _), ok in zip(lanes[k], res):
            local_ok[pos] = ok

    # warm every context's arenas and caches with one untimed verification
    for k in range(n_ctx):
        if lanes[k]:
            i0 = mine[lanes[k][0][0]]
            P.batch_verify([(build_k(i0, k), proofs[i0])])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    import threading

    t0 = time.perf_counter()
    if n_ctx == 1:
        run_lane(0)
    else:
        ths = [threading.Thread(target=run_lane, args=(k,)) for k in range(n_ctx)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    t_local = time.perf_counter() - t0
    res = multi.gather_results(local_ok, n_proofs, device=dev if world > 1 else None)
    t_total = time.perf_counter() - t0
    t = torch.tensor([t_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "n_gpus": world, "contexts_per_gpu": n_ctx, "proofs": n_proofs, "lg_multipliers": lg, "results_correct": res == truth,
            "batch_ms_max_over_ranks": round(float(t.item()) * 1e3, 2), "proofs_per_s": round(n_proofs / float(t.item()), 1),
            "ms_per_proof_per_gpu": round(t_local * 1e3 / max(len(mine), 1), 3), "prove_ms_each_untimed": round(t_prove * 1e3 / max(len(mine), 1), 2),
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
