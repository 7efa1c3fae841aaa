"""The N>1 path on the CPU: two `gloo` ranks run the sharding and exchange logic of
mpc_bulletproof_b200.multi (stride-sharded MSM, MPC open of additive shares, batch-verify
result gather) with a checker engine built on the oracle standing in for the CUDA engine
(the product never imports the oracle; this test supplies it).  SURVEY.md §8e."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import group as G
from tests.util import rand_point, rand_scalar, rng, scalars_bytes

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ext_words(p: G.Point):
    out = []
    for c in (p.X, p.Y, p.Z, p.T):
        out += [(c >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    return out


def _from_words(w):
    cs = [sum(int(w[8 * k + i]) << (32 * i) for i in range(8)) for k in range(4)]
    return G.Point(*cs)


class OracleEngine:
    """Checker stand-in for multi.CudaEngine: same interface, big-integer arithmetic."""

    device = torch.device("cpu")

    def __init__(self, points):
        self.points = points

    def partial(self, scalars: bytes, n_sets: int = 1):
        n = len(self.points)
        ks = [int.from_bytes(scalars[32 * i : 32 * i + 32], "little") for i in range(n * n_sets)]
        words = []
        for s in range(n_sets):
            words += _ext_words(G.msm(ks[s * n : (s + 1) * n], self.points))
        return torch.tensor(words, dtype=torch.int64).to(torch.int32)

    def combine(self, parts, n_parts: int, n_sets: int = 1):
        w = (parts.to(torch.int64) & 0xFFFFFFFF).tolist()
        out = []
        for s in range(n_sets):
            acc = G.IDENTITY
            for p in range(n_parts):
                acc = acc + _from_words(w[(p * n_sets + s) * 32 : (p * n_sets + s + 1) * 32])
            out.append(acc.encode())
        return out


def _worker(rank, port, q):
    from mpc_bulletproof_b200 import multi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        r = rng(2024)
        n, sets = 37, 2  # ragged: 19 + 18 terms
        ps = [rand_point(r) for _ in range(n)]
        ks = [[rand_scalar(r) for _ in range(n)] for _ in range(sets)]
        mine = list(multi.shard_indices(n, rank, WORLD))
        assert mine == list(range(rank, n, WORLD))
        eng = OracleEngine([ps[i] for i in mine])
        local = b"".join(scalars_bytes([k[i] for i in mine]) for k in ks)
        got = multi.sharded_msm(eng, local, n_sets=sets)
        want = [G.msm(k, ps).encode() for k in ks]
        assert got == want, "sharded MSM differs from the single-rank sum"
        assert multi.shard_bytes(scalars_bytes(ks[0]), rank, WORLD) == scalars_bytes([ks[0][i] for i in mine])

        # MPC open: additive shares of the scalars over ALL points (share and MAC as two sets)
        full = OracleEngine(ps)
        x = [rand_scalar(r) for _ in range(n)]
        key = rand_scalar(r)
        share0 = [rand_scalar(r) for _ in range(n)]
        mac0 = [rand_scalar(r) for _ in range(n)]
        share = share0 if rank == 0 else [(a - b) % G.L for a, b in zip(x, share0)]
        mac = mac0 if rank == 0 else [(key * a - b) % G.L for a, b in zip(x, mac0)]
        opened = multi.open_shares(full, scalars_bytes(share) + scalars_bytes(mac), n_sets=2)
        C = G.msm(x, ps)
        assert opened == [C.encode(), (key * C).encode()], "opened commitment / MAC differ from the plain ones"

        # batch verification: whole proofs per rank, result bytes gathered in proof order
        truth = [(i * 7) % 3 != 0 for i in range(11)]
        seen = []

        def verify_one(i):
            seen.append(i)
            return truth[i]

        res = multi.batch_verify_sharded(len(truth), verify_one)
        assert res == truth
        assert seen == list(range(rank, len(truth), WORLD)), "a rank verified proofs it does not own"
        assert multi.gather_results([truth[i] for i in range(rank, 11, WORLD)], 11) == truth
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure to the parent
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_single_rank_paths():
    """Without a process group the same calls degrade to the local result."""
    from mpc_bulletproof_b200 import multi

    r = rng(5)
    ps = [rand_point(r) for _ in range(5)]
    ks = [rand_scalar(r) for _ in range(5)]
    assert multi.sharded_msm(OracleEngine(ps), scalars_bytes(ks)) == [G.msm(ks, ps).encode()]
    assert multi.batch_verify_sharded(3, lambda i: i != 1) == [True, False, True]
