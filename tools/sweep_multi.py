"""BASELINE.json config 3 across GPUs, STRONG scaling: a fixed total of 2^12 .. 2^24 points sharded by stride over
the N ranks (term i on rank i mod N); every rank runs the pipeline on its shard, the 128-byte
partial sums meet in the fused peer exchange, every rank adds and encodes.  Each size is checked against the
expected result ((sum_i s_i k_i) * B: the points are k_i * B) -- the same bytes a single GPU produces.
Run under torchrun (any N, also 1); rank 0 prints one JSON object.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_multi.py [max_lg]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mpc_bulletproof_b200 import Comb, Context, Table  # noqa: E402
from mpc_bulletproof_b200.api import dev_sum_encode  # noqa: E402

BASE = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
L = 2**252 + 27742317777372353535851937790883648493


def main():
    max_lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    comb = Comb(ctx, BASE)
    peer = None
    if world > 1:
        from mpc_bulletproof_b200.multi import PeerExchange

        peer = PeerExchange(ctx, max_sets=1)

    def scalars(n, seed):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        s = torch.randint(-(2**31), 2**31, (n, 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        s[:, 7] &= 0x0FFFFFFF
        return s.contiguous()

    def ints(t):
        raw = t.cpu().numpy().tobytes()
        return [int.from_bytes(raw[i : i + 32], "little") for i in range(0, len(raw), 32)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    rows = []
    for lg in range(12, max_lg + 1, 2):
        n_total = 1 << lg
        n = n_total // world  # this rank's shard (stride sharding: any partition gives the same sum)
        with torch.cuda.stream(stream):
            k = scalars(n, 1000 * lg + rank)
            pts = torch.empty(n * 32, dtype=torch.uint8, device=dev)
            comb.dev_mul(k.data_ptr(), n, pts.data_ptr())
            table = Table(ctx, dev_ptr=pts.data_ptr(), n=n).set_windows(0)
            sc = [scalars(n, 2000 * lg + 10 * rank + i) for i in range(2)]
            part = torch.zeros(32, dtype=torch.int32, device=dev)
            res = torch.zeros(32, dtype=torch.uint8, device=dev)
        stream.synchronize()

        def step(i):
            table.dev_msm(sc[i % 2].data_ptr(), 1, part.data_ptr())
            if peer is not None:
                peer.exchange_sum_encode(part.data_ptr(), 1, res.data_ptr())
            else:
                dev_sum_encode(ctx, part.data_ptr(), 1, 1, res.data_ptr())

        reps = 20 if lg <= 20 else 6
        with torch.cuda.stream(stream):
            for i in range(3):
                step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(reps):
                step(i)
            e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        got = bytes(res.cpu().tolist())
        # expected bytes: exact integer dot products (bounded: the check costs host time, not GPU time)
        ok = None
        if lg <= 22:
            last = sc[(reps - 1) % 2]
            dot = sum(a * b for a, b in zip(ints(last), ints(k))) % L
            if world > 1:
                mine = torch.tensor(list(dot.to_bytes(32, "little")), dtype=torch.uint8, device=dev)
                alld = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(alld, mine)
                dot = sum(int.from_bytes(bytes(d.cpu().tolist()), "little") for d in alld) % L
            ok = comb.mul(dot.to_bytes(32, "little")) == got
        rows.append({"lg_total": lg, "points_per_gpu": n, "window": table.window, "ms": round(float(ms.item()), 4),
                     "mpoints_s": round(n_total / float(ms.item()) / 1e3, 1), "result_ok": ok})
        if rank == 0:
            print(rows[-1], file=sys.stderr, flush=True)
        table.close()
        del pts, sc, k
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"gpus": world, "scaling": "strong", "gpu": torch.cuda.get_device_name(0), "rows": rows}))
    if peer is not None:
        barrier()
        peer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    bad = [r for r in rows if r["result_ok"] is False]
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
