"""MSM latency on structured scalars (what real proofs feed the MSM: bit vectors a_L, a_R = a_L - 1
of range gadgets, reference tests/r1cs.rs:629-632; repeated values) next to uniform ones.
Device-resident, CUDA events, 2^lg points.  Prints JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mpc_bulletproof_b200 import Comb, Context, Table  # noqa: E402

BASE = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
L = 2**252 + 27742317777372353535851937790883648493


def limbs(x):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << lg
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    with torch.cuda.stream(stream):
        uni = torch.randint(-(2**31), 2**31, (n, 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        uni[:, 7] &= 0x0FFFFFFF
        comb = Comb(ctx, BASE)
        pts = torch.empty(n * 32, dtype=torch.uint8, device=dev)
        comb.dev_mul(uni.data_ptr(), n, pts.data_ptr())
        table = Table(ctx, dev_ptr=pts.data_ptr(), n=n).set_windows(0)
        bits = (torch.randint(0, 2, (n,), device=dev, generator=g)).to(torch.int64)
        sc_bits = torch.zeros((n, 8), dtype=torch.int64, device=dev)
        sc_bits[:, 0] = bits
        lm1 = torch.tensor(limbs(L - 1), dtype=torch.int64, device=dev)
        sc_neg = (1 - bits)[:, None] * lm1[None, :]  # a_R = a_L - 1: 0 or l - 1
        sc_same = torch.tensor(limbs(0x1234567 * 2**200 + 12345), dtype=torch.int64, device=dev)[None, :].repeat(n, 1)
        cases = {
            "uniform": uni,
            "bits": sc_bits.to(torch.int32),
            "zero_or_minus_one": sc_neg.to(torch.int32),
            "all_equal": sc_same.to(torch.int32),
            "u64_values": torch.cat([uni[:, :2], torch.zeros((n, 6), dtype=torch.int32, device=dev)], dim=1),
        }
        out_ext = torch.zeros(32, dtype=torch.int32, device=dev)
    stream.synchronize()
    res = {"lg": lg, "window": table.window}
    for name, sc in cases.items():
        sc = sc.contiguous()
        with torch.cuda.stream(stream):
            for _ in range(2):
                table.dev_msm(sc.data_ptr(), 1, out_ext.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                table.dev_msm(sc.data_ptr(), 1, out_ext.data_ptr())
            e1.record(stream)
        stream.synchronize()
        res[name + "_ms"] = round(e0.elapsed_time(e1) / 5, 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
