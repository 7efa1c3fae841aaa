"""Host-side logic of the two-party emulation (mpc_bulletproof_b200/mpc.py) that needs no GPU: additive
sharing with MACs, Beaver multiplication and authenticated opening over a two-thread link, and the
same over two `gloo` processes (world_size 2)."""
import os
import random
import threading

from mpc_bulletproof_b200.mpc import AS, L, Fabric, LocalLink, MacCheckError, MockDealer


def _run_pair(fn):
    la, lb = LocalLink.pair()
    out, err = [None, None], []

    def party(p, link):
        try:
            out[p] = fn(p, Fabric(p, link, MockDealer(4242, p), None))
        except Exception as e:  # noqa: BLE001
            err.append(e)
            # unblock the peer
            try:
                link.tx.put(b"")
            except Exception:
                pass

    ts = [threading.Thread(target=party, args=(0, la)), threading.Thread(target=party, args=(1, lb))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    return out, err


def test_beaver_products_and_authenticated_open():
    r = random.Random(1)
    xs = [r.randrange(L) for _ in range(20)]
    ys = [r.randrange(L) for _ in range(20)]

    def fn(p, f):
        X, Y = f.dealer.share_many(xs), f.dealer.share_many(ys)
        prod = f.mul(X, Y)
        lin = [a.scale(3) + b + f.const(7) for a, b in zip(prod, X)]
        ip = f.inner_product(X, Y)
        return f.open_authenticated(lin + [ip])

    out, err = _run_pair(fn)
    assert not err
    want = [(3 * x * y + x + 7) % L for x, y in zip(xs, ys)] + [sum(x * y for x, y in zip(xs, ys)) % L]
    assert out[0] == want and out[1] == want


def test_mac_check_catches_a_wrong_share():
    def fn(p, f):
        x = f.dealer.share(5)
        if p == 1:
            x = AS((x.s + 1) % L, x.m)  # party 1 lies about its share
        return f.open_authenticated([x])

    out, err = _run_pair(fn)
    assert err and all(isinstance(e, MacCheckError) for e in err)


def _gloo_party(rank, port, q):
    import torch.distributed as dist

    from mpc_bulletproof_b200.mpc import TorchLink

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    f = Fabric(rank, TorchLink(), MockDealer(99, rank), None)
    xs, ys = f.dealer.share_many([3, 4, 5]), f.dealer.share_many([6, 7, 8])
    q.put((rank, f.open_authenticated(f.mul(xs, ys) + [f.inner_product(xs, ys)])))
    dist.destroy_process_group()


def test_two_gloo_processes():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_party, args=(r, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
    assert res[0] == res[1] == [18, 28, 40, 86]
