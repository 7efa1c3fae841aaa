"""profiles/<tag>_prove_round_shares.md from one run of tools/r2_round_shares.sh: the per-stage / per-round wall clock
of an unprofiled proof (BPG_TRACE) next to the ncu launch list of a proof of the same shape (cold-cache, serialised:
shares, not absolute times).  usage: python tools/round_shares_md.py gpurun_out/r2p 16 10 > profiles/r2_prove_round_shares.md"""
import csv
import json
import re
import sys


def proofs(path):
    lines = [l.strip() for l in open(path) if l.startswith("[bpg]")]
    starts = [i for i, l in enumerate(lines) if "prove blindings" in l]
    ends = [i for i, l in enumerate(lines) if "prove ipp rounds" in l]
    return [lines[s : e + 1] for s, e in zip(starts, ends)]


def launches(path):
    rows = []
    lines = [l for l in open(path) if not l.startswith("==")]
    for x in csv.DictReader(lines):
        if x["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1000 if u in ("nsecond", "ns") else (v * 1000 if u in ("msecond", "ms") else v)
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("bpg::", "")
        rows.append((name, v, x["Grid Size"], x["Block Size"]))
    idx = [i for i, r in enumerate(rows) if "k_blind_vectors" in r[0]]
    # the last proof: from its blinding kernel to the verifier's first kernel (k_decode_ext of the prefetch, or the end)
    s = idx[-1]
    e = next((i for i in range(s, len(rows)) if rows[i][0].startswith("k_decode_ext") or rows[i][0].startswith("k_flat_terms") and i > s + 40), len(rows))
    return rows[s:e]


def main():
    base = sys.argv[1]
    lgs = [int(x) for x in sys.argv[2:]]
    out = ["# One R1CS proof, stage by stage and round by round (round 2, final code)", ""]
    out.append("`tools/r2_round_shares.sh` on one B200: `BPG_TRACE=1 python tools/prove_profile.py <lg> 2` (host wall clock of every stage")
    out.append("of an UNPROFILED proof: each stage ends in a device-to-host read, so the wall clock is the device time plus the")
    out.append("host's share) and `ncu --metrics gpu__time_duration.sum --clock-control none` over the same command (per-launch")
    out.append("durations: cold-cache and serialised, so their SHARES are comparable, not their sums).")
    out.append("")
    for lg in lgs:
        d = json.load(open(f"{base}_prove_{lg}.json"))
        P = proofs(f"{base}_trace_{lg}.txt")
        p = P[min(5, len(P) - 1)]
        rounds = [float(l.split()[-2]) for l in p if "round L,R" in l]
        folds = [float(l.split()[-2]) for l in p if "challenge + fold" in l]
        stages = [(re.sub(r"^\[bpg\] prove ", "", l).rsplit(None, 2)[0].strip(), float(l.split()[-2])) for l in p if l.startswith("[bpg] prove")]
        total = sum(v for _, v in stages)
        out.append(f"## 2^{lg} multipliers: prove {min(d['prove_ms_unprofiled']):.2f} ms, verify {min(d['verify_ms_unprofiled']):.2f} ms "
                   f"({d['runs'][0]['launches'] if d['runs'] else '?'} / {d['verify']['launches']} launches)")
        out.append("")
        out.append("| stage (host/protocol.cpp) | ms | share |")
        out.append("|---|---|---|")
        for name, v in stages:
            out.append(f"| {name} | {v:.3f} | {100 * v / total:.1f} % |")
        out.append("")
        out.append("Inner-product rounds (`L,R` = everything up to the two encodings on the host; `fold` = transcript, challenge,")
        out.append("its inverse, and the launch of the fold, which runs under the next round's `L,R`):")
        out.append("")
        out.append("| round j | vector length | L,R ms | challenge + fold (host) ms |")
        out.append("|---|---|---|---|")
        n = 1 << lg
        for j, (r, f) in enumerate(zip(rounds, folds)):
            out.append(f"| {j} | {n >> j} | {r:.3f} | {f:.3f} |")
        out.append("")
        L = launches(f"{base}_launches_{lg}.csv")
        agg = {}
        for name, v, g, b in L:
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v for _, v, _, _ in L)
        out.append(f"ncu launch list of one proof ({len(L)} launches, {tot / 1000:.2f} ms serialised):")
        out.append("")
        out.append("| kernel | launches | total us | share |")
        out.append("|---|---|---|---|")
        for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v / tot < 0.004:
                continue
            out.append(f"| {name} | {c} | {v:.0f} | {100 * v / tot:.1f} % |")
        out.append("")
    print("\n".join(out))


if __name__ == "__main__":
    main()
