"""Measurements of the SURVEY 8f rows at size (one JSON line each):
  late   late-materialisation join (gj_join_aggregate_late; reference join_partitioned_varpayload,
         join-primitives.cu:1420-1557): unique n x n, payload = row id, 2 + 2 side columns; the extra cost over the
         plain join is the gather of 4 values per result pair from side tables of 4 * 4n bytes
  stream out-of-HBM probe side (gj_join_aggregate_stream_host; reference outOfGPU_Join3_payload,
         hash_join_clustered_probe.cu:1684-1984): R resident, S streamed from pinned host memory in chunks;
         overlap = total against max(H2D, compute)
usage: python tools/next_rows_bench.py [late|stream|all]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def late(torch, gj):
    for n in (1 << 22, 1 << 24, 128_000_000):
        with gj.JoinEngine(n, n, 0, nopart_max=0) as eng:
            Rk, Rid, Sk, Sid = (torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4))
            eng.generate_unique(Rk, Rid, 0, n, 4, 40)
            eng.generate_unique(Sk, Sid, 0, n, 5, 50)
            Rid.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
            Sid.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
            g = torch.Generator(device="cuda"); g.manual_seed(1)
            Dr = torch.randint(-2**31, 2**31 - 1, (2, n), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
            Ds = torch.randint(-2**31, 2**31 - 1, (2, n), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
            # every row of both relations finds exactly one partner: the sum is the sum of all side-table values
            want = (int(Dr.to(torch.int64).sum().item()) + int(Ds.to(torch.int64).sum().item())) & 0xFFFFFFFFFFFFFFFF
            torch.cuda.synchronize()
            best = {}
            for name, fn in (("plain", lambda: eng.join_aggregate(Rk, Rid, Sk, Sid)), ("late", lambda: eng.join_aggregate_late(Rk, Rid, Sk, Sid, Dr, Ds))):
                ts = []
                for _ in range(4):
                    r = fn()
                    assert r.matches == n
                    if name == "late":
                        assert r.checksum == want, (r.checksum, want)
                    ts.append((r.timings.total_ms, r.timings.join_ms))
                best[name] = min(ts)
            extra = best["late"][1] - best["plain"][1]
            print(json.dumps({"what": "late materialisation", "n": n, "side_table_MB": 16 * n / 1e6, "plain_total_ms": round(best["plain"][0], 3),
                              "late_total_ms": round(best["late"][0], 3), "plain_join_ms": round(best["plain"][1], 3),
                              "late_join_ms": round(best["late"][1], 3), "gather_useful_GBs": round(16.0 * n / (extra * 1e-3) / 1e9, 1) if extra > 0 else None,
                              "gather_sector_GBs": round(4 * 32.0 * n / (extra * 1e-3) / 1e9, 1) if extra > 0 else None,
                              "checked": "matches == n and sum == sum of all side-table values"}), flush=True)
            del Dr, Ds
        torch.cuda.empty_cache()


def stream(torch, gj):
    nR, nS = 1 << 24, 1 << 28
    pins = [torch.empty(n, dtype=torch.int32).pin_memory() for n in (nR, nR, nS, nS)]
    g = gj.generator
    g.create_relation_unique_parallel(nR, nR, 4, out=pins[0].numpy())
    g.create_relation_unique_parallel(nS, nR, 3, out=pins[2].numpy())
    pins[1].fill_(1); pins[3].fill_(1)
    want = nS - (nS - 1) // nR
    for chunk in (1 << 24, 1 << 25, 1 << 26):
        with gj.JoinEngine(nR, 2 * chunk, 0) as eng:
            ts = []
            for _ in range(3):
                r = eng.join_aggregate_stream_host(pins[0], pins[1], pins[2], pins[3], chunk)
                assert r.matches == want == r.checksum, (r.matches, r.checksum, want)
                ts.append(r.timings.as_dict())
            t = min(ts, key=lambda x: x["total_ms"])
            h2d_GBs = 8.0 * (nR + nS) / (t["h2d_ms"] * 1e-3) / 1e9
            print(json.dumps({"what": "streamed probe side", "nR": nR, "nS": nS, "chunk_tuples": chunk, "total_ms": round(t["total_ms"], 2),
                              "h2d_ms": round(t["h2d_ms"], 2), "build_side_ms": round(t["hist_ms"], 2), "probe_chunks_ms": round(t["join_ms"], 2),
                              "h2d_GBs": round(h2d_GBs, 1), "overlap_total_over_h2d": round(t["total_ms"] / t["h2d_ms"], 3),
                              "G_tuples_s": round((nR + nS) / t["total_ms"] / 1e6, 2),
                              "checked": f"matches == checksum == {want}"}), flush=True)


def main():
    import torch
    import __graft_entry__ as ge
    gj = ge.load_package()
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("late", "all"):
        late(torch, gj)
    if what in ("stream", "all"):
        stream(torch, gj)


if __name__ == "__main__":
    main()
