"""Partitioned radix join vs the non-partitioned baselines (global chained hash table; perfect array) over build
sizes: where does the 126 MB L2 stop carrying the table?  One JSON line per size.
usage: python tools/nopart_crossover.py [--max-log2 27]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=16)
    ap.add_argument("--max-log2", type=int, default=27)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    gj = ge.load_package()
    for lg in range(args.min_log2, args.max_log2 + 1):
        n = 1 << lg
        with gj.JoinEngine(n, n, 0, nopart_max=0) as eng:      # join_aggregate = the partitioned path at every size
            cols = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4)]
            eng.generate_unique(cols[0], cols[1], 0, n, 4, 40)
            eng.generate_unique(cols[2], cols[3], 0, n, 5, 50)
            torch.cuda.synchronize()
            best = {}
            perfect = lambda *c: eng.join_aggregate_perfect(*c, 0, n)      # keys are a permutation of [0, n)  # noqa: E731
            for name, fn in (("partitioned", eng.join_aggregate), ("nopart", eng.join_aggregate_nopart), ("perfect", perfect)):
                ts = []
                for _ in range(args.reps):
                    r = fn(*cols)
                    assert r.matches == n, (name, r.matches)
                    ts.append((r.timings.total_ms, r.timings.wall_ms))
                best[name] = min(ts)
            print(json.dumps({"log2_n": lg, "n": n,
                              "partitioned_ms": round(best["partitioned"][0], 4), "nopart_ms": round(best["nopart"][0], 4),
                              "perfect_ms": round(best["perfect"][0], 4),
                              "partitioned_wall_ms": round(best["partitioned"][1], 4), "nopart_wall_ms": round(best["nopart"][1], 4),
                              "perfect_wall_ms": round(best["perfect"][1], 4),
                              "G_tuples_s_partitioned": round(2 * n / best["partitioned"][0] / 1e6, 2),
                              "G_tuples_s_nopart": round(2 * n / best["nopart"][0] / 1e6, 2)}), flush=True)


if __name__ == "__main__":
    main()
