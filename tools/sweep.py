#!/usr/bin/env python
"""Tuning sweep on a GPU box: per-phase times of the join pipeline for every kernel variant.
Writes JSON lines to stdout.  Usage: python tools/sweep.py [--n 128000000] [--what scatter,join,bits]"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128_000_000)
    ap.add_argument("--nS", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--what", default="scatter,join,bits,unit,combo")
    args = ap.parse_args()
    import torch
    gj = ge.load_package()
    nR, nS = args.n, args.nS or args.n
    eng = gj.JoinEngine(nR, nS, 0)
    mk = lambda n: torch.empty(n, dtype=torch.int32, device="cuda")  # noqa: E731
    Rk, Rp, Sk, Sp = mk(nR), mk(nR), mk(nS), mk(nS)
    eng.generate_unique(Rk, Rp, 0, nR, 4, 40)
    if nS == nR:
        eng.generate_unique(Sk, Sp, 0, nS, 5, 50)
        expect = nS
    else:   # FK: keys of S uniform over R's domain via bijection on nS then mod
        eng.generate_unique(Sk, Sp, 0, nS, 5, 50)
        Sk.remainder_(nR)
        expect = nS
    torch.cuda.synchronize()
    peak = 6531.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def run(tag, **opts):
        for k in ("radix_bits", "pass1_bits", "join_cfg", "unit_tuples", "join_grid"):
            eng.set_option(k, 0)
        eng.set_option("scatter_cfg", 255)
        for k, v in opts.items():
            eng.set_option(k, v)
        ts = []
        try:
            for _ in range(args.reps + 1):
                r = eng.join_aggregate(Rk, Rp, Sk, Sp)
                ts.append(r.timings.as_dict())
        except gj.GJError as e:
            print(json.dumps({"tag": tag, "opts": opts, "error": str(e)}), flush=True)
            return
        ts = ts[1:]
        med = lambda k: statistics.median(t[k] for t in ts)  # noqa: E731
        pm = [statistics.median(t["pass_ms"][i] for t in ts) for i in range(4)]
        n_of = [min(nR, nS), min(nR, nS), max(nR, nS), max(nR, nS)]
        gbs = [16.0 * n_of[i] / (pm[i] * 1e-3) / 1e9 if pm[i] > 0 else 0 for i in range(4)]
        print(json.dumps({"tag": tag, "opts": opts, "ok": r.matches == expect, "bits": [ts[0]["radix_bits"], ts[0]["pass1_bits"], ts[0]["pass2_bits"]],
                          "hist_ms": round(med("hist_ms"), 4), "part_ms": round(med("part_ms"), 4), "join_ms": round(med("join_ms"), 4),
                          "total_ms": round(med("total_ms"), 4), "pass_ms": [round(x, 4) for x in pm],
                          "pass_GBs": [round(x) for x in gbs], "pass_frac": [round(x / peak, 3) for x in gbs],
                          "join_GBs": round(8.0 * (nR + nS) / (med("join_ms") * 1e-3) / 1e9),
                          "hist_GBs": round(4.0 * (nR + nS) / (med("hist_ms") * 1e-3) / 1e9),
                          "Gtuples_s": round((nR + nS) / (med("total_ms") * 1e-3) / 1e9, 2)}), flush=True)

    what = args.what.split(",")
    run("default")
    if "scatter" in what:
        for c in range(eng.get_option("num_scatter_cfgs")):
            run("scatter", scatter_cfg=c)
    if "join" in what:
        for c in range(eng.get_option("num_join_cfgs")):
            run("join", join_cfg=c)
    if "bits" in what:
        for b in (13, 14, 15):
            for p1 in (7, 8):
                if b - p1 <= 8:
                    run("bits", radix_bits=b, pass1_bits=p1)
    if "combo" in what:
        for p1, c1s, c2s in ((7, (0, 1), (2, 3)), (8, (0,), (2, 3))):
            for c1 in c1s:
                for c2 in c2s:
                    run("combo", radix_bits=15, pass1_bits=p1, scatter_cfg1=c1, scatter_cfg2=c2)
    if "unit" in what:
        for u in (4096, 8192, 32768, 65536):
            run("unit", unit_tuples=u)
        for g in (148, 296, 592):
            run("join_grid", join_grid=g)


if __name__ == "__main__":
    main()
