"""One rank's share of the sharded "partition, then push" pipeline on ONE GPU: local pass, fine
histogram, pushing pass (all destination buffers local, so this times the kernels at HBM speed,
without NVLink), for a sweep of first-pass bits and both output variants.
usage: python tools/pp_probe.py [--n 128000000] [--G 8] [--B 15]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128_000_000)
    ap.add_argument("--G", type=int, default=8)
    ap.add_argument("--B", type=int, default=15)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    gj = ge.load_package()
    n, G, B = args.n, args.G, args.B
    g = G.bit_length() - 1
    eng = gj.JoinEngine(n, n, 0)
    cols = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4)]
    N = n * G
    eng.generate_unique(cols[0], cols[1], 0, N, 4, 40)      # rank 0's rows of a G*n-tuple relation
    eng.generate_unique(cols[2], cols[3], 0, N, 5, 50)
    cap = int(n / G * 1.3) + 4096
    own = [[torch.empty(cap + 16, dtype=torch.int64, device="cuda") for _ in range(G)] for _ in range(2)]
    nq = G << B
    hist = [torch.empty(nq, dtype=torch.int32, device="cuda") for _ in range(2)]
    allh = [torch.zeros(G * nq, dtype=torch.int32, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for p1 in ([0] + [b for b in range(max(g, (g + B) - 10), min(10, g + B - 1) + 1)]):
        for out, big in ((0, 0), (1, 0), (1, 1)):
            eng.set_option("pass1_bits", p1)
            eng.set_option("pp_out", out); eng.set_option("pp_tile16k", big)
            best = None
            for _ in range(args.reps):
                eng.pp_begin(N, N, G, 0, B)
                for w in range(2):
                    eng.pp_local(w, cols[2 * w], cols[2 * w + 1], hist[w])
                    torch.cuda.synchronize()                          # the engine runs on its own stream
                    allh[w][:nq].copy_(hist[w])                       # only rank 0 contributes
                    torch.cuda.synchronize()
                    eng.pp_push(w, allh[w], [t.data_ptr() for t in own[w]], cap, n)
                eng.pp_join(own[0][0].data_ptr(), own[1][0].data_ptr(), cap, cap)
                m, c, a, b, ph = eng.pp_finish()
                tot = sum(ph.values())
                if best is None or tot < best[0]:
                    best = (tot, ph, (m, a, b))
            b1, b2 = eng.pp_plan()
            ph = best[1]
            gbs = lambda ms, bytes_per: bytes_per * n / (ms * 1e-3) / 1e9  # noqa: E731
            print(json.dumps({"b1": b1, "b2": b2, "out": out, "big": big, **{k: round(v, 3) for k, v in ph.items()},
                              "local_GBs(4+16+8 B/tuple)": round(gbs(ph["local_R_ms"], 28), 1),
                              "push_GBs(16 B/tuple)": round(gbs(ph["push_R_ms"], 16), 1),
                              "check(matches,recvR,recvS)": best[2]}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
