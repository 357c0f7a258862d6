"""The pcp exchange kernels of ONE process driving TWO GPUs (rank r on device r, peer access enabled), so that a
single-process `ncu` can capture pcp_copy_kernel with its NVLink traffic (nvltx / nvlrx bytes, peer-aperture L2
sectors) -- multi-process captures are off the table (ncu replays kernels).  Runs: coarse histograms, layout + source
pass, staged copy (+ fine counts + flags) on both devices, then the receiver passes + join, and checks the aggregate
unless --no-check (under ncu's kernel replay the copy kernel's counts accumulate).
usage: python tools/ncu_pcp_two_devices.py [--n 64000000] [--stages 1,1] [--no-check] [--copy-only]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64_000_000, help="tuples per relation and GPU")
    ap.add_argument("--stages", default="1,1")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--copy-only", action="store_true")
    ap.add_argument("--ctas", type=int, default=24)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    gj = ge.load_package()
    L = gj.lib()
    G, n = 2, args.n
    assert torch.cuda.device_count() >= 2, "needs 2 GPUs"
    for a, b in ((0, 1), (1, 0)):
        rc = L.gj_enable_peer_access(a, b)
        assert rc == 0, L.gj_last_error()
    st = tuple(int(x) for x in args.stages.split(","))
    N = n * G
    B = gj.distributed.choose_radix_bits(n)
    cap = int(n * 1.1) + 4096
    engs = [gj.JoinEngine(cap, cap, r, gpu_bits=1, pcp_copy_ctas=args.ctas) for r in range(G)]
    dev = [torch.device("cuda", r) for r in range(G)]
    cols = []
    for r in range(G):
        c = [torch.empty(n, dtype=torch.int32, device=dev[r]) for _ in range(4)]
        engs[r].generate_unique(c[0], c[1], r * n, N, 4, 40)
        engs[r].generate_unique(c[2], c[3], r * n, N, 5, 50)
        cols.append(c)
    own = [[torch.zeros(cap + 16, dtype=torch.int64, device=dev[r]) for r in range(G)] for _ in range(2)]
    ctrl = [torch.zeros(gj.pcp_ctrl_bytes(G) // 4, dtype=torch.int32, device=dev[r]) for r in range(G)]
    for r in range(G):
        torch.cuda.synchronize(r)
    for r in range(G):
        engs[r].pcp_begin(N, N, G, r, B)
    g, bl, b2 = engs[0].pcp_plan()
    n1 = 1 << (g + bl)
    hist = [[torch.empty(n1, dtype=torch.int32, device=dev[r]) for r in range(G)] for _ in range(2)]
    for r in range(G):
        for w in range(2):
            engs[r].pcp_hist(w, cols[r][2 * w], hist[w][r])
    for r in range(G):
        torch.cuda.synchronize(r)
    allh = [[torch.stack([h.to(dev[r]) for h in hist[w]]).contiguous() for r in range(G)] for w in range(2)]
    for r in range(G):
        torch.cuda.synchronize(r)
    for w in range(2):
        for r in range(G):
            engs[r].pcp_part(w, cols[r][2 * w], cols[r][2 * w + 1], allh[w][r], own[w][r].data_ptr(), cap)
        for r in range(G):
            torch.cuda.synchronize(r)
        for r in range(G):       # both directions at once, as in the real exchange
            engs[r].pcp_copy(w, [t.data_ptr() for t in own[w]], [t.data_ptr() for t in ctrl], st[w])
        for r in range(G):
            torch.cuda.synchronize(r)
    if args.copy_only:
        print("copy done")
        return
    m = c = 0
    for r in range(G):
        for w in range(2):
            engs[r].pcp_recv(w, own[w][r].data_ptr(), ctrl[r].data_ptr(), cap)
        mm, cc, a, b, ph, bits = engs[r].pcp_finish()
        m += mm
        c = (c + cc) % 2**64
        print(f"rank {r}: received {a} + {b}, phases {ph}")
    if not args.no_check:
        from oracle import oracle
        assert (m, c) == (N, oracle.unique_join_checksum(0, N, 40, 50)), (m, c)
        print(f"ok: matches {m} checksum {c} (closed form)")


if __name__ == "__main__":
    main()
