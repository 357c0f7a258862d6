"""World-size-2/4 CPU tests (gloo) of the multi-GPU host logic in
icde2019-gpu-join_b200/distributed.py: count exchange, receive layout, all-to-all split sizes,
mod-2^64 all-reduce.  The three device steps are replaced by a numpy stand-in whose local join
is the oracle; on a GPU box the same ShardedJoin runs with GpuOps (tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class NumpyOps:
    """Test stand-in for distributed.GpuOps (CPU tensors, oracle as the local join)."""

    def __init__(self, cap, rank, world):
        import torch
        self.torch, self.rank, self.world = torch, rank, world
        self.cap_R = self.cap_S = cap
        self.send = [torch.zeros(cap, dtype=torch.int64) for _ in range(2)]
        self.recv = [torch.zeros(cap, dtype=torch.int64) for _ in range(2)]
        self.bits = self.gpu_bits = None

    def configure(self, radix_bits, gpu_bits):
        self.bits, self.gpu_bits = radix_bits, gpu_bits

    def count(self, keys, G, shift):
        return np.bincount((keys.numpy().view(np.uint32) >> shift) & (G - 1), minlength=G).astype(np.int64)

    def split(self, which, keys, pays, G, shift):
        k, p = keys.numpy(), pays.numpy()
        d = (k.view(np.uint32) >> shift) & (G - 1)
        order = np.argsort(d, kind="stable")
        packed = (k[order].view(np.uint32).astype(np.uint64) | (p[order].view(np.uint32).astype(np.uint64) << 32)).view(np.int64)
        self.send[which][: packed.size] = self.torch.from_numpy(packed)
        return np.bincount(d, minlength=G).astype(np.int64)

    def local_join(self, nR, nS):
        from oracle import oracle
        un = lambda t, n: t[:n].numpy().view(np.uint64)  # noqa: E731
        r, s = un(self.recv[0], nR), un(self.recv[1], nS)
        rk, rp = (r & 0xFFFFFFFF).astype(np.uint32).view(np.int32), (r >> 32).astype(np.uint32).view(np.int32)
        sk, sp = (s & 0xFFFFFFFF).astype(np.uint32).view(np.int32), (s >> 32).astype(np.uint32).view(np.int32)
        for k in (rk, sk):   # every received tuple belongs to this rank
            assert np.all(((k.view(np.uint32) >> self.bits) & (self.world - 1)) == self.rank)
        res = oracle.join_check(rk, rp, sk, sp, 1)
        return res.matches, res.checksum, {}

    def pp_join(self, dist, group, rank, rels, G, B, peers, own_ptrs, n_glob):
        """Mode "pp" without a GPU: fine histograms with numpy, the real all-gather, the layout of
        distributed.pp_layout (the numpy model of pp_cursor_kernel), and the push emulated by an
        all-to-all of (slot, tuple) pairs -- the receiver checks that the slots every source computed
        on its own tile [0, total) exactly and that every partition range holds only its keys."""
        from oracle import oracle
        import __graft_entry__ as ge
        pp_layout = ge.load_package().distributed.pp_layout
        torch = self.torch
        nq, mask = G << B, (G << B) - 1
        cols, local_n = [], []
        for k, p in rels:
            kk, pv = k.numpy(), p.numpy()
            q = (kk.view(np.uint32) & mask).astype(np.int64)
            hist = torch.from_numpy(np.bincount(q, minlength=nq).astype(np.int32))
            allh = torch.empty(G * nq, dtype=torch.int32)
            dist.all_gather_into_tensor(allh, hist, group=group)
            ah = allh.numpy().reshape(G, nq)
            cur, off, cnt = pp_layout(ah, rank, B)
            order = np.argsort(q, kind="stable")
            qs = q[order]
            start = np.concatenate(([0], np.cumsum(np.bincount(qs, minlength=nq))[:-1]))
            slot = cur[qs] + (np.arange(qs.size) - start[qs])
            packed = (kk[order].view(np.uint32).astype(np.uint64) | (pv[order].view(np.uint32).astype(np.uint64) << 32)).view(np.int64)
            send = torch.from_numpy(np.ascontiguousarray(np.stack([slot, packed], axis=1)))
            n_to = np.bincount(qs >> B, minlength=G)
            n_from = ah.reshape(G, G, -1)[:, rank].sum(axis=1)
            recv = torch.empty((int(n_from.sum()), 2), dtype=torch.int64)
            dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in n_from],
                                   input_split_sizes=[int(x) for x in n_to], group=group)
            r = recv.numpy()
            tot = int(off[-1])
            assert r.shape[0] == tot and np.array_equal(np.sort(r[:, 0]), np.arange(tot))
            buf = np.empty(tot, dtype=np.int64)
            buf[r[:, 0]] = r[:, 1]
            u = buf.view(np.uint64)
            keys = (u & 0xFFFFFFFF).astype(np.uint32)
            pid = np.repeat(np.arange(1 << B, dtype=np.int64) + (rank << B), cnt)
            assert np.array_equal((keys & mask).astype(np.int64), pid)
            cols.append((keys.view(np.int32), (u >> 32).astype(np.uint32).view(np.int32)))
            local_n.append(tot)
        res = oracle.join_check(cols[0][0], cols[0][1], cols[1][0], cols[1][1], 1)
        return res.matches, res.checksum, local_n, {}

    def pcp_join(self, dist, group, rank, rels, G, B, peers, own_ptrs, n_glob, flags=None, stages=(2, 4), peer_hist=False):
        """Mode "pcp" without a GPU: coarse histograms with numpy, the real all-gather, the layout of
        distributed.pcp_layout (numpy model of pcp_layout_kernel), the chunk copies emulated by an
        all-to-all of (slot, tuple) pairs.  The receiver checks that the slots tile [0, total) exactly,
        that its buffer is first-pass partitioned, and joins with the oracle."""
        from oracle import oracle
        import __graft_entry__ as ge
        d = ge.load_package().distributed
        torch = self.torch
        g, bl, b2 = d.pcp_plan_bits(G, B)
        n1 = 1 << (g + bl)
        cols, local_n = [], []
        for k, p in rels:
            kk, pv = k.numpy(), p.numpy()
            c = ((kk.view(np.uint32) >> (B - bl)) & (n1 - 1)).astype(np.int64)
            hist = torch.from_numpy(np.bincount(c, minlength=n1).astype(np.int32))
            allh = torch.empty(G * n1, dtype=torch.int32)
            dist.all_gather_into_tensor(allh, hist, group=group)
            ah = allh.numpy().reshape(G, n1)
            dst, src, tots = d.pcp_layout(ah, rank, bl)
            assert not ((dst ^ src) & 1).any()                      # stage slot and destination slot share the 16-byte phase
            order = np.argsort(c, kind="stable")
            cs = c[order]
            start = np.concatenate(([0], np.cumsum(np.bincount(cs, minlength=n1))[:-1]))
            slot = dst[cs] + (np.arange(cs.size) - start[cs])
            packed = (kk[order].view(np.uint32).astype(np.uint64) | (pv[order].view(np.uint32).astype(np.uint64) << 32)).view(np.int64)
            send = torch.from_numpy(np.ascontiguousarray(np.stack([slot, packed], axis=1)))
            n_to = np.bincount(cs >> bl, minlength=G)
            n_from = ah.reshape(G, G, -1)[:, rank].sum(axis=1)
            recv = torch.empty((int(n_from.sum()), 2), dtype=torch.int64)
            dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in n_from],
                                   input_split_sizes=[int(x) for x in n_to], group=group)
            r = recv.numpy()
            tot = int(tots[rank])
            assert r.shape[0] == tot and np.array_equal(np.sort(r[:, 0]), np.arange(tot))
            buf = np.empty(tot, dtype=np.int64)
            buf[r[:, 0]] = r[:, 1]
            u = buf.view(np.uint64)
            keys = (u & 0xFFFFFFFF).astype(np.uint32)
            cnt = ah.sum(axis=0)[rank << bl:(rank + 1) << bl]
            pid = np.repeat(np.arange(1 << bl, dtype=np.int64) + (rank << bl), cnt)
            assert np.array_equal(((keys >> (B - bl)) & (n1 - 1)).astype(np.int64), pid)
            cols.append((keys.view(np.int32), (u >> 32).astype(np.uint32).view(np.int32)))
            local_n.append(tot)
        res = oracle.join_check(cols[0][0], cols[0][1], cols[1][0], cols[1][1], 1)
        return res.matches, res.checksum, local_n, {}

    def result_tensor(self, m, c):
        to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v  # noqa: E731
        return self.torch.tensor([to_i64(m), to_i64(c)], dtype=self.torch.int64)


def _worker(rank, world, port, nR, nS, q, mode="nccl"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle import oracle
    gj = ge.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1234)          # same global relations on every rank
        Rk = rng.integers(-2**31, 2**31, nR, dtype=np.int64).astype(np.int32)
        half = Rk[rng.integers(0, nR, nS // 2)] if nR else np.empty(0, np.int32)
        Sk = np.concatenate([half, rng.integers(-2**31, 2**31, nS - half.size, dtype=np.int64).astype(np.int32)])
        Rp = rng.integers(-2**31, 2**31, nR, dtype=np.int64).astype(np.int32)
        Sp = rng.integers(-2**31, 2**31, nS, dtype=np.int64).astype(np.int32)
        want = oracle.join_check(Rk, Rp, Sk, Sp, 1)
        sl = lambda a: torch.from_numpy(a[len(a) * rank // world: len(a) * (rank + 1) // world].copy())  # noqa: E731
        sj = gj.distributed.ShardedJoin(nR, nS, group=None, mode=mode, ops=NumpyOps(nR + nS, rank, world))
        got = sj.join_aggregate(sl(Rk), sl(Rp), sl(Sk), sl(Sp), nR, nS)
        tot = torch.tensor([got.local_R, got.local_S])
        dist.all_reduce(tot)
        q.put((rank, got.matches == want.matches and got.checksum == want.checksum and tot.tolist() == [nR, nS],
               (got.matches, want.matches, got.checksum, want.checksum, tot.tolist())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nR,nS,mode", [(2, 20_000, 50_000, "nccl"), (4, 30_000, 30_001, "nccl"), (2, 7, 0, "nccl"),
                                              (2, 20_000, 50_000, "pp"), (4, 30_000, 30_001, "pp"), (2, 7, 0, "pp"),
                                              (2, 20_000, 50_000, "pcp"), (4, 30_000, 30_001, "pcp"), (2, 7, 0, "pcp")])
def test_sharded_join_host_logic_gloo(world, nR, nS, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nR, nS, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out


def test_receive_layout_and_bits(gj):
    d = gj.distributed
    counts = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    rc, ro, wa = d.receive_layout(counts, 1)
    assert rc.tolist() == [2, 5, 8] and ro.tolist() == [0, 2, 7] and wa.tolist() == [1, 2, 3]
    rc, ro, wa = d.receive_layout(counts, 0)
    assert wa.tolist() == [0, 0, 0]
    assert d.choose_radix_bits(128_000_000) == 15 and d.choose_radix_bits(1 << 20) == 8
    assert d.choose_radix_bits(250_000_000) == 16 and d.choose_radix_bits(10) == 1


def test_pp_layout_tiles_every_destination(gj):
    """The write cursors every source derives from the gathered histograms partition each
    destination's buffer exactly: partition p of destination d is [off_d[p], off_d[p+1]) and the
    sources' shares follow each other in rank order."""
    d = gj.distributed
    rng = np.random.default_rng(3)
    G, B = 4, 3
    H = rng.integers(0, 7, size=(G, G << B))
    H[:, 5] = 0                                    # an empty partition
    cur = [d.pp_layout(H, r, B)[0] for r in range(G)]
    for dest in range(G):
        _, off, cnt = d.pp_layout(H, dest, B)
        assert off[0] == 0 and np.array_equal(np.diff(off), cnt)
        assert np.array_equal(cnt, H[:, dest << B:(dest + 1) << B].sum(axis=0))
        for p in range(1 << B):
            q = (dest << B) | p
            at = off[p]
            for r in range(G):
                assert cur[r][q] == at
                at += H[r, q]
            assert at == off[p + 1]


def gg_B_too_wide(G, B):
    return (G.bit_length() - 1) + B > 20


def test_pcp_layout_and_plan(gj):
    """pcp: every chunk lands behind the shares of the lower ranks inside its first-pass partition at
    the destination, stage regions do not overlap and share the destination's 16-byte phase; the
    plan keeps g + bl <= 10, bl <= 8, b2 <= 10."""
    d = gj.distributed
    rng = np.random.default_rng(8)
    G, bl = 4, 3
    g = 2
    H = rng.integers(0, 9, size=(G, 1 << (g + bl)))
    H[:, 3] = 0
    tot = H.sum(axis=0)
    for c in range(1 << (g + bl)):
        at = tot[(c >> bl) << bl:c].sum()
        for r in range(G):
            dst, src, tots = d.pcp_layout(H, r, bl)
            assert dst[c] == at and not ((dst[c] ^ src[c]) & 1)
            at += H[r, c]
    for r in range(G):
        dst, src, tots = d.pcp_layout(H, r, bl)
        remote = (np.arange(1 << (g + bl)) >> bl) != r      # the chunks a GPU keeps are never staged
        rs, rc = src[remote], H[r][remote]
        assert np.all(rs[1:] >= (rs + rc)[:-1]) and rs[-1] + rc[-1] <= rc.sum() + remote.sum()
        assert np.array_equal(src[~remote], dst[~remote])
        assert np.array_equal(tots, tot.reshape(G, -1).sum(axis=1))
    for G2 in (2, 4, 8, 16, 64):
        for B in range(1, 17):
            for p1 in (0, 3, 7, 10):
                if gg_B_too_wide(G2, B):
                    with pytest.raises(ValueError):
                        d.pcp_plan_bits(G2, B, p1)
                    continue
                gg, b_l, b2 = d.pcp_plan_bits(G2, B, p1)
                assert gg + b_l <= 10 and 0 <= b_l <= 8 and b_l <= B - 1 and b2 == B - b_l and 1 <= b2 <= 10, (G2, B, p1)
    assert d.pcp_plan_bits(8, 15) == (3, 6, 9) and d.pcp_plan_bits(2, 15) == (1, 7, 8) and d.pcp_plan_bits(8, 16) == (3, 7, 9)


def test_pcp_stage_positions_cover_every_chunk_once(gj):
    """Staged copy: the position ranges of the stages tile [0, 2^(g+bl)) in order, every stage holds whole
    first-pass partitions for all destinations, more stages than partitions collapse; and the copy kernel's
    piece lookup (binary search from the stage's first position, own chunks have no pieces) visits exactly
    the remote chunks of the stage."""
    d = gj.distributed
    for g, bl, K in ((1, 7, 4), (3, 6, 5), (3, 6, 64), (3, 6, 200), (2, 0, 3), (3, 7, 1)):
        pos = d.pcp_stage_positions(K, g, bl)
        assert len(pos) == min(K, 1 << bl) and pos[0][0] == 0 and pos[-1][1] == 1 << (g + bl)
        assert all(a[1] == b[0] for a, b in zip(pos, pos[1:])) and all(lo < hi and lo % (1 << g) == 0 for lo, hi in pos)
    rng = np.random.default_rng(3)
    g, bl, rank = 2, 3, 1
    n1 = 1 << (g + bl)
    perm = lambda k: ((k & ((1 << g) - 1)) << bl) | (k >> g)          # tile_perm: position -> chunk  # noqa: E731
    pieces = np.array([0 if (perm(k) >> bl) == rank else int(rng.integers(0, 4)) for k in range(n1)])
    prefix = np.concatenate(([0], np.cumsum(pieces)))
    seen = []
    for lo, hi in d.pcp_stage_positions(3, g, bl):
        for k in range(int(prefix[lo]), int(prefix[hi])):
            a, b = lo, n1
            while b - a > 1:
                m = (a + b) // 2
                if prefix[m] <= k:
                    a = m
                else:
                    b = m
            assert lo <= a < hi and pieces[a] > 0 and (perm(a) >> bl) != rank
            seen.append((a, k - int(prefix[a])))
    assert sorted(seen) == [(k, s) for k in range(n1) for s in range(pieces[k])]


def test_staged_receiver_prefix_property():
    """What gj_pcp_recv relies on: the receive layout places first-pass partition j at the sum of the counts of all
    j' < j, fine partitions inside it in order, so when the fine counts arrive stage by stage (ascending j) a scan over
    ALL 2^B counters -- later stages still zero -- already gives the final offsets of every partition of the stages
    that have landed, and the final scan leaves them unchanged.  Also: the units of a stage only involve its partitions."""
    rng = np.random.default_rng(12)
    bl, b2, K = 4, 5, 3
    nj, nb = 1 << bl, 1 << (bl + b2)
    fine = rng.integers(0, 40, nb)
    fine[rng.integers(0, nb, 30)] = 0
    final = np.concatenate(([0], np.cumsum(fine)))
    seen = np.zeros(nb, dtype=np.int64)
    for k in range(K):
        j_lo, j_hi = k * nj // K, (k + 1) * nj // K
        p_lo, p_hi = j_lo << b2, j_hi << b2
        seen[p_lo:p_hi] = fine[p_lo:p_hi]                       # pcp_sum_hist_kernel of stage k
        off = np.concatenate(([0], np.cumsum(seen)))            # full-range scan, later stages still zero
        assert np.array_equal(off[:p_hi + 1], final[:p_hi + 1])
        # first-pass partition j of the receive buffer = [off[j << b2], off[(j + 1) << b2]): the tiles of the stage
        coarse = fine.reshape(nj, -1).sum(axis=1)
        recv_off = np.concatenate(([0], np.cumsum(coarse)))
        assert np.array_equal(off[[j << b2 for j in range(j_lo, j_hi + 1)]], recv_off[j_lo:j_hi + 1])
    assert np.array_equal(off, final)


def test_pcp_copy_piece_arithmetic_model():
    """numpy model of pcp_layout_kernel's piece prefix and pcp_copy_kernel's per-piece arithmetic
    (csrc/kernels.cuh section 3d), both piece-to-CTA assignments: every tuple of every chunk is copied
    exactly once, bulk bodies start on even slots on both sides and have even length, heads / tails
    are single tuples."""
    P = 2048
    rng = np.random.default_rng(21)
    for trial in range(20):
        n1 = int(rng.choice([1, 2, 8, 64]))
        cnt = rng.integers(0, 3 * P + 5, n1)
        cnt[rng.integers(0, n1)] = 0
        if n1 > 2:
            cnt[1], cnt[2] = 1, P + 1
        src = np.concatenate(([0], np.cumsum(cnt + 1)[:-1]))
        dst = rng.integers(0, 50, n1) + np.concatenate(([0], np.cumsum(cnt)[:-1])) * 2
        src = src + ((src ^ dst) & 1)                       # phase matched, as pcp_layout_kernel does
        phase = src & 1
        pieces = np.where(cnt > 0, np.maximum(1, (cnt - phase + P - 1) // P), 0)
        prefix = np.concatenate(([0], np.cumsum(pieces)))
        total = int(prefix[-1])
        for grid, contig in ((3, False), (3, True), (7, True), (1, False)):
            copied = [np.zeros(int(c), dtype=np.int32) for c in cnt]
            seen = 0
            for b in range(grid):
                if contig:
                    per = (total + grid - 1) // grid
                    ks = range(min(total, b * per), min(total, b * per + per))
                else:
                    ks = range(b, total, grid)
                for k in ks:
                    seen += 1
                    lo = max(i for i in range(n1) if prefix[i] <= k)     # the kernel's binary search: largest such position
                    assert pieces[lo] > 0
                    c, s = lo, k - int(prefix[lo])
                    ph = int(phase[c])
                    if s == 0 and ph:
                        copied[c][0] += 1
                    body0 = ph + s * P
                    m = min(P, int(cnt[c]) - body0) if cnt[c] > body0 else 0
                    if m & 1:
                        copied[c][body0 + m - 1] += 1
                        m -= 1
                    assert m % 2 == 0 and (src[c] + body0) % 2 == 0 and (dst[c] + body0) % 2 == 0
                    copied[c][body0:body0 + m] += 1
            assert seen == total
            for c in range(n1):
                assert np.all(copied[c] == 1), (trial, grid, contig, c, cnt[c], phase[c])
