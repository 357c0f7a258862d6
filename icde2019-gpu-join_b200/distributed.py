"""Multi-GPU radix join: one process per GPU, torch.distributed for the plumbing.

No reference counterpart exists (the reference only ever calls
cudaSetDevice(1), hash_join_clustered_probe.cu:1001,1685).  Design:

  radix field = [ gpu bits | local bits ]:  destination GPU d = (key >> B) & (G-1) where B is
  the number of radix bits of the LOCAL partitioner, i.e. the top bits of the field pick the
  GPU and co-partitions of R and S meet on one GPU.  Steps per join:
    1. count      per-destination histogram of R and S           (gj_shuffle_count / _split)
    2. exchange   G x G count matrix                              (all_gather, tiny)
    3. shuffle    mode "nccl": split into destination groups, then all_to_all_single
                  mode "p2p" : ONE kernel partitions and stores each tuple run straight into
                               the destination GPU's receive buffer over NVLink (peer memory
                               mapped with CUDA IPC) -- no send buffer, no separate collective
    4. local      partition + build + probe on the received tuples (gj_join_aggregate_tuples)
    5. reduce     all_reduce(SUM) of {matches, checksum} (int64 wrap-around == mod 2^64)

  mode "pp" ("partition, then push") turns the order around: every GPU first partitions its OWN
  shard on all [gpu | local] bits (pass 1 local), the ranks all-gather their fine histograms
  (2^(g+B) counters each -- every rank then derives the same global layout, `pp_layout`), and the
  LAST radix pass stores its runs straight into the destination GPU's final partition buffer over
  NVLink.  The receiver joins what arrives without touching it again, R's push overlaps S's local
  pass, and the shuffle costs no extra pass over the data.

`ops` abstracts the three device steps so the host logic (counts, offsets, split sizes,
collectives, reduction) is testable on CPU with the gloo backend and a stand-in `ops`.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np


def choose_radix_bits(n_build_local: int, part_target: int = 4096, max_bits: int = 16) -> int:
    """Mirror of choose_plan() in csrc/api.cu: smallest B with n >> B <= part_target."""
    b = 0
    while b < max_bits and (n_build_local >> b) > part_target:
        b += 1
    return max(b, 1)


def receive_layout(counts: np.ndarray, rank: int):
    """counts[s][d] = tuples rank s sends to rank d.  Returns (recv_counts[s], recv_offsets[s],
    my_write_offset_at[d]) -- source s writes its tuples for d at sum_{s' < s} counts[s'][d]."""
    counts = np.asarray(counts, dtype=np.int64)
    recv_counts = counts[:, rank].copy()
    recv_offsets = np.concatenate(([0], np.cumsum(recv_counts)[:-1]))
    write_at = np.array([counts[:rank, d].sum() for d in range(counts.shape[1])], dtype=np.int64)
    return recv_counts, recv_offsets, write_at


def pp_layout(all_hist: np.ndarray, rank: int, local_bits: int):
    """numpy model of pp_cursor_kernel (csrc/kernels.cuh): all_hist[s][q] = tuples of source s in
    global partition q = (dest << local_bits) | p.  Destination d lays partition p out at
    sum_{p' < p} C[d][p'] (C = counts over all sources); source `rank` writes its share at
    + sum_{s < rank} H[s][d][p].  Returns (write cursors of `rank` [G << B], own offsets [2^B + 1],
    own counts [2^B])."""
    all_hist = np.asarray(all_hist)
    G = all_hist.shape[0]
    n_p = 1 << local_bits
    H = all_hist.astype(np.int64).reshape(G, G, n_p)          # [source][dest][partition]
    Cn = H.sum(axis=0)                                        # [dest][partition]
    off = np.concatenate((np.zeros((G, 1), dtype=np.int64), np.cumsum(Cn, axis=1)), axis=1)
    cur = (off[:, :-1] + H[:rank].sum(axis=0)).reshape(-1)
    return cur, off[rank], Cn[rank]


def pcp_plan_bits(n_gpus: int, local_bits: int, pass1_bits: int = 0):
    """Mirror of the plan in gj_pcp_begin (csrc/api.cu): (gpu bits g, local bits bl of the source-side
    pass, bits b2 of the receiver-side pass), g + bl <= 10, bl <= 8, b2 = local_bits - bl <= 10."""
    g, B = n_gpus.bit_length() - 1, local_bits
    bl = max(pass1_bits - g, 0) if pass1_bits else ((B - g + 1) // 2 if B > g else 0)
    bl = min(bl, 10 - g, B - 1, 8)
    if B - bl > 10:
        bl = B - 10
    if g + bl > 10 or bl > 8:
        raise ValueError(f"{g} GPU bits + {B} local bits do not fit two passes of <= 10 bits")
    return g, bl, B - bl


def pcp_layout(all_hist: np.ndarray, rank: int, source_local_bits: int):
    """numpy model of pcp_layout_kernel: all_hist[s][c] = tuples of source s in chunk c = (dest <<
    bl) | j (j = first-pass partition at the destination).  Returns (dst_start[c] of `rank`'s
    shares, src_start[c] in `rank`'s stage buffer -- remote chunks only: the chunks a GPU keeps are
    written straight into its receive buffer, there src_start[c] = dst_start[c] --, tuples every
    destination receives)."""
    H = np.asarray(all_hist).astype(np.int64)
    G, n1 = H.shape
    bl = source_local_bits
    tot = H.sum(axis=0)
    ex = np.concatenate(([0], np.cumsum(tot)[:-1]))
    dest = np.arange(n1) >> bl
    dbase = ex[dest << bl]
    dst = ex - dbase + H[:rank].sum(axis=0)
    stays = dest == rank
    region_sz = np.where(stays, 0, H[rank] + 1)
    region = np.concatenate(([0], np.cumsum(region_sz)[:-1]))
    src = np.where(stays, dst, region + ((region ^ dst) & 1))
    return dst, src, tot.reshape(-1, 1 << bl).sum(axis=1)


def pcp_stage_positions(n_stages: int, gpu_bits: int, source_local_bits: int):
    """Copy positions [lo, hi) of every stage (mirror of gj_pcp_copy): positions are (first-pass
    partition j, destination d) with d fastest; stage k covers partitions [k nj / K, (k+1) nj / K)."""
    nj = 1 << source_local_bits
    K = min(n_stages, nj)
    return [(((k * nj) // K) << gpu_bits, (((k + 1) * nj) // K) << gpu_bits) for k in range(K)]


@dataclass
class ShardedResult:
    matches: int
    checksum: int
    local_R: int
    local_S: int
    phases_ms: dict = field(default_factory=dict)


class GpuOps:
    """Device steps backed by libgpujoin.so (the product path)."""

    def __init__(self, max_R: int, max_S: int, device: int, slack: float = 1.3, with_send_buffers: bool = True):
        import torch
        from .engine import JoinEngine
        self.torch = torch
        self.device = device
        self.cap_R = int(max_R * slack) + 4096
        self.cap_S = int(max_S * slack) + 4096
        self.engine = JoinEngine(self.cap_R, self.cap_S, device)
        self.stream = torch.cuda.Stream(device)
        self.engine.use_torch_stream(self.stream)
        # overlap: local radix passes run on a high-priority stream so that their CTAs get SM
        # slots ahead of the (NVLink-bound) peer-scatter kernel of the other relation
        self.stream_local = torch.cuda.Stream(device, priority=-1)
        self.stream_shuffle = torch.cuda.Stream(device)
        dev = torch.device("cuda", device)
        self.dev = dev
        self.send = self.recv = None
        if with_send_buffers:   # only the NCCL shuffle stages tuples; peer stores need neither
            self.send = [torch.empty(self.cap_R, dtype=torch.int64, device=dev),
                         torch.empty(self.cap_S, dtype=torch.int64, device=dev)]
            self.recv = [torch.empty(self.cap_R, dtype=torch.int64, device=dev),
                         torch.empty(self.cap_S, dtype=torch.int64, device=dev)]

    def configure(self, radix_bits: int, gpu_bits: int):
        self.engine.set_option("radix_bits", radix_bits)
        self.engine.set_option("gpu_bits", gpu_bits)

    def count(self, keys, G, shift):
        return self.engine.shuffle_count(keys, G, shift)

    def split(self, which, keys, pays, G, shift):
        return self.engine.shuffle_split(keys, pays, G, shift, self.send[which])

    def scatter_peers(self, keys, pays, G, shift, peer_ptrs, offsets):
        self.engine.shuffle_scatter_peers(keys, pays, G, shift, peer_ptrs, offsets)
        return self.engine.get_option("last_shuffle_us") * 1e-3   # kernel time, ms

    def overlapped_p2p(self, dist, group, rels, G, shift, peers, write_at, n_in, own_ptrs):
        """Shuffle R -> [all ranks done] -> local passes of R  ||  shuffle S -> [done] -> local passes
        of S -> join.  The cross-rank "done" points are 1-element NCCL all-reduces enqueued on the
        streams, so nothing blocks the host until the final result read-back."""
        import os
        torch, eng = self.torch, self.engine
        sA, sB = self.stream_local, self.stream_shuffle
        cur = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_tok"):
            self._tok = [torch.zeros(1, dtype=torch.int32, device=self.dev) for _ in range(2)]
        tok = self._tok
        sA.wait_stream(cur); sB.wait_stream(cur)
        trace = [] if os.environ.get("GJ_TRACE") else None

        def mark(name, stream):
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                trace.append((name, e))
        mark("start", sA)
        eng.stage_begin(n_in[0], n_in[1], sA)
        (Rk, Rp), (Sk, Sp) = rels
        eng.shuffle_scatter_peers_async(0, Rk, Rp, G, shift, peers[0], write_at[0], sA)
        r_sent = torch.cuda.Event()
        r_sent.record(sA)
        mark("R shuffled (local kernel end)", sA)
        with torch.cuda.stream(sA):
            dist.all_reduce(tok[0], group=group)          # every rank's R stores have landed
        mark("R token", sA)
        eng.stage_partition(0, own_ptrs[0], sA)
        mark("R partitioned", sA)
        if not os.environ.get("GJ_CONCURRENT_SHUFFLE"):
            sB.wait_event(r_sent)                         # one relation on NVLink at a time
        eng.shuffle_scatter_peers_async(1, Sk, Sp, G, shift, peers[1], write_at[1], sB)
        mark("S shuffled (local kernel end)", sB)
        with torch.cuda.stream(sB):
            dist.all_reduce(tok[1], group=group)          # every rank's S stores have landed
        mark("S token", sB)
        eng.stage_partition(1, own_ptrs[1], sB)
        mark("S partitioned", sB)
        eng.stage_join(sB)
        mark("joined", sB)
        m, c = eng.stage_finish()
        sA.synchronize()
        ms = eng.shuffle_scatter_ms(0) + eng.shuffle_scatter_ms(1)
        out = {"shuffle_scatter_ms": ms, "pass_ms": eng.stage_pass_ms()}
        if trace is not None:
            out["trace_ms"] = {n: round(trace[0][1].elapsed_time(e), 3) for n, e in trace[1:]}
        return m, c, out

    def dma_shuffle(self, dist, group, rank, rels, G, shift, peers, counts, write_at, n_in, own_ptrs):
        """Shuffle variant "dma": ONE kernel per relation splits the shard by destination -- this
        rank's own share goes straight into its receive buffer, the other shares into a send
        buffer -- and the copy engines move the groups to the peers over NVLink (one stream per
        destination) while the SMs already run the next kernels: split of S, local passes of R
        (after its token), local passes of S, join."""
        torch, eng = self.torch, self.engine
        sK, sT = self.stream_local, self.stream_shuffle
        if not hasattr(self, "_copy_streams"):
            self._copy_streams = [torch.cuda.Stream(self.device) for _ in range(max(G - 1, 1))]
            self._tok = [torch.zeros(1, dtype=torch.int32, device=self.dev) for _ in range(2)]
        cur = torch.cuda.current_stream(self.device)
        sK.wait_stream(cur); sT.wait_stream(cur)
        eng.stage_begin(n_in[0], n_in[1], sK)
        split_done, copies_done = [], []
        for which, (k, p) in enumerate(rels):
            cnt = counts[which][rank]                        # what this rank sends to each destination
            send_off = np.concatenate(([0], np.cumsum(cnt)[:-1]))
            bases = [self.send[which].data_ptr()] * G
            offs = [int(x) for x in send_off]
            bases[rank] = own_ptrs[which]
            offs[rank] = int(write_at[which][rank])
            eng.shuffle_scatter_peers_async(which, k, p, G, shift, bases, offs, sK)
            ev = torch.cuda.Event()
            ev.record(sK)
            split_done.append((ev, cnt, send_off))
        for which in range(2):
            ev, cnt, send_off = split_done[which]
            evs = []
            for j, g in enumerate(x for x in range(G) if x != rank):
                cs = self._copy_streams[j]
                cs.wait_event(ev)
                eng.memcpy_d2d_async(peers[which][g] + int(write_at[which][g]) * 8,
                                     self.send[which].data_ptr() + int(send_off[g]) * 8, int(cnt[g]) * 8, cs)
                e2 = torch.cuda.Event()
                e2.record(cs)
                evs.append(e2)
            copies_done.append(evs)
        for which in range(2):
            for e2 in copies_done[which]:
                sT.wait_event(e2)
            sT.wait_event(split_done[which][0])
            with torch.cuda.stream(sT):
                dist.all_reduce(self._tok[which], group=group)     # every rank's copies have landed
            tok_ev = torch.cuda.Event()
            tok_ev.record(sT)
            sK.wait_event(tok_ev)
            eng.stage_partition(which, own_ptrs[which], sK)
        eng.stage_join(sK)
        m, c = eng.stage_finish()
        for cs in self._copy_streams:
            cs.synchronize()
        return m, c, {"shuffle_scatter_ms": eng.shuffle_scatter_ms(0) + eng.shuffle_scatter_ms(1)}

    def pp_join(self, dist, group, rank, rels, G, B, peers, own_ptrs, n_glob):
        """Mode "pp": local passes -> all-gather of the fine histograms -> pushing last pass ->
        join.  R runs on a normal-priority stream, S on the high-priority one: while R's
        (NVLink-bound) push is in flight, S's local pass gets the SM slots it asks for.  The
        cross-rank points are NCCL collectives enqueued on those streams; the host blocks only
        in pp_finish."""
        import os
        torch, eng = self.torch, self.engine
        streams = (self.stream_shuffle, self.stream_local)      # R: normal priority, S: high
        nq = G << B
        if getattr(self, "_pp_nq", None) != (G, nq):
            self._pp_hist = [torch.empty(nq, dtype=torch.int32, device=self.dev) for _ in range(2)]
            self._pp_all = [torch.empty(G * nq, dtype=torch.int32, device=self.dev) for _ in range(2)]
            self._pp_tok = [torch.zeros(1, dtype=torch.int32, device=self.dev) for _ in range(2)]
            self._pp_nq = (G, nq)
        cur = torch.cuda.current_stream(self.device)
        for s in streams:
            s.wait_stream(cur)
        eng.pp_begin(n_glob[0], n_glob[1], G, rank, B, streams[0])
        caps = (self.cap_R, self.cap_S)
        for which, (k, p) in enumerate(rels):
            eng.pp_local(which, k, p, self._pp_hist[which], streams[which])
            if which == 0:       # S's local pass starts when R's is done: it then runs under R's push
                local_done = torch.cuda.Event()
                local_done.record(streams[0])
                streams[1].wait_event(local_done)
            with torch.cuda.stream(streams[which]):
                dist.all_gather_into_tensor(self._pp_all[which], self._pp_hist[which], group=group)
        pushed = None
        for which, (k, p) in enumerate(rels):
            s = streams[which]
            if pushed is not None and not os.environ.get("GJ_CONCURRENT_SHUFFLE"):
                s.wait_event(pushed)                              # one relation on NVLink at a time
            eng.pp_push(which, self._pp_all[which], peers[which], caps[which], k.numel(), s)
            pushed = torch.cuda.Event()
            pushed.record(s)
            with torch.cuda.stream(s):
                dist.all_reduce(self._pp_tok[which], group=group)   # every rank's pushes have landed
        streams[1].wait_stream(streams[0])
        eng.pp_join(own_ptrs[0], own_ptrs[1], caps[0], caps[1], streams[1])
        m, c, n_r, n_s, ph = eng.pp_finish()
        streams[0].synchronize()
        b1, b2 = eng.pp_plan()
        ph = dict(ph, shuffle_scatter_ms=ph["push_R_ms"] + ph["push_S_ms"], pass1_bits=b1, pass2_bits=b2, radix_bits=B)
        return m, c, (n_r, n_s), ph

    def pcp_join(self, dist, group, rank, rels, G, B, peers, own_ptrs, n_glob, flags=None, stages=(2, 4), peer_hist=False):
        """Mode "pcp": coarse histograms -> all-gather -> first radix pass at the source (own chunks
        straight into the receive buffer) -> staged TMA bulk copies of whole first-pass partitions, a
        flag store into every peer after each stage -> per stage at the receiver: wait, histogram, last
        radix pass and (probing relation) join.  Four streams: source side and receiver side of each
        relation; all source-side work is enqueued before any receiver-side wait.  No collective
        after the histogram all-gathers: arrival is signalled through the peers' flag words."""
        import os
        torch, eng = self.torch, self.engine
        if not hasattr(self, "_pcp_streams"):
            # receiver side of the second (probing) relation on the high-priority stream: it ends the step
            self._pcp_streams = (self.stream_shuffle, torch.cuda.Stream(self.device),
                                 self.stream_local, torch.cuda.Stream(self.device, priority=-1))
        first = 1 if n_glob[0] > n_glob[1] else 0           # the building (smaller) relation travels first
        order = (first, 1 - first)
        src = {order[0]: self._pcp_streams[0], order[1]: self._pcp_streams[2]}
        rcv = {order[0]: self._pcp_streams[1], order[1]: self._pcp_streams[3]}
        nst = {order[0]: stages[0], order[1]: stages[1]}
        trace = [] if os.environ.get("GJ_TRACE") else None        # optional GPU timeline (ms since the start mark)

        def mark(name, stream):
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                trace.append((name, e))
        cur = torch.cuda.current_stream(self.device)
        for s in self._pcp_streams:
            s.wait_stream(cur)
        eng.pcp_begin(n_glob[0], n_glob[1], G, rank, B, src[first])
        g, bl, _ = eng.pcp_plan()
        n1 = 1 << (g + bl)                                        # chunks = first-pass partitions of all destinations
        if getattr(self, "_pcp_key", None) != (G, n1):
            self._pcp_hist = [torch.empty(n1, dtype=torch.int32, device=self.dev) for _ in range(2)]
            self._pcp_all = [torch.empty(G * n1, dtype=torch.int32, device=self.dev) for _ in range(2)]
            self._pcp_key = (G, n1)
        caps = (self.cap_R, self.cap_S)
        names = ("R", "S")
        mark("start", src[first])
        for w in order:
            k, p = rels[w]
            eng.pcp_hist(w, k, self._pcp_hist[w], src[w])
            if peer_hist:      # through the peers' control blocks: no collective on the critical path
                eng.pcp_hist_exchange(w, self._pcp_hist[w], flags, flags[rank], self._pcp_all[w], src[w])
            else:
                with torch.cuda.stream(src[w]):
                    dist.all_gather_into_tensor(self._pcp_all[w], self._pcp_hist[w], group=group)
            mark(f"{names[w]} histograms gathered", src[w])
        ev = {}
        for i, w in enumerate(order):
            k, p = rels[w]
            s = src[w]
            if i == 1:
                s.wait_event(ev["part"])                          # the second source pass runs under the first copy ...
            eng.pcp_part(w, k, p, self._pcp_all[w], own_ptrs[w], caps[w], s)
            ev["part"] = torch.cuda.Event(); ev["part"].record(s)
            mark(f"{names[w]} source pass done", s)
            if i == 1:
                s.wait_event(ev["copy"])                          # ... and one relation crosses NVLink at a time
            eng.pcp_copy(w, peers[w], flags, nst[w], s)
            ev["copy"] = torch.cuda.Event(); ev["copy"].record(s)
            mark(f"{names[w]} copied (local kernels done)", s)
        if not hasattr(self, "_pcp_result"):
            self._pcp_result = torch.zeros(2, dtype=torch.int64, device=self.dev)
        for w in order:                                           # receiver side: only now, behind all source-side work
            eng.pcp_recv(w, own_ptrs[w], flags[rank], caps[w], rcv[w], result_out=self._pcp_result if w != first else None)
            mark(f"{names[w]} received" + (" and joined" if w != first else ""), rcv[w])
        # the global aggregate: all-reduce of the local {matches, checksum} right behind the last join, on its
        # stream (int64 two's-complement add = addition mod 2^64); the host blocks once, in pcp_finish
        with torch.cuda.stream(rcv[order[1]]):
            dist.all_reduce(self._pcp_result, group=group)
            self._pcp_global = self._pcp_result.to("cpu", non_blocking=True)
        m, c, n_r, n_s, ph, bits = eng.pcp_finish()
        for s in self._pcp_streams:
            s.synchronize()
        ph = dict(ph, shuffle_scatter_ms=ph["copy_R_ms"] + ph["copy_S_ms"], radix_bits=B, pass1_bits=bits[0] + bits[1],
                  pass2_bits=bits[2], stages=[nst[0], nst[1]],
                  global_result=[int(x) & 0xFFFFFFFFFFFFFFFF for x in self._pcp_global.tolist()])
        if trace is not None:
            torch.cuda.synchronize(self.device)
            ph["trace_ms"] = {n: round(trace[0][1].elapsed_time(e), 3) for n, e in trace[1:]}
        return m, c, (n_r, n_s), ph

    def exchange_counts(self, dist, group, mine):
        """All ranks' count vectors in ONE small NCCL all-gather (doubles as a barrier)."""
        t = self.torch.tensor([int(x) for x in mine], dtype=self.torch.int64, device=self.dev)
        out = self.torch.empty(dist.get_world_size(group) * t.numel(), dtype=self.torch.int64, device=self.dev)
        dist.all_gather_into_tensor(out, t, group=group)
        return out.cpu().numpy().reshape(dist.get_world_size(group), -1)

    def local_join(self, nR, nS):
        res = self.engine.join_aggregate_tuples(self.recv[0], nR, self.recv[1], nS)
        return res.matches, res.checksum, res.timings.as_dict()

    def local_join_ptrs(self, ptr_R, nR, ptr_S, nS):
        res = self.engine.join_aggregate_ptrs(ptr_R, nR, ptr_S, nS)
        return res.matches, res.checksum, res.timings.as_dict()

    def result_tensor(self, matches, checksum):
        # int64 two's complement add == addition mod 2^64
        to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v  # noqa: E731
        return self.torch.tensor([to_i64(matches), to_i64(checksum)], dtype=self.torch.int64,
                                 device=self.torch.device("cuda", self.device))


class ShardedJoin:
    """R and S are sharded row-wise over the ranks of `group`; every rank calls join_aggregate
    with its local shard (device int32 columns) and all ranks get the global result."""

    def __init__(self, max_local_R: int, max_local_S: int, device: int | None = None, group=None,
                 mode: str = "auto", ops=None, part_target: int = 4096, overlap: bool = True, pcp_stages=(2, 4),
                 slack: float = 1.3, pcp_peer_hist: bool = False):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world & (self.world - 1):
            raise ValueError("world size must be a power of two (GPU id = radix bits)")
        self.gpu_bits = int(math.log2(self.world))
        if mode == "auto":     # measured on B200 / NVLink 5 (profiles/README.md): only whole first-pass
            mode = "pcp"       # partitions cross NVLink, at the bulk-copy rate (684 GB/s out per GPU at 8 GPUs)
        self.mode = mode
        self.overlap = overlap
        self.pcp_stages = tuple(pcp_stages)     # copy / receive stages of the building and of the probing relation
        self.pcp_peer_hist = pcp_peer_hist      # coarse histograms through the peers' control blocks instead of an all-gather
        self.part_target = part_target
        # slack = capacity of a GPU's receive buffers relative to an even split.  A skewed probe side sends one
        # GPU more than its share (Zipf z = 1 at 8 GPUs: 1/8 + 5.2 % of S = 1.42x): the exchange refuses up front
        # (every rank sees the same histograms) when a destination would overflow -- size the slack for the skew.
        self.ops = ops if ops is not None else GpuOps(max_local_R, max_local_S, device, slack=slack,
                                                      with_send_buffers=(mode in ("nccl", "dma")))
        self.max_local = (max_local_R, max_local_S)
        self._peers = None
        self._own = [0, 0]
        if mode in ("p2p", "dma", "pp", "pcp"):
            if ops is None:
                self._setup_peers()
            else:                      # test stand-in: no device buffers to map
                self._peers = [[0] * self.world, [0] * self.world, [0] * self.world]
                self._own = [0, 0, 0]
                self._opened = []
        elif mode != "nccl":
            raise ValueError("mode must be 'auto', 'nccl', 'p2p', 'dma', 'pp' or 'pcp'")

    # -- CUDA IPC mapping of every rank's receive buffers (p2p mode) -------------------------
    def _setup_peers(self):
        from .engine import lib, _check
        L = lib()
        L.gj_ipc_export.argtypes = [C.c_void_p, C.c_char_p]
        L.gj_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.gj_ipc_close.argtypes = [C.c_void_p]
        ops = self.ops
        # receive buffers must be plain cudaMalloc allocations to be exportable
        self._own = []
        handles = []
        from .engine import pcp_ctrl_bytes
        ctrl_bytes = pcp_ctrl_bytes(self.world)   # pcp control block: stage flags + the fine histograms the sources deliver
        for which, nbytes in enumerate(((ops.cap_R + 16) * 8, (ops.cap_S + 16) * 8, ctrl_bytes)):
            p = C.c_void_p()
            _check(L.gj_malloc_device(C.byref(p), nbytes))
            if which == 2:
                L.gj_memset_device.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
                _check(L.gj_memset_device(p, 0, nbytes))
            h = C.create_string_buffer(64)
            _check(L.gj_ipc_export(p, h))
            self._own.append(p.value)
            handles.append(h.raw)
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, handles, group=self.group)
        self._peers = [[0] * self.world, [0] * self.world, [0] * self.world]    # R buffers, S buffers, control blocks
        self._opened = []
        for r in range(self.world):
            for which in range(3):
                if r == self.rank:
                    self._peers[which][r] = self._own[which]
                else:
                    q = C.c_void_p()
                    _check(L.gj_ipc_open(gathered[r][which], C.byref(q)))
                    self._peers[which][r] = q.value
                    self._opened.append(q.value)
        self._L = L

    def close(self):
        if self._peers is not None and self._opened is not None and any(self._own):
            for q in self._opened:
                self._L.gj_ipc_close(C.c_void_p(q))
            self.dist.barrier(group=self.group)
            for p in self._own:
                self._L.gj_free_device(C.c_void_p(p))
            self._peers = None

    # -- helpers ---------------------------------------------------------------------------
    def _all_gather_counts(self, mine: np.ndarray) -> np.ndarray:
        if hasattr(self.ops, "exchange_counts"):
            return self.ops.exchange_counts(self.dist, self.group, mine)
        out = [None] * self.world
        self.dist.all_gather_object(out, [int(x) for x in mine], group=self.group)
        return np.array(out, dtype=np.int64)

    def plan_bits(self, n_build_global: int) -> int:
        return choose_radix_bits(max(1, n_build_global // self.world), self.part_target)

    # -- the join ------------------------------------------------------------------------------
    def join_aggregate(self, Rk, Rp, Sk, Sp, n_R_global: int, n_S_global: int) -> ShardedResult:
        import time
        G, rank, ops, dist = self.world, self.rank, self.ops, self.dist
        t_host = [time.perf_counter()]
        lap = lambda: t_host.append(time.perf_counter())  # noqa: E731
        B = self.plan_bits(min(n_R_global, n_S_global))
        if getattr(self, "_configured", None) != (B, self.gpu_bits):
            ops.configure(B, self.gpu_bits)
            self._configured = (B, self.gpu_bits)
        shift = B
        rels = ((Rk, Rp), (Sk, Sp))
        local_n = [0, 0]
        if self.mode in ("pp", "pcp"):
            if self.mode == "pp":
                m, c, local_n, tm = ops.pp_join(dist, self.group, rank, rels, G, B, self._peers, self._own, (n_R_global, n_S_global))
            else:
                m, c, local_n, tm = ops.pcp_join(dist, self.group, rank, rels, G, B, self._peers, self._own, (n_R_global, n_S_global),
                                                 flags=self._peers[2], stages=self.pcp_stages, peer_hist=self.pcp_peer_hist)
            lap()
        elif self.mode == "nccl":
            import torch
            for which, (k, p) in enumerate(rels):
                cnt = ops.split(which, k, p, G, shift)
                counts = self._all_gather_counts(cnt)
                recv_counts, _, _ = receive_layout(counts, rank)
                n_in = int(recv_counts.sum())
                if n_in > ops.recv[which].numel():
                    raise RuntimeError(f"rank {rank}: receives {n_in} tuples, capacity {ops.recv[which].numel()}")
                n_out = int(cnt.sum())
                with torch.cuda.stream(ops.stream) if hasattr(ops, "stream") else _null():
                    dist.all_to_all_single(ops.recv[which][:n_in], ops.send[which][:n_out],
                                           output_split_sizes=[int(x) for x in recv_counts],
                                           input_split_sizes=[int(x) for x in cnt], group=self.group)
                local_n[which] = n_in
            if hasattr(ops, "stream"):
                ops.stream.synchronize()
        else:
            # both relations' counts travel in one all-gather, which is also the barrier that
            # tells every rank its peers are done reading their receive buffers
            mine = np.concatenate([ops.count(k, G, shift) for k, _ in rels])
            lap()
            both = self._all_gather_counts(mine)
            lap()
            all_counts = [both[:, :G], both[:, G:]]
            shuffle_ms = 0.0
            write_ats = []
            for which in range(2):
                recv_counts, _, write_at = receive_layout(all_counts[which], rank)
                n_in = int(recv_counts.sum())
                cap = ops.cap_R if which == 0 else ops.cap_S
                if n_in > cap:
                    raise RuntimeError(f"rank {rank}: receives {n_in} tuples, capacity {cap}")
                write_ats.append(write_at)
                local_n[which] = n_in
            if self.mode == "dma":
                m, c, tm = ops.dma_shuffle(dist, self.group, rank, rels, G, shift, self._peers, all_counts, write_ats,
                                           local_n, self._own)
            elif self.overlap and hasattr(ops, "overlapped_p2p"):
                m, c, tm = ops.overlapped_p2p(dist, self.group, rels, G, shift, self._peers, write_ats, local_n, self._own)
            else:
                for which, (k, p) in enumerate(rels):
                    shuffle_ms += ops.scatter_peers(k, p, G, shift, self._peers[which], write_ats[which]) or 0.0
                dist.barrier(group=self.group)   # every rank's stores have landed
                m, c, tm = ops.local_join_ptrs(self._own[0], local_n[0], self._own[1], local_n[1])
                tm = dict(tm, shuffle_scatter_ms=shuffle_ms)
        if self.mode == "nccl":
            m, c, tm = ops.local_join(local_n[0], local_n[1])
        if self.mode not in ("pp", "pcp"):
            lap()
        if isinstance(tm, dict) and "global_result" in tm:      # pcp: already all-reduced on the device, behind the last join
            vals = tm.pop("global_result")
        else:
            res = ops.result_tensor(m, c)
            dist.all_reduce(res, op=dist.ReduceOp.SUM, group=self.group)
            vals = [int(x) & 0xFFFFFFFFFFFFFFFF for x in res.tolist()]
        lap()
        if self.mode in ("pp", "pcp"):
            d = [1e3 * (b - a) for a, b in zip(t_host, t_host[1:])]
            tm = dict(tm, host_ms={"pipeline": d[0], "reduce": d[1]})
        elif self.mode != "nccl" and len(t_host) == 5:
            d = [1e3 * (b - a) for a, b in zip(t_host, t_host[1:])]
            tm = dict(tm, host_ms={"count": d[0], "exchange": d[1], "shuffle+local": d[2], "reduce": d[3]})
        return ShardedResult(vals[0], vals[1], local_n[0], local_n[1], tm)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
