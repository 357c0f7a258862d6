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

    def result_tensor(self, m, c):
        to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v  # noqa: E731
        return self.torch.tensor([to_i64(m), to_i64(c)], dtype=self.torch.int64)


def _worker(rank, world, port, nR, nS, q):
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
        sj = gj.distributed.ShardedJoin(nR, nS, group=None, mode="nccl", ops=NumpyOps(nR + nS, rank, world))
        got = sj.join_aggregate(sl(Rk), sl(Rp), sl(Sk), sl(Sp), nR, nS)
        tot = torch.tensor([got.local_R, got.local_S])
        dist.all_reduce(tot)
        q.put((rank, got.matches == want.matches and got.checksum == want.checksum and tot.tolist() == [nR, nS],
               (got.matches, want.matches, got.checksum, want.checksum, tot.tolist())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nR,nS", [(2, 20_000, 50_000), (4, 30_000, 30_001), (2, 7, 0)])
def test_sharded_join_host_logic_gloo(world, nR, nS):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nR, nS, q)) for r in range(world)]
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
