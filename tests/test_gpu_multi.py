"""Multi-GPU parity (needs >= 2 GPUs on the box; `gpurun --gpus N`): the radix-sharded join with
every shuffle variant (NCCL all-to-all, fused peer-store scatter over NVLink, copy-engine transfers,
and the default "partition, then push" pipeline whose last radix pass stores into the peers) against the oracle's
closed form for the device-generated unique relations."""
import os
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_local, mode, overlap, q, peer_hist=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle import oracle
    gj = ge.load_package()
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        N = n_local * world
        sj = gj.distributed.ShardedJoin(n_local, n_local, device=rank, mode=mode, overlap=overlap, pcp_peer_hist=peer_hist)
        eng = sj.ops.engine
        mk = lambda: torch.empty(n_local, dtype=torch.int32, device=f"cuda:{rank}")  # noqa: E731
        Rk, Rp, Sk, Sp = mk(), mk(), mk(), mk()
        eng.generate_unique(Rk, Rp, rank * n_local, N, 4, 40)
        eng.generate_unique(Sk, Sp, rank * n_local, N, 5, 50)
        torch.cuda.synchronize()
        ok = True
        for _ in range(2):     # twice: receive buffers are reused
            res = sj.join_aggregate(Rk, Rp, Sk, Sp, N, N)
            ok = ok and res.matches == N and res.checksum == oracle.unique_join_checksum(0, N, 40, 50)
        tot = torch.tensor([res.local_R, res.local_S], device=f"cuda:{rank}")
        dist.all_reduce(tot)
        q.put((rank, ok and tot.tolist() == [N, N], (res.matches, res.checksum, res.local_R, res.local_S)))
        sj.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,overlap", [("nccl", False), ("p2p", False), ("p2p", True), ("dma", True), ("pp", True), ("pcp", True),
                                          ("pcp-peer-hist", True)])
def test_sharded_join_on_real_gpus(mode, overlap):
    import torch
    import torch.multiprocessing as mp
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (min(ngpu, 8).bit_length() - 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    peer_hist = mode == "pcp-peer-hist"      # coarse histograms through the peers' control blocks, no all-gather
    mode = "pcp" if peer_hist else mode
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6_000_000, mode, overlap, q, peer_hist)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out


def _skew_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle import oracle
    gj = ge.load_package()
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nR, nS = 1_000_000 * world, 3_000_000 * world
        rng = np.random.default_rng(99)                       # the same global relations on every rank
        Rk = rng.permutation(nR).astype(np.int32)
        hot = rng.integers(0, 64, nS // 2)                    # half of the probe side hits 64 keys: they all live on GPU 0
        Sk = np.concatenate((hot, rng.integers(0, nR + 1000, nS - nS // 2))).astype(np.int32)
        rng.shuffle(Sk)
        Rp, Sp = oracle.payload_of_keys(Rk, 40), oracle.payload_of_keys(Sk, 50)
        want = oracle.join_check(Rk, Rp, Sk, Sp)
        sl = lambda a: torch.from_numpy(a[len(a) * rank // world: len(a) * (rank + 1) // world].copy()).cuda()  # noqa: E731
        cols = [sl(a) for a in (Rk, Rp, Sk, Sp)]
        res = {}
        # enough room for the hot destination (1/2 + 1/(2 world) of S): exact result, one GPU holds most of S
        sj = gj.distributed.ShardedJoin(nR // world, nS // world, device=rank, mode="pcp", slack=0.55 * world + 1.0)
        for _ in range(2):
            r = sj.join_aggregate(*cols, nR, nS)
        tot = torch.tensor([r.local_R, r.local_S], device=f"cuda:{rank}")
        mx = tot.clone()
        dist.all_reduce(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        res["ok"] = (r.matches, r.checksum) == (want.matches, want.checksum) and tot.tolist() == [nR, nS]
        res["hot_share"] = mx[1].item() / nS
        sj.close()
        # default slack: the hot destination would overflow -> every rank gets the same error up front, nothing hangs
        sj = gj.distributed.ShardedJoin(nR // world, nS // world, device=rank, mode="pcp")
        try:
            sj.join_aggregate(*cols, nR, nS)
            res["refused"] = False
        except gj.GJError as e:
            res["refused"] = "more tuples than its buffer holds" in str(e)
        sj.close()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_skewed_probe_side_on_real_gpus():
    """Multi-GPU skew: half of the probe side hits 64 keys that live on one GPU.  With receive buffers sized
    for it the streamed exchange gives the oracle's aggregate (the hot GPU holds > 1/2 of S); with the default slack every
    rank refuses the join up front with the same error instead of overflowing or hanging."""
    import torch
    import torch.multiprocessing as mp
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (min(ngpu, 8).bit_length() - 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_skew_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r["ok"] and r["refused"] and r["hot_share"] > 0.5 for _, r in out), out
