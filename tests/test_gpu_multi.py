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


def _worker(rank, world, port, n_local, mode, overlap, q):
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
        sj = gj.distributed.ShardedJoin(n_local, n_local, device=rank, mode=mode, overlap=overlap)
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


@pytest.mark.parametrize("mode,overlap", [("nccl", False), ("p2p", False), ("p2p", True), ("dma", True), ("pp", True), ("pcp", True)])
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
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6_000_000, mode, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out
