"""Full-size GPU runs of the BASELINE.json configurations, checked through size-independent
properties: closed-form match counts implied by the generators, the closed-form
checksum of the device-generated relations, partition sortedness + multiset preservation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    return torch


def gen(torch, eng, n, n_total, seed, pay_seed, row_begin=0):
    k = torch.empty(n, dtype=torch.int32, device="cuda")
    p = torch.empty(n, dtype=torch.int32, device="cuda")
    eng.generate_unique(k, p, row_begin, n_total, seed, pay_seed)
    return k, p


def test_config3_workload_B_128M(gj, orc, torch_cuda):
    """ETHZ workload B: |R|=|S|=128,000,000 unique keys -> exactly n matches; payload = f(key)
    -> checksum = SUM_k f(k,a) f(k,b) mod 2^64 (oracle closed form)."""
    torch = torch_cuda
    n = 128_000_000
    with gj.JoinEngine(n, n, 0) as eng:
        Rk, Rp = gen(torch, eng, n, n, 4, 40)
        Sk, Sp = gen(torch, eng, n, n, 5, 50)
        res = eng.join_aggregate(Rk, Rp, Sk, Sp)
        assert res.matches == n
        assert res.checksum == orc.unique_join_checksum(0, n, 40, 50)
        assert (res.timings.radix_bits, res.timings.pass1_bits, res.timings.pass2_bits) == (15, 7, 8)
        # idempotence: the engine does not disturb its inputs or keep state between calls
        again = eng.join_aggregate(Rk, Rp, Sk, Sp)
        assert (again.matches, again.checksum) == (res.matches, res.checksum)
        # the partitioner at full size: offsets are the exact histogram, output is grouped by
        # partition id (sortedness) and is a permutation of the input (sum / xor-mix preserved)
        L = gj.lib()
        import ctypes as C
        tp, op, b, t = C.c_void_p(), C.c_void_p(), C.c_uint32(), gj.Timings()
        rc = L.gj_partition(eng._ctx, 0, C.c_void_p(Rk.data_ptr()), C.c_void_p(Rp.data_ptr()), n, 0,
                            C.byref(tp), C.byref(op), C.byref(b), C.byref(t))
        assert rc == 0 and b.value == 15
        offs = np.empty((1 << 15) + 1, dtype=np.uint32)
        L.gj_memcpy_d2h(C.c_void_p(offs.ctypes.data), op, offs.nbytes)
        want = np.bincount((Rk & 0x7FFF).cpu().numpy(), minlength=1 << 15)
        assert np.array_equal(np.diff(offs.astype(np.int64)), want) and offs[-1] == n
        host = np.empty((n, 2), dtype=np.int32)
        L.gj_memcpy_d2h(C.c_void_p(host.ctypes.data), tp, host.nbytes)
        pid = host[:, 0] & 0x7FFF
        assert np.all(np.diff(pid) >= 0)
        assert int(host[:, 0].astype(np.int64).sum()) == int(Rk.sum(dtype=torch.int64))
        assert int(host[:, 1].astype(np.int64).sum()) == int(Rp.sum(dtype=torch.int64))
        assert np.array_equal(orc.payload_of_keys(host[:1 << 16, 0].copy(), 40), host[:1 << 16, 1])


def test_config2_workload_A_16M_256M(gj, orc, torch_cuda):
    """ETHZ workload A: |R|=2^24 unique, |S|=2^28 in the reference FK pattern
    (generator_ETHZ.cu:127-149 with maxid=|R|): matches = 2^28 - 15."""
    torch = torch_cuda
    nR, nS = 1 << 24, 1 << 28
    S = gj.generator.create_relation_unique_parallel(nS, nR, 3)
    with gj.JoinEngine(nR, nS, 0) as eng:
        Rk, Rp = gen(torch, eng, nR, nR, 8, 80)
        Sk = torch.from_numpy(S).cuda()
        Sp = torch.ones(nS, dtype=torch.int32, device="cuda")
        res = eng.join_aggregate(Rk, Rp, Sk, Sp)
        assert res.matches == nS - (nS - 1) // nR == 268_435_441
        # checksum: every key k in 1..nR-1 occurs 16 times in S, key 0 once, key nR (15x) has no partner
        pay = orc.payload_of_keys(np.arange(nR, dtype=np.int32), 80).astype(np.int64)
        want = (int(pay[0]) + 16 * int(pay[1:].sum())) % 2**64
        assert res.checksum == want
        assert res.timings.radix_bits == 12


@pytest.mark.parametrize("z", [0.5, 1.0])
def test_config4_zipf_128M(gj, orc, torch_cuda, z):
    """Workload B with a Zipf-skewed probe side (z = 0.5, 1.0): matches = n - #{S == n}
    (alphabet 1..n, key n has no partner); checksum via bincount of S against f(key)."""
    torch = torch_cuda
    n = 128_000_000
    S = gj.generator.create_relation_zipf_parallel(n, n, z, 7)
    with gj.JoinEngine(n, n, 0) as eng:
        Rk, Rp = gen(torch, eng, n, n, 4, 40)
        Sk = torch.from_numpy(S).cuda()
        Sp = torch.ones(n, dtype=torch.int32, device="cuda")
        res = eng.join_aggregate(Rk, Rp, Sk, Sp)
        assert res.matches == n - int((S == n).sum())
        cnt = torch.bincount(Sk.to(torch.int64), minlength=n + 1)[:n].cpu().numpy()
        nz = np.nonzero(cnt)[0]
        pay = orc.payload_of_keys(nz.astype(np.int32), 40).astype(np.int64)
        want = int((pay * cnt[nz]).sum()) % 2**64     # fits: |pay| < 2^31, sum of counts < 2^27
        assert res.checksum == want
