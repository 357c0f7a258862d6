"""SURVEY.md section 8f rows, through the C ABI against the oracle:
  * late-materialisation join (rank 1): gj_join_aggregate_late vs oracle.join_late, the restatement of
    join_partitioned_varpayload (join-primitives.cu:1420-1557);
  * non-partitioned baselines (rank 4): gj_join_aggregate_nopart (build_ht_chains / chains_probing,
    join-primitives.cu:681-742) and gj_join_aggregate_perfect (build_perfect_array / probe_perfect_array,
    join-primitives.cu:628-668) vs the oracle's join checker;
  * out-of-HBM probe side (rank 3): gj_join_aggregate_stream_host (outOfGPU_Join3_payload,
    hash_join_clustered_probe.cu:1684-1984) vs the oracle's join checker."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    return torch


def dev(torch, *arrs):
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda() for a in arrs]


@pytest.mark.parametrize("nR,nS,keys,cr,cs", [(300_000, 700_000, 1 << 18, 3, 2), (700_000, 300_000, 1 << 18, 1, 4),
                                              (1 << 20, 1 << 20, 1 << 20, 2, 0), (50_000, 2_000_000, 1 << 12, 0, 1),
                                              (1000, 1000, 1 << 30, 2, 2), (0, 1000, 16, 1, 1)])
def test_late_materialisation_matches_oracle(gj, orc, torch_cuda, nR, nS, keys, cr, cs):
    """Row-id payloads, column-major side tables with 0-4 columns per side, build side on R or on S
    (the engine builds on the smaller relation), N:M matches, negative side-table values."""
    rng = np.random.default_rng(nR + 3 * nS + cr)
    Rk = rng.integers(-keys // 2, keys // 2, nR).astype(np.int32)
    Sk = rng.integers(-keys // 2, keys // 2, nS).astype(np.int32)
    Rid, Sid = rng.permutation(nR).astype(np.int32), rng.permutation(nS).astype(np.int32)
    Dr = rng.integers(-2**31, 2**31, (cr, max(nR, 1))).astype(np.int32)
    Ds = rng.integers(-2**31, 2**31, (cs, max(nS, 1))).astype(np.int32)
    want_n, want_sum = orc.join_late(Rk, Rid, Sk, Sid, Dr, Ds) if nR and nS else (0, 0)
    with gj.JoinEngine(max(nR, 1), max(nS, 1), 0) as eng:
        dRk, dRid, dSk, dSid = dev(torch_cuda, Rk, Rid, Sk, Sid)
        dDr = torch_cuda.from_numpy(Dr).cuda()
        dDs = torch_cuda.from_numpy(Ds).cuda()
        got = eng.join_aggregate_late(dRk, dRid, dSk, dSid, dDr, dDs)
        assert (got.matches, got.checksum) == (want_n, want_sum)
        # with the payload product instead, the same engine still gives the plain aggregate
        plain = eng.join_aggregate(dRk, dRid, dSk, dSid)
        assert plain.matches == want_n


@pytest.mark.parametrize("nR,nS,lo,hi", [(1 << 20, 1 << 20, 0, 1 << 20), (300_000, 700_000, -(1 << 17), 1 << 17),
                                         (700_000, 300_000, 0, 1 << 30), (5, 1_000_000, 0, 4), (1, 1, 7, 8), (0, 10, 0, 4)])
def test_nonpartitioned_baseline_matches_oracle(gj, orc, torch_cuda, nR, nS, lo, hi):
    """Global chained hash table, no radix pass: same matches / checksum as the oracle (and as the
    partitioned path); dense, sparse, signed, heavily duplicated keys; build side on R or on S."""
    rng = np.random.default_rng(nR + 7 * nS)
    Rk = rng.integers(lo, hi, nR).astype(np.int32)
    Sk = rng.integers(lo, hi, nS).astype(np.int32)
    Rp = rng.integers(-2**31, 2**31, nR).astype(np.int32)
    Sp = rng.integers(-2**31, 2**31, nS).astype(np.int32)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    with gj.JoinEngine(max(nR, 1), max(nS, 1), 0) as eng:
        d = dev(torch_cuda, Rk, Rp, Sk, Sp)
        got = eng.join_aggregate_nopart(*d)
        assert (got.matches, got.checksum) == (want.matches, want.checksum)
        assert got.timings.kernel_launches == (2 if nR and nS else 0)
        part = eng.join_aggregate(*d)
        assert (part.matches, part.checksum) == (want.matches, want.checksum)


@pytest.mark.parametrize("nR,nS,key_min,spread", [(1 << 20, 1 << 20, 0, 1), (300_000, 2_000_000, -150_000, 1),
                                                   (2_000_000, 300_000, 1 << 30, 3), (1, 1000, -5, 1), (0, 10, 0, 1)])
def test_perfect_array_matches_oracle(gj, orc, torch_cuda, nR, nS, key_min, spread):
    """Perfect array: unique build keys in [key_min, key_min + range) (dense or every `spread`-th value),
    probe keys partly outside the range, every int32 payload incl. -1 (which the reference's payload+1
    sentinel loses, join-primitives.cu:640); the build side is the smaller relation."""
    rng = np.random.default_rng(nR + 11 * nS)
    nb, npr = min(nR, nS), max(nR, nS)
    rangev = max(nb * spread, 1)
    bk = (key_min + spread * rng.permutation(nb).astype(np.int64)).astype(np.int32)
    pk = rng.integers(key_min - rangev // 8, key_min + rangev + rangev // 8 + 1, npr).astype(np.int64)
    pk = np.clip(pk, -2**31, 2**31 - 1).astype(np.int32)
    bp = rng.integers(-2**31, 2**31, nb).astype(np.int32)
    if nb:
        bp[:: max(1, nb // 7)] = -1
    pp = rng.integers(-2**31, 2**31, npr).astype(np.int32)
    Rk, Rp, Sk, Sp = (bk, bp, pk, pp) if nR <= nS else (pk, pp, bk, bp)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    with gj.JoinEngine(max(nR, 1), max(nS, 1), 0) as eng:
        d = dev(torch_cuda, Rk, Rp, Sk, Sp)
        got = eng.join_aggregate_perfect(*d, key_min, rangev)
        assert (got.matches, got.checksum) == (want.matches, want.checksum)
        assert got.timings.kernel_launches == (2 if nR and nS else 0)


def test_perfect_array_rejects_violated_precondition(gj, torch_cuda):
    """Duplicate or out-of-range build keys are detected on the device and reported, never silently joined."""
    Rk = np.array([0, 1, 2, 2, 4], dtype=np.int32)          # duplicate key 2
    Sk = np.arange(10, dtype=np.int32)
    ones = lambda n: np.ones(n, dtype=np.int32)  # noqa: E731
    with gj.JoinEngine(16, 16, 0) as eng:
        d = dev(torch_cuda, Rk, ones(5), Sk, ones(10))
        with pytest.raises(gj.GJError, match="duplicate"):
            eng.join_aggregate_perfect(*d, 0, 8)
        d = dev(torch_cuda, np.array([0, 1, 9], dtype=np.int32), ones(3), Sk, ones(10))
        with pytest.raises(gj.GJError, match="outside"):
            eng.join_aggregate_perfect(*d, 0, 8)             # key 9 outside [0, 8)
        d = dev(torch_cuda, np.array([3, 1, 7], dtype=np.int32), ones(3), Sk, ones(10))
        r = eng.join_aggregate_perfect(*d, 0, 8)              # the context stays usable
        assert (r.matches, r.checksum) == (3, 3)


# ------------------------------------------------------------------------------- out-of-HBM probe side
@pytest.mark.parametrize("nR,nS,chunk", [(200_000, 1_750_000, 500_000), (1 << 20, 1 << 20, 1 << 19), (300_000, 100_000, 250_000),
                                         (50_000, 1_000_001, 1), (1000, 0, 100)])
def test_streamed_probe_side_matches_oracle(gj, orc, torch_cuda, nR, nS, chunk):
    """gj_join_aggregate_stream_host (reference outOfGPU_Join3_payload,
    hash_join_clustered_probe.cu:1684-1984): R resident, S streamed from host memory through a double
    buffer; ragged last chunk, a single chunk, S smaller than one chunk, empty S."""
    if chunk == 1:
        chunk = 333_333
    rng = np.random.default_rng(nR + nS + chunk)
    Rk = rng.integers(0, 1 << 19, nR).astype(np.int32)
    Sk = rng.integers(0, 1 << 19, nS).astype(np.int32)
    Rp = rng.integers(-2**31, 2**31, nR).astype(np.int32)
    Sp = rng.integers(-2**31, 2**31, nS).astype(np.int32)
    want = orc.join_check(Rk, Rp, Sk, Sp) if nS else None
    with gj.JoinEngine(nR, 2 * chunk, 0) as eng:
        got = eng.join_aggregate_stream_host(Rk, Rp, Sk, Sp, chunk)
        assert (got.matches, got.checksum) == ((want.matches, want.checksum) if nS else (0, 0))
        if nS:
            again = eng.join_aggregate_stream_host(Rk, Rp, Sk, Sp, chunk)      # buffers and events are reused
            assert (again.matches, again.checksum) == (want.matches, want.checksum)
