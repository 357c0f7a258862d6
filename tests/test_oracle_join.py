"""Pins the multithreaded host radix-join checker (oracle/oracle_join.c section 4): brute force on
small inputs, closed-form known answers implied by the reference generator."""
import numpy as np
import pytest


def rnd(rng, n, lo, hi):
    return rng.integers(lo, hi, size=n, dtype=np.int64).astype(np.int32)


@pytest.mark.parametrize("nR,nS,lo,hi", [
    (0, 0, 0, 10), (0, 17, 0, 10), (17, 0, 0, 10), (1, 1, 5, 6),
    (300, 500, 0, 64),                 # heavy duplicates on both sides (N:M)
    (1000, 3000, -2**31, 2**31),       # full range incl. negative keys
    (5000, 5000, -50, 50),
    (40000, 9000, 0, 20000),           # crosses the 1-pass/2-pass boundary (bits > 11)
])
def test_checker_equals_bruteforce(orc, nR, nS, lo, hi):
    rng = np.random.default_rng(nR * 7919 + nS)
    Rk, Sk = rnd(rng, nR, lo, hi), rnd(rng, nS, lo, hi)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    if nR * nS > 6e7:
        Rk, Rp = Rk[:6000], Rp[:6000]
    want = orc.join_naive(Rk, Rp, Sk, Sp)
    for threads in (1, 0):
        assert orc.join_check(Rk, Rp, Sk, Sp, threads) == want


def test_numpy_crosscheck_counts(orc):
    rng = np.random.default_rng(5)
    Rk, Sk = rnd(rng, 20000, 0, 3000), rnd(rng, 50000, 0, 3000)
    ones_r, ones_s = np.ones_like(Rk), np.ones_like(Sk)
    cr, cs = np.bincount(Rk, minlength=3000), np.bincount(Sk, minlength=3000)
    res = orc.join_check(Rk, ones_r, Sk, ones_s)
    assert res.matches == int((cr * cs).sum()) == res.checksum
    assert res.ref_results_int32 == res.matches


def test_known_answer_unique_unique(orc):
    # two permutations of 0..n-1 -> exactly n matches; all-ones payload -> "n results"
    n = 1 << 16
    R, S = orc.random_unique_gen(n, n, 1), orc.random_unique_gen(n, n, 2)
    res = orc.join_check(R, np.ones(n, np.int32), S, np.ones(n, np.int32))
    assert (res.matches, res.checksum, res.ref_results_int32) == (n, n, n)
    # row-id payloads: checksum = sum over keys of rowR(k)*rowS(k)
    rid = np.arange(n, dtype=np.int32)
    invR, invS = np.empty(n, np.int64), np.empty(n, np.int64)
    invR[R], invS[S] = rid, rid
    res = orc.join_check(R, rid, S, rid)
    assert res.matches == n and res.checksum == int((invR * invS).sum()) % 2**64


def test_known_answer_fk_pattern(orc):
    # S = create_relation_unique(nS, maxid=nR): matches = nS - floor((nS-1)/nR)
    nR, nS = 1 << 12, (1 << 16)
    R, S = orc.random_unique_gen(nR, nR, 3), orc.random_unique_gen(nS, nR, 4)
    res = orc.join_check(R, np.ones(nR, np.int32), S, np.ones(nS, np.int32))
    assert res.matches == nS - (nS - 1) // nR


def test_known_answer_zipf(orc):
    # zipf alphabet is 1..nR -> matches = nS - #{S == nR}
    nR, nS = 4096, 30000
    R = orc.random_unique_gen(nR, nR, 6)
    orc.seed_generator(9)
    S = orc.gen_zipf(nS, nR, 1.0)
    assert S.min() >= 1 and S.max() <= nR
    res = orc.join_check(R, np.ones(nR, np.int32), S, np.ones(nS, np.int32))
    assert res.matches == nS - int((S == nR).sum())


def test_int32_wrap_rule(orc):
    # reference accumulates in int32 (join-primitives.cu:914,1092): low 32 bits of checksum64
    n = 3000
    k = np.arange(n, dtype=np.int32)
    p = np.full(n, 2**31 - 1, np.int32)
    res = orc.join_check(k, p, k, p)
    want = (n * (2**31 - 1) ** 2) % 2**64
    assert res.checksum == want
    lo = want & 0xFFFFFFFF
    assert res.ref_results_int32 == (lo - 2**32 if lo >= 2**31 else lo)


def test_materialize_and_pairhash(orc):
    rng = np.random.default_rng(11)
    Rk, Sk = rnd(rng, 2000, 0, 500), rnd(rng, 3000, 0, 500)
    Rp, Sp = np.arange(2000, dtype=np.int32), np.arange(3000, dtype=np.int32) + 10000
    n, orp, osp, res = orc.join_materialize(Rk, Rp, Sk, Sp, cap=1 << 20)
    assert n == res.matches == len(orp)
    assert orc.pairs_hash(orp, osp) == res.pairhash
    # pairs are (row in R, row in S+10000): verify each is a true match, and all are distinct
    assert np.array_equal(Rk[orp], Sk[osp - 10000])
    assert len(set(zip(orp.tolist(), osp.tolist()))) == n
    # capped: exact count is still returned
    n2, orp2, _, _ = orc.join_materialize(Rk, Rp, Sk, Sp, cap=100)
    assert n2 == n and len(orp2) == 100


@pytest.mark.parametrize("shift,bits", [(0, 1), (0, 8), (5, 8), (24, 8), (3, 11)])
def test_partition_oracle(orc, shift, bits):
    rng = np.random.default_rng(shift * 31 + bits)
    k, p = rnd(rng, 10000, -2**31, 2**31), rnd(rng, 10000, -2**31, 2**31)
    off, ko, po = orc.partition(k, p, shift, bits)
    d = (k.view(np.uint32) >> np.uint32(shift)) & np.uint32((1 << bits) - 1)
    order = np.argsort(d, kind="stable")
    assert np.array_equal(ko, k[order]) and np.array_equal(po, p[order])
    assert np.array_equal(off[1:] - off[:-1], np.bincount(d, minlength=1 << bits).astype(np.uint64))
    cnt, hsh = orc.partition_fingerprint(k, p, shift, bits)
    cnt2, hsh2 = orc.partition_fingerprint(ko, po, shift, bits)
    assert np.array_equal(cnt, cnt2) and np.array_equal(hsh, hsh2)


def test_late_materialisation_oracle_against_brute_force(orc):
    """oracle.join_late (restating join_partitioned_varpayload, join-primitives.cu:1420-1557: payload = row id,
    side-table values of both rows added per result pair) against a numpy nested loop; N:M keys,
    negative values, zero columns on either side."""
    rng = np.random.default_rng(12)
    nR, nS = 2500, 4000
    Rk = rng.integers(-300, 300, nR).astype(np.int32)
    Sk = rng.integers(-300, 300, nS).astype(np.int32)
    Rid, Sid = rng.permutation(nR).astype(np.int32), rng.permutation(nS).astype(np.int32)
    Dr = rng.integers(-2**31, 2**31, (3, nR)).astype(np.int32)
    Ds = rng.integers(-2**31, 2**31, (2, nS)).astype(np.int32)
    ri, si = np.nonzero(Rk[:, None] == Sk[None, :])
    for dr, ds in ((Dr, Ds), (Dr[:0], Ds), (Dr, Ds[:0]), (Dr[:1], Ds[:1])):
        want = (int(dr[:, Rid[ri]].astype(object).sum()) + int(ds[:, Sid[si]].astype(object).sum())) % (1 << 64)
        n, tot = orc.join_late(Rk, Rid, Sk, Sid, dr, ds)
        assert n == len(ri) and tot == want
    # the reference accumulates an int32 (:1460): its value is the low 32 bits of ours
    n, tot = orc.join_late(Rk, Rid, Sk, Sid, Dr, Ds)
    ref32 = (Dr[:, Rid[ri]].astype(np.int64).sum() + Ds[:, Sid[si]].astype(np.int64).sum()) & 0xFFFFFFFF
    assert tot & 0xFFFFFFFF == int(ref32)
