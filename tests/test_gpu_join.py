"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI against the CPU
oracle on identical seeded inputs.  Bit-exact: match count, 64-bit checksum, pair multiset hash,
partition offsets and per-partition multisets.  Reference semantics: join-primitives.cu:885-1095
(aggregate), :1107-1416 (materialise), :58-535 (partition passes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    return torch


def dev(torch, *arrs):
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda() for a in arrs]


def rnd(rng, n, lo, hi):
    return rng.integers(lo, hi, size=n, dtype=np.int64).astype(np.int32)


@pytest.fixture(scope="module")
def eng(gj, torch_cuda):
    e = gj.JoinEngine(1 << 22, 1 << 23, 0)
    yield e
    e.close()


def reset(eng):
    for k in ("radix_bits", "pass1_bits", "join_cfg", "unit_tuples", "gpu_bits"):
        eng.set_option(k, 0)
    eng.set_option("scatter_cfg", 255)   # auto
    eng.set_option("part_target", 4096)
    eng.set_option("nopart_max", 0)      # these tests exercise the partitioned path at every size


# ------------------------------------------------------------------------------- partitioner
@pytest.mark.parametrize("n,bits", [(0, 4), (1, 1), (5, 3), (4096, 8), (4097, 8), (100_003, 0),
                                    (1 << 20, 8), (300_000, 11), (1_000_000, 15), (777_777, 13), (2_000_003, 16)])
def test_partition_matches_oracle(gj, orc, eng, torch_cuda, n, bits):
    reset(eng)
    rng = np.random.default_rng(n + bits)
    keys, pays = rnd(rng, n, -2**31, 2**31), rnd(rng, n, -2**31, 2**31)
    dk, dp = dev(torch_cuda, keys, pays)
    tup, off, B, t = eng.partition(dk, dp, bits)
    if bits:
        assert B == bits
    want_off, _, _ = orc.partition(keys, pays, 0, B)
    assert np.array_equal(off, want_off.astype(np.int64))
    ok, op = np.ascontiguousarray(tup[:, 0]), np.ascontiguousarray(tup[:, 1])
    # every tuple sits inside the range of its own partition ...
    pid = np.repeat(np.arange(1 << B), np.diff(off))
    assert np.array_equal((ok.view(np.uint32) & ((1 << B) - 1)).astype(np.int64), pid)
    # ... and each partition holds exactly the oracle's multiset
    c1, h1 = orc.partition_fingerprint(keys, pays, 0, B)
    c2, h2 = orc.partition_fingerprint(ok, op, 0, B)
    assert np.array_equal(c1, c2) and np.array_equal(h1, h2)


@pytest.mark.parametrize("cfg", range(4))
def test_partition_every_scatter_variant(gj, orc, eng, torch_cuda, cfg):
    reset(eng)
    eng.set_option("scatter_cfg", cfg)
    rng = np.random.default_rng(cfg)
    n = 600_011
    keys, pays = rnd(rng, n, 0, 1 << 20), np.arange(n, dtype=np.int32)
    dk, dp = dev(torch_cuda, keys, pays)
    for bits in (7, 12):
        tup, off, B, _ = eng.partition(dk, dp, bits)
        want_off, _, _ = orc.partition(keys, pays, 0, B)
        assert np.array_equal(off, want_off.astype(np.int64))
        c1, h1 = orc.partition_fingerprint(keys, pays, 0, B)
        c2, h2 = orc.partition_fingerprint(np.ascontiguousarray(tup[:, 0]), np.ascontiguousarray(tup[:, 1]), 0, B)
        assert np.array_equal(c1, c2) and np.array_equal(h1, h2)
    reset(eng)


def test_partition_skewed_digit(gj, orc, eng, torch_cuda):
    """Most of a tile falls into one digit (Zipf-like): stresses same-address shared atomics; with
    16 radix bits the histogram uses packed 16-bit counters and the hot bin overflows its field
    (> 32768 hits per CTA) many times."""
    reset(eng)
    rng = np.random.default_rng(9)
    for n, bits, slot in ((500_000, 10, 0), (8_000_000, 16, 1)):
        keys = np.where(rng.random(n) < 0.8, 77, rng.integers(0, 1 << 20, n)).astype(np.int32)
        pays = np.arange(n, dtype=np.int32)
        tup, off, B, _ = eng.partition(*dev(torch_cuda, keys, pays), bits, slot=slot)
        want_off, _, _ = orc.partition(keys, pays, 0, B)
        assert np.array_equal(off, want_off.astype(np.int64))
        c1, h1 = orc.partition_fingerprint(keys, pays, 0, B)
        c2, h2 = orc.partition_fingerprint(np.ascontiguousarray(tup[:, 0]), np.ascontiguousarray(tup[:, 1]), 0, B)
        assert np.array_equal(c1, c2) and np.array_equal(h1, h2)


# ------------------------------------------------------------------------------- aggregate join
CASES = [
    (0, 0, 0, 10), (0, 17, 0, 10), (17, 0, 0, 10), (1, 1, 5, 6), (3, 2, 0, 2),
    (300, 500, 0, 64),                       # heavy duplicates on both sides (N:M)
    (1000, 3000, -2**31, 2**31),             # full key range incl. negative keys
    (5000, 5000, -50, 50),
    (40_000, 9_000, 0, 20_000),              # build side becomes S (smaller)
    (200_000, 700_000, 0, 150_000),
    (1 << 20, 1 << 20, 0, 1 << 20),
    (3_000_000, 5_000_000, -2**31, 2**31),   # two passes
]


@pytest.mark.parametrize("nR,nS,lo,hi", CASES)
def test_join_aggregate_matches_oracle(gj, orc, eng, torch_cuda, nR, nS, lo, hi):
    reset(eng)
    rng = np.random.default_rng(nR * 7919 + nS)
    Rk, Sk = rnd(rng, nR, lo, hi), rnd(rng, nS, lo, hi)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    got = eng.join_aggregate(*dev(torch_cuda, Rk, Rp, Sk, Sp))
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    assert got.ref_results_int32 == want.ref_results_int32
    if nR and nS:
        assert got.timings.kernel_launches >= 5


def test_config1_reference_workload(gj, orc, eng, torch_cuda):
    """BASELINE config 1: |R|=|S|=2^20 unique keys (generator_ETHZ.cu:127-149), FK join.
    Known answer: 2^20 matches; with all-ones payloads the reference prints `1048576 results`."""
    reset(eng)
    n = 1 << 20
    R = gj.generator.create_relation_unique(n, n, 1)
    S = gj.generator.create_relation_unique(n, n, 2)
    assert np.array_equal(R, orc.random_unique_gen(n, n, 1))      # product generator == oracle
    ones = np.ones(n, np.int32)
    got = eng.join_aggregate(*dev(torch_cuda, R, ones, S, ones))
    assert (got.matches, got.checksum, got.ref_results_int32) == (n, n, n)
    rid = np.arange(n, dtype=np.int32)
    want = orc.join_check(R, rid, S, rid)
    got = eng.join_aggregate(*dev(torch_cuda, R, rid, S, rid))
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    assert got.timings.radix_bits == 8 and got.timings.pass2_bits == 0


def test_small_build_sides_take_the_nonpartitioned_path_by_default(gj, orc, torch_cuda):
    """Default dispatch (option "nopart_max" = 2^21, the measured crossover): a small build side joins through the
    global hash table (2 launches, no radix plan), a larger one or a forced plan through the radix passes --
    same aggregate either way."""
    rng = np.random.default_rng(31)
    with gj.JoinEngine(1 << 22, 1 << 22, 0) as e:
        assert e.get_option("nopart_max") == 1 << 21
        for nR, nS, nopart in ((1 << 20, 1 << 20, True), (1000, 3_000_000, True), ((1 << 21) + 1, (1 << 21) + 1, False)):
            Rk, Sk = rnd(rng, nR, 0, nR), rnd(rng, nS, 0, nR)
            Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
            want = orc.join_check(Rk, Rp, Sk, Sp)
            d = dev(torch_cuda, Rk, Rp, Sk, Sp)
            got = e.join_aggregate(*d)
            assert (got.matches, got.checksum) == (want.matches, want.checksum)
            assert (got.timings.kernel_launches == 2 and got.timings.radix_bits == 0) == nopart
            e.set_option("radix_bits", 9)
            forced = e.join_aggregate(*d)
            assert (forced.matches, forced.checksum) == (want.matches, want.checksum) and forced.timings.radix_bits == 9
            e.set_option("radix_bits", 0)


def test_fk_pattern_known_answer(gj, orc, eng, torch_cuda):
    """S = create_relation_unique(nS, maxid=nR) -> matches = nS - floor((nS-1)/nR)."""
    reset(eng)
    nR, nS = 1 << 18, 1 << 22
    R = gj.generator.create_relation_unique(nR, nR, 3)
    S = gj.generator.create_relation_unique(nS, nR, 4)
    got = eng.join_aggregate(*dev(torch_cuda, R, np.ones(nR, np.int32), S, np.ones(nS, np.int32)))
    assert got.matches == nS - (nS - 1) // nR == got.checksum


@pytest.mark.parametrize("z", [0.5, 1.0])
def test_zipf_probe_side(gj, orc, eng, torch_cuda, z):
    """BASELINE config 4 (scaled down): Zipf-skewed probe side -> unit splitting of hot partitions."""
    reset(eng)
    nR, nS = 1 << 21, 1 << 22
    R = gj.generator.create_relation_unique_parallel(nR, nR, 6)
    S = gj.generator.create_relation_zipf_parallel(nS, nR, z, 7)
    Rp, Sp = np.arange(nR, dtype=np.int32), np.arange(nS, dtype=np.int32) * 3
    want = orc.join_check(R, Rp, S, Sp)
    got = eng.join_aggregate(*dev(torch_cuda, R, Rp, S, Sp))
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    assert got.matches == nS - int((S == nR).sum())


def test_build_partition_larger_than_shared_memory(gj, orc, eng, torch_cuda):
    """Heavy-hitter BUILD keys: one partition exceeds the shared-memory table and is joined in
    rounds (the reference's block-nested branch, join-primitives.cu:929-1003)."""
    reset(eng)
    rng = np.random.default_rng(3)
    nR, nS = 60_000, 90_000
    Rk = np.where(rng.random(nR) < 0.5, 42, rng.integers(0, 1000, nR)).astype(np.int32)
    Sk = np.where(rng.random(nS) < 0.01, 42, rng.integers(0, 1000, nS)).astype(np.int32)
    Rp, Sp = rnd(rng, nR, -1000, 1000), rnd(rng, nS, -1000, 1000)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    for jc in (0, 1, 2):
        eng.set_option("join_cfg", jc)
        got = eng.join_aggregate(*dev(torch_cuda, Rk, Rp, Sk, Sp))
        assert (got.matches, got.checksum) == (want.matches, want.checksum), jc
    reset(eng)


def test_every_radix_split_and_join_variant(gj, orc, eng, torch_cuda):
    reset(eng)
    rng = np.random.default_rng(11)
    nR, nS = 400_000, 900_000
    Rk, Sk = rnd(rng, nR, 0, 300_000), rnd(rng, nS, 0, 300_000)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    d = dev(torch_cuda, Rk, Rp, Sk, Sp)
    for bits in (1, 5, 8, 9, 12, 15, 16):
        eng.set_option("radix_bits", bits)
        got = eng.join_aggregate(*d)
        assert (got.matches, got.checksum) == (want.matches, want.checksum), bits
        assert got.timings.radix_bits == bits
    eng.set_option("radix_bits", 14)
    for p1 in (6, 7, 8):
        eng.set_option("pass1_bits", p1)
        got = eng.join_aggregate(*d)
        assert (got.matches, got.checksum) == (want.matches, want.checksum), p1
        assert (got.timings.pass1_bits, got.timings.pass2_bits) == (p1, 14 - p1)
    reset(eng)
    for jc in range(eng.get_option("num_join_cfgs")):
        eng.set_option("join_cfg", jc)
        for unit in (0, 1024, 5000):
            eng.set_option("unit_tuples", unit)
            got = eng.join_aggregate(*d)
            assert (got.matches, got.checksum) == (want.matches, want.checksum), (jc, unit)
    reset(eng)


@pytest.mark.parametrize("bits", [17, 19, 21])
def test_three_pass_partitioning(gj, orc, eng, torch_cuda, bits):
    """More than 16 radix bits (what build sides beyond 2^28 tuples get): third pass on the low bits
    of every second-level partition.  Forced here on small inputs, incl. skew and both input layouts."""
    reset(eng)
    rng = np.random.default_rng(bits)
    nR, nS = 1_500_000, 3_000_000
    Rk = rnd(rng, nR, -2**31, 2**31)
    Sk = np.concatenate([Rk[rng.integers(0, nR, nS // 2)], rnd(rng, nS - nS // 2, -2**31, 2**31)])
    Sk[: 200_000] = Rk[7]                        # one hot key
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    eng.set_option("radix_bits", bits)
    got = eng.join_aggregate(*dev(torch_cuda, Rk, Rp, Sk, Sp))
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    assert (got.timings.radix_bits, got.timings.pass1_bits, got.timings.pass2_bits, got.timings.pass3_bits) == (bits, 8, 8, bits - 16)
    Rt = torch_cuda.from_numpy(np.stack([Rk, Rp], axis=1).copy()).cuda()
    St = torch_cuda.from_numpy(np.stack([Sk, Sp], axis=1).copy()).cuda()
    got = eng.join_aggregate_tuples(Rt, nR, St, nS)
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    out_r = torch_cuda.empty(int(want.matches), dtype=torch_cuda.int32, device="cuda")
    out_s = torch_cuda.empty(int(want.matches), dtype=torch_cuda.int32, device="cuda")
    n, res = eng.join_materialize(*dev(torch_cuda, Rk, Rp, Sk, Sp), out_r, out_s)
    assert n == want.matches and orc.pairs_hash(out_r.cpu().numpy(), out_s.cpu().numpy()) == want.pairhash
    reset(eng)


def test_packed_tuple_entry(gj, orc, eng, torch_cuda):
    reset(eng)
    rng = np.random.default_rng(12)
    nR, nS = 333_333, 1_000_001
    Rk, Sk = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    Sk[: nR] = Rk    # guarantee matches
    Rp, Sp = rnd(rng, nR, -9, 9), rnd(rng, nS, -9, 9)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    Rt = torch_cuda.from_numpy(np.stack([Rk, Rp], axis=1).copy()).cuda()
    St = torch_cuda.from_numpy(np.stack([Sk, Sp], axis=1).copy()).cuda()
    got = eng.join_aggregate_tuples(Rt, nR, St, nS)
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    # packed input that is only 8-byte aligned (odd tuple offset inside a larger buffer)
    Rt1 = torch_cuda.zeros((nR + 1, 2), dtype=torch_cuda.int32, device="cuda")
    St1 = torch_cuda.zeros((nS + 3, 2), dtype=torch_cuda.int32, device="cuda")
    Rt1[1:] = Rt
    St1[3:] = St
    got = eng.join_aggregate_tuples(Rt1[1:], nR, St1[3:], nS)
    assert (got.matches, got.checksum) == (want.matches, want.checksum)


def test_host_entry_end_to_end(gj, orc, eng, torch_cuda):
    reset(eng)
    rng = np.random.default_rng(13)
    nR, nS = 1_500_000, 2_500_000
    Rk, Sk = rnd(rng, nR, 0, 1 << 21), rnd(rng, nS, 0, 1 << 21)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    eng.set_option("h2d_chunk", 300_000)    # many chunks: histogram chases the copies
    got = eng.join_aggregate_host(Rk, Rp, Sk, Sp)
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    assert got.timings.h2d_ms > 0
    pin = [torch_cuda.from_numpy(a).pin_memory() for a in (Sk, Sp, Rk, Rp)]   # swapped roles too
    got = eng.join_aggregate_host(*pin)
    assert (got.matches, got.checksum) == (want.matches, want.checksum)
    eng.set_option("h2d_chunk", 8 << 20)


def test_staged_pipeline_equals_monolithic(gj, orc, eng, torch_cuda):
    """gj_stage_begin/partition/join/finish (the multi-GPU overlap entry points): the two sides are
    partitioned on different streams, in either order, and give the oracle's result."""
    reset(eng)
    torch = torch_cuda
    rng = np.random.default_rng(41)
    for nR, nS in ((600_000, 900_001), (1_200_000, 300_000), (5, 0)):
        Rk, Sk = rnd(rng, nR, 0, 1 << 19), rnd(rng, nS, 0, 1 << 19)
        Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
        want = orc.join_check(Rk, Rp, Sk, Sp)
        Rt = torch.from_numpy(np.stack([Rk, Rp], axis=1).copy()).cuda()
        St = torch.from_numpy(np.stack([Sk, Sp], axis=1).copy()).cuda()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
        torch.cuda.synchronize()
        for order in ((0, 1), (1, 0)):
            eng.stage_begin(nR, nS, s1)
            for side in order:
                eng.stage_partition(side, (Rt if side == 0 else St).data_ptr(), s2 if side == 0 else s1)
            eng.stage_join(s1)
            assert eng.stage_finish() == (want.matches, want.checksum)
    with pytest.raises(gj.GJError):
        eng.stage_join(torch.cuda.Stream())        # no stage_begin


# ------------------------------------------------------------------------------- materialise
@pytest.mark.parametrize("nR,nS,hi", [(0, 5, 4), (2000, 3000, 50), (1 << 18, 1 << 19, 1 << 18), (900_000, 300_000, 1 << 19)])
def test_materialize_pairs_match_oracle(gj, orc, eng, torch_cuda, nR, nS, hi):
    reset(eng)
    rng = np.random.default_rng(nR + nS)
    Rk, Sk = rnd(rng, nR, 0, hi), rnd(rng, nS, 0, hi)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    cap = max(int(want.matches), 1)
    out_r = torch_cuda.full((cap,), 123456789, dtype=torch_cuda.int32, device="cuda")
    out_s = torch_cuda.full((cap,), 123456789, dtype=torch_cuda.int32, device="cuda")
    n, res = eng.join_materialize(*dev(torch_cuda, Rk, Rp, Sk, Sp), out_r, out_s)
    assert n == want.matches and res.checksum == want.checksum
    k = int(want.matches)
    assert orc.pairs_hash(out_r.cpu().numpy()[:k], out_s.cpu().numpy()[:k]) == want.pairhash


def test_materialize_capped_output_keeps_exact_count(gj, orc, eng, torch_cuda):
    """The reference's ring overwrites itself (join-primitives.cu:1097-1099); here pairs beyond
    `cap` are counted but not written, and nothing is written past the buffer."""
    reset(eng)
    rng = np.random.default_rng(21)
    nR, nS = 50_000, 80_000
    Rk, Sk = rnd(rng, nR, 0, 5000), rnd(rng, nS, 0, 5000)
    Rp, Sp = rnd(rng, nR, 1, 100), rnd(rng, nS, 1, 100)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    cap = 10_000
    guard = 4096
    out_r = torch_cuda.zeros(cap + guard, dtype=torch_cuda.int32, device="cuda")
    out_s = torch_cuda.zeros(cap + guard, dtype=torch_cuda.int32, device="cuda")
    n, res = eng.join_materialize(*dev(torch_cuda, Rk, Rp, Sk, Sp), out_r[:cap], out_s[:cap])
    assert n == want.matches > cap and res.checksum == want.checksum
    assert int((out_r[:cap] > 0).sum()) == cap and int(out_r[cap:].abs().sum()) == 0
    assert int(out_s[cap:].abs().sum()) == 0


# ------------------------------------------------------------------------------- multi-GPU step
@pytest.mark.parametrize("G", [1, 2, 8])
def test_virtual_shards_on_one_gpu(gj, orc, eng, torch_cuda, G):
    """the shuffle logic with G virtual shards on one GPU -- split both
    relations by destination, join every destination's tuples with gpu_bits set, add up."""
    reset(eng)
    rng = np.random.default_rng(G)
    nR, nS = 700_000, 1_300_000
    Rk, Sk = rnd(rng, nR, 0, 1 << 20), rnd(rng, nS, 0, 1 << 20)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    B = gj.distributed.choose_radix_bits(nR // G)
    gbits = G.bit_length() - 1
    dRk, dRp, dSk, dSp = dev(torch_cuda, Rk, Rp, Sk, Sp)
    outR = torch_cuda.empty(nR, dtype=torch_cuda.int64, device="cuda")
    outS = torch_cuda.empty(nS, dtype=torch_cuda.int64, device="cuda")
    cR = eng.shuffle_split(dRk, dRp, G, B, outR)
    cS = eng.shuffle_split(dSk, dSp, G, B, outS)
    assert np.array_equal(cR, np.bincount((Rk.view(np.uint32) >> B) & (G - 1), minlength=G))
    assert np.array_equal(cS, eng.shuffle_count(dSk, G, B))
    eng.set_option("radix_bits", B)
    eng.set_option("gpu_bits", gbits)
    oR = np.concatenate(([0], np.cumsum(cR)))
    oS = np.concatenate(([0], np.cumsum(cS)))
    m = c = 0
    for g in range(G):
        grp = outR[oR[g]:oR[g + 1]].cpu().numpy().view(np.uint32).reshape(-1, 2)
        assert np.all(((grp[:, 0] >> B) & (G - 1)) == g)
        r = eng.join_aggregate_tuples(outR[oR[g]:], int(cR[g]), outS[oS[g]:], int(cS[g]))
        m += r.matches
        c = (c + r.checksum) % 2**64
    assert (m, c) == (want.matches, want.checksum)
    reset(eng)


def test_peer_store_scatter_into_local_buffers(gj, orc, eng, torch_cuda):
    """gj_shuffle_scatter_peers with every 'peer' being a local buffer: same kernel, same
    per-destination pointer table as over NVLink."""
    reset(eng)
    rng = np.random.default_rng(31)
    n, G, shift = 900_000, 4, 12
    k, p = rnd(rng, n, 0, 1 << 20), rnd(rng, n, -2**31, 2**31)
    dk, dp = dev(torch_cuda, k, p)
    cnt = eng.shuffle_count(dk, G, shift)
    pad = 1000
    bufs = [torch_cuda.zeros(int(cnt[g]) + pad, dtype=torch_cuda.int64, device="cuda") for g in range(G)]
    eng.shuffle_scatter_peers(dk, dp, G, shift, [b.data_ptr() for b in bufs], [pad // 2] * G)
    c_all, h_all = orc.partition_fingerprint(k, p, shift, 2)
    for g in range(G):
        got = bufs[g].cpu().numpy()
        assert not got[: pad // 2].any() and not got[pad // 2 + int(cnt[g]):].any()
        t = got[pad // 2: pad // 2 + int(cnt[g])].view(np.int32).reshape(-1, 2)
        c, h = orc.partition_fingerprint(np.ascontiguousarray(t[:, 0]), np.ascontiguousarray(t[:, 1]), shift, 2)
        assert c[g] == cnt[g] == c_all[g] and h[g] == h_all[g] and c.sum() == c[g]


# ------------------------------------------------------------------------------- sharded "partition, then push"
def _pp_virtual(gj, orc, torch, G, B, Rk, Rp, Sk, Sp, splits=None, slack=1.6, opts=None, check_layout=True):
    """Runs the gj_pp_* pipeline with G virtual ranks on ONE GPU (one engine context per rank, every
    'peer' buffer local): same kernels, same cursor arithmetic, same call sequence as over NVLink;
    the all-gather of the fine histograms is a torch.stack."""
    rels = [(Rk, Rp), (Sk, Sp)]
    n = [len(Rk), len(Sk)]
    if splits is None:
        splits = [np.linspace(0, n[w], G + 1).astype(np.int64) for w in range(2)]
    shard_n = [[int(splits[w][r + 1] - splits[w][r]) for r in range(G)] for w in range(2)]
    nq = G << B
    mask = nq - 1
    # worst destination decides the buffer size (all ranks must agree on it)
    caps = []
    for w in range(2):
        d = ((rels[w][0].view(np.uint32) >> B) & (G - 1)) if n[w] else np.zeros(0, dtype=np.int64)
        caps.append(int(max(np.bincount(d, minlength=G).max() if n[w] else 0, 1) * slack) + 64)
    engs = [gj.JoinEngine(max(max(shard_n[0]), 1), max(max(shard_n[1]), 1), 0, **(opts or {})) for _ in range(G)]
    try:
        own = [[torch.zeros(caps[w] + 16, dtype=torch.int64, device="cuda") for _ in range(G)] for w in range(2)]
        cols = [[dev(torch, rels[w][0][splits[w][r]:splits[w][r + 1]], rels[w][1][splits[w][r]:splits[w][r + 1]])
                 for r in range(G)] for w in range(2)]
        hist = [[torch.empty(nq, dtype=torch.int32, device="cuda") for _ in range(G)] for w in range(2)]
        torch.cuda.synchronize()
        for r in range(G):
            engs[r].pp_begin(n[0], n[1], G, r, B)
            for w in range(2):
                engs[r].pp_local(w, cols[w][r][0], cols[w][r][1], hist[w][r])
        torch.cuda.synchronize()
        allh = [torch.stack(hist[w]).contiguous() for w in range(2)]
        for w in range(2):      # the fine histogram of every shard is exact
            for r in range(G):
                k = rels[w][0][splits[w][r]:splits[w][r + 1]].view(np.uint32)
                assert np.array_equal(hist[w][r].cpu().numpy(), np.bincount(k & mask, minlength=nq))
        for r in range(G):
            for w in range(2):
                engs[r].pp_push(w, allh[w], [t.data_ptr() for t in own[w]], caps[w], shard_n[w][r])
        torch.cuda.synchronize()
        if check_layout:
            for w in range(2):
                ah = allh[w].cpu().numpy()
                c_all, h_all = orc.partition_fingerprint(rels[w][0], rels[w][1], 0, G.bit_length() - 1 + B) if n[w] else (None, None)
                for d in range(G):
                    _, off, cnt = gj.distributed.pp_layout(ah, d, B)
                    tot = int(off[-1])
                    got = own[w][d].cpu().numpy()
                    assert not got[tot:].any()                    # nothing beyond the received tuples
                    t = got[:tot].view(np.int32).reshape(-1, 2)
                    keys = np.ascontiguousarray(t[:, 0])
                    pid = np.repeat(np.arange(1 << B, dtype=np.int64) + (d << B), cnt)
                    assert np.array_equal((keys.view(np.uint32) & mask).astype(np.int64), pid)
                    if tot:
                        c, h = orc.partition_fingerprint(keys, np.ascontiguousarray(t[:, 1]), 0, G.bit_length() - 1 + B)
                        sl = slice(d << B, (d + 1) << B)
                        assert np.array_equal(c[sl], c_all[sl]) and np.array_equal(h[sl], h_all[sl])
        m = c = 0
        got_n = [0, 0]
        for r in range(G):
            engs[r].pp_join(own[0][r].data_ptr(), own[1][r].data_ptr(), caps[0], caps[1])
            mm, cc, a, b, ph = engs[r].pp_finish()
            m += mm
            c = (c + cc) % 2**64
            got_n[0] += a
            got_n[1] += b
        assert got_n == n
        return m, c, engs[0].pp_plan()
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("G,B", [(1, 6), (2, 7), (4, 9), (8, 8), (8, 13)])
def test_pp_virtual_shards(gj, orc, torch_cuda, G, B):
    """gpu bits + local bits <= 16: 8-bit (or smaller) passes."""
    rng = np.random.default_rng(100 * G + B)
    nR, nS = 600_000, 1_100_000
    Rk, Sk = rnd(rng, nR, 0, 1 << 19), rnd(rng, nS, 0, 1 << 19)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    m, c, _ = _pp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, Sk, Sp)
    assert (m, c) == (want.matches, want.checksum)


@pytest.mark.parametrize("G,B,p1,out,big", [(8, 15, 0, 0, 0), (8, 15, 0, 1, 0), (4, 16, 0, 0, 0), (16, 16, 0, 0, 0), (16, 16, 0, 1, 0),
                                            (2, 12, 4, 0, 0), (2, 12, 4, 1, 0), (8, 11, 6, 1, 0), (8, 14, 10, 0, 0),
                                            (8, 15, 0, 1, 1), (8, 15, 10, 0, 1), (16, 16, 0, 1, 1), (16, 16, 0, 0, 1), (2, 9, 0, 1, 1)])
def test_pp_wide_passes_and_tma_output(gj, orc, torch_cuda, G, B, p1, out, big):
    """9- and 10-bit passes (512 / 1024-way tiles), explicit first-pass bits, 8-byte-store and TMA
    bulk-store runs, 8 K and 16 K-tuple push tiles;
    signed keys, N:M matches, ragged shards (one of them empty)."""
    rng = np.random.default_rng(7 * G + B + p1 + out + 3 * big)
    nR, nS = 1_500_000, 2_500_000
    Rk = rnd(rng, nR, -(1 << 21), 1 << 21)
    Sk = rnd(rng, nS, -(1 << 21), 1 << 21)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    splits = []
    for n in (nR, nS):
        cuts = np.sort(rng.integers(0, n, size=G - 1)) if G > 1 else np.zeros(0, dtype=np.int64)
        if G > 2:
            cuts[1] = cuts[0]                   # an empty shard
        splits.append(np.concatenate(([0], cuts, [n])).astype(np.int64))
    opts = {"pp_out": out, "pp_tile16k": big}
    if p1:
        opts["pass1_bits"] = p1
    m, c, plan = _pp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, Sk, Sp, splits=splits, opts=opts)
    assert (m, c) == (want.matches, want.checksum)
    g = G.bit_length() - 1
    assert sum(plan) == g + B and plan[0] >= g and max(plan) <= 10
    if p1:
        assert plan[0] == max(p1, g)


def test_pp_skew_and_empty_relation(gj, orc, torch_cuda):
    """Zipf-like probe side (one destination and one partition far heavier than the rest) and an
    empty relation."""
    rng = np.random.default_rng(5)
    nR, nS, G, B = 400_000, 1_600_000, 4, 8
    Rk = rng.permutation(nR).astype(np.int32)
    hot = rng.integers(0, 50, size=nS // 2)
    Sk = np.concatenate((hot, rng.integers(0, nR, size=nS - nS // 2))).astype(np.int32)
    rng.shuffle(Sk)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    m, c, _ = _pp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, Sk, Sp, slack=1.2)
    assert (m, c) == (want.matches, want.checksum)
    e = np.zeros(0, dtype=np.int32)
    assert _pp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, e, e)[:2] == (0, 0)
    assert _pp_virtual(gj, orc, torch_cuda, G, B, e, e, Sk, Sp)[:2] == (0, 0)


def test_pp_destination_overflow_is_reported(gj, orc, torch_cuda):
    """A destination that would receive more than its buffer holds: nothing is pushed anywhere and
    gj_pp_finish fails (every rank computes the same totals from the gathered histograms)."""
    rng = np.random.default_rng(9)
    n, G, B = 300_000, 4, 8
    k = (rng.integers(0, 1 << 8, size=n) | (2 << 8)).astype(np.int32)     # every key goes to GPU 2
    p = rnd(rng, n, -2**31, 2**31)
    with pytest.raises(gj.GJError):
        _pp_virtual(gj, orc, torch_cuda, G, B, k, p, k, p, slack=0.5, check_layout=False)


# ------------------------------------------------------------------------------- sharded "partition, copy, partition"
def _pcp_virtual(gj, orc, torch, G, B, Rk, Rp, Sk, Sp, splits=None, slack=1.6, opts=None, check_layout=True, stages=(1, 1),
                 peer_hist=False):
    """gj_pcp_* with G virtual ranks on ONE GPU (one engine context per rank, every 'peer' buffer and
    flag word local; the all-gather is a torch.stack).  stages = copy / receive stages of the building
    and of the probing relation."""
    rels = [(Rk, Rp), (Sk, Sp)]
    n = [len(Rk), len(Sk)]
    if splits is None:
        splits = [np.linspace(0, n[w], G + 1).astype(np.int64) for w in range(2)]
    shard_n = [[int(splits[w][r + 1] - splits[w][r]) for r in range(G)] for w in range(2)]
    caps = []
    for w in range(2):
        d = ((rels[w][0].view(np.uint32) >> B) & (G - 1)) if n[w] else np.zeros(0, dtype=np.int64)
        caps.append(int(max(np.bincount(d, minlength=G).max() if n[w] else 0, 1) * slack) + 64)
    # engine capacity: the receive capacity, and the shard + one spare stage slot per chunk
    mx = [max(caps[w], max(shard_n[w]) + 1024) for w in range(2)]
    engs = [gj.JoinEngine(mx[0], mx[1], 0, **(opts or {})) for _ in range(G)]
    try:
        own = [[torch.zeros(caps[w] + 16, dtype=torch.int64, device="cuda") for _ in range(G)] for w in range(2)]
        flags = [torch.zeros(gj.pcp_ctrl_bytes(G) // 4, dtype=torch.int32, device="cuda") for _ in range(G)]   # control blocks
        cols = [[dev(torch, rels[w][0][splits[w][r]:splits[w][r + 1]], rels[w][1][splits[w][r]:splits[w][r + 1]])
                 for r in range(G)] for w in range(2)]
        order = (1, 0) if n[0] > n[1] else (0, 1)          # the building (smaller) relation travels first
        nst = {order[0]: stages[0], order[1]: stages[1]}
        torch.cuda.synchronize()
        for r in range(G):
            engs[r].pcp_begin(n[0], n[1], G, r, B)
        g, bl, b2 = engs[0].pcp_plan()
        assert g == G.bit_length() - 1 and bl + b2 == B
        assert (g, bl, b2) == gj.distributed.pcp_plan_bits(G, B, (opts or {}).get("pass1_bits", 0))
        n1 = 1 << (g + bl)
        hist = [[torch.empty(n1, dtype=torch.int32, device="cuda") for _ in range(G)] for w in range(2)]
        for r in range(G):
            for w in range(2):
                engs[r].pcp_hist(w, cols[w][r][0], hist[w][r])
        if peer_hist:      # gj_pcp_hist_exchange instead of a collective: every rank pushes, waits, compacts
            allr = [[torch.full((G, n1), -1, dtype=torch.int32, device="cuda") for _ in range(G)] for w in range(2)]
            for r in range(G):
                for w in range(2):
                    engs[r].pcp_hist_exchange(w, hist[w][r], [f.data_ptr() for f in flags], flags[r].data_ptr(), allr[w][r])
        torch.cuda.synchronize()
        allh = [torch.stack(hist[w]).contiguous() for w in range(2)]
        if peer_hist:
            for w in range(2):
                for r in range(G):
                    assert torch.equal(allr[w][r], allh[w])
        for w in range(2):
            for r in range(G):
                k = rels[w][0][splits[w][r]:splits[w][r + 1]].view(np.uint32)
                assert np.array_equal(hist[w][r].cpu().numpy(), np.bincount((k >> (B - bl)) & (n1 - 1), minlength=n1))
        for r in range(G):
            for w in order:
                engs[r].pcp_part(w, cols[w][r][0], cols[w][r][1], allr[w][r] if peer_hist else allh[w], own[w][r].data_ptr(), caps[w])
                engs[r].pcp_copy(w, [t.data_ptr() for t in own[w]], [f.data_ptr() for f in flags], nst[w])
        torch.cuda.synchronize()
        # every receive buffer is first-pass partitioned: partition j of destination d holds exactly the
        # tuples with (key >> (B - bl)) & (2^(g+bl) - 1) == (d << bl) | j, nothing beyond the total
        for w in range(2):
            ah = allh[w].cpu().numpy()
            if not n[w] or not check_layout:
                continue
            c_all, h_all = orc.partition_fingerprint(rels[w][0], rels[w][1], B - bl, g + bl)
            for d in range(G):
                _, _, tots = gj.distributed.pcp_layout(ah, d, bl)
                tot = int(tots[d])
                got = own[w][d].cpu().numpy()
                assert not got[tot:].any()
                t = got[:tot].view(np.int32).reshape(-1, 2)
                keys = np.ascontiguousarray(t[:, 0])
                cnt = ah.sum(axis=0)[d << bl:(d + 1) << bl]
                pid = np.repeat(np.arange(1 << bl, dtype=np.int64) + (d << bl), cnt)
                assert np.array_equal(((keys.view(np.uint32) >> (B - bl)) & (n1 - 1)).astype(np.int64), pid)
                if tot:
                    c, h = orc.partition_fingerprint(keys, np.ascontiguousarray(t[:, 1]), B - bl, g + bl)
                    sl = slice(d << bl, (d + 1) << bl)
                    assert np.array_equal(c[sl], c_all[sl]) and np.array_equal(h[sl], h_all[sl])
        m = c = 0
        got_n = [0, 0]
        for r in range(G):
            for w in order:
                engs[r].pcp_recv(w, own[w][r].data_ptr(), flags[r].data_ptr(), caps[w])
            mm, cc, a, b, ph, bits = engs[r].pcp_finish()
            m += mm
            c = (c + cc) % 2**64
            got_n[0] += a
            got_n[1] += b
        assert got_n == n
        return m, c, (g, bl, b2)
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("G,B,p1", [(2, 7, 0), (4, 9, 0), (8, 8, 0), (8, 13, 0), (8, 15, 0), (8, 16, 0), (2, 15, 0), (16, 14, 0),
                                    (8, 15, 10), (8, 15, 3), (4, 12, 6), (2, 1, 0)])
def test_pcp_virtual_shards(gj, orc, torch_cuda, G, B, p1):
    """Source-side pass on [gpu | top local bits] (up to 1024 chunks), TMA bulk copies of whole chunks
    (odd head / tail tuples, empty chunks, one-tuple chunks), receiver-side last pass of up to 10 bits;
    signed keys, N:M matches, ragged shards (one empty)."""
    rng = np.random.default_rng(11 * G + B + p1)
    nR, nS = 900_000, 1_700_000
    Rk = rnd(rng, nR, -(1 << 21), 1 << 21)
    Sk = rnd(rng, nS, -(1 << 21), 1 << 21)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    splits = []
    for n in (nR, nS):
        cuts = np.sort(rng.integers(0, n, size=G - 1))
        if G > 2:
            cuts[1] = cuts[0]                   # an empty shard
        splits.append(np.concatenate(([0], cuts, [n])).astype(np.int64))
    stages = [(1, 1), (2, 4), (3, 5), (64, 64), (1, 7)][(G + B + p1) % 5]
    m, c, bits = _pcp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, Sk, Sp, splits=splits, opts={"pass1_bits": p1} if p1 else None,
                              stages=stages, peer_hist=(G + B) % 2 == 0 and G <= 8)   # (one stream per virtual rank: beyond the 8 hardware
                                                                                     # queues a spinning gather would block another rank's push)
    assert (m, c) == (want.matches, want.checksum)
    # the larger relation first in the argument list: the engine builds on (and ships first) the smaller one
    m2, c2, _ = _pcp_virtual(gj, orc, torch_cuda, G, B, Sk, Sp, Rk, Rp, splits=splits[::-1], opts={"pass1_bits": p1} if p1 else None,
                             stages=stages, check_layout=False)
    assert (m2, c2) == (want.matches, want.checksum)
    g = G.bit_length() - 1
    assert bits[0] == g and bits[0] + bits[1] <= 10 and bits[2] <= 10 and bits[1] <= 8
    if p1:
        assert bits[1] == max(min(max(p1 - g, 0), 10 - g, B - 1, 8), B - 10)


def test_pcp_more_tiles_than_resident_ctas(gj, orc, torch_cuda):
    """The receiver's last pass runs a persistent grid (the tile count of a stage lives on the device): with
    ~10 M tuples per receiver a stage has more tiles than CTAs fit the GPU, so every CTA loops."""
    rng = np.random.default_rng(77)
    nR, nS, G, B = 4_000_000, 20_000_000, 2, 11
    Rk = rng.permutation(nR).astype(np.int32)
    Sk = rng.integers(0, nR, nS).astype(np.int32)
    Rp, Sp = orc.payload_of_keys(Rk, 40), orc.payload_of_keys(Sk, 50)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    for stages in ((1, 1), (2, 3)):
        assert _pcp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, Sk, Sp, slack=1.1, stages=stages, check_layout=False)[:2] == (want.matches, want.checksum)


def test_pcp_skew_tiny_and_overflow(gj, orc, torch_cuda):
    rng = np.random.default_rng(6)
    nR, nS, G, B = 300_000, 1_200_000, 4, 8
    Rk = rng.permutation(nR).astype(np.int32)
    hot = rng.integers(0, 50, size=nS // 2)
    Sk = np.concatenate((hot, rng.integers(0, nR, size=nS - nS // 2))).astype(np.int32)
    rng.shuffle(Sk)
    Rp, Sp = rnd(rng, nR, -2**31, 2**31), rnd(rng, nS, -2**31, 2**31)
    want = orc.join_check(Rk, Rp, Sk, Sp)
    assert _pcp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, Sk, Sp, slack=1.2, stages=(2, 4))[:2] == (want.matches, want.checksum)
    # a handful of tuples: most chunks empty, some with a single tuple
    tk = np.array([5, 5, 7, 300, -1, 1 << 20], dtype=np.int32)
    tp = np.arange(6, dtype=np.int32) + 1
    w2 = orc.join_check(tk, tp, tk, tp)
    assert _pcp_virtual(gj, orc, torch_cuda, 4, 3, tk, tp, tk, tp, slack=8.0)[:2] == (w2.matches, w2.checksum)
    e = np.zeros(0, dtype=np.int32)
    assert _pcp_virtual(gj, orc, torch_cuda, G, B, Rk, Rp, e, e)[:2] == (0, 0)
    k = (rng.integers(0, 1 << 8, size=200_000) | (2 << 8)).astype(np.int32)     # every key goes to GPU 2
    with pytest.raises(gj.GJError):
        _pcp_virtual(gj, orc, torch_cuda, G, B, k, k, k, k, slack=0.5, check_layout=False)


# ------------------------------------------------------------------------------- device generator
def test_device_generator_is_the_host_bijection(gj, orc, eng, torch_cuda):
    n = 1_000_003
    k = torch_cuda.empty(n, dtype=torch_cuda.int32, device="cuda")
    p = torch_cuda.empty(n, dtype=torch_cuda.int32, device="cuda")
    eng.generate_unique(k, p, 0, n, 5, 9)
    hk, hp = k.cpu().numpy(), p.cpu().numpy()
    assert np.array_equal(np.sort(hk), np.arange(n))
    for row in (0, 1, 77, n - 1):
        assert hk[row] == gj.bijection(row, n, 5)
        assert hp[row] == gj.payload_of_key(int(hk[row]), 9)
    # sharded generation == whole generation
    k2 = torch_cuda.empty(1000, dtype=torch_cuda.int32, device="cuda")
    p2 = torch_cuda.empty(1000, dtype=torch_cuda.int32, device="cuda")
    eng.generate_unique(k2, p2, 5000, n, 5, 9)
    assert np.array_equal(k2.cpu().numpy(), hk[5000:6000]) and np.array_equal(p2.cpu().numpy(), hp[5000:6000])
    # oracle restates the payload function independently
    assert np.array_equal(orc.payload_of_keys(hk[:4096], 9), hp[:4096])


def test_error_behaviour(gj, eng, torch_cuda):
    """Errors come back as codes + message, never exit() (reference: common.h:132-141)."""
    big = torch_cuda.zeros((1 << 22) + 1, dtype=torch_cuda.int32, device="cuda")
    small = torch_cuda.zeros(4, dtype=torch_cuda.int32, device="cuda")
    with pytest.raises(gj.GJError) as ei:
        eng.join_aggregate(big, big, small, small)
    assert ei.value.code == -1 and "capacity" in str(ei.value)
    with pytest.raises(gj.GJError):
        eng.set_option("no_such_option", 1)
    with pytest.raises(gj.GJError):
        eng.set_option("radix_bits", 22)
    with pytest.raises(gj.GJError):
        gj.JoinEngine(16, 16, device=99)
