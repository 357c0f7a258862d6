"""Pins the oracle's generator restatement (oracle/oracle_join.c section 1) against golden vectors
produced by the REFERENCE's generator object code (tests/golden/make_golden.py) and, when
oracle/_ref/ exists, against that object code live.  Reference: generator_ETHZ.cu:115-348."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "generator_vectors.npz"))


@pytest.mark.parametrize("seed", [1, 12345, 0xDEADBEEF])
def test_shuffle48_matches_reference(orc, seed):
    got = orc.knuth_shuffle48(np.arange(1000, dtype=np.int32), orc.state48(seed))
    assert np.array_equal(got, GOLD[f"shuffle48_seed{seed}"])


def test_shuffle_rand_matches_reference(orc):
    orc.seed_generator(7)
    assert np.array_equal(orc.knuth_shuffle(np.arange(1000, dtype=np.int32)), GOLD["shuffle_srand7"])


def test_random_gen_matches_reference(orc):
    orc.seed_generator(3)
    assert np.array_equal(orc.random_gen(1000, 500), GOLD["random_gen_srand3_n1000_max500"])


def test_fk_from_pk_matches_reference(orc):
    pk = orc.knuth_shuffle48(np.arange(64, dtype=np.int32), orc.state48(5))
    orc.seed_generator(11)
    assert np.array_equal(orc.fk_from_pk(200, pk), GOLD["fk_pk64_n200_srand11"])


@pytest.mark.parametrize("z", [0.5, 1.0])
def test_zipf_matches_reference(orc, z):
    orc.seed_generator(42)
    assert np.array_equal(orc.gen_zipf(2000, 1000, z), GOLD[f"zipf_srand42_n2000_a1000_z{z}"])


def test_unique_sequence_multiset_matches_reference(orc):
    # n=40, maxid=16 -> key 0 once, 1..7 three times, 8..16 twice
    assert np.array_equal(np.sort(orc.random_unique_gen(40, 16, 99)), GOLD["unique_n40_max16_sorted"])
    assert np.array_equal(np.sort(orc.random_unique_gen(64, 64, 99)), GOLD["unique_n64_max64_sorted"])
    cnt = np.bincount(orc.unique_sequence(40, 16), minlength=17)
    assert cnt[0] == 1 and all(cnt[1:8] == 3) and all(cnt[8:17] == 2)


def test_unique_is_cyclic_permutation(orc):
    # knuth_shuffle48 draws j in [0,i): Sattolo -> a permutation of 0..n-1 (when maxid==n)
    r = orc.random_unique_gen(5000, 5000, 1)
    assert np.array_equal(np.sort(r), np.arange(5000))
    # Sattolo applied to the identity yields no fixed points
    assert not np.any(r == np.arange(5000))


def test_live_reference_generator_if_present(orc):
    if not os.path.exists(orc.REF_GEN_PATH):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    g = orc.RefGenerator()
    for seed in (2, 77):
        a = orc.knuth_shuffle48(np.arange(4097, dtype=np.int32), orc.state48(seed))
        b = g.knuth_shuffle48(np.arange(4097, dtype=np.int32), orc.state48(seed))
        assert np.array_equal(a, b)
    orc.seed_generator(5)
    a = orc.gen_zipf(500, 300, 0.75)
    g.seed_generator(5)
    b = g.gen_zipf(500, 300, 0.75)
    assert np.array_equal(a, b)
