"""Generates tests/golden/generator_vectors.npz from the REFERENCE's own generator object
code (oracle/_ref/libref_generator.so, built from /root/reference/src/generator_ETHZ.cu).

Run in the build container only (the reference is not present on the GPU box):
    python tests/golden/make_golden.py
The committed .npz pins oracle/oracle_join.c's generator restatement bit-exactly.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def main():
    orc.build(ref=True)
    g = orc.RefGenerator()
    v = {}
    for seed in (1, 12345, 0xDEADBEEF):
        v[f"shuffle48_seed{seed}"] = g.knuth_shuffle48(np.arange(1000, dtype=np.int32), orc.state48(seed))
    g.seed_generator(7)
    v["shuffle_srand7"] = g.knuth_shuffle(np.arange(1000, dtype=np.int32))
    g.seed_generator(3)
    v["random_gen_srand3_n1000_max500"] = g.random_gen(1000, 500)
    pk = g.knuth_shuffle48(np.arange(64, dtype=np.int32), orc.state48(5))
    g.seed_generator(11)
    v["fk_pk64_n200_srand11"] = g.fk_from_pk(200, pk, "/tmp/_gj_fk_golden.bin")
    for z in (0.5, 1.0):
        g.seed_generator(42)
        v[f"zipf_srand42_n2000_a1000_z{z}"] = g.gen_zipf(2000, 1000, z)
    # time(NULL)-seeded in the reference: only the multiset is reproducible
    v["unique_n40_max16_sorted"] = np.sort(g.random_unique_gen_timeseeded(40, 16))
    v["unique_n64_max64_sorted"] = np.sort(g.random_unique_gen_timeseeded(64, 64))
    out = os.path.join(ROOT, "tests", "golden", "generator_vectors.npz")
    np.savez_compressed(out, **v)
    print("wrote", out, {k: a.shape for k, a in v.items()})


if __name__ == "__main__":
    main()
