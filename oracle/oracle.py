"""ctypes front-end for the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.

`RefGenerator` additionally loads the reference's own generator object code
(oracle/_ref/libref_generator.so = /root/reference/src/generator_ETHZ.cu compiled as host
C++ by oracle/Makefile).  It exists only where /root/reference was present at build time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_GEN_PATH = os.path.join(REF_DIR, "libref_generator.so")
REF_BENCH_PATH = os.path.join(REF_DIR, "bench_ref")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(ref: bool | None = None) -> None:
    """Compile the oracle (and the rebuilt reference when its sources are present)."""
    targets = ["all"]
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        targets.append("ref")
        # the reference's driver objects linked against the product library
        if os.path.exists(os.path.join(HERE, "..", "icde2019-gpu-join_b200", "lib", "libgpujoin.so")):
            targets.append("dropin")
    subprocess.run(["make", "-s", "-C", HERE, "-j8"] + targets, check=True)


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        build(ref=False)
    lib = C.CDLL(LIB_PATH)
    lib.orc_seed_generator.argtypes = [C.c_uint]
    lib.orc_knuth_shuffle48.argtypes = [_i32p, C.c_uint64, C.POINTER(C.c_ushort)]
    lib.orc_knuth_shuffle.argtypes = [_i32p, C.c_uint64]
    lib.orc_random_gen.argtypes = [_i32p, C.c_uint64, C.c_int64]
    lib.orc_unique_sequence.argtypes = [_i32p, C.c_uint64, C.c_int64]
    lib.orc_random_unique_gen.argtypes = [_i32p, C.c_uint64, C.c_int64, C.c_uint]
    lib.orc_fk_from_pk.argtypes = [_i32p, C.c_uint64, _i32p, C.c_uint64]
    lib.orc_gen_zipf.argtypes = [C.c_uint64, C.c_uint, C.c_double, _i32p]
    lib.orc_pair_mix.argtypes = [C.c_int32, C.c_int32]
    lib.orc_pair_mix.restype = C.c_uint64
    lib.orc_join_naive.argtypes = [_i32p, _i32p, C.c_uint64, _i32p, _i32p, C.c_uint64, _u64p]
    lib.orc_join_check.argtypes = [_i32p, _i32p, C.c_uint64, _i32p, _i32p, C.c_uint64, C.c_int, _u64p]
    lib.orc_join_check.restype = C.c_double
    lib.orc_join_materialize.argtypes = [_i32p, _i32p, C.c_uint64, _i32p, _i32p, C.c_uint64, C.c_int,
                                         _i32p, _i32p, C.c_uint64, _u64p]
    lib.orc_join_materialize.restype = C.c_uint64
    lib.orc_pairs_hash.argtypes = [_i32p, _i32p, C.c_uint64]
    lib.orc_pairs_hash.restype = C.c_uint64
    lib.orc_late_sum.argtypes = [_i32p, _i32p, C.c_uint64, _i32p, C.c_uint32, C.c_uint64, _i32p, C.c_uint32, C.c_uint64]
    lib.orc_late_sum.restype = C.c_uint64
    lib.orc_partition.argtypes = [_i32p, _i32p, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, _i32p, _i32p]
    lib.orc_partition_fingerprint.argtypes = [_i32p, _i32p, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, _u64p]
    lib.orc_max_threads.restype = C.c_int
    lib.orc_payload_of_keys.argtypes = [_i32p, C.c_uint64, C.c_uint32, _i32p]
    lib.orc_bijection.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    lib.orc_bijection.restype = C.c_uint32
    lib.orc_unique_join_checksum.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
    lib.orc_unique_join_checksum.restype = C.c_uint64
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------- generators
def seed_generator(seed: int) -> None:
    lib().orc_seed_generator(seed)


def state48(seed: int):
    """nrand48 state as the reference builds it: 4 seed bytes over a zeroed short[3]."""
    return (C.c_ushort * 3)(seed & 0xFFFF, (seed >> 16) & 0xFFFF, 0)


def knuth_shuffle48(rel: np.ndarray, state) -> np.ndarray:
    lib().orc_knuth_shuffle48(rel, rel.size, state)
    return rel


def knuth_shuffle(rel: np.ndarray) -> np.ndarray:
    lib().orc_knuth_shuffle(rel, rel.size)
    return rel


def random_gen(n: int, maxid: int) -> np.ndarray:
    out = np.empty(n, np.int32)
    lib().orc_random_gen(out, n, maxid)
    return out


def unique_sequence(n: int, maxid: int) -> np.ndarray:
    out = np.empty(n, np.int32)
    lib().orc_unique_sequence(out, n, maxid)
    return out


def random_unique_gen(n: int, maxid: int, seed: int) -> np.ndarray:
    out = np.empty(n, np.int32)
    lib().orc_random_unique_gen(out, n, maxid, seed)
    return out


def fk_from_pk(nfk: int, pk: np.ndarray) -> np.ndarray:
    out = np.empty(nfk, np.int32)
    lib().orc_fk_from_pk(out, nfk, _c(pk), pk.size)
    return out


def gen_zipf(n: int, alphabet: int, z: float) -> np.ndarray:
    out = np.empty(n, np.int32)
    lib().orc_gen_zipf(n, alphabet, z, out)
    return out


# ---------------------------------------------------------------- join checkers
class JoinResult(tuple):
    matches = property(lambda s: s[0])
    checksum = property(lambda s: s[1])
    pairhash = property(lambda s: s[2])

    @property
    def ref_results_int32(self) -> int:
        """What the reference prints as `%d results` (hash_join_clustered_probe.cu:984-986):
        the int32-wrapped SUM(Pr*Ps) = low 32 bits of checksum, as a signed int."""
        v = self[1] & 0xFFFFFFFF
        return v - (1 << 32) if v >= (1 << 31) else v


def join_naive(Rk, Rp, Sk, Sp) -> JoinResult:
    out = np.zeros(3, np.uint64)
    lib().orc_join_naive(_c(Rk), _c(Rp), len(Rk), _c(Sk), _c(Sp), len(Sk), out)
    return JoinResult(int(x) for x in out)


def join_check(Rk, Rp, Sk, Sp, threads: int = 0, with_time: bool = False):
    out = np.zeros(3, np.uint64)
    secs = lib().orc_join_check(_c(Rk), _c(Rp), len(Rk), _c(Sk), _c(Sp), len(Sk), threads, out)
    res = JoinResult(int(x) for x in out)
    return (res, secs) if with_time else res


def join_materialize(Rk, Rp, Sk, Sp, cap: int, threads: int = 0):
    out = np.zeros(3, np.uint64)
    orp = np.empty(max(cap, 1), np.int32)
    osp = np.empty(max(cap, 1), np.int32)
    n = lib().orc_join_materialize(_c(Rk), _c(Rp), len(Rk), _c(Sk), _c(Sp), len(Sk), threads,
                                   orp, osp, cap, out)
    k = min(int(n), cap)
    return int(n), orp[:k], osp[:k], JoinResult(int(x) for x in out)


def join_late(Rk, Rid, Sk, Sid, Dr, Ds, threads: int = 0):
    """Late-materialisation join on the CPU (reference join_partitioned_varpayload,
    join-primitives.cu:1420-1557): joins (key, row id) relations with the checker, then adds, for every
    result pair, all side-table values of both rows.  Dr / Ds: int32 arrays [cols, rows] (cols may be
    0).  Returns (matches, sum mod 2^64)."""
    cap = max(1, int(join_check(Rk, Rid, Sk, Sid, threads).matches))
    n, rid, sid, _ = join_materialize(Rk, Rid, Sk, Sid, cap, threads)
    assert n <= cap
    Dr = np.ascontiguousarray(Dr, dtype=np.int32).reshape(len(Dr), -1) if len(Dr) else np.zeros((0, 1), np.int32)
    Ds = np.ascontiguousarray(Ds, dtype=np.int32).reshape(len(Ds), -1) if len(Ds) else np.zeros((0, 1), np.int32)
    dr = Dr.reshape(-1) if Dr.size else np.zeros(1, np.int32)
    ds = Ds.reshape(-1) if Ds.size else np.zeros(1, np.int32)
    total = lib().orc_late_sum(_c(rid), _c(sid), int(n), dr, Dr.shape[0], Dr.shape[1], ds, Ds.shape[0], Ds.shape[1])
    return int(n), int(total)


def pairs_hash(rp, sp) -> int:
    rp, sp = _c(rp), _c(sp)
    return int(lib().orc_pairs_hash(rp, sp, rp.size))


def partition(keys, pays, shift: int, bits: int):
    keys, pays = _c(keys), _c(pays)
    off = np.zeros((1 << bits) + 1, np.uint64)
    ko, po = np.empty_like(keys), np.empty_like(pays)
    lib().orc_partition(keys, pays, keys.size, shift, bits, off, ko, po)
    return off, ko, po


def partition_fingerprint(keys, pays, shift: int, bits: int):
    keys, pays = _c(keys), _c(pays)
    cnt = np.zeros(1 << bits, np.uint64)
    hsh = np.zeros(1 << bits, np.uint64)
    lib().orc_partition_fingerprint(keys, pays, keys.size, shift, bits, cnt, hsh)
    return cnt, hsh


def payload_of_keys(keys, pay_seed: int) -> np.ndarray:
    keys = _c(keys)
    out = np.empty_like(keys)
    lib().orc_payload_of_keys(keys, keys.size, pay_seed, out)
    return out


def bijection(row: int, n_total: int, seed: int) -> int:
    return int(lib().orc_bijection(row, n_total, seed))


def unique_join_checksum(k_begin: int, k_end: int, seed_a: int, seed_b: int) -> int:
    return int(lib().orc_unique_join_checksum(k_begin, k_end, seed_a, seed_b))


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---------------------------------------------------------------- the reference's own generator
class RefGenerator:
    """The reference's generator_ETHZ.cu object code (C++-mangled symbols)."""

    def __init__(self):
        if not os.path.exists(REF_GEN_PATH):
            raise FileNotFoundError(REF_GEN_PATH)
        g = C.CDLL(REF_GEN_PATH)
        self.seed_generator = g._Z14seed_generatorj
        self.seed_generator.argtypes = [C.c_uint]
        self._shuffle48 = g._Z15knuth_shuffle48PimPt
        self._shuffle48.argtypes = [_i32p, C.c_uint64, C.POINTER(C.c_ushort)]
        self._shuffle = g._Z13knuth_shufflePim
        self._shuffle.argtypes = [_i32p, C.c_uint64]
        self._random_gen = g._Z10random_genPiml
        self._random_gen.argtypes = [_i32p, C.c_uint64, C.c_int64]
        self._zipf = g._Z8gen_zipfmjdPi
        self._zipf.argtypes = [C.c_uint64, C.c_uint, C.c_double, _i32p]
        self._unique = g._Z17random_unique_genPiml
        self._unique.argtypes = [_i32p, C.c_uint64, C.c_int64]
        self._fk = g._Z26create_relation_fk_from_pkPKcPimS1_m
        self._fk.argtypes = [C.c_char_p, _i32p, C.c_uint64, _i32p, C.c_uint64]

    def knuth_shuffle48(self, rel, state):
        self._shuffle48(rel, rel.size, state)
        return rel

    def knuth_shuffle(self, rel):
        self._shuffle(rel, rel.size)
        return rel

    def random_gen(self, n, maxid):
        out = np.empty(n, np.int32)
        self._random_gen(out, n, maxid)
        return out

    def gen_zipf(self, n, alphabet, z):
        out = np.empty(n, np.int32)
        self._zipf(n, alphabet, z, out)  # prints "live k" lines to stdout
        return out

    def random_unique_gen_timeseeded(self, n, maxid):
        """Seeded from time(NULL) inside the reference -> only its multiset is checkable."""
        out = np.empty(n, np.int32)
        self._unique(out, n, maxid)
        return out

    def fk_from_pk(self, nfk, pk, tmpfile: str):
        out = np.empty(nfk, np.int32)
        if os.path.exists(tmpfile):
            os.remove(tmpfile)
        self._fk(tmpfile.encode(), out, nfk, _c(pk), pk.size)
        return out
