"""ETHZ-style workload generator (host) -- ctypes binding of include/gpujoin_generator.h.

Function names follow the reference's generator_ETHZ.cuh:11-23; seeds are explicit.
Returns numpy int32 arrays.  The C++ implementation lives in csrc/generator.cpp inside
libgpujoin.so; nothing here touches the oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import lib

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_ready = False


def _L():
    global _ready
    L = lib()
    if not _ready:
        L.gj_seed_generator.argtypes = [C.c_uint]
        L.gj_read_relation.argtypes = [C.c_char_p, _i32p, C.c_uint64]
        L.gj_write_relation.argtypes = [C.c_char_p, _i32p, C.c_uint64]
        L.gj_random_gen.argtypes = [_i32p, C.c_uint64, C.c_int64]
        L.gj_random_unique_gen.argtypes = [_i32p, C.c_uint64, C.c_int64, C.c_uint]
        L.gj_knuth_shuffle.argtypes = [_i32p, C.c_uint64]
        L.gj_knuth_shuffle48.argtypes = [_i32p, C.c_uint64, C.POINTER(C.c_ushort)]
        L.gj_gen_zipf.argtypes = [C.c_uint64, C.c_uint, C.c_double, _i32p]
        L.gj_create_relation_unique.argtypes = [C.c_char_p, _i32p, C.c_uint64, C.c_int64, C.c_uint]
        L.gj_create_relation_nonunique.argtypes = [C.c_char_p, _i32p, C.c_uint64, C.c_int64]
        L.gj_create_relation_fk_from_pk.argtypes = [C.c_char_p, _i32p, C.c_uint64, _i32p, C.c_uint64]
        L.gj_create_relation_zipf.argtypes = [C.c_char_p, _i32p, C.c_uint64, C.c_int64, C.c_double]
        L.gj_create_relation_n.argtypes = [_i32p, _i32p, C.c_uint64, C.c_uint64]
        L.gj_create_relation_unique_parallel.argtypes = [_i32p, C.c_uint64, C.c_int64, C.c_uint, C.c_int]
        L.gj_create_relation_zipf_parallel.argtypes = [_i32p, C.c_uint64, C.c_uint, C.c_double, C.c_uint, C.c_int]
        _ready = True
    return L


def _fn(filename):
    return filename.encode() if filename else None


def seed_generator(seed: int) -> None:
    _L().gj_seed_generator(seed)


def state48(seed: int):
    return (C.c_ushort * 3)(seed & 0xFFFF, (seed >> 16) & 0xFFFF, 0)


def knuth_shuffle48(rel: np.ndarray, state) -> np.ndarray:
    _L().gj_knuth_shuffle48(rel, rel.size, state)
    return rel


def knuth_shuffle(rel: np.ndarray) -> np.ndarray:
    _L().gj_knuth_shuffle(rel, rel.size)
    return rel


def random_gen(n: int, maxid: int) -> np.ndarray:
    out = np.empty(n, np.int32)
    _L().gj_random_gen(out, n, maxid)
    return out


def gen_zipf(n: int, alphabet: int, z: float) -> np.ndarray:
    out = np.empty(n, np.int32)
    _L().gj_gen_zipf(n, alphabet, z, out)
    return out


def create_relation_unique(n: int, maxid: int, seed: int, filename: str | None = None, out=None) -> np.ndarray:
    out = np.empty(n, np.int32) if out is None else out
    if _L().gj_create_relation_unique(_fn(filename), out, n, maxid, seed):
        raise OSError(f"could not write {filename}")
    return out


def create_relation_nonunique(n: int, maxid: int, filename: str | None = None) -> np.ndarray:
    out = np.empty(n, np.int32)
    if _L().gj_create_relation_nonunique(_fn(filename), out, n, maxid):
        raise OSError(f"could not write {filename}")
    return out


def create_relation_fk_from_pk(nfk: int, pk: np.ndarray, filename: str | None = None) -> np.ndarray:
    out = np.empty(nfk, np.int32)
    pk = np.ascontiguousarray(pk, np.int32)
    if _L().gj_create_relation_fk_from_pk(_fn(filename), out, nfk, pk, pk.size):
        raise OSError(f"could not write {filename}")
    return out


def create_relation_zipf(n: int, maxid: int, z: float, filename: str | None = None) -> np.ndarray:
    out = np.empty(n, np.int32)
    if _L().gj_create_relation_zipf(_fn(filename), out, n, maxid, z):
        raise OSError(f"could not write {filename}")
    return out


def create_relation_n(rel: np.ndarray, copies: int) -> np.ndarray:
    rel = np.ascontiguousarray(rel, np.int32)
    out = np.empty(rel.size * copies, np.int32)
    _L().gj_create_relation_n(rel, out, rel.size, copies)
    return out


def create_relation_unique_parallel(n: int, maxid: int, seed: int, threads: int = 0, out=None) -> np.ndarray:
    out = np.empty(n, np.int32) if out is None else out
    if _L().gj_create_relation_unique_parallel(out, n, maxid, seed, threads):
        raise ValueError("bad arguments")
    return out


def create_relation_zipf_parallel(n: int, alphabet: int, z: float, seed: int, threads: int = 0, out=None) -> np.ndarray:
    out = np.empty(n, np.int32) if out is None else out
    if _L().gj_create_relation_zipf_parallel(out, n, alphabet, z, seed, threads):
        raise ValueError("bad arguments")
    return out


def read_relation(filename: str, n: int) -> np.ndarray:
    out = np.empty(n, np.int32)
    if _L().gj_read_relation(filename.encode(), out, n):
        raise OSError(f"could not read {n} keys from {filename}")
    return out


def write_relation(filename: str, rel: np.ndarray) -> None:
    rel = np.ascontiguousarray(rel, np.int32)
    if _L().gj_write_relation(filename.encode(), rel, rel.size):
        raise OSError(f"could not write {filename}")
