"""ctypes binding of the C ABI in include/gpujoin.h (libgpujoin.so).

`JoinEngine` mirrors the reference operator outOfGPU_Join1_payload
(/root/reference/src/hash_join_clustered_probe.cu:802-994): relations are (keys, payloads)
int32 column pairs; `join_aggregate` returns what the reference prints as "%d results"
(widened to 64 bit) plus the exact match count.  Device inputs are torch CUDA tensors (their
data_ptr() crosses the ABI as a plain pointer); host inputs are numpy arrays / CPU tensors.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# every symbol include/gpujoin.h declares (tests check that the library exports all of them)
C_ABI_SYMBOLS = [
    "gj_create", "gj_destroy", "gj_last_error", "gj_version", "gj_set_stream", "gj_set_option",
    "gj_get_option", "gj_join_aggregate", "gj_join_aggregate_tuples", "gj_join_aggregate_host",
    "gj_join_materialize", "gj_join_aggregate_late", "gj_join_aggregate_nopart", "gj_join_aggregate_perfect", "gj_join_aggregate_stream_host", "gj_partition", "gj_shuffle_split", "gj_shuffle_scatter_peers",
    "gj_shuffle_count", "gj_shuffle_scatter_peers_async", "gj_shuffle_scatter_ms", "gj_memcpy_d2d_async", "gj_stage_begin",
    "gj_stage_partition", "gj_stage_join", "gj_stage_finish", "gj_stage_pass_ms", "gj_pp_begin", "gj_pp_local", "gj_pp_push", "gj_pp_join",
    "gj_pp_finish", "gj_pp_plan", "gj_pcp_begin", "gj_pcp_plan", "gj_pcp_hist", "gj_pcp_part", "gj_pcp_copy", "gj_pcp_recv", "gj_pcp_ctrl_bytes", "gj_pcp_hist_exchange",
    "gj_pcp_finish", "gj_ipc_export", "gj_ipc_open", "gj_ipc_close", "gj_enable_peer_access", "gj_generate_unique", "gj_bijection", "gj_payload_of_key",
    "gj_device_count", "gj_malloc_device", "gj_free_device", "gj_malloc_pinned", "gj_free_pinned",
    "gj_memcpy_h2d", "gj_memcpy_d2h", "gj_memset_device", "gj_device_synchronize", "gj_flush_l2",
    "gj_kernel_launch_count",
]


class GJError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgpujoin error {code}: {msg}")
        self.code = code


class Timings(C.Structure):
    _fields_ = [("hist_ms", C.c_float), ("part_ms", C.c_float), ("join_ms", C.c_float),
                ("total_ms", C.c_float), ("h2d_ms", C.c_float), ("wall_ms", C.c_float),
                ("pass_ms", C.c_float * 4),
                ("radix_bits", C.c_uint32), ("pass1_bits", C.c_uint32), ("pass2_bits", C.c_uint32),
                ("pass3_bits", C.c_uint32),
                ("kernel_launches", C.c_uint32)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "pass_ms" else getattr(self, k)) for k, _ in self._fields_}


@dataclass
class JoinResult:
    matches: int
    checksum: int
    timings: Timings

    @property
    def ref_results_int32(self) -> int:
        """The int32 the reference prints as `%d results` (hash_join_clustered_probe.cu:984-986)."""
        v = self.checksum & 0xFFFFFFFF
        return v - (1 << 32) if v >= (1 << 31) else v


def lib_path() -> str:
    return os.path.join(HERE, "lib", "libgpujoin.so")


_lib = None


def lib() -> C.CDLL:
    """Load libgpujoin.so.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise GJError(-2, f"{path} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C icde2019-gpu-join_b200/csrc); there is no CPU fallback")
    L = C.CDLL(path)
    vp, u64, u32, i32p = C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p
    L.gj_create.argtypes = [C.POINTER(vp), C.c_int, u64, u64]
    L.gj_destroy.argtypes = [vp]
    L.gj_destroy.restype = None
    L.gj_last_error.restype = C.c_char_p
    L.gj_set_stream.argtypes = [vp, vp]
    L.gj_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.gj_get_option.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
    L.gj_join_aggregate.argtypes = [vp, i32p, i32p, u64, i32p, i32p, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(Timings)]
    L.gj_join_aggregate_host.argtypes = L.gj_join_aggregate.argtypes
    L.gj_join_aggregate_nopart.argtypes = L.gj_join_aggregate.argtypes
    L.gj_join_aggregate_perfect.argtypes = [vp, i32p, i32p, u64, i32p, i32p, u64, C.c_int32, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(Timings)]
    L.gj_join_aggregate_stream_host.argtypes = [vp, i32p, i32p, u64, i32p, i32p, u64, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(Timings)]
    L.gj_join_aggregate_tuples.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(Timings)]
    L.gj_join_materialize.argtypes = [vp, i32p, i32p, u64, i32p, i32p, u64, i32p, i32p, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(Timings)]
    L.gj_join_aggregate_late.argtypes = [vp, i32p, i32p, u64, i32p, i32p, u64, i32p, u32, u64, i32p, u32, u64,
                                         C.POINTER(u64), C.POINTER(u64), C.POINTER(Timings)]
    L.gj_partition.argtypes = [vp, C.c_int, i32p, i32p, u64, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(u32), C.POINTER(Timings)]
    L.gj_shuffle_split.argtypes = [vp, i32p, i32p, u64, u32, u32, vp, C.POINTER(u64)]
    L.gj_shuffle_scatter_peers.argtypes = [vp, i32p, i32p, u64, u32, u32, C.POINTER(vp), C.POINTER(u64)]
    L.gj_shuffle_count.argtypes = [vp, i32p, u64, u32, u32, C.POINTER(u64)]
    L.gj_shuffle_scatter_peers_async.argtypes = [vp, C.c_int, i32p, i32p, u64, u32, u32, C.POINTER(vp), C.POINTER(u64), vp]
    L.gj_shuffle_scatter_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.gj_memcpy_d2d_async.argtypes = [vp, vp, u64, vp]
    L.gj_stage_begin.argtypes = [vp, u64, u64, vp]
    L.gj_stage_partition.argtypes = [vp, C.c_int, vp, vp]
    L.gj_stage_join.argtypes = [vp, vp]
    L.gj_stage_finish.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.gj_stage_pass_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.gj_pp_begin.argtypes = [vp, u64, u64, u32, u32, u32, vp]
    L.gj_pp_local.argtypes = [vp, C.c_int, i32p, i32p, u64, vp, vp]
    L.gj_pp_push.argtypes = [vp, C.c_int, vp, C.POINTER(vp), u64, u64, vp]
    L.gj_pp_join.argtypes = [vp, vp, vp, u64, u64, vp]
    L.gj_pp_finish.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_float)]
    L.gj_pp_plan.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.gj_pcp_begin.argtypes = [vp, u64, u64, u32, u32, u32, vp]
    L.gj_pcp_plan.argtypes = [vp, C.POINTER(u32)]
    L.gj_pcp_hist.argtypes = [vp, C.c_int, i32p, u64, vp, vp]
    L.gj_pcp_part.argtypes = [vp, C.c_int, i32p, i32p, vp, vp, u64, vp]
    L.gj_pcp_copy.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), u32, vp]
    L.gj_pcp_recv.argtypes = [vp, C.c_int, vp, vp, u64, vp, vp]
    L.gj_pcp_hist_exchange.argtypes = [vp, C.c_int, vp, C.POINTER(vp), vp, vp, vp]
    L.gj_pcp_ctrl_bytes.argtypes = [u32]
    L.gj_pcp_ctrl_bytes.restype = u64
    L.gj_pcp_finish.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_float), C.POINTER(u32)]
    L.gj_generate_unique.argtypes = [vp, i32p, i32p, u64, u64, u64, u32, u32]
    L.gj_bijection.argtypes = [u64, u64, u32]
    L.gj_bijection.restype = u32
    L.gj_payload_of_key.argtypes = [u32, u32]
    L.gj_payload_of_key.restype = C.c_int32
    L.gj_enable_peer_access.argtypes = [C.c_int, C.c_int]
    L.gj_device_count.argtypes = [C.POINTER(C.c_int)]
    L.gj_malloc_device.argtypes = [C.POINTER(vp), u64]
    L.gj_free_device.argtypes = [vp]
    L.gj_malloc_pinned.argtypes = [C.POINTER(vp), u64]
    L.gj_free_pinned.argtypes = [vp]
    L.gj_memcpy_h2d.argtypes = [vp, vp, u64]
    L.gj_memcpy_d2h.argtypes = [vp, vp, u64]
    L.gj_memset_device.argtypes = [vp, C.c_int, u64]
    L.gj_flush_l2.argtypes = [vp]
    L.gj_kernel_launch_count.restype = u64
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise GJError(rc, lib().gj_last_error().decode(errors="replace"))


def pcp_ctrl_bytes(n_gpus: int) -> int:
    """Size of one GPU's pcp control block (stage flags + delivered fine histograms)."""
    return int(lib().gj_pcp_ctrl_bytes(n_gpus))


def kernel_launch_count() -> int:
    return int(lib().gj_kernel_launch_count())


def bijection(row: int, n_total: int, seed: int) -> int:
    return int(lib().gj_bijection(row, n_total, seed))


def payload_of_key(key: int, pay_seed: int) -> int:
    return int(lib().gj_payload_of_key(key & 0xFFFFFFFF, pay_seed))


def _dev_ptr(t, n_expected=None, what="tensor"):
    """data_ptr of a contiguous int32 CUDA tensor."""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{what}: expected a CUDA torch tensor")
    if t.dtype not in (torch.int32, torch.uint32) or not t.is_contiguous():
        raise TypeError(f"{what}: expected a contiguous int32 tensor")
    if n_expected is not None and t.numel() != n_expected:
        raise ValueError(f"{what}: expected {n_expected} elements, got {t.numel()}")
    return C.c_void_p(t.data_ptr() if t.numel() else 0)


def _host_ptr(a, what="array"):
    import torch
    if isinstance(a, torch.Tensor):
        if a.is_cuda or a.dtype != torch.int32 or not a.is_contiguous():
            raise TypeError(f"{what}: expected a contiguous int32 CPU tensor")
        return C.c_void_p(a.data_ptr() if a.numel() else 0), a.numel()
    a = np.ascontiguousarray(a, dtype=np.int32)
    return C.c_void_p(a.ctypes.data if a.size else 0), a.size, a


class JoinEngine:
    """One engine context per GPU (gj_create / gj_destroy)."""

    def __init__(self, max_R: int, max_S: int, device: int = 0, **options):
        self._L = lib()
        self._ctx = C.c_void_p()
        self.device = device
        self.max_R, self.max_S = int(max_R), int(max_S)
        _check(self._L.gj_create(C.byref(self._ctx), device, self.max_R, self.max_S))
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.gj_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- options / streams --------------------------------------------------------------
    def set_option(self, name: str, value: int):
        _check(self._L.gj_set_option(self._ctx, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int64()
        _check(self._L.gj_get_option(self._ctx, name.encode(), C.byref(v)))
        return int(v.value)

    def use_torch_stream(self, stream=None):
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _check(self._L.gj_set_stream(self._ctx, C.c_void_p(s.cuda_stream)))

    def _sync_inputs(self):
        """The engine runs on its own stream: make sure torch has finished producing the inputs."""
        import torch
        torch.cuda.current_stream(self.device).synchronize()

    def flush_l2(self):
        _check(self._L.gj_flush_l2(self._ctx))

    # -- joins ----------------------------------------------------------------------------
    def join_aggregate(self, Rk, Rp, Sk, Sp) -> JoinResult:
        self._sync_inputs()
        nR, nS = Rk.numel(), Sk.numel()
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate(self._ctx, _dev_ptr(Rk, nR, "Rk"), _dev_ptr(Rp, nR, "Rp"), nR,
                                         _dev_ptr(Sk, nS, "Sk"), _dev_ptr(Sp, nS, "Sp"), nS,
                                         C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    def join_aggregate_nopart(self, Rk, Rp, Sk, Sp) -> JoinResult:
        """Non-partitioned baseline: global-memory chained hash table, no radix pass."""
        self._sync_inputs()
        nR, nS = Rk.numel(), Sk.numel()
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate_nopart(self._ctx, _dev_ptr(Rk, nR, "Rk"), _dev_ptr(Rp, nR, "Rp"), nR,
                                                _dev_ptr(Sk, nS, "Sk"), _dev_ptr(Sp, nS, "Sp"), nS,
                                                C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    def join_aggregate_perfect(self, Rk, Rp, Sk, Sp, key_min: int, key_range: int) -> JoinResult:
        """Non-partitioned perfect-array join: unique build keys in [key_min, key_min + key_range)."""
        self._sync_inputs()
        nR, nS = Rk.numel(), Sk.numel()
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate_perfect(self._ctx, _dev_ptr(Rk, nR, "Rk"), _dev_ptr(Rp, nR, "Rp"), nR,
                                                 _dev_ptr(Sk, nS, "Sk"), _dev_ptr(Sp, nS, "Sp"), nS,
                                                 int(key_min), int(key_range), C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    def join_aggregate_tuples(self, Rt, nR, St, nS) -> JoinResult:
        """Rt/St: int32 CUDA tensors of shape [n, 2] (or int64 [n]) holding packed {key,payload}."""
        self._sync_inputs()
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate_tuples(self._ctx, C.c_void_p(Rt.data_ptr() if nR else 0), nR,
                                                C.c_void_p(St.data_ptr() if nS else 0), nS,
                                                C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    def join_aggregate_ptrs(self, ptr_R: int, nR: int, ptr_S: int, nS: int) -> JoinResult:
        """Packed-tuple join over raw device pointers (peer-store receive buffers)."""
        self._sync_inputs()
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate_tuples(self._ctx, C.c_void_p(ptr_R), nR, C.c_void_p(ptr_S), nS,
                                                C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    def join_aggregate_host(self, Rk, Rp, Sk, Sp) -> JoinResult:
        """End-to-end entry: host columns in, H2D copies inside the call."""
        keep = [_host_ptr(x, n) for x, n in ((Rk, "Rk"), (Rp, "Rp"), (Sk, "Sk"), (Sp, "Sp"))]
        if keep[0][1] != keep[1][1] or keep[2][1] != keep[3][1]:
            raise ValueError("key and payload columns differ in length")
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate_host(self._ctx, keep[0][0], keep[1][0], keep[0][1],
                                              keep[2][0], keep[3][0], keep[2][1],
                                              C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    def join_aggregate_stream_host(self, Rk, Rp, Sk, Sp, chunk_tuples: int) -> JoinResult:
        """Out-of-HBM probe side: host columns in; R resident, S streamed in chunks of chunk_tuples."""
        keep = [_host_ptr(x, n) for x, n in ((Rk, "Rk"), (Rp, "Rp"), (Sk, "Sk"), (Sp, "Sp"))]
        if keep[0][1] != keep[1][1] or keep[2][1] != keep[3][1]:
            raise ValueError("key and payload columns differ in length")
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate_stream_host(self._ctx, keep[0][0], keep[1][0], keep[0][1],
                                                     keep[2][0], keep[3][0], keep[2][1], chunk_tuples,
                                                     C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    def join_materialize(self, Rk, Rp, Sk, Sp, out_Rp, out_Sp):
        """Returns (n_pairs, JoinResult); at most out_Rp.numel() pairs are written."""
        self._sync_inputs()
        nR, nS, cap = Rk.numel(), Sk.numel(), out_Rp.numel()
        if out_Sp.numel() != cap:
            raise ValueError("output columns differ in length")
        n, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_materialize(self._ctx, _dev_ptr(Rk, nR, "Rk"), _dev_ptr(Rp, nR, "Rp"), nR,
                                           _dev_ptr(Sk, nS, "Sk"), _dev_ptr(Sp, nS, "Sp"), nS,
                                           _dev_ptr(out_Rp, cap, "out_Rp"), _dev_ptr(out_Sp, cap, "out_Sp"), cap,
                                           C.byref(n), C.byref(c), C.byref(t)))
        return int(n.value), JoinResult(int(n.value), int(c.value), t)

    def join_aggregate_late(self, Rk, Rid, Sk, Sid, Dr, Ds) -> JoinResult:
        """Late materialisation: Rid / Sid are row ids into the column-major side tables Dr / Ds (int32
        CUDA tensors of shape [cols, rows], possibly with zero columns).  `checksum` of the result is the
        sum over result pairs of all side-table values of both rows."""
        self._sync_inputs()
        nR, nS = Rk.numel(), Sk.numel()

        def side(D, what):
            if D is None or D.numel() == 0:
                return C.c_void_p(0), 0, 0
            if D.dim() != 2 or not D.is_contiguous():
                raise TypeError(f"{what}: expected a contiguous [cols, rows] tensor")
            return _dev_ptr(D, None, what), D.shape[0], D.shape[1]
        pr, cr, sr = side(Dr, "Dr")
        ps, cs, ss = side(Ds, "Ds")
        m, c, t = C.c_uint64(), C.c_uint64(), Timings()
        _check(self._L.gj_join_aggregate_late(self._ctx, _dev_ptr(Rk, nR, "Rk"), _dev_ptr(Rid, nR, "Rid"), nR,
                                              _dev_ptr(Sk, nS, "Sk"), _dev_ptr(Sid, nS, "Sid"), nS,
                                              pr, cr, sr, ps, cs, ss, C.byref(m), C.byref(c), C.byref(t)))
        return JoinResult(int(m.value), int(c.value), t)

    # -- partitioner ----------------------------------------------------------------------
    def partition(self, keys, pays, radix_bits: int = 0, slot: int = 0):
        """Returns (tuples [n,2] int32 numpy copy, offsets [2^B+1] int64 numpy, B, timings)."""
        self._sync_inputs()
        n = keys.numel()
        tp, op, b, t = C.c_void_p(), C.c_void_p(), C.c_uint32(), Timings()
        _check(self._L.gj_partition(self._ctx, slot, _dev_ptr(keys, n, "keys"), _dev_ptr(pays, n, "pays"), n,
                                    radix_bits, C.byref(tp), C.byref(op), C.byref(b), C.byref(t)))
        B = int(b.value)
        offs = np.empty((1 << B) + 1, dtype=np.uint32)
        if n:
            _check(self._L.gj_memcpy_d2h(C.c_void_p(offs.ctypes.data), op, offs.nbytes))
            host = np.empty((n, 2), dtype=np.int32)
            _check(self._L.gj_memcpy_d2h(C.c_void_p(host.ctypes.data), tp, host.nbytes))
            tuples = host
        else:
            offs[:] = 0
            tuples = np.empty((0, 2), dtype=np.int32)
        return tuples, offs.astype(np.int64), B, t

    # -- multi-GPU shuffle step -------------------------------------------------------------
    def shuffle_count(self, keys, n_gpus: int, gpu_shift: int) -> np.ndarray:
        self._sync_inputs()
        n = keys.numel()
        cnt = (C.c_uint64 * n_gpus)()
        _check(self._L.gj_shuffle_count(self._ctx, _dev_ptr(keys, n, "keys"), n, n_gpus, gpu_shift, cnt))
        return np.array(list(cnt), dtype=np.int64)

    def shuffle_split(self, keys, pays, n_gpus: int, gpu_shift: int, out_tuples) -> np.ndarray:
        """out_tuples: int64 CUDA tensor with >= n elements (packed tuples grouped by destination)."""
        self._sync_inputs()
        n = keys.numel()
        if out_tuples.numel() * out_tuples.element_size() < n * 8:
            raise ValueError("out_tuples too small")
        cnt = (C.c_uint64 * n_gpus)()
        _check(self._L.gj_shuffle_split(self._ctx, _dev_ptr(keys, n, "keys"), _dev_ptr(pays, n, "pays"), n,
                                        n_gpus, gpu_shift, C.c_void_p(out_tuples.data_ptr()), cnt))
        return np.array(list(cnt), dtype=np.int64)

    def shuffle_scatter_peers(self, keys, pays, n_gpus: int, gpu_shift: int, peer_ptrs, peer_offsets):
        self._sync_inputs()
        n = keys.numel()
        bases = (C.c_void_p * n_gpus)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        offs = (C.c_uint64 * n_gpus)(*[int(o) for o in peer_offsets])
        _check(self._L.gj_shuffle_scatter_peers(self._ctx, _dev_ptr(keys, n, "keys"), _dev_ptr(pays, n, "pays"), n,
                                                n_gpus, gpu_shift, bases, offs))

    # -- asynchronous / staged entry points (multi-GPU overlap) ------------------------------
    def shuffle_scatter_peers_async(self, which: int, keys, pays, n_gpus: int, gpu_shift: int, peer_ptrs,
                                    peer_offsets, stream):
        n = keys.numel()
        bases = (C.c_void_p * n_gpus)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        offs = (C.c_uint64 * n_gpus)(*[int(o) for o in peer_offsets])
        _check(self._L.gj_shuffle_scatter_peers_async(self._ctx, which, _dev_ptr(keys, n, "keys"), _dev_ptr(pays, n, "pays"),
                                                      n, n_gpus, gpu_shift, bases, offs, C.c_void_p(stream.cuda_stream)))

    def shuffle_scatter_ms(self, which: int) -> float:
        ms = C.c_float()
        _check(self._L.gj_shuffle_scatter_ms(self._ctx, which, C.byref(ms)))
        return float(ms.value)

    def memcpy_d2d_async(self, dst: int, src: int, nbytes: int, stream):
        _check(self._L.gj_memcpy_d2d_async(C.c_void_p(dst), C.c_void_p(src), nbytes, C.c_void_p(stream.cuda_stream)))

    def stage_begin(self, nR: int, nS: int, stream):
        _check(self._L.gj_stage_begin(self._ctx, nR, nS, C.c_void_p(stream.cuda_stream)))

    def stage_partition(self, side: int, ptr: int, stream):
        _check(self._L.gj_stage_partition(self._ctx, side, C.c_void_p(ptr), C.c_void_p(stream.cuda_stream)))

    def stage_join(self, stream):
        _check(self._L.gj_stage_join(self._ctx, C.c_void_p(stream.cuda_stream)))

    def stage_finish(self):
        m, c = C.c_uint64(), C.c_uint64()
        _check(self._L.gj_stage_finish(self._ctx, C.byref(m), C.byref(c)))
        return int(m.value), int(c.value)

    def stage_pass_ms(self):
        ms = (C.c_float * 4)()
        _check(self._L.gj_stage_pass_ms(self._ctx, ms))
        return [float(x) for x in ms]

    # -- sharded "partition, then push" pipeline (gj_pp_*) -----------------------------------
    @staticmethod
    def _sptr(stream):
        return C.c_void_p(stream.cuda_stream if stream is not None else 0)

    def pp_begin(self, n_R_global: int, n_S_global: int, n_gpus: int, rank: int, local_bits: int, stream=None):
        _check(self._L.gj_pp_begin(self._ctx, n_R_global, n_S_global, n_gpus, rank, local_bits, self._sptr(stream)))

    def pp_local(self, which: int, keys, pays, fine_hist, stream=None):
        """fine_hist: int32/uint32 CUDA tensor of 2^(gpu bits + local bits) counters (overwritten)."""
        n = keys.numel()
        _check(self._L.gj_pp_local(self._ctx, which, _dev_ptr(keys, n, "keys"), _dev_ptr(pays, n, "pays"), n,
                                   C.c_void_p(fine_hist.data_ptr()), self._sptr(stream)))

    def pp_push(self, which: int, all_hist, peer_ptrs, cap_tuples: int, n: int, stream=None):
        bases = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        _check(self._L.gj_pp_push(self._ctx, which, C.c_void_p(all_hist.data_ptr()), bases, cap_tuples, n, self._sptr(stream)))

    def pp_join(self, own_R: int, own_S: int, cap_R: int, cap_S: int, stream=None):
        _check(self._L.gj_pp_join(self._ctx, C.c_void_p(own_R), C.c_void_p(own_S), cap_R, cap_S, self._sptr(stream)))

    def pp_finish(self):
        """Returns (matches, checksum, tuples received of R, of S, phase_ms dict)."""
        m, c, a, b = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        ph = (C.c_float * 5)()
        _check(self._L.gj_pp_finish(self._ctx, C.byref(m), C.byref(c), C.byref(a), C.byref(b), ph))
        names = ("local_R_ms", "push_R_ms", "local_S_ms", "push_S_ms", "join_ms")
        return int(m.value), int(c.value), int(a.value), int(b.value), dict(zip(names, (float(x) for x in ph)))

    def pp_plan(self):
        a, b = C.c_uint32(), C.c_uint32()
        _check(self._L.gj_pp_plan(self._ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    # -- sharded "partition, copy, partition" pipeline (gj_pcp_*) ------------------------------
    def pcp_begin(self, n_R_global: int, n_S_global: int, n_gpus: int, rank: int, local_bits: int, stream=None):
        _check(self._L.gj_pcp_begin(self._ctx, n_R_global, n_S_global, n_gpus, rank, local_bits, self._sptr(stream)))

    def pcp_plan(self):
        """(gpu bits, local bits of the source-side pass, bits of the receiver-side pass)."""
        bits = (C.c_uint32 * 3)()
        _check(self._L.gj_pcp_plan(self._ctx, bits))
        return tuple(int(x) for x in bits)

    def pcp_hist(self, which: int, keys, coarse_hist, stream=None):
        n = keys.numel()
        _check(self._L.gj_pcp_hist(self._ctx, which, _dev_ptr(keys, n, "keys"), n, C.c_void_p(coarse_hist.data_ptr()), self._sptr(stream)))

    def pcp_hist_exchange(self, which: int, coarse_hist, peer_ctrl_ptrs, own_ctrl_ptr: int, all_hist, stream=None):
        """Coarse histograms through the peers' control blocks (no collective): fills all_hist [G][2^(g+bl)]."""
        ctrl = (C.c_void_p * len(peer_ctrl_ptrs))(*[C.c_void_p(int(p)) for p in peer_ctrl_ptrs])
        _check(self._L.gj_pcp_hist_exchange(self._ctx, which, C.c_void_p(coarse_hist.data_ptr()), ctrl, C.c_void_p(own_ctrl_ptr),
                                            C.c_void_p(all_hist.data_ptr()), self._sptr(stream)))

    def pcp_part(self, which: int, keys, pays, all_hist, own_ptr: int, cap_tuples: int, stream=None):
        n = keys.numel()
        _check(self._L.gj_pcp_part(self._ctx, which, _dev_ptr(keys, n, "keys"), _dev_ptr(pays, n, "pays"),
                                   C.c_void_p(all_hist.data_ptr()), C.c_void_p(own_ptr), cap_tuples, self._sptr(stream)))

    def pcp_copy(self, which: int, peer_ptrs, peer_ctrl_ptrs, n_stages: int = 1, stream=None):
        bases = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        ctrl = (C.c_void_p * len(peer_ctrl_ptrs))(*[C.c_void_p(int(p)) for p in peer_ctrl_ptrs])
        _check(self._L.gj_pcp_copy(self._ctx, which, bases, ctrl, n_stages, self._sptr(stream)))

    def pcp_recv(self, which: int, own_ptr: int, own_ctrl_ptr: int, cap_tuples: int, stream=None, result_out=None):
        """result_out: optional int64 CUDA tensor of >= 2 elements receiving the local {matches, checksum}."""
        _check(self._L.gj_pcp_recv(self._ctx, which, C.c_void_p(own_ptr), C.c_void_p(own_ctrl_ptr), cap_tuples,
                                   C.c_void_p(result_out.data_ptr() if result_out is not None else 0), self._sptr(stream)))

    def pcp_finish(self, phases: bool = True):
        """Returns (matches, checksum, tuples received of R, of S, phase_ms dict, (gpu bits, source bits, receiver bits))."""
        m, c, a, b = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        ph, bits = (C.c_float * 7)(), (C.c_uint32 * 3)()
        _check(self._L.gj_pcp_finish(self._ctx, C.byref(m), C.byref(c), C.byref(a), C.byref(b), ph if phases else None, bits))
        names = ("part_R_ms", "copy_R_ms", "recv_R_ms", "part_S_ms", "copy_S_ms", "recv_S_ms", "tail_ms")
        return (int(m.value), int(c.value), int(a.value), int(b.value), dict(zip(names, (float(x) for x in ph))),
                tuple(int(x) for x in bits))

    # -- synthetic data ---------------------------------------------------------------------
    def generate_unique(self, keys, pays, row_begin: int, n_total: int, seed: int, pay_seed: int):
        n = keys.numel()
        _check(self._L.gj_generate_unique(self._ctx, _dev_ptr(keys, n, "keys"), _dev_ptr(pays, n, "pays"),
                                          row_begin, n, n_total, seed, pay_seed))
