"""bench.py's own payload / closed-form arithmetic (torch, independent of the engine and the oracle)
against the product's host helper gj_payload_of_key and the oracle's join checker."""
import numpy as np
import torch

import bench


def test_payload_matches_product_and_oracle(gj, orc):
    rng = np.random.default_rng(3)
    keys = np.concatenate((rng.integers(-2**31, 2**31, 2000), [0, 1, -1, 2**31 - 1, -2**31])).astype(np.int32)
    for seed in (bench.PAY_SEED_R, bench.PAY_SEED_S, 7):
        got = bench.payload_i64(torch, torch.from_numpy(keys), seed).numpy()
        want = np.array([gj.payload_of_key(int(k), seed) for k in keys], dtype=np.int64)
        assert (got == want).all()
        assert (got == orc.payload_of_keys(keys, seed)).all()


def test_closed_form_matches_oracle_join(orc):
    n_r = 5000
    rng = np.random.default_rng(5)
    Rk = rng.permutation(n_r).astype(np.int32)
    Sk = rng.integers(-3, n_r + 40, 20000).astype(np.int32)           # some keys outside [0, n_r)
    want = orc.join_check(Rk, orc.payload_of_keys(Rk, bench.PAY_SEED_R), Sk, orc.payload_of_keys(Sk, bench.PAY_SEED_S))
    assert bench.closed_form(torch, torch.from_numpy(Sk), n_r, chunk=3000) == (want.matches, want.checksum)
    assert want.checksum == sum(int(a) * int(b) for a, b in zip(orc.payload_of_keys(Sk[(Sk >= 0) & (Sk < n_r)], 40),
                                                                orc.payload_of_keys(Sk[(Sk >= 0) & (Sk < n_r)], 50))) % 2**64


def test_unique_closed_form_equals_oracle_closed_form(orc):
    n = 100_000
    keys = torch.from_numpy(np.random.default_rng(1).permutation(n).astype(np.int32))
    assert bench.closed_form(torch, keys, n) == (n, orc.unique_join_checksum(0, n, bench.PAY_SEED_R, bench.PAY_SEED_S))


def test_config_is_a_function_of_workload_and_n():
    a, b = bench.config_of("B", 8), bench.config_of("B", 8)
    assert a == b and a["global_R"] == 8 * 128_000_000 and "8 GPUs" in a["parallelism"]
    assert bench.config_of("cfg5", 8)["per_gpu_R"] == 250_000_000
    assert bench.config_of("B", 1)["parallelism"] == "1 GPU"


def test_cpu_sample_is_the_whole_workload_b_and_a_bounded_sample_beyond(monkeypatch):
    monkeypatch.setitem(bench.WORKLOADS, "tinyB", (4096, 4096, "unique", 0.0))
    Rk, Sk, what = bench.cpu_join_sample("tinyB", 1)
    assert Rk.size == Sk.size == 4096 and "whole workload" in what and sorted(Rk.tolist()) == list(range(4096))
    monkeypatch.setattr(bench, "CPU_SAMPLE_MAX", 5000)
    Rk, Sk, what = bench.cpu_join_sample("tinyB", 8)          # 8 GPUs, weak scaling: 32768 tuples per side in total
    assert Rk.size == Sk.size == 5000 and "sample" in what


def test_config5_record_speedup_uses_the_cached_one_gpu_value(tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "CFG5_CACHE", str(tmp_path / "cfg5.json"))
    one = bench.cfg5_record(50.0, 1)
    assert one["speedup_vs_1gpu"] == 1.0 and abs(one["value"] - 4e9 / 0.05) < 1
    eight = bench.cfg5_record(8.0, 8, shuffle="pcp")
    assert abs(eight["speedup_vs_1gpu"] - 50.0 / 8.0) < 1e-9 and "this box" in eight["one_gpu_source"] and eight["per_gpu_R"] == 250_000_000
    monkeypatch.setattr(bench, "CFG5_CACHE", str(tmp_path / "missing.json"))
    fallback = bench.cfg5_record(8.0, 8)
    assert fallback["one_gpu_value"] == bench.CFG5_1GPU_MEASURED["value"] and "profiles/" in fallback["one_gpu_source"]
