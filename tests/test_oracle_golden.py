"""The oracle port (oracle/rspt_oracle.c) against the committed golden vectors, which were
produced by the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import zlib

import numpy as np
import pytest


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def test_golden_inputs_intact(golden, golden_inputs):
    for k, meta in golden["inputs"].items():
        assert golden_inputs[k].size == meta["bytes"], k
        assert sha(golden_inputs[k]) == meta["sha256"], k  # also pins the synthetic generator


def _cases(golden):
    return golden["cases"]


def test_port_reproduces_every_golden_case(golden, golden_inputs, oracle):
    for case in _cases(golden):
        if case["kind"] == "dct" and case["ns"] >= 4096 and case["ch"] > 3:
            continue  # covered by the slow test below
        data = golden_inputs[case["input"]]
        fb = case["bps"] * case["ch"] * case["ns"]
        p = oracle.OraclePacker(case["kind"], case["bps"], case["ch"], case["ns"], case["nb"])
        for f, want in enumerate(case["frames"]):
            src = data[f * fb:(f + 1) * fb]
            comp = p.compress(src)
            assert len(comp) == want["len"], case["name"]
            assert "%08x" % zlib.crc32(comp) == want["crc32"], case["name"]
            assert sha(comp) == want["sha256"], case["name"]
            dec, used = p.decompress(comp)
            assert used == len(comp)
            assert sha(dec) == want["dec_sha256"], case["name"]
            if "prdn" in want:
                assert abs(oracle.prdn(src, dec, case["bps"], case["ch"], case["ns"]) - want["prdn"]) < 1e-9
            else:
                assert dec == src.tobytes()
        assert p.nb == case["final_nb"], case["name"]


def test_port_dct_4096_12ch(golden, golden_inputs, oracle):
    for case in _cases(golden):
        if not (case["kind"] == "dct" and case["ns"] >= 4096 and case["ch"] > 3):
            continue
        data = golden_inputs[case["input"]]
        fb = case["bps"] * case["ch"] * case["ns"]
        p = oracle.OraclePacker("dct", case["bps"], case["ch"], case["ns"])
        comp = p.compress(data[:fb])
        assert sha(comp) == case["frames"][0]["sha256"], case["name"]


def test_readme_known_answer(oracle):
    """sine 8192 x 32-bit, xdelta_hzr nb=3 -> 2028 B (README's 2022 is stale)."""
    import math
    sine = np.array([int(math.sin(i / 100.0) * 1000.0) for i in range(8192)], np.int32).view(np.uint8)
    comp = oracle.OraclePacker("xdelta_hzr", 4, 1, 8192, 3).compress(sine)
    assert len(comp) == 2028 and "%08x" % zlib.crc32(comp) == "672647f0"


def test_crc32c_known_answers(oracle):
    # standard CRC-32C check values
    assert oracle.crc32c(b"123456789") == 0xE3069283
    assert oracle.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert oracle.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert oracle.crc32c(b"") == 0


def test_average_32_unsigned_division_quirk(oracle):
    """utils.cpp:30-40: floor for power-of-two lengths, garbage for negative sums otherwise."""
    L = oracle.oracle_lib()
    a = np.array([-6, -5, -5, -5], np.int32)
    assert L.oracle_average_32(a.ctypes.data, 4) == -6  # -21/4 floors
    b = np.array([-5, -5, -6], np.int32)
    assert L.oracle_average_32(b.ctypes.data, 3) == 1431655760


def test_hzr_modes_and_bounds(oracle):
    rng = np.random.default_rng(1)
    fill = np.full(1000, 7, np.uint8)
    enc = oracle.hzr_encode(fill)
    assert len(enc) == 4 + 8 and enc[10] == 2 and enc[11] == 7
    zeros = np.zeros(70000, np.uint8)
    enc = oracle.hzr_encode(zeros)
    assert len(enc) == 4 + 8 + 8 and enc[10] == 2
    noise = rng.integers(0, 256, 65536, dtype=np.uint8)
    enc = oracle.hzr_encode(noise)
    assert enc[10] == 0 and len(enc) == 4 + 7 + 65536  # incompressible -> COPY
    for data in (fill, zeros, noise):
        dec, ok = oracle.hzr_decode(oracle.hzr_encode(data), data.size)
        assert ok and dec == data.tobytes()
    assert oracle.oracle_lib().oracle_hzr_max_compressed_size(65537) == 4 + 14 + 65537
