"""Differential pinning of the oracle port against the UNMODIFIED reference compiled into
oracle/_ref/libref.so.  Skipped where that library was not built.  CPU only."""
import numpy as np
import pytest

from conftest import needs_ref

pytestmark = needs_ref


def _buffers(rng):
    """hzr inputs across all three block modes, zero-run classes and block boundaries."""
    out = []
    for n in (1, 2, 3, 7, 22, 23, 278, 279, 1000, 16662, 16663, 33324, 65535, 65536, 65537, 140000):
        out.append(np.zeros(n, np.uint8))
        out.append(rng.integers(0, 256, n, dtype=np.uint8))
        out.append(rng.integers(0, 4, n, dtype=np.uint8))
        sparse = np.zeros(n, np.uint8)
        k = max(1, n // 50)
        sparse[rng.integers(0, n, k)] = rng.integers(1, 256, k, dtype=np.uint8)
        out.append(sparse)
        lap = np.clip(np.rint(rng.laplace(0, 6, n)), -120, 120).astype(np.int8).view(np.uint8)
        out.append(lap)
        out.append(np.full(n, 255, np.uint8))
    # near the COPY/HUFF boundary: skewed two-symbol data and nearly uniform data
    for p in (0.5, 0.9, 0.99):
        out.append((rng.random(65536) < p).astype(np.uint8) * 3 + 1)
    out.append(rng.integers(0, 200, 65536, dtype=np.uint8))
    out.append(rng.integers(1, 256, 300, dtype=np.uint8))
    return out


def test_hzr_stream_bit_exact(oracle):
    rng = np.random.default_rng(7)
    for buf in _buffers(rng):
        a = oracle.hzr_encode(buf, "port")
        b = oracle.hzr_encode(buf, "reference")
        assert a == b, (buf.size, buf[:8])
        for impl in ("port", "reference"):
            dec, ok = oracle.hzr_decode(b, buf.size, impl)
            assert ok and dec == buf.tobytes()


def test_crc32c_matches(oracle):
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 3, 4, 5, 63, 64, 65, 1000, 65536):
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        assert oracle.crc32c(buf, "port") == oracle.crc32c(buf, "reference")


@pytest.mark.parametrize("kind", ["xdelta_hzr", "hzr", "hadamard", "dct"])
def test_packers_random_shapes(oracle, kind):
    rng = np.random.default_rng(hash(kind) & 0xFFFF)
    for trial in range(14):
        bps = int(rng.integers(1, 5))
        ch = int(rng.integers(1, 5))
        if kind in ("hadamard", "dct"):
            ns = 1 << int(rng.integers(3, 10 if kind == "dct" else 13))
        else:
            ns = int(rng.integers(1, 30000))
        nb = int(rng.integers(1, 5))
        amp = 1 << int(rng.integers(2, 8 * bps))
        walk = np.cumsum(rng.integers(-amp // 16 - 1, amp // 16 + 2, (ns, ch)), axis=0)
        x = np.clip(walk, -(1 << (8 * bps - 1)), (1 << (8 * bps - 1)) - 1).astype(np.int32)
        raw = x.astype("<i4").view(np.uint8).reshape(ns, ch, 4)[:, :, :bps].copy().reshape(-1)
        o = oracle.OraclePacker(kind, bps, ch, ns, nb)
        r = oracle.RefPacker(kind, bps, ch, ns, nb)
        for rep in range(2):  # twice: the xdelta plane-count state carries over
            co, cr = o.compress(raw), r.compress(raw)
            assert co == cr, (kind, bps, ch, ns, nb, rep)
            do, uo = o.decompress(cr)
            dr, ur = r.decompress(cr)
            assert do == dr and uo == ur == len(cr)


def test_xdelta_escalation_matches_bit_range_rule(oracle):
    """the plane count ends at max(nb0, planes needed by the post-xor words)."""
    rng = np.random.default_rng(11)
    for trial in range(40):
        bps = int(rng.integers(2, 5))
        ch = int(rng.integers(1, 4))
        ns = int(rng.integers(16, 400))
        nb0 = int(rng.integers(1, bps + 1))
        amp = 1 << int(rng.integers(2, 8 * bps - 1))
        x = rng.integers(-amp, amp, (ns, ch)).astype(np.int32)
        raw = x.astype("<i4").view(np.uint8).reshape(ns, ch, 4)[:, :, :bps].copy().reshape(-1)
        o = oracle.OraclePacker("xdelta_hzr", bps, ch, ns, nb0)
        r = oracle.RefPacker("xdelta_hzr", bps, ch, ns, nb0)
        assert o.compress(raw) == r.compress(raw)
        words, _ = oracle.OraclePacker("xdelta_hzr", bps, ch, ns, 4).transform(raw)
        need = nb0
        for nb in range(nb0, bps):
            top = words.astype(np.int64) >> (8 * nb - 1)
            lim = (1 << (8 * (bps - nb) + 1)) - 1
            ok = np.all(((top & lim) == 0) | ((top & lim) == lim))
            if ok:
                break
            need = nb + 1
        assert o.nb == need, (bps, ch, ns, nb0, o.nb, need)


def test_full_12ch_fixture_known_answers(oracle):
    """BASELINE.md section 3 rows for the whole 34199-sample fixture (needs /root/reference)."""
    import json, os, sys, zlib
    path = "/root/reference/lib_rspt_test/12_chan_32bit_34199_samples_r00000135fghd8.raw.7z"
    if not os.path.exists(path):
        pytest.skip("reference fixtures not present")
    here = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, here)
    from extract_fixtures import unpack_7z_single_lzma2
    data = np.frombuffer(unpack_7z_single_lzma2(path), np.uint8)
    want = json.load(open(os.path.join(here, "golden_full_fixture.json")))
    for c in want["cases"]:
        comp = oracle.OraclePacker(c["kind"], 4, 12, 34199, c["nb"]).compress(data)
        assert len(comp) == c["len"] and "%08x" % zlib.crc32(comp) == c["crc32"]


# the band-pass the reference's own pipeline uses (rspt_test.cpp:123-125) and two shorter filters
IIR5_N = [1.00000000000, -3.14332095199, 3.70064088865, -1.97083923944, 0.41351972908]
IIR5_D = [0.06722876941, 0.00000000000, -0.13445753881, 0.00000000000, 0.06722876941]


@needs_ref
def test_prefilter_restatement_equals_reference(oracle):
    """oracle_prefilter_iir / _fir vs the compiled i_filter (lib_filter/*.cpp) driven the way
    rspt_test.cpp:116-136 drives it: same bytes, including the state that leaks from one channel
    into the next through the single filter object."""
    rng = np.random.default_rng(3)
    for bps, ch, ns in ((3, 12, 8192), (4, 3, 1000), (2, 2, 500), (1, 1, 64)):
        raws = oracle.synth_ecg(5, 2, bps, ch, ns, amplitude=20000 if bps >= 3 else (3000 if bps == 2 else 50))
        for f in raws:
            for nc, init in ((5, 2000), (3, 100), (2, 0), (4, 7)):
                n, d = IIR5_N[:nc], IIR5_D[:nc]
                assert np.array_equal(oracle.prefilter_iir(f, bps, ch, ns, n, d, init, "port"),
                                      oracle.prefilter_iir(f, bps, ch, ns, n, d, init, "reference")), (bps, ch, ns, nc)
            for K in (1, 5, 31):
                k = rng.normal(size=K)
                k /= np.abs(k).sum()
                assert np.array_equal(oracle.prefilter_fir(f, bps, ch, ns, k, "port"),
                                      oracle.prefilter_fir(f, bps, ch, ns, k, "reference")), (bps, ch, ns, K)
