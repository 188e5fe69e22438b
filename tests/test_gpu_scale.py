"""BASELINE.json configs at their full sizes on one B200, through size-independent properties:
encode -> decode is the identity on EVERY frame, frame offsets are a strictly increasing scan, and
the streams of a sample of frames (first and last of every batch) are byte-identical to the oracle.
Frames are generated on the device batch by batch (1 M frames of config 2 are 295 GB)."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def R():
    from rspt_b200 import packer
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return packer


def _run_config(R, oracle, kind, bps, ch, ns, total_frames, batch, lossless, check_every=1):
    fb = bps * ch * ns
    p = R.SignalPacker(kind, bps, ch, ns, 3, max_batch_frames=batch)
    o = oracle.OraclePacker(kind, bps, ch, ns, 3)
    out = p.alloc_output(batch)
    raw = torch.empty(batch * fb, dtype=torch.uint8, device="cuda")
    dec = torch.empty_like(raw)
    comp_total = 0
    nb = (total_frames + batch - 1) // batch
    for bi in range(nb):
        first = bi * batch
        n = min(batch, total_frames - first)
        R.synth_ecg(first, n, bps, ch, ns, out=raw)
        b = p.compress_batch(raw[: n * fb], out=out)
        p.decompress_batch(b, out=dec)
        offs = b.offsets[: n + 1]
        assert bool((offs[1:] > offs[:-1]).all()) and int(offs[0].item()) == 0
        comp_total += int(offs[n].item())
        if lossless:
            assert torch.equal(raw[: n * fb], dec[: n * fb]), (kind, bi)
        if bi % check_every == 0:
            for i in (0, n - 1):
                lo, hi = int(offs[i].item()), int(offs[i + 1].item())
                got = bytes(b.stream[lo:hi].cpu().numpy())
                host = raw[i * fb:(i + 1) * fb].cpu().numpy()
                want = o.compress(host)
                assert got == want, (kind, first + i)
                if not lossless:
                    assert dec[i * fb:(i + 1) * fb].cpu().numpy().tobytes() == o.decompress(want)[0], (kind, first + i)
    p.close()
    return total_frames * fb / comp_total


def test_config2_xdelta_hzr_one_million_frames(R, oracle):
    """configs[1]: xdelta_hzr, 12 ch x 3 B x 8192 samples, 1 M frames (64 batches of 15 625)."""
    frames = int(os.environ.get("RSPT_SCALE_FRAMES", "1000000"))
    cr = _run_config(R, oracle, "xdelta_hzr", 3, 12, 8192, frames, 15625, lossless=True, check_every=4)
    assert 3.0 < cr < 4.5, cr


def test_config3_hadamard_one_million_frames(R, oracle):
    """configs[2]: hadamard, 12 ch x 4 B x 4096 samples, 1 M frames (197 GB, 64 batches of 15 625);
    integer transform, so the stream of the sampled frames is bit-exact and their decode equals the
    reference's decode of the same stream."""
    frames = int(os.environ.get("RSPT_SCALE_FRAMES", "1000000"))
    cr = _run_config(R, oracle, "hadamard", 4, 12, 4096, frames, 15625, lossless=False, check_every=4)
    assert 3.5 < cr < 6.0, cr


def test_config4_dct_sampled_against_the_reference(R, oracle):
    """configs[3]: dct, 12 ch x 4 B x 4096 samples.  The reference's dct costs ~2 s per frame on one
    core, so the GPU runs all 1 M frames (197 GB, 64 batches of 15 625) and a sample of them -- first
    and last frame of the first and of the last batch -- is compared: streams that the reference
    decodes to within the stated tolerance of its own (PRDN within 0.01 percentage points), identical
    channel means, CR in the expected range, and the GPU's own decode of every frame stays close to
    its input (PRDN over all frames, reduced on the device)."""
    bps, ch, ns, batch = 4, 12, 4096, 15625
    total = int(os.environ.get("RSPT_SCALE_FRAMES", "1000000")) // batch * batch
    fb = bps * ch * ns
    p = R.SignalPacker.new_dct(bps, ch, ns, max_batch_frames=batch)
    o = oracle.OraclePacker("dct", bps, ch, ns)
    out = p.alloc_output(batch)
    raw = torch.empty(batch * fb, dtype=torch.uint8, device="cuda")
    dec = torch.empty_like(raw)
    comp_total = 0
    num = den = 0.0
    nbatches = total // batch
    sample_batches = sorted({0, nbatches // 3, (2 * nbatches) // 3, nbatches - 1})
    per_batch = -(-64 // len(sample_batches))      # >= 64 frames in all
    sampled = []                                    # (host frame, GPU byte planes) of the sampled frames
    for bi in range(nbatches):
        R.synth_ecg(bi * batch, batch, bps, ch, ns, out=raw)
        if bi in sample_batches:
            for i in np.linspace(0, batch - 1, per_batch).astype(int):
                fr = raw[i * fb:(i + 1) * fb]
                planes, _ = p.debug_planes(fr)
                sampled.append((fr.cpu().numpy(), planes[0]))
        b = p.compress_batch(raw, out=out)
        p.decompress_batch(b, out=dec)
        offs = b.offsets[: batch + 1]
        assert bool((offs[1:] > offs[:-1]).all())
        comp_total += int(offs[batch].item())
        a, c = R.prdn_terms(raw, dec, batch, bps, ch, ns)
        num += a
        den += c
        for i in ((0, batch - 1) if bi in (0, total // batch - 1) else ()):
            lo, hi = int(offs[i].item()), int(offs[i + 1].item())
            got = bytes(b.stream[lo:hi].cpu().numpy())
            host = raw[i * fb:(i + 1) * fb].cpu().numpy()
            want = o.compress(host)
            assert got[:1 + 3 * ch] == want[:1 + 3 * ch]              # method byte + channel means: exact
            assert abs(len(got) - len(want)) <= max(8, len(want) // 100)
            wd, _ = o.decompress(want)
            gd, _ = o.decompress(got)                                  # the reference decodes the GPU's stream
            assert abs(oracle.prdn(host, wd, bps, ch, ns) - oracle.prdn(host, gd, bps, ch, ns)) < 0.01
            own = dec[i * fb:(i + 1) * fb].cpu().numpy().tobytes()
            assert abs(oracle.prdn(host, gd, bps, ch, ns) - oracle.prdn(host, own, bps, ch, ns)) < 0.01
    # the +-1 LSB / stated-fraction clause on >= 64 frames spread over the run: quantised coefficients
    # (stencil undone) against the reference's O(n^2) transform, one reference instance per host thread
    from concurrent.futures import ThreadPoolExecutor
    from test_gpu_parity import DCT_EQUAL_FRACTION, dct_coefficient_diff
    assert len(sampled) >= 64

    def ref_words(host):
        return oracle.OraclePacker("dct", bps, ch, ns).transform(host)[0]

    with ThreadPoolExecutor(max_workers=max(1, min(32, os.cpu_count() or 1))) as ex:
        words = list(ex.map(ref_words, [h for h, _ in sampled]))
    n_all = n_bad = 0
    for ww, (_, planes) in zip(words, sampled):
        diff = dct_coefficient_diff(ww, planes)
        assert np.abs(diff).max() <= 1
        n_all += diff.size
        n_bad += int((diff != 0).sum())
    assert n_bad <= int(n_all * (1.0 - DCT_EQUAL_FRACTION)), (n_bad, n_all)
    cr = total * fb / comp_total
    prdn = 100.0 * (num / den) ** 0.5
    assert 15.0 < cr < 40.0, cr
    assert prdn < 3.0, prdn
    p.close()
