"""Unpack the reference's two .7z test fixtures with the Python stdlib only (no 7z tool in the
image).  Both archives are a single LZMA2 folder with an unencoded header.
Used by make_golden.py in the build container; never at test time on the GPU box."""
import lzma
import struct
import zlib


def _num(b, i):
    """7z variable-length number."""
    first = b[i]
    i += 1
    mask, val = 0x80, 0
    for k in range(8):
        if not (first & mask):
            val |= (first & (mask - 1)) << (8 * k)
            return val, i
        val |= b[i] << (8 * k)
        i += 1
        mask >>= 1
    return val, i


def unpack_7z_single_lzma2(path: str) -> bytes:
    raw = open(path, "rb").read()
    assert raw[:6] == b"7z\xbc\xaf\x27\x1c"
    nh_off, nh_size = struct.unpack_from("<QQ", raw, 12)
    h = raw[32 + nh_off: 32 + nh_off + nh_size]
    i = 0
    assert h[i] == 0x01 and h[i + 1] == 0x04 and h[i + 2] == 0x06
    i += 3
    packpos, i = _num(h, i)
    npack, i = _num(h, i)
    assert npack == 1 and h[i] == 0x09
    i += 1
    packsize, i = _num(h, i)
    assert h[i] == 0x00
    i += 1
    # folder: 07 0b 01 00 01 21 21 01 <prop> 0c <unpacksize>
    assert h[i:i + 8] == bytes([0x07, 0x0B, 0x01, 0x00, 0x01, 0x21, 0x21, 0x01])
    i += 8
    prop = h[i]
    i += 1
    assert h[i] == 0x0C
    i += 1
    unpack, i = _num(h, i)
    dict_size = 0xFFFFFFFF if prop == 40 else (2 | (prop & 1)) << (prop // 2 + 11)
    dec = lzma.LZMADecompressor(format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA2, "dict_size": dict_size}])
    out = dec.decompress(raw[32 + packpos: 32 + packpos + packsize])
    assert len(out) == unpack
    return out


if __name__ == "__main__":
    for name in ("data_stream.7z", "12_chan_32bit_34199_samples_r00000135fghd8.raw.7z"):
        d = unpack_7z_single_lzma2("/root/reference/lib_rspt_test/" + name)
        print(name, len(d), "%08x" % zlib.crc32(d))
