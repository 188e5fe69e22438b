// hzr decoder for sm_100a.  Replaces decompress_i32's chunk walk
// (lib_signalpacker/signal_packer_base.cpp:98-121) and lib_hzr/hzr_decode.c (hzr_decode :626-674,
// DecodeSingleBlock :335-567, RecoverTree :263-333).  Like the reference, the block CRC is not
// checked on decode (hzr_decode.c:343).
//
// Parallelism: frames and hzr blocks are located by a cheap header walk (one thread per frame);
// each block is decoded by one CTA.  Inside a block the token stream has no sync points, so the
// decoder is seeded from the encoder's out-of-band index (per 256 output bytes: the bit offset of
// the first token that starts there + the leading bytes covered by a zero run that started
// earlier; rspt_gpu_compress_batch's d_sidecar) and every thread decodes one segment
// through a 10-bit lookup table.  Streams without an index (produced by the CPU reference) are
// decoded by a single thread per block.
#pragma once

#include "common.cuh"
#include "hzr_encode.cuh"

namespace rspt {

constexpr int kDecodeThreads = 256;
constexpr int kLutBits = 10;
constexpr uint32_t kModeZero = 3;      // frame failed to parse: emit zeros
constexpr uint32_t kModeInactive = 255;

struct DecBlk {
    unsigned long long payload_off;  // byte offset of the payload in the stream buffer
    uint32_t payload_len;
    uint32_t out_n;
    uint32_t mode;
    uint32_t pad;
};

__device__ __forceinline__ uint32_t ld_le32(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// one thread per frame: method byte, header, nb chunks, block headers
__global__ void __launch_bounds__(128) k_frame_parse(const uint8_t* __restrict__ src, const uint64_t* __restrict__ offsets,
                                                      Shape s, const uint8_t* __restrict__ frame_nb_in,
                                                      const uint32_t* __restrict__ nb_state, uint32_t n_frames,
                                                      DecBlk* __restrict__ dec, uint8_t* __restrict__ headers,
                                                      uint8_t* __restrict__ dec_nb, int32_t* __restrict__ status,
                                                      Counters* __restrict__ ctr)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f == 0) atomicAdd(&ctr->frames_decompressed, (unsigned long long)n_frames);
    if (f >= n_frames) return;
    const unsigned long long base = offsets[f], end = offsets[f + 1];
    uint32_t nb = frame_nb_in ? frame_nb_in[f] : *nb_state;
    int err = 0;
    if (nb < 1 || nb > s.nb_alloc) {
        err = 1;
        nb = nb < 1 ? 1 : s.nb_alloc;
    }
    unsigned long long pos = base;
    if (end < base + 1 + s.hdr_bytes) err = 1;
    if (!err) {
        // "ERROR: compression method unsupported." in the reference (e.g. xdelta.cpp:78-79)
        if (src[pos] != s.method) err = 1;
        ++pos;
        for (uint32_t i = 0; i < s.hdr_bytes; ++i) headers[(size_t)f * s.hdr_bytes + i] = src[pos + i];
        pos += s.hdr_bytes;
    }
    DecBlk* row = dec + (size_t)f * s.nb_alloc * s.nblk;
    for (uint32_t k = 0; k < s.nb_alloc; ++k) {
        for (uint32_t b = 0; b < s.nblk; ++b) {
            DecBlk d;
            d.payload_off = 0; d.payload_len = 0; d.out_n = blk_len(s, b); d.pad = 0;
            d.mode = k < nb ? kModeZero : kModeInactive;
            row[k * s.nblk + b] = d;
        }
    }
    for (uint32_t k = 0; k < nb && !err; ++k) {
        if (pos + 8 > end) { err = 1; break; }
        const uint32_t len = ld_le32(src + pos);            // chunk length (base.cpp:103)
        const unsigned long long cend = pos + 4 + len;
        if (cend > end || len < 4) { err = 1; break; }
        if (ld_le32(src + pos + 4) != s.N) { err = 1; break; }  // hzr master header (dec:644)
        unsigned long long q = pos + 8;
        for (uint32_t b = 0; b < s.nblk; ++b) {
            if (q + 7 > cend) { err = 1; break; }
            const uint32_t plen = ((uint32_t)src[q] | ((uint32_t)src[q + 1] << 8)) + 1u;  // dec:342
            const uint32_t mode = src[q + 6];
            if (mode > MODE_FILL || q + 7 + plen > cend) { err = 1; break; }
            DecBlk& d = row[k * s.nblk + b];
            d.payload_off = q + 7;
            d.payload_len = plen;
            d.mode = mode;
            q += 7 + plen;
        }
        if (!err && q != cend) err = 1;
        pos = cend;
    }
    if (!err && pos != end) err = 1;
    if (err) {
        for (uint32_t k = 0; k < nb; ++k)
            for (uint32_t b = 0; b < s.nblk; ++b) row[k * s.nblk + b].mode = kModeZero;
    }
    dec_nb[f] = (uint8_t)nb;
    status[f] = err ? -4 : 0;
}

// sequential bit reader over global memory, LSB-first, 32-bit aligned refills
struct BitReader {
    const uint32_t* words;   // aligned base
    uint32_t nwords;         // aligned words that overlap the payload
    uint32_t widx;
    unsigned long long buf;
    uint32_t cnt;
    __device__ __forceinline__ void init(const uint8_t* payload, uint32_t plen, uint32_t bitpos)
    {
        const uintptr_t a = (uintptr_t)payload;
        words = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t lead = (uint32_t)(a & 3u);
        nwords = (lead + plen + 3u) >> 2;
        const uint32_t abs_bit = lead * 8u + bitpos;
        widx = abs_bit >> 5;
        buf = 0;
        cnt = 0;
        refill();
        const uint32_t drop = abs_bit & 31u;
        buf >>= drop;
        cnt -= drop;
    }
    __device__ __forceinline__ void refill()
    {
        while (cnt <= 32) {
            const uint32_t w = widx < nwords ? __ldg(words + widx) : 0u;
            buf |= (unsigned long long)w << cnt;
            cnt += 32;
            ++widx;
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void skip(uint32_t n)
    {
        buf >>= n;
        cnt -= n;
    }
    __device__ __forceinline__ uint32_t take(uint32_t n)
    {
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
};

constexpr uint32_t kDecStageWords = kBlock / 4;
constexpr size_t kDecodeSmem = (size_t)kDecStageWords * 4;

__device__ __forceinline__ uint32_t dec_swizzle(uint32_t w) { return w ^ ((w >> 6) & 31u); }

__global__ void __launch_bounds__(kDecodeThreads) k_hzr_decode(const uint8_t* __restrict__ src, Shape s,
                                                                const DecBlk* __restrict__ dec,
                                                                const uint32_t* __restrict__ sc_bit,
                                                                const uint16_t* __restrict__ sc_skip,
                                                                uint8_t* __restrict__ planes, int32_t* __restrict__ status)
{
    extern __shared__ __align__(16) uint32_t stg[];  // decoded block, word-swizzled per 256-byte segment
    __shared__ uint16_t s_lut[1 << kLutBits];        // <0x8000: sym | len << 9 ; >=0x8000: node index
    __shared__ uint16_t s_child[2 * kNumSymbols][2];
    __shared__ int16_t s_nsym[2 * kNumSymbols];      // >= 0 leaf symbol, -1 branch
    __shared__ uint32_t s_leaf_code[kNumSymbols];
    __shared__ uint16_t s_leaf_info[kNumSymbols];    // sym | len << 9
    __shared__ uint32_t s_meta[4];                   // n_leaves, tree_end_bit, error, single-leaf flag

    const uint32_t blk = blockIdx.x, tid = threadIdx.x;
    const DecBlk d = dec[blk];
    if (d.mode == kModeInactive) return;
    uint32_t f, k, b;
    blk_decode(s, blk, f, k, b);
    uint8_t* out = planes + ((size_t)f * s.nb_alloc + k) * s.plane_stride + (size_t)b * kBlock;
    uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
    const uint32_t n = d.out_n, nw = (n + 3u) >> 2;
    const uint8_t* pay = src + d.payload_off;

    if (d.mode == MODE_FILL || d.mode == kModeZero) {
        const uint32_t v = d.mode == MODE_FILL ? pay[0] * 0x01010101u : 0u;  // memset (dec:362-370)
        for (uint32_t i = tid; i < nw; i += blockDim.x) out32[i] = v;
        return;
    }
    if (d.mode == MODE_COPY) {
        if (d.payload_len != n) {  // "Encoded / decoded size mismatch (COPY)" dec:351-355
            if (tid == 0) status[f] = -4;
            for (uint32_t i = tid; i < nw; i += blockDim.x) out32[i] = 0;
            return;
        }
        const uintptr_t a = (uintptr_t)pay;
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t lead = (uint32_t)(a & 3u), sh = lead * 8u;
        const uint32_t naw = (lead + n + 3u) >> 2;
        for (uint32_t i = tid; i < nw; i += blockDim.x) {
            const uint32_t lo = __ldg(aw + i), hi = (i + 1 < naw) ? __ldg(aw + i + 1) : 0u;
            out32[i] = __funnelshift_r(lo, hi, sh);
        }
        return;
    }

    // ---- MODE_HUFF
    for (uint32_t i = tid; i < kDecStageWords; i += blockDim.x) stg[i] = 0;
    for (uint32_t i = tid; i < (1u << kLutBits); i += blockDim.x) s_lut[i] = 0;
    if (tid == 0) {
        // RecoverTree (dec:263-333), iteratively: pre-order, 0 = branch, 1 + 9-bit symbol = leaf
        BitReader r;
        r.init(pay, d.payload_len, 0);
        uint32_t nodes = 0, leaves = 0, err = 0, bits_used = 0;
        uint32_t st_parent[40], st_code[40], st_depth[40];
        int sp = 0;
        uint32_t parent = 0xFFFFu, which = 0, code = 0, depth = 0;
        for (;;) {
            if (nodes >= 2 * kNumSymbols - 1 || depth > 31) { err = 1; break; }
            const uint32_t me = nodes++;
            if (parent != 0xFFFFu) s_child[parent][which] = (uint16_t)me;
            r.refill();
            const uint32_t leaf = r.take(1);
            ++bits_used;
            if (leaf) {
                const uint32_t sym = r.take(9);
                bits_used += 9;
                if (sym >= kNumSymbols || leaves >= kNumSymbols) { err = 1; break; }
                s_nsym[me] = (int16_t)sym;
                s_leaf_code[leaves] = code;
                s_leaf_info[leaves] = (uint16_t)(sym | (depth << 9));
                ++leaves;
                if (sp == 0) break;
                --sp;
                parent = st_parent[sp]; which = 1; code = st_code[sp]; depth = st_depth[sp];
            } else {
                s_nsym[me] = -1;
                if (depth == kLutBits) s_lut[code] = (uint16_t)(0x8000u | me);
                if (sp >= 40) { err = 1; break; }
                st_parent[sp] = me; st_code[sp] = code | (1u << depth); st_depth[sp] = depth + 1; ++sp;
                parent = me; which = 0; depth = depth + 1;
            }
        }
        if (bits_used > d.payload_len * 8u) err = 1;
        s_meta[0] = leaves;
        s_meta[1] = bits_used;
        s_meta[2] = err;
        s_meta[3] = (nodes == 1);
    }
    __syncthreads();
    if (s_meta[2]) {
        if (tid == 0) status[f] = -4;
        for (uint32_t i = tid; i < nw; i += blockDim.x) out32[i] = 0;
        return;
    }
    const bool single = s_meta[3] != 0;
    for (uint32_t l = tid; l < s_meta[0]; l += blockDim.x) {
        const uint32_t info = s_leaf_info[l];
        uint32_t len = info >> 9;
        const uint32_t code = s_leaf_code[l];
        if (single) len = 1;  // lone leaf: 1-bit code (dec:306 `hzr_max(bits, 1)`)
        if (len <= (uint32_t)kLutBits) {
            const uint16_t e = (uint16_t)((info & 511u) | (len << 9));
            for (uint32_t i = 0; i < (1u << (kLutBits - len)); ++i) s_lut[(i << len) | code] = e;
        }
    }
    __syncthreads();

    const uint32_t nseg = (n + kSegBytes - 1) / kSegBytes;
    uint32_t my_err = 0;
    const bool indexed = sc_bit != nullptr;
    if (indexed ? tid < nseg : tid == 0) {
        uint32_t bitpos, outpos, end_bit;
        if (indexed) {
            bitpos = sc_bit[(size_t)blk * kMaxSegs + tid];
            outpos = tid * kSegBytes + sc_skip[(size_t)blk * kMaxSegs + tid];
            end_bit = tid + 1 < nseg ? sc_bit[(size_t)blk * kMaxSegs + tid + 1] : 0xFFFFFFFFu;
        } else {
            bitpos = s_meta[1];
            outpos = 0;
            end_bit = 0xFFFFFFFFu;
        }
        const uint32_t limit_bits = d.payload_len * 8u;
        if (bitpos > limit_bits || outpos > n) my_err = 1;
        BitReader r;
        r.init(pay, d.payload_len, my_err ? 0 : bitpos);
        uint8_t* sb = reinterpret_cast<uint8_t*>(stg);
        while (!my_err && bitpos < end_bit && outpos < n) {
            r.refill();
            const uint32_t e = s_lut[r.peek(kLutBits)];
            uint32_t sym;
            if (!(e & 0x8000u)) {
                const uint32_t len = (e >> 9) & 15u;
                if (len == 0) { my_err = 1; break; }
                sym = e & 511u;
                r.skip(len);
                bitpos += len;
            } else {
                uint32_t node = e & 0x3FFu;
                r.skip(kLutBits);
                bitpos += kLutBits;
                while (s_nsym[node] < 0) {  // codes longer than the table: walk the tree (dec:418-431)
                    if (r.cnt == 0) r.refill();
                    node = s_child[node][r.take(1)];
                    ++bitpos;
                }
                sym = (uint32_t)s_nsym[node];
            }
            if (sym < 256u) {
                const uint32_t w = outpos >> 2;
                sb[(dec_swizzle(w) << 2) | (outpos & 3u)] = (uint8_t)sym;
                ++outpos;
            } else {
                uint32_t z = 2;
                if (sym > 256u) {
                    const uint32_t eb = sym_extra_bits(sym);
                    r.refill();
                    const uint32_t ev = r.take(eb);
                    bitpos += eb;
                    z = ev + (sym == 257u ? 3u : sym == 258u ? 7u : sym == 259u ? 23u : 279u);
                }
                if (outpos + z > n) { my_err = 1; break; }  // "Output buffer full" dec:473-476
                outpos += z;
            }
            if (bitpos > limit_bits) my_err = 1;
        }
    }
    if (my_err) status[f] = -4;
    __syncthreads();
    for (uint32_t i = tid; i < nw; i += blockDim.x) out32[i] = stg[dec_swizzle(i)];
}

}  // namespace rspt
