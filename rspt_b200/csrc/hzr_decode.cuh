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

#include <type_traits>

#include "common.cuh"
#include "hzr_encode.cuh"

namespace rspt {

constexpr int kDecodeThreads = kMaxSegs;  // one thread per decode segment
constexpr int kLutBits = 12;
constexpr int kPairBits = 11;  // index width of the pair table (32-bit entries in the same 8 KB)
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
            // A legitimate block stream is capped at header + in_size (hzr_encode.c:377-382), COPY carries
            // exactly the block (:307-339) and FILL one byte (:352-364).  Anything else is rejected here:
            // the decode / index / verify kernels stage payload_len bytes in shared memory sized for the
            // largest legitimate payload.
            const uint32_t bl = blk_len(s, b);
            if ((mode == MODE_HUFF && plen > bl) || (mode == MODE_COPY && plen != bl) || (mode == MODE_FILL && plen != 1u)) { err = 1; break; }
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

// sequential bit reader over the payload staged in shared memory (word 0 = payload bytes 0..3),
// LSB-first, 32-bit refills into a 64-bit window
struct BitReader {
    const uint32_t* words;
    uint32_t widx;
    unsigned long long buf;
    uint32_t cnt;
    __device__ __forceinline__ void init(const uint32_t* pay_words, uint32_t bitpos)
    {
        words = pay_words;
        widx = bitpos >> 5;
        buf = (unsigned long long)words[widx] | ((unsigned long long)words[widx + 1] << 32);
        widx += 2;
        const uint32_t drop = bitpos & 31u;
        buf >>= drop;
        cnt = 64u - drop;
    }
    __device__ __forceinline__ void refill()
    {
        if (cnt <= 32u) {
            buf |= (unsigned long long)words[widx] << cnt;
            cnt += 32u;
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

constexpr uint32_t kDecPayWords = kBlock / 4 + 16;  // cap of the payload staging: 16-byte chunks at any alignment + slack
constexpr size_t kDecodeSmem = (size_t)kDecPayWords * 4;
constexpr uint32_t kLongFlag = 0x8000u;
constexpr uint32_t kLongEnd = 0x1FFu;  // flagged table entries: index of the first long symbol of the chain, kLongEnd = none

// RecoverTree (dec:263-333), iteratively, by ONE thread: pre-order, 0 = branch, 1 + 9-bit symbol =
// leaf.  Only the code word of every leaf is needed: a stack of (code, depth) of pending right
// children.  cw[sym] = code | len << 27 (must be zeroed by the caller).  Returns the number of
// tree bits, or 0xFFFFFFFF on a malformed tree.
__device__ __forceinline__ uint32_t recover_tree(const uint32_t* payw, uint32_t plen, uint32_t* cw)
{
    BitReader r;
    r.init(payw, 0);
    uint32_t nodes = 0, leaves = 0, bits_used = 0;
    uint32_t st_code[40], st_depth[40];
    int sp = 0;
    uint32_t code = 0, depth = 0;
    for (;;) {
        if (nodes >= 2 * kNumSymbols - 1 || depth > 27u) return 0xFFFFFFFFu;
        ++nodes;
        r.refill();
        const uint32_t leaf = r.take(1);
        ++bits_used;
        if (leaf) {
            const uint32_t sym = r.take(9);
            bits_used += 9;
            if (sym >= (uint32_t)kNumSymbols || leaves >= (uint32_t)kNumSymbols || cw[sym] != 0u) return 0xFFFFFFFFu;
            // lone leaf: 1-bit code (dec:306 `hzr_max(bits, 1)`)
            cw[sym] = code | (max(depth, 1u) << 27);
            ++leaves;
            if (sp == 0) break;
            --sp;
            code = st_code[sp]; depth = st_depth[sp];
        } else {
            if (sp >= 40) return 0xFFFFFFFFu;
            st_code[sp] = code | (1u << depth); st_depth[sp] = depth + 1; ++sp;
            depth = depth + 1;
        }
    }
    return bits_used > plen * 8u ? 0xFFFFFFFFu : bits_used;
}

// look-up table on the next kLutBits bits (whole CTA): a warp per symbol, lanes over the
// 2^(12 - len) entries that end in its code; symbols with longer codes go to the list `longs`
// (*nlong must be 0 on entry).  The caller initialises lut to kLongFlag and synchronises after.
template <class T, int BITS = kLutBits>
__device__ __forceinline__ void build_lut(const uint32_t* cw_tab, T* lut, uint16_t* longs, uint32_t* nlong)
{
    const uint32_t lane = lane_id();
    for (uint32_t sym = warp_id(); sym < (uint32_t)kNumSymbols; sym += (blockDim.x >> 5)) {
        const uint32_t cw = cw_tab[sym];
        if (cw == 0u) continue;
        const uint32_t len = cw >> 27, code = cw & 0x07FFFFFFu;
        if (len <= (uint32_t)BITS) {
            const T e = (T)(sym | (len << 9));
            for (uint32_t i = lane; i < (1u << (BITS - len)); i += 32) lut[(i << len) | code] = e;
        } else if (lane == 0) {
            longs[atomicAdd(nlong, 1u)] = (uint16_t)sym;
        }
    }
}

// Long codes: the table entry of their first kLutBits bits (kLongFlag | index) heads a chain through
// `next` of the long symbols that share those bits, so a long code costs a compare or two instead of
// a scan of all long symbols (a lane on this path stalls its whole warp).  Whole CTA; the table must have
// been initialised to kLongFlag | kLongEnd; synchronise after.
template <class T, int BITS = kLutBits>
__device__ __forceinline__ void chain_long_codes(const uint32_t* cw_tab, T* lut, const uint16_t* longs, uint16_t* next,
                                                 uint32_t nlong)
{
    for (uint32_t j = threadIdx.x; j < nlong; j += blockDim.x) {
        const uint32_t prefix = cw_tab[longs[j]] & ((1u << BITS) - 1u);
        if (sizeof(T) == 4) {  // 32-bit entries: one exchange
            next[j] = (uint16_t)(atomicExch(reinterpret_cast<uint32_t*>(lut) + prefix, kLongFlag | j) & kLongEnd);
            continue;
        }
        uint32_t* word = reinterpret_cast<uint32_t*>(lut) + (prefix >> 1);
        const uint32_t shift = (prefix & 1u) * 16u;
        uint32_t old = *word, seen;
        do {
            seen = old;
            old = atomicCAS(word, seen, (seen & ~(0xFFFFu << shift)) | ((kLongFlag | j) << shift));
        } while (old != seen);
        next[j] = (uint16_t)((seen >> shift) & kLongEnd);
    }
}

// One CTA per hzr block.  The payload is staged in shared memory (coalesced, re-aligned); the
// code table comes from the decode index (sc_codes); a 12-bit look-up table maps the next bits to
// (symbol, length), longer codes are matched against the short list of long code words.  Every
// thread decodes the tokens that start in its 128-byte segment and writes exactly that segment.
// Streams that arrive without an index (CPU reference) get one from k_hzr_build_index first.
__global__ void __launch_bounds__(kDecodeThreads, 3) k_hzr_decode(const uint8_t* __restrict__ src, Shape s,
                                                                   const DecBlk* __restrict__ dec,
                                                                   const uint32_t* __restrict__ sc_bit,
                                                                   const uint16_t* __restrict__ sc_skip,
                                                                   const uint32_t* __restrict__ sc_codes,
                                                                   uint8_t* __restrict__ planes, int32_t* __restrict__ status,
                                                                   uint8_t* __restrict__ seg_xor, uint32_t segs_per_plane,
                                                                   uint32_t pair_max_bits)
{
    extern __shared__ __align__(16) uint32_t payw[];  // payload words (+ zero slack)
    // 8 KB of look-up table in one of two shapes, chosen per block:
    //   kLutBits bits -> 16-bit entries: sym | len << 9, or kLongFlag | chain head
    //   kPairBits bits -> 32-bit entries: the same in the low half; high half: a second literal whose whole
    //   code the same bits also hold -- bit 31, both lengths << 24, its byte << 16
    __shared__ __align__(16) uint16_t s_lut[1 << kLutBits];
    uint32_t* s_lut32 = reinterpret_cast<uint32_t*>(s_lut);
    __shared__ uint32_t s_cw[kSymStride];            // code | len << 27 per symbol, 0 = unused
    __shared__ uint16_t s_long[kSymStride];          // symbols whose code is longer than the table
    __shared__ uint16_t s_next[kSymStride];          // chains of the long symbols that share their first kLutBits bits
    __shared__ uint32_t s_meta[4];                   // tree_end_bit, error, long count

    const uint32_t blk = blockIdx.x, tid = threadIdx.x;
    const DecBlk d = dec[blk];
    if (d.mode == kModeInactive) return;
    uint32_t f, k, b;
    blk_decode(s, blk, f, k, b);
    uint8_t* out = planes + ((size_t)f * s.nb_alloc + k) * s.plane_stride + (size_t)b * kBlock;
    uint4* out4 = reinterpret_cast<uint4*>(out);
    const uint32_t n = d.out_n, nq = (n + 15u) >> 4;
    const uint8_t* pay = src + d.payload_off;
    // xor of the bytes of every 128-byte output segment, for the inverse transform's first scan
    // (k_planes_to_samples_fast: a segment is one of its pieces), so that it need not read the planes for it
    uint8_t* my_xor = seg_xor ? seg_xor + ((size_t)f * s.nb_alloc + k) * segs_per_plane + (size_t)b * kMaxSegs : nullptr;
    const uint32_t nseg_all = (n + kSegBytes - 1) / kSegBytes;

    if (d.mode == MODE_FILL || d.mode == kModeZero) {
        const uint32_t v = d.mode == MODE_FILL ? pay[0] * 0x01010101u : 0u;  // memset (dec:362-370)
        for (uint32_t i = tid; i < nq; i += blockDim.x) out4[i] = make_uint4(v, v, v, v);
        if (my_xor)
            for (uint32_t i = tid; i < nseg_all; i += blockDim.x)
                my_xor[i] = (uint8_t)((min((uint32_t)kSegBytes, n - i * kSegBytes) & 1u) ? (v & 0xFFu) : 0u);
        return;
    }
    const uintptr_t pa = (uintptr_t)pay;
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(pa & ~(uintptr_t)3);
    const uint32_t lead = (uint32_t)(pa & 3u), sh = lead * 8u;
    if (d.mode == MODE_COPY) {
        if (d.payload_len != n) {  // "Encoded / decoded size mismatch (COPY)" dec:351-355
            if (tid == 0) status[f] = -4;
            for (uint32_t i = tid; i < nq; i += blockDim.x) out4[i] = make_uint4(0, 0, 0, 0);
            if (my_xor)
                for (uint32_t i = tid; i < nseg_all; i += blockDim.x) my_xor[i] = 0;
            return;
        }
        const uint32_t naw = (lead + n + 3u) >> 2, nw = (n + 3u) >> 2;
        uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
        for (uint32_t i = tid; i < nw; i += blockDim.x) {
            const uint32_t lo = __ldg(aw + i), hi = (i + 1 < naw) ? __ldg(aw + i + 1) : 0u;
            out32[i] = __funnelshift_r(lo, hi, sh);
        }
        if (my_xor) {
            for (uint32_t sg = tid; sg < nseg_all; sg += blockDim.x) {
                const uint32_t w0 = sg * (kSegBytes / 4), w1 = min(nw, w0 + kSegBytes / 4);
                uint32_t x = 0;
                for (uint32_t i = w0; i < w1; ++i) {
                    const uint32_t lo = __ldg(aw + i), hi = (i + 1 < naw) ? __ldg(aw + i + 1) : 0u;
                    uint32_t v = __funnelshift_r(lo, hi, sh);
                    if (4u * i + 4u > n) v &= (1u << (8u * (n - 4u * i))) - 1u;  // bytes beyond the block
                    x ^= v;
                }
                x ^= x >> 16;
                x ^= x >> 8;
                my_xor[sg] = (uint8_t)x;
            }
        }
        return;
    }

    // ---- MODE_HUFF: stage the payload, asynchronously (cp.async) and as it lies in the stream: the 16-byte
    // chunks that hold it go to shared memory unshifted, the payload starts lead16 bytes into the staging and
    // the bit reader starts that much later.  The copies land while the tables are built.
    const uint32_t plen = d.payload_len;
    const uint32_t lead16 = (uint32_t)(pa & 15u), n16 = (lead16 + plen + 15u) >> 4;
    {
        const uint4* a16 = reinterpret_cast<const uint4*>(pa & ~(uintptr_t)15);
        const uint32_t pay_s = smem_addr(payw);
        for (uint32_t i = tid; i < n16; i += blockDim.x)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pay_s + 16u * i), "l"(a16 + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (tid < 8u) payw[4u * n16 + tid] = 0u;  // the bit reader looks two words ahead
    }
    // Pairs are worth it where codes are short and tokens many (1/2 .. pair_max_bits payload bits per output byte):
    // quantised coefficient planes, smooth upper planes.  Sparse blocks would only pay for the extra pass, and
    // planes with ~6-bit codes rarely hold two codes in the window and want the longer single-symbol table.
    const bool use_pairs = plen * 16u >= d.out_n && plen * 8u <= pair_max_bits * d.out_n;
    for (uint32_t i = tid; i < kSymStride; i += blockDim.x) s_cw[i] = i < (uint32_t)kNumSymbols ? sc_codes[(size_t)blk * kSymStride + i] : 0u;
    if (tid == 0) { s_meta[0] = 0; s_meta[1] = 0; s_meta[2] = 0; }
    if (!use_pairs) {
        for (uint32_t i = tid; i < (1u << kLutBits) / 2; i += blockDim.x) s_lut32[i] = (kLongFlag | kLongEnd) * 0x00010001u;
        __syncthreads();
        build_lut(s_cw, s_lut, s_long, &s_meta[2]);
        __syncthreads();
        chain_long_codes(s_cw, s_lut, s_long, s_next, s_meta[2]);
    } else {
        for (uint32_t i = tid; i < (1u << kPairBits); i += blockDim.x) s_lut32[i] = kLongFlag | kLongEnd;
        __syncthreads();
        build_lut<uint32_t, kPairBits>(s_cw, s_lut32, s_long, &s_meta[2]);
        __syncthreads();
        chain_long_codes<uint32_t, kPairBits>(s_cw, s_lut32, s_long, s_next, s_meta[2]);
        __syncthreads();
        // pair pass: only the high half of an entry changes, so the look-ups of the other threads into the
        // low halves stay valid while it runs
        for (uint32_t i = tid; i < (1u << kPairBits); i += blockDim.x) {
            const uint32_t e1 = s_lut32[i], len1 = (e1 >> 9) & 15u;
            if (!(e1 & kLongFlag) && (e1 & 511u) < 256u && len1 < (uint32_t)kPairBits) {
                const uint32_t e2 = s_lut32[i >> len1] & 0xFFFFu, len2 = (e2 >> 9) & 15u;
                if (!(e2 & kLongFlag) && (e2 & 511u) < 256u && len1 + len2 <= (uint32_t)kPairBits)
                    s_lut32[i] = e1 | 0x80000000u | ((len1 + len2) << 24) | ((e2 & 255u) << 16);
            }
        }
    }
    // The whole output block is cleared first (coalesced 128-bit stores that the L2 merges with the word
    // stores below), so a zero run only advances the write position and the token loop is the same
    // straight-line code for literals and runs.
    for (uint32_t i = tid; i < nq; i += blockDim.x) out4[i] = make_uint4(0, 0, 0, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const uint32_t nseg = (n + kSegBytes - 1) / kSegBytes;
    uint32_t my_err = 0;
    if (tid < nseg) {
        uint32_t bitpos = sc_bit[(size_t)blk * kMaxSegs + tid];
        const uint32_t skip = sc_skip[(size_t)blk * kMaxSegs + tid];  // bytes covered by a zero run that started earlier
        uint32_t end_bit = tid + 1 < nseg ? sc_bit[(size_t)blk * kMaxSegs + tid + 1] : 0xFFFFFFFFu;
        const uint32_t seg0 = tid * kSegBytes, seg_len = min((uint32_t)kSegBytes, n - seg0);
        uint32_t xb = 0;
        if (skip < seg_len) {
            const uint32_t limit_bits = plen * 8u;
            if (bitpos > limit_bits) { my_err = 1; bitpos = 0; }
            end_bit = min(end_bit, limit_bits);
            // the thread's bytes are gathered into the open word w (bytes of the word at and beyond pos are
            // zero); a word goes out with one 32-bit store when the position leaves it, unless it is empty
            uint8_t* dst = out + seg0;
            uint32_t pos = skip, w = 0, xacc = 0;
            BitReader r;
            r.init(payw, bitpos + 8u * lead16);
            const uint32_t lut_s = smem_addr(s_lut);
            auto token_loop = [&](auto pairs_t) {
            constexpr bool PAIRS = decltype(pairs_t)::value;
            while (!my_err && bitpos < end_bit && pos < seg_len) {
                r.refill();
                uint32_t e = PAIRS ? lds_u32(lut_s + 4u * r.peek(kPairBits)) : lds_u16(lut_s + 2u * r.peek(kLutBits));
                if (e & kLongFlag) {
                    // code longer than the table: match the few long code words (dec:418-431 walks the tree)
                    uint32_t j = e & kLongEnd;
                    e = 0;
                    while (j != kLongEnd) {
                        const uint32_t sym = s_long[j], cw = s_cw[sym], len = cw >> 27;
                        if (((uint32_t)r.buf & ((1u << len) - 1u)) == (cw & 0x07FFFFFFu)) {
                            e = sym | (len << 9);
                            break;
                        }
                        j = s_next[j];
                    }
                    if (e == 0u) { my_err = 1; break; }
                    // consume it here and top the window up, so that the extra bits of a run are there
                    r.skip(e >> 9);
                    bitpos += e >> 9;
                    r.refill();
                    e &= 511u;
                }
                // two literals at once when the entry has them and both belong to this segment
                const bool two = PAIRS && (e >> 31) != 0u && pos + 2u <= seg_len;
                const uint32_t len = (two ? e >> 24 : e >> 9) & 15u, sym = e & 511u;
                r.skip(len);  // <= kLutBits bits: at least 21 are left in the window
                bitpos += len;
                const bool run = sym >= 256u;  // symbol 0 (a zero run of one) is handled as a literal
                uint32_t adv = two ? 2u : 1u;
                if (run) {
                    const uint32_t kk = sym - 256u;                              // run class 0..4 (hzr_internal.h:117-121)
                    const uint32_t eb = (0xE8420u >> (4u * kk)) & 15u;             // 0, 2, 4, 8, 14 extra bits
                    const uint32_t ev = (uint32_t)r.buf & ((1u << eb) - 1u);
                    r.skip(eb);
                    bitpos += eb;
                    const uint32_t z = ev + (kk == 4u ? 279u : (0x17070302u >> (8u * kk)) & 255u);
                    if (seg0 + pos + z > n) { my_err = 1; break; }  // "Output buffer full" dec:473-476
                    adv = min(z, seg_len - pos);  // the rest of the run is the next segment's `skip`
                }
                const uint32_t np = pos + adv, sh = (pos & 3u) * 8u;
                const uint32_t val = two ? sym | ((e >> 8) & 0xFF00u) : sym;  // the literal byte(s)
                const uint32_t wv = run ? w : w | (val << sh);
                const bool cross = (np >> 2) != (pos >> 2);
                if (cross) {
                    if (wv) *reinterpret_cast<uint32_t*>(dst + (pos & ~3u)) = wv;
                    xacc ^= wv;
                }
                // a pair that starts in a word's last byte leaves its second byte in the next word
                w = cross ? ((!PAIRS || run) ? 0u : __funnelshift_l(val, 0u, sh)) : wv;
                pos = np;
            }
            };
            if (use_pairs) token_loop(std::true_type{});
            else token_loop(std::false_type{});
            if (bitpos > limit_bits) my_err = 1;
            if (w) *reinterpret_cast<uint32_t*>(dst + (pos & ~3u)) = w;
            xb = xacc ^ w;  // xor of the segment's bytes
            xb ^= xb >> 16;
            xb ^= xb >> 8;
        }
        if (my_xor) my_xor[tid] = (uint8_t)xb;
    }
    if (my_err) status[f] = -4;
}

// ------------------------------------------------------------------------------------------
// Decode index for streams that arrive without one (written by the CPU reference): one CTA per
// HUFF block.  The token stream has no sync points, but a prefix code re-synchronises by itself
// after a few tokens, so the block's payload bits are cut into equal sub-sequences, one per
// thread, and every thread decodes from a guessed start (the sub-sequence boundary; thread 0
// from the true start behind the tree) to the first token boundary inside the next
// sub-sequence, which becomes that neighbour's start.  Threads whose start moved decode again;
// the starts are exact once nothing moves (each round fixes at least one more sub-sequence,
// in practice two or three rounds do).  A scan of the bytes every sub-sequence produces gives
// the output position of every token, and a last walk writes, for every 128-byte output
// segment, the bit offset of the first token that starts in it and the bytes an earlier zero run
// still covers -- the same index k_hzr_encode emits -- plus the block's code table.
// (hzr_decode.c has no counterpart: DecodeSingleBlock :335-567 is a sequential walk.)
// ------------------------------------------------------------------------------------------
constexpr int kIndexThreads = 512;
constexpr uint32_t kSubMinBits = 64;  // sub-sequences are longer than the longest token (27 + 14 bits)

struct TokenDecoder {
    const uint16_t* lut;
    const uint32_t* cw;
    const uint16_t* longs;
    const uint16_t* chain;
    // one token at the reader's position: bits consumed (0 = no code word matches) and bytes produced
    __device__ __forceinline__ uint32_t next(BitReader& r, uint32_t& out_bytes) const
    {
        r.refill();
        uint32_t e = lut[r.peek(kLutBits)];
        if (e & kLongFlag) {
            uint32_t j = e & kLongEnd;
            e = 0;
            while (j != kLongEnd) {
                const uint32_t sym = longs[j], c = cw[sym], len = c >> 27;
                if (((uint32_t)r.buf & ((1u << len) - 1u)) == (c & 0x07FFFFFFu)) {
                    e = sym | (len << 9);
                    break;
                }
                j = chain[j];
            }
            if (e == 0u) return 0u;
        }
        uint32_t len = e >> 9;
        const uint32_t sym = e & 511u;
        r.skip(len);
        out_bytes = 1u;
        if (sym >= 256u) {
            out_bytes = 2u;
            if (sym > 256u) {
                const uint32_t eb = sym_extra_bits(sym);
                r.refill();
                out_bytes = r.take(eb) + (sym == 257u ? 3u : sym == 258u ? 7u : sym == 259u ? 23u : 279u);
                len += eb;
            }
        }
        return len;
    }
};

__global__ void __launch_bounds__(kIndexThreads, 2) k_hzr_build_index(const uint8_t* __restrict__ src, Shape s,
                                                                       const DecBlk* __restrict__ dec,
                                                                       uint32_t* __restrict__ sc_bit,
                                                                       uint16_t* __restrict__ sc_skip,
                                                                       uint32_t* __restrict__ sc_codes,
                                                                       int32_t* __restrict__ status)
{
    extern __shared__ __align__(16) uint32_t payw[];  // payload words (+ zero slack)
    __shared__ __align__(16) uint16_t s_lut[1 << kLutBits];
    __shared__ uint32_t s_cw[kSymStride];
    __shared__ uint16_t s_long[kSymStride];
    __shared__ uint16_t s_next[kSymStride];
    __shared__ uint32_t s_meta[4];                    // tree bits, error, long count
    __shared__ uint32_t s_start[kIndexThreads + 1];   // first token boundary of every sub-sequence
    __shared__ uint32_t s_wsum[kIndexThreads / 32];

    const uint32_t blk = blockIdx.x, tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    const DecBlk d = dec[blk];
    if (d.mode != MODE_HUFF) return;  // COPY / FILL / unparsed frames need no index
    uint32_t f, k, b;
    blk_decode(s, blk, f, k, b);
    const uint32_t n = d.out_n, nseg = (n + kSegBytes - 1) / kSegBytes;
    uint32_t* my_bit = sc_bit + (size_t)blk * kMaxSegs;
    uint16_t* my_skip = sc_skip + (size_t)blk * kMaxSegs;
    const uint8_t* pay = src + d.payload_off;
    const uintptr_t pa = (uintptr_t)pay;
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(pa & ~(uintptr_t)3);
    const uint32_t lead = (uint32_t)(pa & 3u), sh = lead * 8u;
    const uint32_t plen = d.payload_len, pwords = (plen + 3u) >> 2, naw = (lead + plen + 3u) >> 2;
    for (uint32_t i = tid; i < pwords + 4u; i += blockDim.x) {
        uint32_t v = 0;
        if (i < pwords) {
            const uint32_t lo = __ldg(aw + i), hi = (i + 1 < naw) ? __ldg(aw + i + 1) : 0u;
            v = __funnelshift_r(lo, hi, sh);
        }
        payw[i] = v;
    }
    for (uint32_t i = tid; i < (1u << kLutBits) / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(s_lut)[i] = (kLongFlag | kLongEnd) * 0x00010001u;
    for (uint32_t i = tid; i < kSymStride; i += blockDim.x) s_cw[i] = 0;
    __syncthreads();
    if (tid == 0) {
        const uint32_t tb = recover_tree(payw, plen, s_cw);
        s_meta[0] = tb;
        s_meta[1] = tb == 0xFFFFFFFFu;
        s_meta[2] = 0;
    }
    __syncthreads();
    bool bad = s_meta[1] != 0u;
    if (!bad) {
        build_lut(s_cw, s_lut, s_long, &s_meta[2]);
        __syncthreads();
        chain_long_codes(s_cw, s_lut, s_long, s_next, s_meta[2]);
        __syncthreads();
    }
    const uint32_t limit = plen * 8u, t0 = bad ? 0u : s_meta[0];
    // sub-sequences: nsub equal pieces of [t0, limit), each longer than any token
    const uint32_t span = limit - t0;
    uint32_t nsub = span / kSubMinBits;
    nsub = nsub < 1u ? 1u : (nsub > (uint32_t)kIndexThreads ? (uint32_t)kIndexThreads : nsub);
    const uint32_t sub = (span + nsub - 1u) / nsub;
    const uint32_t lo = t0 + tid * sub, hi = min(limit, lo + sub);  // tokens that START in [lo, hi) are mine
    const bool live = !bad && tid < nsub && lo < limit;
    const TokenDecoder td{s_lut, s_cw, s_long, s_next};
    if (tid <= nsub) s_start[tid] = min(lo, limit);
    __syncthreads();
    uint32_t my_start = 0xFFFFFFFFu, land = 0, cnt = 0;
    if (!bad) {
        for (;;) {
            const uint32_t st = live ? s_start[tid] : 0u;
            const bool redo = live && st != my_start;
            if (redo) {
                my_start = st;
                BitReader r;
                uint32_t pos = st;
                cnt = 0;
                if (pos < hi) r.init(payw, pos);
                while (pos < hi) {
                    uint32_t ob = 0;
                    uint32_t l = td.next(r, ob);
                    if (l == 0u) {  // no code word here (only from a wrong start, or a corrupt stream): slip one bit
                        l = 1u;
                        ob = 0u;
                        r.skip(1);
                    }
                    pos += l;
                    cnt += ob;
                }
                land = min(pos, limit);
            }
            __syncthreads();
            bool moved = false;
            if (redo && tid + 1 < nsub && s_start[tid + 1] != land) {
                s_start[tid + 1] = land;
                moved = true;
            }
            if (!__syncthreads_or(moved)) break;
        }
    }
    // output position of every sub-sequence: exclusive scan of the byte counts
    uint32_t v = live ? cnt : 0u, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += y;
    }
    if (lane == 31) s_wsum[wid] = inc;
    __syncthreads();
    uint32_t opos = inc - v, total = 0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
        const uint32_t t = s_wsum[w];
        if (w < wid) opos += t;
        total += t;
    }
    // the tokens must produce at least n bytes (zero padding of the last byte may decode into a
    // few more: they are ignored, like the reference stops at the output size, dec:440)
    if (total < n) bad = true;
    if (bad) {
        for (uint32_t i = tid; i < nseg; i += blockDim.x) {
            my_bit[i] = 0xFFFFFFFFu;  // k_hzr_decode reports the frame and writes zeros
            my_skip[i] = 0;
        }
        for (uint32_t i = tid; i < kSymStride; i += blockDim.x) sc_codes[(size_t)blk * kSymStride + i] = 0u;
        if (tid == 0) status[f] = -4;
        return;
    }
    for (uint32_t i = tid; i < kSymStride; i += blockDim.x) sc_codes[(size_t)blk * kSymStride + i] = s_cw[i];
    if (live) {
        BitReader r;
        uint32_t pos = my_start, op = opos;
        if (pos < hi) r.init(payw, pos);
        while (pos < hi && op < n) {
            uint32_t ob = 0;
            uint32_t l = td.next(r, ob);
            if (l == 0u) {  // corrupt stream: the sequential decoder would fail here (dec:431)
                status[f] = -4;
                l = 1u;
                ob = 0u;
                r.skip(1);
            }
            // segment boundaries B in [op, op + ob): B == op -> this token starts the segment;
            // B inside a zero run -> resume behind the token, with the rest of the run to skip
            const uint32_t end = min(op + ob, n);
            for (uint32_t B = (op + kSegBytes - 1u) & ~(uint32_t)(kSegBytes - 1); B < end; B += kSegBytes) {
                my_bit[B / kSegBytes] = B == op ? pos : pos + l;
                my_skip[B / kSegBytes] = (uint16_t)(B == op ? 0u : op + ob - B);
            }
            pos += l;
            op += ob;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Integrity check without decoding: hzr_verify (lib_hzr/hzr_decode.c:569-624) for every hzr
// stream of every frame.  The header walk is k_frame_parse's; here one CTA per block stages the
// payload in shared memory and compares its CRC-32C with the block header's
// (hzr_decode.c:606-611: "CRC32 check failed").  status[f] = RSPT_E_CRC (-6) on a mismatch.
// ------------------------------------------------------------------------------------------
constexpr int kVerifyThreads = 256;
constexpr int kVerifyZtSel = 1;  // log2(kVerifyThreads / 128)

__global__ void __launch_bounds__(kVerifyThreads) k_hzr_verify(const uint8_t* __restrict__ src, Shape s,
                                                                const DecBlk* __restrict__ dec,
                                                                const CrcConst* __restrict__ cc,
                                                                int32_t* __restrict__ status, Counters* __restrict__ ctr)
{
    extern __shared__ __align__(16) uint32_t payw[];
    __shared__ __align__(16) uint32_t s_zt[1024];
    __shared__ uint32_t s_red[33];
    const uint32_t blk = blockIdx.x, tid = threadIdx.x;
    const DecBlk d = dec[blk];
    if (d.mode == kModeInactive || d.mode == kModeZero) return;  // not there / frame already flagged
    for (uint32_t i = tid; i < 256; i += blockDim.x)
        reinterpret_cast<uint4*>(s_zt)[i] = __ldg(reinterpret_cast<const uint4*>(&cc->zt[kVerifyZtSel][0][0]) + i);
    const uint8_t* pay = src + d.payload_off;
    const uintptr_t pa = (uintptr_t)pay;
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(pa & ~(uintptr_t)3);
    const uint32_t lead = (uint32_t)(pa & 3u), sh = lead * 8u;
    const uint32_t plen = d.payload_len, pwords = (plen + 3u) >> 2, naw = (lead + plen + 3u) >> 2;
    for (uint32_t i = tid; i < pwords; i += blockDim.x) {
        const uint32_t lo = __ldg(aw + i), hi = (i + 1 < naw) ? __ldg(aw + i + 1) : 0u;
        payw[i] = __funnelshift_r(lo, hi, sh);
    }
    __syncthreads();
    const uint32_t crc = block_crc32c(payw, plen, s_zt, cc, s_red);
    if (tid == 0) {
        const uint32_t want = ld_le32(pay - 5);  // block header: size u16, crc u32, mode u8
        if (crc != want) {
            uint32_t f, k, b;
            blk_decode(s, blk, f, k, b);
            status[f] = -6;
            atomicAdd(&ctr->crc_failures, 1ull);
        }
    }
}

}  // namespace rspt
