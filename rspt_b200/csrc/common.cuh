// Shared definitions for the sm_100a signal-packer kernels.
//
// Vocabulary (follows the reference): a FRAME is one compress() call's input,
// [ns][ch][bps] interleaved little-endian samples; its N = ch*ns sample words are split into
// `nb` byte PLANES (signal_packer_base.cpp:40-68); each plane is an hzr stream cut into BLOCKS
// of <= 65536 bytes (hzr_encode.c:528-539).  Inside a block, work is divided into 64-byte
// STRIPS, one per thread.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rspt {

constexpr int kNumSymbols = 261;       // hzr_internal.h:114
constexpr int kSymStride = 264;        // padded row for per-block tables
constexpr int kTreeWords = 92;         // >= ceil((11*261-1)/32)
constexpr uint32_t kBlock = 65536;     // HZR_MAX_BLOCK_SIZE, hzr_internal.h:109
constexpr uint32_t kRunCap = 16662;    // hzr_encode.c:149
constexpr int kStrip = 64;             // bytes per thread-strip
constexpr int kSegStrips = 2;          // strips per decode segment (decode index granularity: 128 B)
constexpr int kSegBytes = kStrip * kSegStrips;
constexpr int kMaxSegs = kBlock / kSegBytes;  // 512 per block

enum : uint32_t { MODE_COPY = 0, MODE_HUFF = 1, MODE_FILL = 2 };  // hzr_internal.h:98-101

// Per-block plan written by the tree kernel and consumed by layout + encode.
struct BlkInfo {
    uint32_t payload_len;  // bytes after the 7-byte block header
    uint32_t total_bits;   // tree bits + token bits (HUFF)
    uint16_t tree_nbits;
    uint8_t mode;
    uint8_t fill;
    uint16_t n_used;       // symbols with a non-zero count
    uint16_t n_tokens;     // tokens of the block, saturated at 65535 (selects the encoder path)
};

struct Shape {
    int kind;
    int bps, ch, ns;
    uint32_t N;             // ch * ns
    uint32_t nblk;          // hzr blocks per plane
    uint32_t nb_init;       // planes the instance starts with
    uint32_t nb_alloc;      // planes reserved per frame in scratch (max reachable)
    uint32_t hdr_bytes;     // 3*ch for hadamard/dct
    uint32_t plane_stride;  // N rounded up to 16
    uint32_t frame_bytes;   // bps * N
    uint32_t method;        // frame method byte (0 hzr/xdelta, 1 dct, 2 hadamard)
};

__host__ __device__ __forceinline__ uint32_t blk_len(const Shape& s, uint32_t b)
{
    uint32_t off = b * kBlock;
    return s.N - off < kBlock ? s.N - off : kBlock;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// Extra bits carried by the zero-run symbols 256..260 (hzr_internal.h:117-121).
__device__ __forceinline__ uint32_t sym_extra_bits(uint32_t sym)
{
    // 256:0 257:2 258:4 259:8 260:14
    return sym < 257 ? 0u : (sym == 257 ? 2u : (sym == 258 ? 4u : (sym == 259 ? 8u : 14u)));
}

// Classify a zero-run chunk z in [1, 16662] -> (symbol, extra value, extra bits); hzr_encode.c:152-166.
// Branch-free: class index 0..5 for 1 | 2 | 3-6 | 7-22 | 23-278 | 279-16662.
__device__ __forceinline__ void run_token(uint32_t z, uint32_t& sym, uint32_t& ev, uint32_t& eb)
{
    const uint32_t idx = (z >= 2u) + (z >= 3u) + (z >= 7u) + (z >= 23u) + (z >= 279u);
    sym = idx ? 255u + idx : 0u;
    eb = (0xE84200u >> (4u * idx)) & 0xFu;                      // 0 0 2 4 8 14
    const uint32_t base = idx == 5u ? 279u : (0x17070300u >> (8u * (idx < 2u ? 0u : idx - 1u))) & 0xFFu;  // 3 7 23
    ev = idx < 2u ? 0u : z - base;
}

// ---- strip loading -----------------------------------------------------------------------
// blk is 16-byte aligned (plane rows are padded to 16 and blocks start at multiples of 65536).
// Bytes at or beyond `valid` are forced to zero so that az/tz can be computed on whole words;
// the token walk itself never looks past `valid`.
__device__ __forceinline__ int load_strip(const uint8_t* __restrict__ blk, uint32_t n, uint32_t t,
                                          uint32_t (&w)[16])
{
    int valid = (int)n - (int)(t * kStrip);
    valid = valid < 0 ? 0 : (valid > kStrip ? kStrip : valid);
    const uint4* p = reinterpret_cast<const uint4*>(blk + (size_t)t * kStrip);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q * 16 < valid) v = __ldg(p + q);
        w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    if (valid < kStrip) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int nbv = valid - 4 * j;
            uint32_t m = nbv >= 4 ? 0xFFFFFFFFu : (nbv <= 0 ? 0u : (0xFFFFFFFFu >> (32 - 8 * nbv)));
            w[j] &= m;
        }
    }
    return valid;
}

// trailing zero bytes of a 64-byte strip (64 when it is all zero)
__device__ __forceinline__ uint32_t strip_trailing_zeros(const uint32_t (&w)[16])
{
    uint32_t tz = 64;
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (w[j]) tz = 4u * (15 - j) + ((uint32_t)__clz((int)w[j]) >> 3);
    return tz;
}

// Zero bytes immediately preceding this thread's strip inside the block (the pending run that
// the first non-zero byte of the strip will have to emit).  Block-wide; s_wtz/s_waz are
// shared arrays of 32 words each.  Contains one __syncthreads().
__device__ __forceinline__ uint32_t strip_carry_in(uint32_t tz, bool az, uint32_t* s_wtz, uint32_t* s_waz)
{
    const uint32_t lane = lane_id(), wid = warp_id();
    const uint32_t azm = __ballot_sync(0xFFFFFFFFu, az);
    const uint32_t lower = ~azm & ((1u << lane) - 1u);
    const int p = 31 - __clz((int)lower);  // nearest preceding lane with a non-zero byte, -1 = none
    const uint32_t tz_p = __shfl_sync(0xFFFFFFFFu, tz, p < 0 ? 0 : p);
    uint32_t carry = p >= 0 ? tz_p + (uint32_t)kStrip * (lane - 1 - p) : (uint32_t)kStrip * lane;
    // warp summary: zeros at the end of the warp's 2048-byte range
    const int q = 31 - __clz((int)~azm);
    const uint32_t tz_q = __shfl_sync(0xFFFFFFFFu, tz, q < 0 ? 0 : q);
    if (lane == 0) {
        s_wtz[wid] = q >= 0 ? tz_q + (uint32_t)kStrip * (31 - q) : 32u * kStrip;
        s_waz[wid] = q < 0;
    }
    __syncthreads();
    if (p < 0) {
        for (int v = (int)wid - 1; v >= 0; --v) {
            carry += s_wtz[v];
            if (!s_waz[v]) break;
        }
    }
    return carry;
}

// ---- token walk --------------------------------------------------------------------------
// Emits, in stream order, the tokens OWNED by this strip: every literal in the strip, each
// preceded by the zero run that ends right before it (which may have started in earlier
// strips: `carry`), plus -- for the last strip of the block -- the run that reaches the block
// end.  Concatenated over strips this is exactly the reference token sequence
// (hzr_encode.c:410-457: greedy chunks of <= 16662 zeros).
template <class Sink>
__device__ __forceinline__ void emit_run(uint32_t z, Sink& sink)
{
    while (z > kRunCap) {
        sink.token(260u, kRunCap - 279u, 14u);
        z -= kRunCap;
    }
    uint32_t sym, ev, eb;
    run_token(z, sym, ev, eb);
    sink.token(sym, ev, eb);
}

template <class Sink>
__device__ __forceinline__ void walk_strip(const uint32_t (&w)[16], int valid, uint32_t carry,
                                           bool last_strip, Sink& sink)
{
    uint32_t zrun = carry;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int nbv = valid - 4 * j;
        if (nbv > 0) {
            const uint32_t x = w[j];
            if (x == 0) {
                zrun += nbv >= 4 ? 4u : (uint32_t)nbv;
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (b < nbv) {
                        const uint32_t v = (x >> (8 * b)) & 0xFFu;
                        if (v) {
                            if (zrun) {
                                emit_run(zrun, sink);
                                zrun = 0;
                            }
                            sink.token(v, 0u, 0u);
                        } else {
                            ++zrun;
                        }
                    }
                }
            }
        }
    }
    if (last_strip && zrun) emit_run(zrun, sink);
}

// ---- shared-memory staging of a whole block ---------------------------------------------------
// The block's bytes live in shared memory as 16-word strips; word j of strip t is stored at
// t*16 + (j ^ ((t >> 1) & 15)) so that the 32 lanes of a warp, each reading word j of its own
// strip, hit 32 different banks.
__device__ __forceinline__ uint32_t strip_word_index(uint32_t t, uint32_t j) { return (t << 4) | (j ^ ((t >> 1) & 15u)); }

// Walk of one staged strip.  `valid` bytes are meaningful; the rest of the strip is zero.
// Fast path: a zero byte whose neighbours are non-zero is simply the literal token 0
// (hzr_encode.c:152-153, run of one), so a word in which every byte is either non-zero or such
// an isolated zero -- and no run is pending -- emits four tokens with no run bookkeeping.  In
// dense planes that is ~99 % of the words, which keeps the warp converged; genuine runs
// (two or more zeros, or a zero in the strip's last byte, whose run may continue in the next
// strip) take the general path.
template <class Sink>
__device__ __forceinline__ void walk_strip_staged(const uint32_t* in_sw, uint32_t t, int valid, uint32_t carry,
                                                  bool last_strip, Sink& sink)
{
    uint32_t zrun = carry;
    const uint32_t base = t << 4, sw = (t >> 1) & 15u;
    const int nwords = (valid + 3) >> 2;
    uint32_t x = in_sw[base | sw];  // word 0
#pragma unroll 1
    for (int j = 0; j < nwords; ++j) {
        const uint32_t xn = j + 1 < nwords ? in_sw[base | ((uint32_t)(j + 1) ^ sw)] : 0u;
        const int nbv = valid - 4 * j;
        if (x == 0) {
            zrun += nbv >= 4 ? 4u : (uint32_t)nbv;
        } else {
            const uint32_t nzf = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;  // bit 7 of every non-zero byte
            const uint32_t nxt = (nzf >> 8) | ((xn & 0xFFu) ? 0x80000000u : 0u);        // ... of every byte's successor
            if (zrun == 0 && nbv >= 4 && (nzf | nxt) == 0x80808080u) {
                sink.token(x & 0xFFu, 0u, 0u);
                sink.token((x >> 8) & 0xFFu, 0u, 0u);
                sink.token((x >> 16) & 0xFFu, 0u, 0u);
                sink.token(x >> 24, 0u, 0u);
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (b < nbv) {
                        const uint32_t v = (x >> (8 * b)) & 0xFFu;
                        if (v) {
                            if (zrun) {
                                emit_run(zrun, sink);
                                zrun = 0;
                            }
                            sink.token(v, 0u, 0u);
                        } else {
                            ++zrun;
                        }
                    }
                }
            }
        }
        x = xn;
    }
    if (last_strip && zrun) emit_run(zrun, sink);
}

// ---- block-wide exclusive scan of one uint32 per thread (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /*[33]*/, uint32_t* total)
{
    const uint32_t lane = lane_id(), wid = warp_id(), nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += y;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t x = lane < nw ? s_warp[lane] : 0u;
        uint32_t xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, xi, o);
            if (lane >= (uint32_t)o) xi += y;
        }
        s_warp[lane] = xi - x;  // exclusive warp offsets
        if (lane == 31) s_warp[32] = xi;
    }
    __syncthreads();
    if (total) *total = s_warp[32];
    return s_warp[wid] + inc - v;
}

// Load block [src, src+n) into swizzled shared memory (thread t <-> strip t; blockDim.x must be
// >= the number of strips), compute every strip's pending zero run and the compact list of
// strips that own at least one token.  Returns this thread's position-independent facts through
// s_carry / s_list and the number of active strips.  Ends with a __syncthreads().
__device__ __forceinline__ uint32_t stage_block(const uint8_t* __restrict__ src, uint32_t n, uint32_t* in_sw,
                                                uint16_t* s_carry, uint16_t* s_list, uint32_t* s_wtz, uint32_t* s_waz,
                                                uint32_t* s_scan)
{
    const uint32_t t = threadIdx.x;
    const uint32_t nstrips = (n + kStrip - 1) / kStrip;
    uint32_t w[16];
    const int valid = load_strip(src, n, t, w);
    const uint32_t tz = strip_trailing_zeros(w);
    if (t < nstrips) {
        const uint32_t sw = (t >> 1) & 15u;
#pragma unroll
        for (int j = 0; j < 16; ++j) in_sw[(t << 4) | ((uint32_t)j ^ sw)] = w[j];
    }
    const uint32_t carry = strip_carry_in(tz, tz == 64, s_wtz, s_waz);
    const bool last = t + 1 == nstrips;
    const bool active = valid > 0 && (tz != 64 || last);
    uint32_t n_active;
    const uint32_t pos = block_exclusive_scan(active ? 1u : 0u, s_scan, &n_active);
    if (active) s_list[pos] = (uint16_t)t;
    if (t < nstrips) s_carry[t] = (uint16_t)carry;
    __syncthreads();
    return n_active;
}

#define RSPT_CUDA_CHECK(call)                                      \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return rspt::fail_cuda(p, e_, #call); \
    } while (0)

}  // namespace rspt
