// Fused front end of the xdelta_hzr / hzr packers for sm_100a: ONE pass over the raw frames that
// produces everything the tree builder and the two encoders need.
//
//   raw frame tile --bulk async copy (TMA unit) + mbarrier, double buffered--> shared memory
//   phase A  de-interleave + sign-extend (convert_native_to_i32, utils.cpp:123-191), the
//            delta / offset / xor stencil (utils.cpp:193-230, flat chain crossing channel rows),
//            byte-plane split (signal_packer_base.cpp:40-68): the planes of the tile land in a
//            shared-memory tile in flat (channel-major) order, one 512-byte STEP per plane and channel
//   phase B  a warp per step.  Dense planes: hzr token histogram (Histogram, hzr_encode.c:133-173:
//            literals + zero-run classes) and the step's leading / trailing zero counts.  Planes
//            that are mostly zero (decided per frame and plane from the first tile): only their
//            non-zero bytes, appended in order to the (plane, channel) sub-list `position | value << 16`
//            that lives where the plane bytes of that channel row would have been.
//   dense planes go to HBM with bulk shared -> global copies (one 512-byte row per channel and tile);
//   sparse planes never do.
//   frame epilogue: zero runs that cross step boundaries are chained in flat order; histograms,
//   step_lz, sub-list lengths and the list decision per block go out.
// The tree kernel concatenates the sub-lists of a block's channels into the sorted list
// k_hzr_encode_sparse consumes and takes the block's token histogram from it.
//
// A persistent CTA (128 threads) walks whole frames, tile by tile, so the histograms stay in shared
// memory and need no global atomics.  Replaces k_xdelta_planes_fast + k_hzr_hist<1> + k_hzr_hist<2>
// for the eligible shapes (ch in {4, 8, 12}, ns % 512 == 0, hzr blocks that start on channel rows, no
// plane escalation); every other shape keeps those kernels.  A frame whose sparse-looking plane
// overflows its sub-list (or whose listed block the tree kernel finds unfit for the list encoder)
// is flagged and runs again through this kernel with every plane forced dense (`only` + force_dense).
#pragma once

#include <type_traits>
#include <utility>

#include "bulk.cuh"
#include "common.cuh"
#include "hzr_hist.cuh"
#include "transforms.cuh"

namespace rspt {

constexpr int kFrontThreads = 128;
constexpr int kFrontQuads = 128;                 // sample quads per tile: 512 samples = one 512-byte step per channel and plane
constexpr uint32_t kFrontMaxHist = 16;           // (plane, block) histograms kept in shared memory
constexpr uint32_t kFrontMaxTiles = 64;

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// sign-extended sample whose BPS bytes start at the compile-time byte offset OFF of the word array w
template <int BPS, int OFF>
__device__ __forceinline__ uint32_t unpack_at(const uint32_t* w)
{
    constexpr int k = OFF / 4, s = OFF % 4;
    if constexpr (BPS == 4) {
        return w[k];
    } else if constexpr (BPS == 3) {
        if constexpr (s == 0) return prmt(w[k], w[k], 0xA210u);
        else if constexpr (s == 1) return prmt(w[k], w[k], 0xB321u);
        else if constexpr (s == 2) return prmt(w[k], w[k + 1], 0xC432u);
        else return prmt(w[k], w[k + 1], 0xD543u);
    } else {
        static_assert(BPS == 2, "k_front: 2, 3 or 4 bytes per sample");
        if constexpr (s == 0) return prmt(w[k], w[k], 0x9910u);
        else return prmt(w[k], w[k], 0xBB32u);
    }
}

// ------------------------------------------------------------------------------------------
// Forward sample transform, TMA-fed (default for the eligible shapes; k_xdelta_planes_fast for the others).
// One CTA (128 threads) per tile of 128 sample quads of one frame.  The raw tile (+ the quad in front of it, whose
// last two sample rows are the stencil's predecessors) arrives with ONE bulk asynchronous copy global -> shared,
// completion on an mbarrier; it lies UNPADDED in shared memory and thread t reads the whole quad t + 1 with 128-bit
// loads (lanes are 4 * ROW bytes apart: 144 bytes for 12 ch x 3 B, conflict-free for 128-bit accesses), so that all
// CH channels of its four samples are in registers at once: one PRMT per sample to unpack, the stencil, the 4 x 4
// byte transpose (8 PRMT), one 32-bit store per plane and channel (lanes run along the samples: coalesced).
// About 9 instructions per sample against 25 in k_xdelta_planes_fast (whose staging loop, padded layout and 32-bit
// shared-memory loads are gone).  Same bytes as k_xdelta_planes (tests: STREAM_CASES, test_transform_planes_bit_exact).
// ------------------------------------------------------------------------------------------
template <int BPS, int CH, bool STENCIL>
__global__ void __launch_bounds__(kFrontThreads, 4) k_xdelta_planes_tma(const uint8_t* __restrict__ src, Shape s, uint32_t tiles_per_frame,
                                                                        uint8_t* __restrict__ planes)
{
    constexpr int ROW = CH * BPS;          // bytes per sample row == words per quad
    constexpr int QB = 4 * ROW;            // bytes per quad
    constexpr int QW = ROW;                // words per quad
    constexpr int HSTART = (QW / 2) & ~3;  // first word of the 16-byte chunks that hold rows 2, 3 of a quad
    constexpr int HW = QW - HSTART;
    extern __shared__ __align__(128) uint8_t smem[];   // quads t*128 - 1 .. t*128 + 127, then the frame's last quad
    __shared__ __align__(8) uint64_t s_full;
    const uint32_t f = blockIdx.x / tiles_per_frame, t = blockIdx.x % tiles_per_frame, tid = threadIdx.x;
    const uint32_t nb = s.nb_alloc, ns = (uint32_t)s.ns, nq = ns >> 2;
    const uint8_t* frame = src + (size_t)f * s.frame_bytes;
    uint8_t* raw = smem;
    uint8_t* tail = smem + (size_t)(kFrontQuads + 1) * QB;
    if (tid == 0) {
        mbar_init(&s_full, 1);
        mbar_fence_init();
        if (t == 0) {
            mbar_arrive_expect_tx(&s_full, (uint32_t)(kFrontQuads * QB + (STENCIL ? QB : 0)));
            bulk_g2s(raw + QB, frame, (uint32_t)(kFrontQuads * QB), &s_full);
            if (STENCIL) bulk_g2s(tail, frame + (size_t)(nq - 1) * QB, (uint32_t)QB, &s_full);
        } else {
            mbar_arrive_expect_tx(&s_full, (uint32_t)((kFrontQuads + 1) * QB));
            bulk_g2s(raw, frame + ((size_t)t * kFrontQuads - 1) * QB, (uint32_t)((kFrontQuads + 1) * QB), &s_full);
        }
    }
    __syncthreads();   // the barrier is initialised before anyone waits on it
    mbar_wait(&s_full, 0);
    if (STENCIL && t == 0) {
        // the flat chain crosses channel rows (signal_packer_xdelta_hzr.cpp:55-57): the two predecessors of channel
        // c's first sample are the last two samples of channel c - 1 (zero for channel 0): rows 2, 3 of the halo quad
        for (uint32_t i = tid; i < 2u * ROW; i += kFrontThreads) {
            const uint32_t r = 2u + i / ROW, bb = i % ROW;
            raw[r * ROW + bb] = bb < (uint32_t)BPS ? (uint8_t)0 : tail[r * ROW + bb - BPS];
        }
        __syncthreads();
    }
    uint32_t w[QW];
    const uint4* q4 = reinterpret_cast<const uint4*>(raw + (size_t)(tid + 1) * QB);
#pragma unroll
    for (int i = 0; i < QW / 4; ++i) {
        const uint4 v = q4[i];
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    uint32_t hw[STENCIL ? HW : 4];
    if (STENCIL) {
        const uint4* h4 = reinterpret_cast<const uint4*>(raw + (size_t)tid * QB + HSTART * 4);
#pragma unroll
        for (int i = 0; i < HW / 4; ++i) {
            const uint4 v = h4[i];
            hw[4 * i] = v.x; hw[4 * i + 1] = v.y; hw[4 * i + 2] = v.z; hw[4 * i + 3] = v.w;
        }
    }
    const bool first = t == 0 && tid == 0;
    const uint32_t ps = s.plane_stride >> 2;
    uint32_t* out0 = reinterpret_cast<uint32_t*>(planes + (size_t)f * nb * s.plane_stride) + (t * (uint32_t)kFrontQuads + tid);
    static_for<CH>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        uint32_t y[4];
        const uint32_t x0 = unpack_at<BPS, 0 * ROW + c * BPS>(w), x1 = unpack_at<BPS, 1 * ROW + c * BPS>(w);
        const uint32_t x2 = unpack_at<BPS, 2 * ROW + c * BPS>(w), x3 = unpack_at<BPS, 3 * ROW + c * BPS>(w);
        if constexpr (STENCIL) {
            constexpr int HB = (QW / 2 - HSTART) * 4;  // byte offset of row 2 inside hw
            const uint32_t xm2 = unpack_at<BPS, HB + c * BPS>(hw), xm1 = unpack_at<BPS, HB + ROW + c * BPS>(hw);
            uint32_t d0 = xm1 - xm2 - 128u;
            // the very first word of the frame has no predecessor delta: y[0] = x[0] - 128
            if (c == 0 && first) d0 = 0;
            const uint32_t d1 = x0 - xm1 - 128u, d2 = x1 - x0 - 128u, d3 = x2 - x1 - 128u, d4 = x3 - x2 - 128u;
            y[0] = d1 ^ d0; y[1] = d2 ^ d1; y[2] = d3 ^ d2; y[3] = d4 ^ d3;
        } else {
            y[0] = x0; y[1] = x1; y[2] = x2; y[3] = x3;
        }
        const uint32_t t01 = prmt(y[0], y[1], 0x5140u), t23 = prmt(y[2], y[3], 0x5140u);
        uint32_t* out = out0 + (size_t)c * (ns >> 2);
        out[0] = prmt(t01, t23, 0x5410u);
        if (nb > 1) out[ps] = prmt(t01, t23, 0x7632u);
        if (nb > 2) {
            const uint32_t u01 = prmt(y[0], y[1], 0x7362u), u23 = prmt(y[2], y[3], 0x7362u);
            out[2 * ps] = prmt(u01, u23, 0x5410u);
            if (nb > 3) out[3 * ps] = prmt(u01, u23, 0x7632u);
        }
    });
}

struct FrontOut {
    uint8_t* planes;       // [F][nb_alloc][plane_stride]; sparse planes: sub-list storage, one channel row each
    uint32_t* hist;        // [blocks][kSymStride] (dense blocks only; the tree kernel derives the others)
    uint16_t* step_lz;     // [blocks][kMaxSteps]
    uint32_t* sub_n;       // [F][nb_alloc][ch]: entries in the (plane, channel) sub-list
    uint32_t* list_n;      // [blocks]: list length the tree kernel will find, kNoList = dense block
    uint8_t* redo;         // [F]: first pass: 1 = a sparse-mode plane overflowed, run the frame again forced dense
    uint8_t* redo2;        // [F]: cleared here, set by the tree kernel
};

// dynamic shared memory of k_front for a shape (bytes)
__host__ __device__ inline size_t front_smem_bytes(int bps, int ch, uint32_t nb, uint32_t nblk, uint32_t tiles)
{
    const size_t qb = 4u * (size_t)ch * bps;
    size_t b = (size_t)(kFrontQuads + 1) * qb;           // raw tile (+ halo quad)
    b += qb;                                         // last quad of the frame (predecessors of the first samples)
    b += (size_t)nb * ch * kStepBytes;                   // plane tile
    b += (size_t)nb * nblk * (256 + 8) * 4;              // histograms
    b += (size_t)nb * ch * 4;                            // sub-list lengths
    b += 2 * (size_t)nb * ch * 4 * 4;                    // per tile parity, (plane, channel), warp: non-zero bytes of the warp's 128 samples
    b += 2 * (size_t)nb * ch * tiles * 2;                // per step: lz, tz (u16)
    return b + 64;
}

// the r-th (0-based) set bit of a 16-bit mask that has more than r bits set
__device__ __forceinline__ uint32_t select_bit16(uint32_t m, uint32_t r)
{
    uint32_t pos = 0;
    uint32_t c = __popc(m & 0xFFu);
    if (r >= c) { pos = 8; r -= c; m >>= 8; }
    c = __popc(m & 0xFu);
    if (r >= c) { pos += 4; r -= c; m >>= 4; }
    c = __popc(m & 3u);
    if (r >= c) { pos += 2; r -= c; m >>= 2; }
    if (r >= (m & 1u)) pos += 1;
    return pos;
}

template <int BPS, int CH, bool STENCIL>
__global__ void __launch_bounds__(kFrontThreads, 4) k_front(const uint8_t* __restrict__ src, Shape s, uint32_t n_frames, FrontOut o,
                                                            const uint8_t* __restrict__ only, int force_dense)
{
    constexpr int ROW = CH * BPS;          // bytes per sample row == words per quad
    constexpr int QB = 4 * ROW;            // bytes per quad
    constexpr int QW = ROW;                // words per quad
    constexpr int HSTART = (QW / 2) & ~3;  // first word of the 16-byte chunks that hold rows 2, 3 of a quad
    constexpr int HW = QW - HSTART;
    constexpr uint32_t SW = kStepBytes / 4;  // words per step
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[1];
    __shared__ uint32_t s_probe[4];
    __shared__ uint32_t s_over;

    const uint32_t nb = s.nb_alloc, nblk = s.nblk, ns = (uint32_t)s.ns;
    const uint32_t T = ns / kStepBytes, NH = nb * nblk, NSEG = nb * CH * T;
    uint8_t* raw0 = smem;
    constexpr uint32_t RAWB = (kFrontQuads + 1) * QB;
    uint8_t* tail0 = smem + RAWB;
    uint32_t* planeT = reinterpret_cast<uint32_t*>(tail0 + QB);
    uint32_t* s_lit = planeT + nb * CH * SW;
    uint32_t* s_run = s_lit + NH * 256;
    uint32_t* s_subn = s_run + NH * 8;
    uint32_t* s_wcnt = s_subn + nb * CH;   // [2][nb * CH][4]
    uint16_t* s_lz = reinterpret_cast<uint16_t*>(s_wcnt + 2 * nb * CH * 4);
    uint16_t* s_tz = s_lz + NSEG;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t nq = ns >> 2;
    const uint32_t sub_cap = ns / 4u;      // entries that fit in a channel row of a plane

    // frames of this CTA: blockIdx.x, + gridDim.x, ... restricted to the flagged ones when `only` is given
    auto next_frame = [&](uint32_t f) {
        while (f < n_frames && only && !only[f]) f += gridDim.x;
        return f;
    };
    // ---- producer side (thread 0): one bulk copy per tile (+ the frame's last quad with tile 0)
    uint32_t pf_f = 0, pf_t = 0, pf_it = 0;   // next item to fetch
    auto fetch = [&]() {
        if (pf_f >= n_frames) return;
        const uint32_t buf = 0;
        const uint8_t* frame = src + (size_t)pf_f * s.frame_bytes;
        uint8_t* dst = raw0 + buf * RAWB;
        if (pf_t == 0) {
            mbar_arrive_expect_tx(&s_full[buf], (uint32_t)(kFrontQuads * QB + QB));
            bulk_g2s(dst + QB, frame, (uint32_t)(kFrontQuads * QB), &s_full[buf]);
            bulk_g2s(tail0 + buf * QB, frame + (size_t)(nq - 1) * QB, (uint32_t)QB, &s_full[buf]);
        } else {
            mbar_arrive_expect_tx(&s_full[buf], (uint32_t)RAWB);
            bulk_g2s(dst, frame + (size_t)(pf_t * kFrontQuads - 1) * QB, (uint32_t)RAWB, &s_full[buf]);
        }
        ++pf_it;
        if (++pf_t == T) {
            pf_t = 0;
            pf_f = next_frame(pf_f + gridDim.x);
        }
    };
    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_fence_init();
    }
    for (uint32_t i = tid; i < NH * 264 + nb * CH * 9; i += kFrontThreads) s_lit[i] = 0;  // s_lit, s_run, s_subn, s_wcnt are contiguous
    if (tid < 4) s_probe[tid] = 0;
    if (tid == 0) s_over = 0;
    __syncthreads();
    if (tid == 0) {
        pf_f = next_frame(blockIdx.x);
        fetch();
    }

    uint32_t it = 0;
    for (uint32_t f = next_frame(blockIdx.x); f < n_frames; f = next_frame(f + gridDim.x)) {
        uint8_t* fplanes = o.planes + (size_t)f * nb * s.plane_stride;
        uint32_t sparse_mask = 0;
        for (uint32_t t = 0; t < T; ++t, ++it) {
            const uint32_t buf = 0;
            uint8_t* raw = raw0;
            mbar_wait(&s_full[0], it & 1u);
            if (STENCIL && t == 0) {
                // the flat chain crosses channel rows (signal_packer_xdelta_hzr.cpp:55-57): the two
                // predecessors of channel c's first sample are the last two samples of channel c - 1
                // (zero for channel 0).  Build that as rows 2, 3 of the (unused) halo quad.
                const uint8_t* tl = tail0 + buf * QB;
                for (uint32_t i = tid; i < 2u * ROW; i += kFrontThreads) {
                    const uint32_t r = 2u + i / ROW, bb = i % ROW;
                    raw[r * ROW + bb] = bb < (uint32_t)BPS ? (uint8_t)0 : tl[r * ROW + bb - BPS];
                }
                __syncthreads();
            }
            // ---- phase A: my quad of every channel
            uint32_t busy = 0;   // channels in which my warp has list entries in this tile
            {
                uint32_t w[QW];
                const uint4* q4 = reinterpret_cast<const uint4*>(raw + (size_t)(tid + 1) * QB);
#pragma unroll
                for (int i = 0; i < QW / 4; ++i) {
                    const uint4 v = q4[i];
                    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                }
                uint32_t hw[STENCIL ? HW : 4];
                if (STENCIL) {
                    const uint4* h4 = reinterpret_cast<const uint4*>(raw + (size_t)tid * QB + HSTART * 4);
#pragma unroll
                    for (int i = 0; i < HW / 4; ++i) {
                        const uint4 v = h4[i];
                        hw[4 * i] = v.x; hw[4 * i + 1] = v.y; hw[4 * i + 2] = v.z; hw[4 * i + 3] = v.w;
                    }
                }
                const bool first = t == 0 && tid == 0;
                uint32_t pcnt[4] = {0, 0, 0, 0};
                // planes that may be kept as lists (all upper planes while the frame's first tile is still being probed)
                const uint32_t cand = t == 0 ? (force_dense ? 0u : ((1u << nb) - 2u)) : sparse_mask;
                const uint32_t pm1 = (cand & 2u) ? 0xFFFFFFFFu : 0u, pm2 = (cand & 4u) ? 0xFFFFFFFFu : 0u, pm3 = (cand & 8u) ? 0xFFFFFFFFu : 0u;
                static_for<CH>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    uint32_t y[4];
                    const uint32_t x0 = unpack_at<BPS, 0 * ROW + c * BPS>(w), x1 = unpack_at<BPS, 1 * ROW + c * BPS>(w);
                    const uint32_t x2 = unpack_at<BPS, 2 * ROW + c * BPS>(w), x3 = unpack_at<BPS, 3 * ROW + c * BPS>(w);
                    if constexpr (STENCIL) {
                        constexpr int HB = (QW / 2 - HSTART) * 4;  // byte offset of row 2 inside hw
                        const uint32_t xm2 = unpack_at<BPS, HB + c * BPS>(hw), xm1 = unpack_at<BPS, HB + ROW + c * BPS>(hw);
                        uint32_t d0 = xm1 - xm2 - 128u;
                        // the very first word of the frame has no predecessor delta: y[0] = x[0] - 128
                        if (c == 0 && first) d0 = 0;
                        const uint32_t d1 = x0 - xm1 - 128u, d2 = x1 - x0 - 128u, d3 = x2 - x1 - 128u, d4 = x3 - x2 - 128u;
                        y[0] = d1 ^ d0; y[1] = d2 ^ d1; y[2] = d3 ^ d2; y[3] = d4 ^ d3;
                    } else {
                        y[0] = x0; y[1] = x1; y[2] = x2; y[3] = x3;
                    }
                    const uint32_t t01 = prmt(y[0], y[1], 0x5140u), t23 = prmt(y[2], y[3], 0x5140u);
                    uint32_t p[4];
                    p[0] = prmt(t01, t23, 0x5410u);
                    p[1] = prmt(t01, t23, 0x7632u);
                    if (nb > 2) {
                        const uint32_t u01 = prmt(y[0], y[1], 0x7362u), u23 = prmt(y[2], y[3], 0x7362u);
                        p[2] = prmt(u01, u23, 0x5410u);
                        p[3] = prmt(u01, u23, 0x7632u);
                    }
                    uint32_t* dstw = planeT + c * SW + tid;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((uint32_t)k < nb) {
                            dstw[k * CH * SW] = p[k];
                            if (t == 0) pcnt[k] += __popc(nz_nibble(p[k]));
                        }
                    uint32_t acc = p[1] & pm1;
                    if (nb > 2) acc |= (p[2] & pm2) | (p[3] & pm3);
                    if (__ballot_sync(0xFFFFFFFFu, acc != 0u)) busy |= 1u << c;
                });
                // for every channel in which my warp's 128 samples have a non-zero byte in a list plane, publish per
                // plane how many: the warps behind it in the tile place their entries after these
                {
                    uint32_t* wc = s_wcnt + (t & 1u) * nb * CH * 4;
                    for (uint32_t bm = busy; bm; bm &= bm - 1u) {
                        const uint32_t c = (uint32_t)__ffs(bm) - 1u;
                        for (uint32_t k = 1; k < nb; ++k) {
                            if (!((cand >> k) & 1u)) continue;
                            const uint32_t u = k * CH + c;
                            const uint32_t n1 = __reduce_add_sync(0xFFFFFFFFu, __popc(nz_nibble(planeT[u * SW + tid])));
                            if (lane == 0) wc[u * 4 + wid] = n1;
                        }
                    }
                }
                if (t == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((uint32_t)k < nb) {
                            const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, pcnt[k]);
                            if (lane == 0) atomicAdd(&s_probe[k], v);
                        }
                }
            }
            fence_async_smem();   // my plane words, before the bulk stores read them
            __syncthreads();      // plane tile complete; raw[buf] is free
            if (t == 0) {
                // a plane is kept as lists when <= 1/8 of its bytes in the frame's first tile are non-zero
                sparse_mask = 0;
                if (!force_dense)
                    for (uint32_t k = 1; k < nb; ++k)
                        if (s_probe[k] * 8u <= (uint32_t)(kStepBytes * CH)) sparse_mask |= 1u << k;
            }
            if (tid == 0) {
                fetch();
                for (uint32_t k = 0; k < nb; ++k) {
                    if ((sparse_mask >> k) & 1u) continue;
                    uint8_t* prow = fplanes + (size_t)k * s.plane_stride + (size_t)t * kStepBytes;
                    for (uint32_t c = 0; c < (uint32_t)CH; ++c) bulk_s2g(prow + (size_t)c * ns, planeT + (k * CH + c) * SW, kStepBytes);
                }
                bulk_commit();
            }
            // ---- list planes: a busy warp appends the non-zero bytes of its 128 samples of every channel, in
            // order, to the (plane, channel) sub-list -- behind the entries of earlier tiles and earlier warps
            {
                const uint32_t* wc = s_wcnt + (t & 1u) * nb * CH * 4;
                const uint32_t lt = (1u << lane) - 1u;
                for (uint32_t bm = busy; bm; bm &= bm - 1u) {
                    const uint32_t c = (uint32_t)__ffs(bm) - 1u;
                    for (uint32_t k = 1; k < nb; ++k) {
                        if (!((sparse_mask >> k) & 1u)) continue;
                        const uint32_t u = k * CH + c;
                        const uint32_t word = planeT[u * SW + tid];
                        const uint32_t nzn = nz_nibble(word);
                        if (__ballot_sync(0xFFFFFFFFu, nzn != 0u) == 0u) continue;
                        const uint32_t c0 = wc[u * 4], c1 = wc[u * 4 + 1], c2 = wc[u * 4 + 2], c3 = wc[u * 4 + 3];
                        const uint32_t have = s_subn[u];
                        if (have + c0 + c1 + c2 + c3 > sub_cap) {
                            if (lane == 0) atomicOr(&s_over, 1u << k);
                            continue;
                        }
                        const uint32_t cnt = __popc(nzn);
                        const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, cnt & 1u), b1 = __ballot_sync(0xFFFFFFFFu, cnt & 2u),
                                       b2 = __ballot_sync(0xFFFFFFFFu, cnt & 4u);
                        uint32_t at = have + (wid > 0 ? c0 : 0u) + (wid > 1 ? c1 : 0u) + (wid > 2 ? c2 : 0u) + __popc(b0 & lt) +
                                      2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
                        uint32_t* sub = reinterpret_cast<uint32_t*>(fplanes + (size_t)k * s.plane_stride + (size_t)c * ns);
                        const uint32_t pos = ((c * ns + t * kStepBytes) & (kBlock - 1u)) + 4u * tid;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if ((nzn >> j) & 1u) sub[at++] = (pos + j) | (((word >> (8 * j)) & 0xFFu) << 16);
                    }
                }
            }
            // ---- phase B: one 512-byte step (dense plane k, channel c) per warp pass
            for (uint32_t u = wid; u < nb * CH; u += kFrontThreads / 32) {
                const uint32_t k = u / CH, c = u - k * CH;
                if ((sparse_mask >> k) & 1u) continue;   // list planes were handled above
                const uint4 v = reinterpret_cast<const uint4*>(planeT + u * SW)[lane];
                const bool any = (v.x | v.y | v.z | v.w) != 0u;
                const uint32_t anyt = __ballot_sync(0xFFFFFFFFu, any);
                const uint32_t flat = c * ns + t * kStepBytes;   // flat position of the step in the plane
                // dense plane
                const uint32_t si = u * T + t;
                if (anyt == 0u) {
                    if (lane == 0) {
                        s_lz[si] = (uint16_t)kStepAllZero;
                        s_tz[si] = 0;
                    }
                    continue;
                }
                const uint32_t hb = k * nblk + (flat >> 16);
                uint32_t* lit = s_lit + hb * 256;
                uint32_t* run = s_run + hb * 8;
                {
                    // literals: one shared-memory atomic per byte; zero bytes land in the scratch bin 0
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        atomicAdd(&lit[w4[j] & 0xFFu], 1u);
                        atomicAdd(&lit[(w4[j] >> 8) & 0xFFu], 1u);
                        atomicAdd(&lit[(w4[j] >> 16) & 0xFFu], 1u);
                        atomicAdd(&lit[w4[j] >> 24], 1u);
                    }
                }
                // zero runs closed inside the step (same bit-mask classification as k_hzr_hist<2>)
                const uint32_t nz = any ? nz_mask16(v) : 0u;
                const uint32_t z = ~nz & 0xFFFFu;
                uint32_t prevz, nextz;
                neighbour_zero(z, prevz, nextz);
                const uint32_t starts = z & ~((z << 1) | prevz);
                const uint32_t cont = (z >> 1) | (nextz << 15);
                const uint32_t rs2 = starts & cont;
                const uint32_t a2 = z & (z >> 1), a3 = a2 & (z >> 2), a4 = a2 & (a2 >> 2), a7 = a4 & (a3 >> 4);
                const uint32_t top = nz ? 32u - (uint32_t)__clz((int)nz) : 0u;
                const uint32_t open_bit = (z >> 15) << top;
                const uint32_t closed = rs2 & ~open_bit;
                uint32_t cnt_a = __popc(starts & ~cont) | (__popc(closed & ~a3) << 16);
                uint32_t cnt_b = __popc(closed & a3 & ~a7) | (__popc(closed & a7) << 16);
                const uint32_t first_stop = nz ? (uint32_t)__ffs(nz) - 1u : 16u;
                const bool open_run = (rs2 & open_bit) != 0u;
                if (__any_sync(0xFFFFFFFFu, open_run)) {
                    const uint32_t above = lane < 31 ? anyt & ~((2u << lane) - 1u) : 0u;
                    const uint32_t q = above ? (uint32_t)__ffs(above) - 1u : 0u;
                    const uint32_t fq = __shfl_sync(0xFFFFFFFFu, first_stop, q);
                    if (open_run && above) hist_run(16u - top + 16u * (q - lane - 1u) + fq, run);
                }
                cnt_a = __reduce_add_sync(0xFFFFFFFFu, cnt_a);
                cnt_b = __reduce_add_sync(0xFFFFFFFFu, cnt_b);
                const uint32_t qf = (uint32_t)__ffs(anyt) - 1u, ql = 31u - (uint32_t)__clz((int)anyt);
                const uint32_t lz = 16u * qf + __shfl_sync(0xFFFFFFFFu, first_stop, qf);
                const uint32_t last_stop = nz ? 31u - (uint32_t)__clz((int)nz) : 0u;
                const uint32_t tz = 16u * (31u - ql) + 15u - __shfl_sync(0xFFFFFFFFu, last_stop, ql);
                if (lane == 0) {
                    if (cnt_a & 0xFFFFu) atomicAdd(&run[0], cnt_a & 0xFFFFu);
                    if (cnt_a >> 16) atomicAdd(&run[1], cnt_a >> 16);
                    if (cnt_b & 0xFFFFu) atomicAdd(&run[2], cnt_b & 0xFFFFu);
                    if (cnt_b >> 16) atomicAdd(&run[3], cnt_b >> 16);
                    s_lz[si] = (uint16_t)lz;
                    s_tz[si] = (uint16_t)tz;
                }
            }
            if (tid == 0) bulk_wait_read<0>();   // the plane tile is about to be overwritten
            __syncthreads();
            if (tid >= CH && tid < nb * CH) {
                uint32_t* wc = s_wcnt + (t & 1u) * nb * CH * 4 + tid * 4;
                if ((sparse_mask >> (tid / CH)) & 1u) s_subn[tid] = min(s_subn[tid] + wc[0] + wc[1] + wc[2] + wc[3], sub_cap);
                wc[0] = 0; wc[1] = 0; wc[2] = 0; wc[3] = 0;
            }
        }

        // ---- frame epilogue
        __syncthreads();   // the last tile's counts are in s_subn
        // (1) dense blocks: chain the zero runs across step boundaries in flat order; one warp per (plane, block).
        //     sparse planes: the list decision per block.
        for (uint32_t hb = wid; hb < NH; hb += kFrontThreads / 32) {
            const uint32_t k = hb / nblk, b = hb - k * nblk;
            const uint32_t n = blk_len(s, b), nsteps = n / kStepBytes;
            const size_t blk = (size_t)f * NH + hb;
            if ((sparse_mask >> k) & 1u) {
                // channels of the block: its flat range starts and ends on channel rows
                const uint32_t c0 = (b * kBlock) / ns, c1 = (b * kBlock + n) / ns;
                uint32_t m = 0;
                for (uint32_t c = c0 + lane; c < c1; c += 32) m += s_subn[k * CH + c];
                m = __reduce_add_sync(0xFFFFFFFFu, m);
                const bool listed = !((s_over >> k) & 1u) && m <= kListCap && m <= n / 4u;
                if (lane == 0) {
                    if (!listed) atomicOr(&s_over, 0x100u << k);  // this block needs its plane bytes after all
                    o.list_n[blk] = listed ? m : kNoList;
                }
                continue;
            }
            uint32_t pending = 0;
            uint32_t* run = s_run + hb * 8;
            for (uint32_t j0 = 0; j0 < nsteps; j0 += 32) {
                const uint32_t j = j0 + lane;
                const bool live = j < nsteps;
                const uint32_t flat = b * kBlock + j * kStepBytes;
                const uint32_t c = flat / ns, tt = (flat - c * ns) / kStepBytes;
                const uint32_t si = (k * CH + c) * T + tt;
                const uint32_t lz = live ? s_lz[si] : 0u, tz = live ? s_tz[si] : 0u;
                const bool stop = live && lz != kStepAllZero;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, stop);
                // the run that ends in a step = its leading zeros + the all-zero steps before it + the trailing
                // zeros of the last step that had a stop (or what was pending when this group began)
                const uint32_t below = m & ((1u << lane) - 1u);
                const uint32_t pl = below ? 31u - (uint32_t)__clz((int)below) : 0u;
                const uint32_t tzp = __shfl_sync(0xFFFFFFFFu, tz, pl);
                if (stop) {
                    const uint32_t r = below ? lz + kStepBytes * (lane - pl - 1u) + tzp : lz + kStepBytes * lane + pending;
                    if (r) hist_run(r, run);
                }
                const uint32_t nlive = min(32u, nsteps - j0);
                if (m) {
                    const uint32_t last = 31u - (uint32_t)__clz((int)m);
                    pending = kStepBytes * (nlive - 1u - last) + __shfl_sync(0xFFFFFFFFu, tz, last);
                } else {
                    pending += kStepBytes * nlive;
                }
            }
            if (lane == 0) {
                if (pending) hist_run(pending, run);  // the block's last run
                o.list_n[blk] = kNoList;
            }
        }
        __syncthreads();
        // (2) histograms and per-step leading-zero counts of the dense blocks, sub-list lengths
        for (uint32_t i = tid; i < NH * kSymStride; i += kFrontThreads) {
            const uint32_t hb = i / kSymStride, sym = i - hb * kSymStride;
            if ((sparse_mask >> (hb / nblk)) & 1u) continue;
            const uint32_t v = sym == 0 ? s_run[hb * 8] : (sym < 256 ? s_lit[hb * 256 + sym] : (sym < (uint32_t)kNumSymbols ? s_run[hb * 8 + sym - 255] : 0u));
            o.hist[((size_t)f * NH + hb) * kSymStride + sym] = v;
        }
        for (uint32_t i = tid; i < NH * kMaxSteps; i += kFrontThreads) {
            const uint32_t hb = i / kMaxSteps, j = i - hb * kMaxSteps;
            const uint32_t k = hb / nblk, b = hb - k * nblk;
            if ((sparse_mask >> k) & 1u) continue;
            if (j < blk_len(s, b) / kStepBytes) {
                const uint32_t flat = b * kBlock + j * kStepBytes;
                const uint32_t c = flat / ns, tt = (flat - c * ns) / kStepBytes;
                o.step_lz[((size_t)f * NH + hb) * kMaxSteps + j] = s_lz[(k * CH + c) * T + tt];
            }
        }
        for (uint32_t i = tid; i < nb * CH; i += kFrontThreads) o.sub_n[(size_t)f * nb * CH + i] = s_subn[i];
        __syncthreads();
        if (tid == 0) {
            if (!only) {
                o.redo[f] = s_over ? 1 : 0;
                o.redo2[f] = 0;
            }
            s_over = 0;
        }
        if (tid < 4) s_probe[tid] = 0;
        for (uint32_t i = tid; i < NH * 264 + nb * CH * 9; i += kFrontThreads) s_lit[i] = 0;
        __syncthreads();
    }
    if (tid == 0) bulk_wait<0>();   // the plane rows have reached global memory
}

}  // namespace rspt
