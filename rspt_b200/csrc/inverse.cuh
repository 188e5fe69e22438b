// planes -> samples in ONE pass over the planes: a thread-block CLUSTER per frame.
//
// Replaces, for the lossless packers, decompress_i32's reassembly (signal_packer_base.cpp:122-138),
// xor_decode_32 (utils.cpp:232-236), offset_32(+128) and delta_decode (:204-219) and
// convert_i32_to_native (:51-121).  The two decode-side chains are scans over the FLAT [ch * ns]
// order that cross channel rows: d = prefix-xor(y), x = prefix-sum(d + 128).
//
//   * CTA r of the cluster owns the SAMPLE range [r * S, (r + 1) * S), S = ns / C, of every channel.
//     A warp takes a PIECE: 128 consecutive samples of four neighbouring channels, one sample quad
//     (4 samples x 4 channels) per lane, loaded once from the planes and kept in registers.
//   * Round 1: xor of every (channel, piece) -> shared memory; the CTA's total per channel is written
//     into every CTA of the cluster (distributed shared memory, st.shared::cluster); after a cluster
//     barrier a warp xor-reduces what lies in front of its pieces in flat order: the channels before,
//     the CTAs before in the same channel, the pieces before in the same CTA.  Round 2: the same for
//     the sums of (d + 128).
//   * The lane packs its 4 x 4 samples with PRMT into `bps` words per sample row, in a shared-memory
//     tile that holds the CTA's whole (contiguous) share of the output; the tile leaves with bulk
//     shared -> global copies (TMA unit).
// HBM traffic: the planes once, the samples once.  (k_planes_to_samples_fast reads the planes three times.)
#pragma once

#include "bulk.cuh"
#include "common.cuh"
#include "transforms.cuh"

namespace rspt {

constexpr uint32_t kInvTileQuads = 64;   // sample quads per output tile

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
// the same shared-memory offset in the CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t caddr, uint32_t v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cluster_v4(uint32_t caddr)
{
    uint4 v;
    asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(caddr));
    return v;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// dynamic shared memory: the CTA's share of the output
__host__ __device__ inline size_t inverse_cluster_smem(int bps, int ch, uint32_t S) { return (size_t)S * ch * bps + 128; }

constexpr uint32_t kInvMaxPieces = 32;   // pieces of 128 samples per channel in one CTA
constexpr uint32_t kInvMaxCh = 32;
constexpr uint32_t kInvMaxSeg = 768;     // (channel, CTA) segments of a frame: ch * C

// NBT = planes held (>= the frame's plane count); `scan` = the xdelta chain (0: plain hzr packer, y is the
// sample).  blockDim = 32 * (ch / 4) * (S / 128); gridDim = frames * cluster size.
// CHAIN = false: the CTAs of a frame are a cluster and exchange their totals through distributed shared memory.
// CHAIN = true: plain CTAs (blockIdx = frame * C + r, no co-scheduling); a CTA publishes its totals in global
// memory behind a release flag and waits for the other CTAs of its frame.  CTAs are dispatched in block-index
// order, so when one CTA of a frame is resident every CTA of every earlier frame has been dispatched and can
// finish: the frame's remaining CTAs always get their turn.  `epoch` distinguishes the launches, so the flags
// need no clearing.
struct InvChain {
    uint32_t* tot;     // [frames][2 rounds][C][ch]
    uint32_t* flag;    // [frames][2 rounds][C]
    uint32_t epoch;
    uint32_t C;
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int BPS, int NBT, bool CHAIN>
__global__ void __launch_bounds__(768, 2) k_inverse_cluster(const uint8_t* __restrict__ planes, Shape s, const uint8_t* __restrict__ dec_nb,
                                                          uint8_t* __restrict__ dst_raw, int scan, uint32_t S, InvChain chain)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_px[kInvMaxCh][kInvMaxPieces];   // per (channel, piece) of this CTA: xor / sum
    __shared__ uint32_t s_cx[kInvMaxSeg];                 // per (channel, CTA) of the frame: totals, written by their owners
    __shared__ uint32_t s_cs[kInvMaxSeg];

    const uint32_t C = CHAIN ? chain.C : cluster_nctarank();
    const uint32_t r = CHAIN ? blockIdx.x % C : cluster_ctarank(), f = CHAIN ? blockIdx.x / C : cluster_id_x();
    const uint32_t ns = (uint32_t)s.ns, ch = (uint32_t)s.ch, G = ch >> 2, P = S >> 7;
    const uint32_t nb = dec_nb[f], nba = s.nb_alloc;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t g = wid % G, p = wid / G;                  // my channel group and piece
    const uint32_t row = ch * BPS, rw = row >> 2;             // bytes per sample row, words per row
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem);

    // ---- my quad of four channels: plane words -> the sign-extended 32-bit words y
    uint32_t y[4][4];
    {
        const uint8_t* fp = planes + (size_t)f * nba * s.plane_stride + (size_t)r * S + p * 128u + 4u * lane;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            uint32_t q[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < NBT; ++k)
                if ((uint32_t)k < nb) q[k] = __ldg(reinterpret_cast<const uint32_t*>(fp + (size_t)k * s.plane_stride + (size_t)(4 * g + cc) * ns));
            planes_to_words(q[0], q[1], q[2], q[3], nb, y[cc]);
        }
    }
    if (scan) {
        // what lies in front of the pieces of channel c in this CTA: channels before, then CTAs before
        auto front = [&](const uint32_t* tot, uint32_t c, bool add) {
            const uint32_t me = c * C + r;
            uint32_t v = 0;
            for (uint32_t j = lane; j < me; j += 32) v = add ? v + tot[j] : v ^ tot[j];
            return add ? __reduce_add_sync(0xFFFFFFFFu, v) : __reduce_xor_sync(0xFFFFFFFFu, v);
        };
        // the CTA's total per channel (from the pieces in s_px) goes to every CTA of the frame that needs it; on
        // return tot[c * C + r'] holds the totals of every CTA r' of the frame
        auto exchange = [&](uint32_t round, uint32_t* tot, bool add) {
            if (!CHAIN) {
                if (tid < ch) {
                    uint32_t t = 0;
                    for (uint32_t pp = 0; pp < P; ++pp) t = add ? t + s_px[tid][pp] : t ^ s_px[tid][pp];
                    for (uint32_t q2 = 0; q2 < C; ++q2) st_cluster_u32(mapa(smem_u32(&tot[tid * C + r]), q2), t);
                }
                cluster_sync_all();
                return;
            }
            uint32_t* gt = chain.tot + ((size_t)(f * 2 + round) * C) * ch;
            uint32_t* gf = chain.flag + (size_t)(f * 2 + round) * C;
            if (tid < ch) {
                uint32_t t = 0;
                for (uint32_t pp = 0; pp < P; ++pp) t = add ? t + s_px[tid][pp] : t ^ s_px[tid][pp];
                gt[r * ch + tid] = t;
                tot[tid * C + r] = t;
                __threadfence();
            }
            __syncthreads();
            if (tid == 0) st_release(gf + r, chain.epoch);
            // wait for the CTAs in front of mine (one lane per CTA), then fetch their totals
            if (wid == 0) {
                if (lane < C && lane != r) while (ld_acquire(gf + lane) != chain.epoch) { }
                __syncwarp();
            }
            __syncthreads();
            for (uint32_t i = tid; i < C * ch; i += blockDim.x) {
                const uint32_t rr = i / ch, c = i - rr * ch;
                if (rr != r) tot[c * C + rr] = __ldcg(gt + rr * ch + c);
            }
            __syncthreads();
        };
        // ---- round 1: xor
        uint32_t inc[4], x[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            x[cc] = y[cc][0] ^ y[cc][1] ^ y[cc][2] ^ y[cc][3];
            inc[cc] = warp_xor_inclusive(x[cc]);
        }
        if (lane == 31) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) s_px[4 * g + cc][p] = inc[cc];
        }
        __syncthreads();
        exchange(0, s_cx, false);
        uint32_t sum[4], sinc[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const uint32_t c = 4 * g + cc;
            uint32_t run = front(s_cx, c, false);
            for (uint32_t pp = 0; pp < p; ++pp) run ^= s_px[c][pp];
            run ^= inc[cc] ^ x[cc];                           // xor of everything in front of my first sample
            sum[cc] = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                run ^= y[cc][j];
                y[cc][j] = run + 128u;                        // d + 128
                sum[cc] += y[cc][j];
            }
            sinc[cc] = warp_add_inclusive(sum[cc]);
        }
        __syncthreads();                                       // everybody has read the xor pieces
        // ---- round 2: sums of (d + 128)
        if (lane == 31) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) s_px[4 * g + cc][p] = sinc[cc];
        }
        __syncthreads();
        exchange(1, s_cs, true);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const uint32_t c = 4 * g + cc;
            uint32_t acc = front(s_cs, c, true);
            for (uint32_t pp = 0; pp < p; ++pp) acc += s_px[c][pp];
            acc += sinc[cc] - sum[cc];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc += y[cc][j];
                y[cc][j] = acc;
            }
        }
    }
    // ---- pack the 4 channels of every sample row into BPS words (convert_i32_to_native, utils.cpp:51-121)
    {
        uint32_t* out = tile + (size_t)(p * 32u + lane) * row + g * BPS;   // a quad of rows = 4 * row bytes = `row` words
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t a = y[0][i], b = y[1][i], c2 = y[2][i], d = y[3][i];
            if (BPS == 4) {
                out[i * rw + 0] = a; out[i * rw + 1] = b; out[i * rw + 2] = c2; out[i * rw + 3] = d;
            } else if (BPS == 3) {
                out[i * rw + 0] = prmt(a, b, 0x4210u);
                out[i * rw + 1] = prmt(b, c2, 0x5421u);
                out[i * rw + 2] = prmt(c2, d, 0x6542u);
            } else if (BPS == 2) {
                out[i * rw + 0] = prmt(a, b, 0x5410u);
                out[i * rw + 1] = prmt(c2, d, 0x5410u);
            } else {
                out[i * rw + 0] = prmt(prmt(a, b, 0x0040u), prmt(c2, d, 0x0040u), 0x5410u);
            }
        }
    }
    fence_async_smem();
    __syncthreads();
    // the CTA's share of the output is contiguous: one bulk copy per piece row block, issued by the warps' first lanes
    if (lane == 0 && g == 0) {
        uint8_t* fout = dst_raw + (size_t)f * s.frame_bytes + ((size_t)r * S + p * 128u) * row;
        bulk_s2g(fout, tile + (size_t)p * 32u * row, 128u * row);
        bulk_commit();
        bulk_wait<0>();
    }
}

}  // namespace rspt
