// Bit-packing and CRC-32C helpers shared by the kernels that produce hzr payloads
// (k_hzr_encode_sparse for blocks with a list of non-zero bytes, k_hzr_encode for the rest).
// WriteBits / FlushBitCache: lib_hzr/hzr_encode.c:63-113; CRC-32C: lib_hzr/hzr_crc32c.c:77-97.
#pragma once

#include "bulk.cuh"
#include "common.cuh"

namespace rspt {

// CRC-32C constants, built on the host at library load (crc_tables.cpp) and kept in global
// memory: byte table, the per-lane multipliers x^(32(j+1)) and, for every supported CTA width
// T, the 4x256 table of Z^(4T) (advance the register by 4T zero bytes).
struct CrcConst {
    uint32_t byte_tab[256];
    uint32_t lane_mul[1024];
    uint32_t zt[4][4][256];  // [log2(T/128)][byte][value]
    uint32_t zt32[4][256];   // the same table for T = 32 (warp_crc32c)
};

__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
    // product of two polynomials mod P in the reflected representation (bit 31 = x^0)
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        p ^= b & (0u - ((a >> (31 - i)) & 1u));
        b = (b >> 1) ^ (0x82F63B78u & (0u - (b & 1u)));
    }
    return p;
}

// CRC-32C of `len` bytes that start `lead` (0..3) bytes into the WORD-ALIGNED shared-memory array `words`; the
// `lead` bytes in front of them must be zero (leading zeros leave a zero CRC register unchanged, so the message
// can be folded word by word from words[0]).  hzr_crc32c.c:77-84 semantics: init ~0, final ~.  All threads of the
// CTA must call it; the result is returned to every thread.  s_zt is the CTA's copy of zt[log2(T/128)],
// s_red holds 33 words.  blockDim.x must be 128, 256, 512 or 1024.
__device__ __forceinline__ uint32_t block_crc32c(const uint32_t* words, uint32_t len, const uint32_t* s_zt,
                                                 const CrcConst* __restrict__ cc, uint32_t* s_red, uint32_t lead = 0)
{
    const uint32_t T = blockDim.x, j = threadIdx.x;
    const uint32_t tot = lead + len;
    // init 0xFFFFFFFF == complement of the first four bytes of the message: bytes lead .. lead + 3
    const uint32_t m0 = 0xFFFFFFFFu << (8u * lead), m1 = ~m0;
    uint32_t part = 0;
    if (tot >= 8) {
        const uint32_t W = tot >> 2;
        if (j < W) {
            uint32_t S = 0;
            for (uint32_t i = (W - 1 - j) % T; i < W; i += T) {
                uint32_t w = words[i];
                if (i == 0) w ^= m0;
                if (i == 1) w ^= m1;
                S = s_zt[S & 255u] ^ s_zt[256 + ((S >> 8) & 255u)] ^ s_zt[512 + ((S >> 16) & 255u)] ^
                    s_zt[768 + (S >> 24)] ^ w;
            }
            part = crc_mulmod(__ldg(&cc->lane_mul[j]), S);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, o);
    __syncthreads();
    if (lane_id() == 0) s_red[warp_id()] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (uint32_t w = 0; w < (T >> 5); ++w) s ^= s_red[w];
        const uint8_t* bytes = reinterpret_cast<const uint8_t*>(words);
        uint32_t q = tot & ~3u;
        if (tot < 8) {
            s = 0xFFFFFFFFu;
            q = lead;
        }
        for (; q < tot; ++q) s = (s >> 8) ^ __ldg(&cc->byte_tab[(s ^ bytes[q]) & 255u]);
        s_red[32] = ~s;
    }
    __syncthreads();
    return s_red[32];
}

// CRC-32C of `len` bytes at the WORD-ALIGNED shared-memory address `words`, computed by ONE warp
// (all 32 lanes must call it; every lane gets the result).  Same scheme as block_crc32c with
// T = 32; the advance table is read through the read-only cache instead of shared memory.
__device__ __forceinline__ uint32_t warp_crc32c(const uint32_t* words, uint32_t len, const CrcConst* __restrict__ cc)
{
    const uint32_t j = lane_id();
    const uint32_t* zt = &cc->zt32[0][0];
    uint32_t part = 0;
    if (len >= 8) {
        const uint32_t W = len >> 2;
        if (j < W) {
            uint32_t S = 0;
            for (uint32_t i = (W - 1 - j) & 31u; i < W; i += 32) {
                uint32_t w = words[i];
                if (i == 0) w = ~w;  // init 0xFFFFFFFF == complement of the first four bytes
                S = __ldg(zt + (S & 255u)) ^ __ldg(zt + 256 + ((S >> 8) & 255u)) ^ __ldg(zt + 512 + ((S >> 16) & 255u)) ^
                    __ldg(zt + 768 + (S >> 24)) ^ w;
            }
            part = crc_mulmod(__ldg(&cc->lane_mul[j]), S);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, o);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(words);
    uint32_t s = part, q = len & ~3u;
    if (len < 8) {
        s = 0xFFFFFFFFu;
        q = 0;
    }
    for (; q < len; ++q) s = (s >> 8) ^ __ldg(&cc->byte_tab[(s ^ bytes[q]) & 255u]);
    return ~s;
}

struct LenSink {
    const uint32_t* sc;
    uint32_t bits;
    __device__ __forceinline__ void token(uint32_t sym, uint32_t, uint32_t eb) { bits += (sc[sym] >> 27) + eb; }
};

struct EmitSink {
    const uint32_t* sc;
    uint32_t* stg;  // staging words; bit position 0 = payload bit 0
    unsigned long long acc;
    uint32_t nacc, wptr;
    bool first;
    __device__ __forceinline__ void flush()
    {
        if (first) {
            atomicOr(&stg[wptr], (uint32_t)acc);
            first = false;
        } else {
            stg[wptr] = (uint32_t)acc;
        }
        ++wptr;
        acc >>= 32;
        nacc -= 32;
    }
    __device__ __forceinline__ void append(uint32_t v, uint32_t nbits)
    {
        acc |= (unsigned long long)v << nacc;
        nacc += nbits;
        if (nacc >= 32) flush();
    }
    __device__ __forceinline__ void token(uint32_t sym, uint32_t ev, uint32_t eb)
    {
        const uint32_t c = sc[sym];
        append(c & 0x07FFFFFFu, c >> 27);
        if (eb) append(ev, eb);
    }
    __device__ __forceinline__ void finish()
    {
        if (nacc) atomicOr(&stg[wptr], (uint32_t)acc);
    }
};

// A token slot: value | nbits << 27.  A literal or single-token run symbol fills one slot with
// its code word; a run's extra bits fill the next slot.
__device__ __forceinline__ uint32_t slot_bits(uint32_t cw) { return cw >> 27; }

// bit length of the tokens of one zero run of z >= 1 bytes (hzr_encode.c:146-166)
__device__ __forceinline__ uint32_t run_bits(uint32_t z, const uint32_t* sc)
{
    LenSink ls{sc, 0};
    emit_run(z, ls);
    return ls.bits;
}

// copy `len` bytes from shared memory (byte offset `soff` into the word array `sw`) to an
// arbitrarily aligned global address, 4 bytes per thread-step
__device__ __forceinline__ void copy_smem_to_global(uint8_t* __restrict__ dst, const uint32_t* sw, uint32_t soff, uint32_t len)
{
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(sw);
    uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);
    if (head > len) head = len;
    if (threadIdx.x < head) dst[threadIdx.x] = sb[soff + threadIdx.x];
    const uint32_t nw = (len - head) >> 2;
    const uint32_t so = soff + head;
    const uint32_t sh = (so & 3u) * 8u, wbase = so >> 2;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
    for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x)
        dw[i] = __funnelshift_r(sw[wbase + i], sw[wbase + i + 1], sh);
    const uint32_t done = head + (nw << 2);
    if (threadIdx.x < len - done) dst[done + threadIdx.x] = sb[soff + done + threadIdx.x];
}

// The same when the shared-memory bytes sit at an offset congruent to the global address mod 16 (the caller staged
// them that way): the 16-byte-aligned middle leaves with ONE bulk asynchronous store (cp.async.bulk shared -> global,
// issued by thread 0), the <= 15 bytes on either side with byte stores.  Every thread must have fenced its writes
// (fence_async_smem) and the CTA synchronised before the call; thread 0 waits for the store to complete before
// returning, so the CTA may exit.
__device__ __forceinline__ void copy_smem_to_global_bulk(uint8_t* __restrict__ dst, const uint8_t* sb, uint32_t len)
{
    uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > len) head = len;
    const uint32_t mid = (len - head) & ~15u, tail = len - head - mid;
    if (threadIdx.x < head) dst[threadIdx.x] = sb[threadIdx.x];
    if (threadIdx.x < tail) dst[head + mid + threadIdx.x] = sb[head + mid + threadIdx.x];
    if (threadIdx.x == 0 && mid) {
        bulk_s2g(dst + head, sb + head, mid);
        bulk_commit();
        bulk_wait<0>();
    }
}

}  // namespace rspt
