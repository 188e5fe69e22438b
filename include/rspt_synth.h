/* Synthetic ECG-like multichannel input -- the workload generator shared by the CUDA library
 * (device kernel + host table builder), the CPU oracle and the tests.  Integer-only per sample
 * so host and device produce identical bytes.
 *
 * sample(f, c, s) = dc + gain * beat[(s + phase) mod rr] + baseline wander + noise
 *   f = GLOBAL frame index (sharding a stream over GPUs does not change its bytes)
 *   per (f, c): gain in [0.4, 1.0) * A, dc in +-A/2, rr in [600, 980] samples, phase in [0, rr)
 *   beat     = P + QRS + T Gaussian template, 1024 entries, Q14
 *   baseline = A/8 * sin(2 pi s / 4096 + c * 77/1024 turns), 1024-entry Q14 sine table
 *   noise    = (sum of four 16-bit fields of a splitmix64 hash - 2*65535) * sigma / 37837
 *              (Irwin-Hall, approximately Gaussian with standard deviation sigma LSB)
 * The value is clamped to the signed range of 8*bps bits and stored little-endian, interleaved
 * [ns][ch][bps] -- the layout convert_native_to_i32 expects (lib_signalpacker/utils.cpp:123).
 */
#ifndef RSPT_SYNTH_H_
#define RSPT_SYNTH_H_

#include <stdint.h>

#ifdef __CUDACC__
#define RSPT_HD __host__ __device__ __forceinline__
#else
#define RSPT_HD static inline
#endif

#define RSPT_SYNTH_TABLE 1024

typedef struct {
    uint64_t seed;      /* 42 in the benchmarks */
    int32_t amplitude;  /* A: 20000 for the 24/32-bit shapes */
    int32_t sigma;      /* noise standard deviation in LSB: 3 */
} rspt_synth_params;

RSPT_HD uint64_t rspt_splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct {
    int32_t gain_q16;
    int32_t dc;
    int32_t rr;
    int32_t phase;
} rspt_synth_chan;

RSPT_HD rspt_synth_chan rspt_synth_channel(const rspt_synth_params* p, uint64_t frame, uint32_t c)
{
    uint64_t h = rspt_splitmix64(p->seed ^ rspt_splitmix64(131ull * frame + c));
    rspt_synth_chan k;
    k.gain_q16 = 26214 + (int32_t)(((h & 0xFFFFu) * 39322u) >> 16);
    k.dc = (int32_t)((((int64_t)((h >> 16) & 0xFFFFu) - 32768) * (int64_t)p->amplitude) >> 16);
    k.rr = 600 + (int32_t)((((h >> 32) & 0xFFFFu) * 381u) >> 16);
    k.phase = (int32_t)(((h >> 48) & 0xFFFFu) % (uint32_t)k.rr);
    return k;
}

RSPT_HD int32_t rspt_synth_sample(const rspt_synth_params* p, const rspt_synth_chan* k,
                                  const int32_t* beat, const int32_t* sine, uint64_t frame,
                                  uint32_t c, uint32_t s, int bps)
{
    uint32_t pos = (s + (uint32_t)k->phase) % (uint32_t)k->rr;
    uint32_t idx = (pos * RSPT_SYNTH_TABLE) / (uint32_t)k->rr;
    int64_t wave = ((int64_t)p->amplitude * beat[idx]) >> 14;
    wave = (wave * k->gain_q16) >> 16;
    uint32_t bidx = ((s >> 2) + c * 77u) & (RSPT_SYNTH_TABLE - 1);
    int64_t base = ((int64_t)(p->amplitude >> 3) * sine[bidx]) >> 14;
    uint64_t h = rspt_splitmix64(p->seed + 0x1234567ull * frame + 1000003ull * c + s);
    int32_t u = (int32_t)(h & 0xFFFFu) + (int32_t)((h >> 16) & 0xFFFFu) +
                (int32_t)((h >> 32) & 0xFFFFu) + (int32_t)((h >> 48) & 0xFFFFu) - 2 * 65535;
    int64_t noise = ((int64_t)u * p->sigma) / 37837;
    int64_t v = (int64_t)k->dc + wave + base + noise;
    const int64_t hi = ((int64_t)1 << (8 * bps - 1)) - 1, lo = -hi - 1;
    if (v > hi) v = hi;
    if (v < lo) v = lo;
    return (int32_t)v;
}

#include <math.h>
/* Host-side table construction (double math, rounded once). */
static inline void rspt_synth_build_tables(int32_t* beat, int32_t* sine)
{
    /* centre, width (fraction of the beat), height: P, Q, R, S, T */
    static const double g[5][3] = {{0.18, 0.025, 0.12}, {0.285, 0.008, -0.10}, {0.30, 0.010, 1.00},
                                   {0.318, 0.009, -0.22}, {0.55, 0.045, 0.28}};
    for (int i = 0; i < RSPT_SYNTH_TABLE; ++i) {
        double t = (double)i / RSPT_SYNTH_TABLE, v = 0;
        for (int k = 0; k < 5; ++k) {
            double d = (t - g[k][0]) / g[k][1];
            v += g[k][2] * exp(-0.5 * d * d);
        }
        beat[i] = (int32_t)lrint(v * 16384.0);
        sine[i] = (int32_t)lrint(sin(2.0 * 3.14159265358979323846 * t) * 16384.0);
    }
}

#endif
