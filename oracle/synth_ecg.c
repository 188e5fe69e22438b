/* TEST INFRASTRUCTURE ONLY -- CPU side of the synthetic workload generator. */
#include "synth_ecg.h"
#include "../include/rspt_synth.h"

void oracle_synth_ecg(uint8_t* dst, uint64_t first_frame, size_t n, int bps, int ch, int ns,
                      uint64_t seed, int32_t amplitude, int32_t sigma)
{
    static int32_t beat[RSPT_SYNTH_TABLE], sine[RSPT_SYNTH_TABLE];
    static int ready = 0;
    if (!ready) {
        rspt_synth_build_tables(beat, sine);
        ready = 1;
    }
    rspt_synth_params p = {seed, amplitude, sigma};
    const size_t frame_bytes = (size_t)bps * ch * ns;
    for (size_t f = 0; f < n; ++f)
        for (int c = 0; c < ch; ++c) {
            rspt_synth_chan k = rspt_synth_channel(&p, first_frame + f, (uint32_t)c);
            for (int s = 0; s < ns; ++s) {
                uint32_t v = (uint32_t)rspt_synth_sample(&p, &k, beat, sine, first_frame + f,
                                                          (uint32_t)c, (uint32_t)s, bps);
                uint8_t* q = dst + f * frame_bytes + ((size_t)s * ch + c) * bps;
                for (int b = 0; b < bps; ++b)
                    q[b] = (uint8_t)(v >> (8 * b));
            }
        }
}
