/* TEST INFRASTRUCTURE ONLY -- CPU side of the synthetic workload generator (include/rspt_synth.h). */
#ifndef ORACLE_SYNTH_ECG_H_
#define ORACLE_SYNTH_ECG_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* frames [first, first+n) of shape (bps, ch, ns) -> dst[n][ns][ch][bps] */
void oracle_synth_ecg(uint8_t* dst, uint64_t first_frame, size_t n, int bps, int ch, int ns,
                      uint64_t seed, int32_t amplitude, int32_t sigma);
#ifdef __cplusplus
}
#endif
#endif
