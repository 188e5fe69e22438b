/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the rspt signal-packer hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (rspt_b200/) never links, imports or executes it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against the
 * unmodified reference compiled into oracle/_ref/libref.so (same inputs -> same bytes), and
 * tests/test_oracle_golden.py checks it against the committed known answers in tests/golden/
 * (generated from the reference by tests/golden/make_golden.py).
 *
 * All file:line citations are relative to /root/reference/lib_rspt/.
 */
#ifndef RSPT_ORACLE_H_
#define RSPT_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_XDELTA_HZR = 0, ORACLE_HZR = 1, ORACLE_HADAMARD = 2, ORACLE_DCT = 3 };

#define ORACLE_HZR_NSYM 261
#define ORACLE_HZR_BLOCK 65536

typedef struct oracle_packer oracle_packer;

/* i_signal_packer factories, signal_packer.h:59-69 */
oracle_packer* oracle_new(int kind, size_t bps, size_t ch, size_t ns, size_t nb);
void oracle_delete(oracle_packer* p);
/* compress / decompress, signal_packer.h:44,57.  Return 0 on success. */
int oracle_compress(oracle_packer* p, const uint8_t* src, uint8_t* dst, size_t dst_cap, size_t* dst_len);
int oracle_decompress(oracle_packer* p, const uint8_t* src, size_t* src_len, uint8_t* dst);
/* current number of byte planes (xdelta_hzr escalates it, signal_packer_xdelta_hzr.cpp:63-69) */
unsigned oracle_nb(const oracle_packer* p);
/* number of "needs one more byte" escalations so far */
unsigned oracle_escalations(const oracle_packer* p);
/* tight bound: 1 + hdr + nb*(4 + hzr_max(ch*ns)) */
size_t oracle_max_compressed_size(const oracle_packer* p);
size_t oracle_header_bytes(const oracle_packer* p);

/* Stage outputs for kernel-level parity checks: the int32 words handed to compress_i32
 * (signal_packer_base.cpp:38) in flat channel-major order, and the 3-byte-per-channel header. */
int oracle_transform(oracle_packer* p, const uint8_t* src, int32_t* words /*[ch*ns]*/, uint8_t* header);
/* inverse of the above: words (already sign-extended from 8*nb bits) + header -> native bytes */
int oracle_inverse(oracle_packer* p, const int32_t* words, const uint8_t* header, uint8_t* dst);

/* loops for the CPU baseline */
size_t oracle_compress_many(oracle_packer* p, const uint8_t* src, size_t frame_bytes, size_t n,
                            uint8_t* dst, size_t dst_stride, uint32_t* sizes);
size_t oracle_decompress_many(oracle_packer* p, const uint8_t* src, size_t src_stride, size_t n,
                              uint8_t* dst, size_t frame_bytes);

/* hzr codec, lib_hzr/hzr_encode.c, hzr_decode.c */
uint32_t oracle_crc32c(const void* data, size_t n);
size_t oracle_hzr_max_compressed_size(size_t n);
int oracle_hzr_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* enc);
int oracle_hzr_decode(const uint8_t* in, size_t n, uint8_t* out, size_t out_size);
int oracle_hzr_verify(const uint8_t* in, size_t n, size_t* decoded);

/* hzr stages (one block, n <= 65536) */
void oracle_hzr_histogram(const uint8_t* in, size_t n, uint32_t hist[ORACLE_HZR_NSYM]);
/* returns the number of used symbols; code/len per symbol; tree bits (LSB-first) into tree[] and
 * their count into *tree_nbits.  tree[] must hold >= 360 bytes. */
int oracle_hzr_build_codes(const uint32_t hist[ORACLE_HZR_NSYM], uint32_t code[ORACLE_HZR_NSYM],
                           uint8_t len[ORACLE_HZR_NSYM], uint8_t* tree, uint32_t* tree_nbits);
/* mode (0 COPY,1 HUFF,2 FILL) and payload length a block of n bytes will get */
int oracle_hzr_block_plan(const uint8_t* in, size_t n, uint32_t* payload_len);

/* misc reference arithmetic */
int32_t oracle_average_32(const int32_t* a, size_t len);                 /* utils.cpp:30-40 */
void oracle_fwht(int n, const int32_t* src, int32_t* dst);              /* lib_fwht/fwht.c:4-28 */
/* PRDN as printed by lib_rspt_test/rspt_test.cpp:98-111 (computed in double throughout) */
double oracle_prdn(const uint8_t* orig, const uint8_t* dec, size_t bps, size_t ch, size_t ns);

/* Pre-filter step in front of the packers, in place on one native frame, as the reference's test
 * harness applies it (lib_rspt_test/rspt_test.cpp:116-136): one filter object walked over the
 * channels, init_history_values(first sample), filter_opt per sample, truncation to int32.
 * IIR: lib_filter/iir_filter.cpp:46-116 (2..5 coefficients); FIR: lib_filter/fir_filter.cpp:26-68. */
int oracle_prefilter_iir(uint8_t* frame, size_t bps, size_t ch, size_t ns, const double* n, const double* d,
                         int nr_coefficients, int init_nr_samples);
int oracle_prefilter_fir(uint8_t* frame, size_t bps, size_t ch, size_t ns, const double* kernel, int kernel_size);

#ifdef __cplusplus
}
#endif
#endif
