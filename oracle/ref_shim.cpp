// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" shim around the UNMODIFIED reference (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libref.so).  It exposes the reference's own
// i_signal_packer factories (lib_rspt/signal_packer.h:29-73) and the vendored hzr codec
// (lib_rspt/lib_hzr/libhzr.h) to ctypes so that tests and bench.py --impl reference can
// run the real thing.  Nothing here restates any algorithm (the pre-filter helper repeats the
// reference test harness's own calling sequence around i_filter).
#include <cstddef>
#include <cstdint>
#include "signal_packer.h"
#include "lib_hzr/libhzr.h"
#include <vector>
using namespace std;  // filter.h names vector unqualified, like the reference's own sources that include it
#include "filter.h"
#include "lib_signalpacker/utils.h"
extern "C" uint32_t _hzr_crc32(const void* data, size_t length);

extern "C" {

// kind: 0 xdelta_hzr, 1 hzr, 2 hadamard, 3 dct  (same numbering as include/rspt_gpu.h)
void* ref_new(int kind, size_t bps, size_t ch, size_t ns, size_t nb)
{
    switch (kind) {
    case 0: return i_signal_packer::new_xdelta_hzr(bps, ch, ns, nb);
    case 1: return i_signal_packer::new_hzr(bps, ch, ns);
    case 2: return i_signal_packer::new_hadamard(bps, ch, ns);
    case 3: return i_signal_packer::new_dct(bps, ch, ns);
    }
    return nullptr;
}

void ref_delete(int kind, void* p)
{
    i_signal_packer* q = (i_signal_packer*)p;
    switch (kind) {
    case 0: i_signal_packer::delete_xdelta_hzr(q); break;
    case 1: i_signal_packer::delete_hzr(q); break;
    case 2: i_signal_packer::delete_hadamard(q); break;
    case 3: i_signal_packer::delete_dct(q); break;
    }
}

size_t ref_compress(void* p, const unsigned char* src, unsigned char* dst, size_t dst_max_len)
{
    size_t n = 0;
    ((i_signal_packer*)p)->compress(src, dst, dst_max_len, n);
    return n;
}

size_t ref_decompress(void* p, const unsigned char* src, unsigned char* dst)
{
    size_t n = 0;
    ((i_signal_packer*)p)->decompress(src, n, dst);
    return n;
}

// Loop helpers for the CPU baseline: n frames back to back, one packer instance.
size_t ref_compress_many(void* p, const unsigned char* src, size_t frame_bytes, size_t n,
                         unsigned char* dst, size_t dst_stride, uint32_t* sizes)
{
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t len = 0;
        ((i_signal_packer*)p)->compress(src + i * frame_bytes, dst + i * dst_stride, dst_stride, len);
        if (sizes) sizes[i] = (uint32_t)len;
        total += len;
    }
    return total;
}

size_t ref_decompress_many(void* p, const unsigned char* src, size_t src_stride, size_t n,
                           unsigned char* dst, size_t frame_bytes)
{
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t len = 0;
        ((i_signal_packer*)p)->decompress(src + i * src_stride, len, dst + i * frame_bytes);
        total += len;
    }
    return total;
}

size_t ref_hzr_max_compressed_size(size_t n) { return hzr_max_compressed_size(n); }

int ref_hzr_encode(const void* in, size_t in_size, void* out, size_t out_size, size_t* enc)
{
    return (int)hzr_encode(in, in_size, out, out_size, enc);
}

int ref_hzr_decode(const void* in, size_t in_size, void* out, size_t out_size)
{
    return (int)hzr_decode(in, in_size, out, out_size);
}

int ref_hzr_verify(const void* in, size_t in_size, size_t* dec)
{
    return (int)hzr_verify(in, in_size, dec);
}

uint32_t ref_crc32c(const void* data, size_t n) { return _hzr_crc32(data, n); }

// The reference's pre-filter step exactly as its test harness applies it before packing
// (lib_rspt_test/rspt_test.cpp:116-136): native -> int32 matrix, ONE i_filter instance walked
// over the channels (init_history_values(first sample, init_nr_samples), then filter_opt per
// sample, result truncated to int32), int32 matrix -> native.  In place on one frame.
static void prefilter(i_filter* filter, uint8_t* frame, int bps, int ch, int ns, int init_nr_samples)
{
    std::vector<int32_t> buf((size_t)ch * ns);
    std::vector<int32_t*> rows(ch);
    for (int j = 0; j < ch; ++j) rows[j] = buf.data() + (size_t)j * ns;
    convert_native_to_i32(rows.data(), frame, ns, ch, bps, false);
    for (int j = 0; j < ch; ++j) {
        filter->init_history_values(rows[j][0], init_nr_samples);
        for (int i = 0; i < ns; ++i) rows[j][i] = filter->filter_opt(rows[j][i]);
    }
    convert_i32_to_native(frame, rows.data(), ns, ch, bps, false);
}

void ref_prefilter_iir(uint8_t* frame, int bps, int ch, int ns, const double* n, const double* d, int nr_coefficients,
                       int init_nr_samples)
{
    i_filter* f = i_filter::new_iir(n, d, (size_t)nr_coefficients);
    prefilter(f, frame, bps, ch, ns, init_nr_samples);
    i_filter::delete_iir(f);
}

void ref_prefilter_fir(uint8_t* frame, int bps, int ch, int ns, const double* kernel, int kernel_size)
{
    i_filter* f = i_filter::new_fir(kernel, (size_t)kernel_size);
    prefilter(f, frame, bps, ch, ns, kernel_size);
    i_filter::delete_fir(f);
}

}  // extern "C"
