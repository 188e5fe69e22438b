// The reference's README example / test_5 (lib_rspt_test/rspt_test.cpp:225-256), compiled against
// THIS repo's include/signal_packer.h and linked with librspt_packer.so.  Prints the compressed
// size and a round-trip verdict; tests/test_dropin_cxx.py checks the output.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <vector>

#include "signal_packer.h"

int main()
{
    const int bytes_per_sample = 4, nr_samples = 8192, nr_channels = 1;
    std::vector<int32_t> data_stream(nr_samples);
    for (int i = 0; i < nr_samples; ++i) data_stream[i] = sin(i / 100.0) * 1000.0;

    i_signal_packer* c = i_signal_packer::new_xdelta_hzr(bytes_per_sample, nr_channels, nr_samples, 3);
    if (!c) return 2;
    size_t dst_max_len = nr_samples * nr_channels * bytes_per_sample * 2;
    std::vector<unsigned char> dst(dst_max_len), decdst(dst_max_len);
    size_t compressed_size = 0, cmpr_size = 0;
    c->compress((uint8_t*)data_stream.data(), dst.data(), dst_max_len, compressed_size);
    int rc = c->decompress(dst.data(), cmpr_size, decdst.data());
    const bool same = std::memcmp(decdst.data(), data_stream.data(), nr_samples * 4) == 0;
    std::cout << "compressed_size: " << compressed_size << " consumed: " << cmpr_size << " rc: " << rc
              << " roundtrip: " << (same ? "ok" : "MISMATCH")
              << " CR = " << (double)(nr_channels * bytes_per_sample * nr_samples) / cmpr_size << std::endl;
    i_signal_packer::delete_xdelta_hzr(c);

    // the lossy packers through the same interface
    i_signal_packer* h = i_signal_packer::new_hadamard(bytes_per_sample, nr_channels, nr_samples);
    i_signal_packer* d = i_signal_packer::new_dct(bytes_per_sample, nr_channels, 4096);
    i_signal_packer* z = i_signal_packer::new_hzr(bytes_per_sample, nr_channels, nr_samples);
    if (!h || !d || !z) return 3;
    size_t n1 = 0, n2 = 0, n3 = 0;
    h->compress((uint8_t*)data_stream.data(), dst.data(), dst_max_len, n1);
    d->compress((uint8_t*)data_stream.data(), dst.data(), dst_max_len, n2);
    z->compress((uint8_t*)data_stream.data(), dst.data(), dst_max_len, n3);
    std::cout << "hadamard: " << n1 << " dct: " << n2 << " hzr: " << n3 << std::endl;
    i_signal_packer::delete_hadamard(h);
    i_signal_packer::delete_dct(d);
    i_signal_packer::delete_hzr(z);
    return same ? 0 : 1;
}
