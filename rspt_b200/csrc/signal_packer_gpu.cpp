// C++ host side of the drop-in: i_signal_packer (include/signal_packer.h, mirroring the
// reference's lib_rspt/signal_packer.h:29-73) implemented on the C ABI of include/rspt_gpu.h.
// No CUDA headers here; no CPU implementation of any stage either -- if the GPU library cannot
// create a handle the factories print why and return nullptr.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "rspt_gpu.h"
#include "signal_packer.h"

namespace {

class gpu_packer : public i_signal_packer
{
    rspt_gpu_packer* h_ = nullptr;
    int kind_;

public:
    gpu_packer(int kind, size_t bps, size_t ch, size_t ns, size_t nb) : kind_(kind)
    {
        const char* dev = std::getenv("RSPT_GPU_DEVICE");
        const int rc = rspt_gpu_create(kind, bps, ch, ns, nb, dev ? std::atoi(dev) : 0, nullptr, 1, &h_);
        if (rc != RSPT_OK) {
            std::cout << "ERROR: rspt_gpu_create failed (" << rc << ")." << std::endl;
            h_ = nullptr;
        }
    }
    ~gpu_packer()
    {
        if (h_) rspt_gpu_destroy(h_);
    }
    bool ok() const { return h_ != nullptr; }

    void compress(const unsigned char* src, unsigned char* dst, size_t dst_max_len, size_t& dst_len) override
    {
        unsigned before = 0, after = 0;
        rspt_gpu_nb(h_, &before);
        const int rc = rspt_gpu_compress_host(h_, src, dst, dst_max_len, &dst_len);
        if (rc != RSPT_OK) {
            // the reference ignores hzr's status codes (signal_packer_base.cpp:72); we do not
            std::cout << "ERROR: compression failed: " << rspt_gpu_last_error(h_) << std::endl;
            dst_len = 0;
            return;
        }
        rspt_gpu_nb(h_, &after);
        for (unsigned k = before; k < after; ++k)
            std::cout << "Compression needs one more byte to encode." << std::endl;  // xdelta.cpp:65
    }

    int decompress(const unsigned char* src, size_t& src_len, unsigned char* dst) override
    {
        const int rc = rspt_gpu_decompress_host(h_, src, &src_len, dst);
        if (rc == RSPT_E_STREAM)
            std::cout << "ERROR: compression method unsupported." << std::endl;  // xdelta.cpp:79 etc.
        else if (rc != RSPT_OK)
            std::cout << "ERROR: decompression failed: " << rspt_gpu_last_error(h_) << std::endl;
        return 0;  // the reference always returns 0 (xdelta.cpp:84, hadamard.cpp:103, dct.cpp:152)
    }
};

i_signal_packer* make(int kind, size_t bps, size_t ch, size_t ns, size_t nb)
{
    gpu_packer* p = new gpu_packer(kind, bps, ch, ns, nb);
    if (!p->ok()) {
        delete p;
        return nullptr;
    }
    return p;
}

}  // namespace

i_signal_packer* i_signal_packer::new_xdelta_hzr(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel, size_t nr_bytes_to_encode)
{
    return make(RSPT_XDELTA_HZR, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, nr_bytes_to_encode);
}
void i_signal_packer::delete_xdelta_hzr(i_signal_packer* instance) { delete static_cast<gpu_packer*>(instance); }

i_signal_packer* i_signal_packer::new_hzr(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel)
{
    return make(RSPT_HZR, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, 4);
}
void i_signal_packer::delete_hzr(i_signal_packer* instance) { delete static_cast<gpu_packer*>(instance); }

i_signal_packer* i_signal_packer::new_dct(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel)
{
    return make(RSPT_DCT, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, 2);
}
void i_signal_packer::delete_dct(i_signal_packer* instance) { delete static_cast<gpu_packer*>(instance); }

i_signal_packer* i_signal_packer::new_hadamard(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel)
{
    return make(RSPT_HADAMARD, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, 3);
}
void i_signal_packer::delete_hadamard(i_signal_packer* instance) { delete static_cast<gpu_packer*>(instance); }
