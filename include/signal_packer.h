// signal_packer.h -- drop-in for rspt's packer interface, backed by the B200 library.
//
// Same class name, method signatures and factory/deleter pairs as the reference's
// lib_rspt/signal_packer.h:29-73, so a caller written against rspt (the README example, the
// test harness lib_rspt_test/rspt_test.cpp:58-112) recompiles against this header and links
// librspt_packer.so unchanged.  The implementation (rspt_b200/csrc/signal_packer_gpu.cpp)
// calls CUDA only through the C ABI in rspt_gpu.h.
//
// Behavioural notes carried over from the reference:
//  * compress() returns nothing and decompress() always returns 0 (signal_packer_xdelta_hzr.cpp:84);
//    problems are printed to stdout with the reference's own messages.
//  * an instance is not re-entrant; xdelta_hzr's plane count grows permanently when a frame
//    needs one more byte (signal_packer_xdelta_hzr.cpp:63-69).
//  * new_lala / delete_lala are declared but never defined, exactly like the reference (:71-72).
// Additions (non-virtual, so the vtable layout of the two reference methods is unchanged):
//  * the environment variable RSPT_GPU_DEVICE selects the CUDA device (default 0).
#ifndef RSPT_B200_SIGNAL_PACKER_H_
#define RSPT_B200_SIGNAL_PACKER_H_

#include <cstddef>

class i_signal_packer
{
public:
    /// Compress one frame of bytes_per_channel * nr_of_channels * nr_of_samples bytes from `src`
    /// into `dst` (capacity `dst_max_len`); the produced length is returned through `dst_len`.
    virtual void compress(const unsigned char* src, unsigned char* dst, size_t dst_max_len, size_t& dst_len) = 0;

    /// Decompress one frame from `src` into `dst`; the number of compressed bytes consumed is
    /// returned through `src_len` (the caller does not have to know it beforehand).
    virtual int decompress(const unsigned char* src, size_t& src_len, unsigned char* dst) = 0;

    static i_signal_packer* new_xdelta_hzr(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel, size_t nr_bytes_to_encode);
    static void delete_xdelta_hzr(i_signal_packer* instance);

    static i_signal_packer* new_hzr(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel);
    static void delete_hzr(i_signal_packer* instance);

    static i_signal_packer* new_dct(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel);
    static void delete_dct(i_signal_packer* instance);

    static i_signal_packer* new_hadamard(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel);
    static void delete_hadamard(i_signal_packer* instance);

    static i_signal_packer* new_lala(size_t bytes_per_channel, size_t nr_of_channels, size_t nr_of_samples_in_each_channel);
    static void delete_lala(i_signal_packer* instance);
};

#endif
