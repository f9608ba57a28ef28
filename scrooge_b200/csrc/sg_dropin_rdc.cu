// sg_dropin_rdc.cu -- the one device-side symbol of the reference's GPU header (src/genasm_gpu.hpp:9):
//     __global__ void genasm_gpu::ascii_to_twobit_strings(int count, long long *string_lengths,
//                                                         char **ascii_strings, char **twobit_strings)
// The reference's own test binary launches it directly (src/tests.cu:626,703), so a build of the unmodified
// src/tests.cu against this library needs the kernel as RELOCATABLE device code: the reference compiles everything
// with -rdc=true (Makefile:9) and nvlink resolves the launch stub's kernel across translation units.  This file is
// compiled with -rdc=true into scrooge_b200/lib/libscrooge_b200_rdc.a, which replaces src/genasm_gpu.cu on the
// reference's link line together with -lscrooge_b200.
//
// Output layout = the reference's (src/genasm_gpu.cu:640-673, checked byte for byte by src/tests.cu:583-650): four
// bases per byte, base k of a byte in bits 7-2k:6-2k, A=0 C=1 G=2 T=3, case-insensitive, the last byte zero-padded;
// string i goes to twobit_strings[i], ceil(len/4) bytes.  This is NOT the layout the aligner uses internally
// (little-endian 16-base words, sg_dev_pack_2bit) -- it exists for source compatibility only.
//
// Mapping: the reference gives a string to a block and a byte to a thread with byte loads.  Here a block still takes
// strings grid-stride (the launch shape is the caller's: <<<32,32>>> and <<<256,32>>> in src/tests.cu), each thread
// builds whole output bytes from one 4-byte load when the source is aligned, byte loads otherwise and for the tail.
#include <cassert>
#include <cstdint>

namespace genasm_gpu {

namespace {

__device__ __forceinline__ unsigned code_of(unsigned c)
{
    // bits 2:1 of the letter: A 00, C 01, T 10, G 11 -> swap the last two to get A0 C1 G2 T3
    const unsigned u = c & 0xDFu;   // fold case
    assert(u == 'A' || u == 'C' || u == 'G' || u == 'T');   // the reference asserts too (src/genasm_gpu.cu:636)
    const unsigned v = (c >> 1) & 3u;
    return v ^ (v >> 1);
}

}  // namespace

__global__ void ascii_to_twobit_strings(int count, long long *string_lengths, char **ascii_strings, char **twobit_strings)
{
    for (int s = blockIdx.x; s < count; s += gridDim.x) {
        const long long len = string_lengths[s];
        const unsigned char *src = reinterpret_cast<const unsigned char *>(ascii_strings[s]);
        unsigned char *dst = reinterpret_cast<unsigned char *>(twobit_strings[s]);
        const long long full = len / 4;   // bytes that hold four bases
        const bool aligned = (reinterpret_cast<uintptr_t>(src) & 3u) == 0;
        for (long long q = threadIdx.x; q < full; q += blockDim.x) {
            unsigned c0, c1, c2, c3;
            if (aligned) {
                const unsigned w = *reinterpret_cast<const unsigned *>(src + 4 * q);
                c0 = w & 0xFFu; c1 = (w >> 8) & 0xFFu; c2 = (w >> 16) & 0xFFu; c3 = w >> 24;
            } else {
                c0 = src[4 * q]; c1 = src[4 * q + 1]; c2 = src[4 * q + 2]; c3 = src[4 * q + 3];
            }
            dst[q] = (unsigned char)((code_of(c0) << 6) | (code_of(c1) << 4) | (code_of(c2) << 2) | code_of(c3));
        }
        if (threadIdx.x == 0 && full * 4 < len) {   // 1..3 bases left: high bits first, the rest of the byte stays zero
            unsigned b = 0;
            for (long long k = 0; full * 4 + k < len; k++) b |= code_of(src[full * 4 + k]) << (6 - 2 * (int)k);
            dst[full] = (unsigned char)b;
        }
    }
}

}  // namespace genasm_gpu
