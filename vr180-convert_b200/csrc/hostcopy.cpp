// hostcopy.cpp -- the byte mover of the host-buffer pipeline's copy threads (csrc/pipeline.cu).
//
// Replaces nothing in the reference (its arrays never leave host memory); it is what stands between a caller's pageable
// NumPy arrays (remapper.py:373-379 cv.imread results, :388-398 the frames handed to cv.remap) and the page-locked ring
// the GPU's copy engines read.  Packing is bandwidth bound with ~8 threads at work: a plain memcpy of 1 MB units stays
// below glibc's non-temporal threshold, so every destination line is first READ into the cache (write-allocate) before
// it is overwritten -- three DRAM transfers per byte moved.  The destination here is never read by the CPU again (the DMA
// engine reads it), so large copies use streaming stores: two transfers per byte.
// Compiled by g++ with -mavx2 for this file only; the caller checks the CPU at run time (VR180_NT_COPY=0 disables it).
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace vr180 {

void stream_copy_avx2(uint8_t* d, const uint8_t* s, size_t n) {
#if !defined(__x86_64__)
    memcpy(d, s, n);  // (never selected: the caller's run-time check is x86 only)
#else
    size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;  // streaming stores need a 32-byte aligned destination
    if (head > n) head = n;
    if (head) {
        memcpy(d, s, head);
        d += head;
        s += head;
        n -= head;
    }
    const size_t blocks = n / 128;
    for (size_t i = 0; i < blocks; ++i) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 96), e);
        s += 128;
        d += 128;
    }
    _mm_sfence();  // the DMA engine, not this core, is the next reader
    const size_t tail = n - blocks * 128;
    if (tail) memcpy(d, s, tail);
#endif
}

}  // namespace vr180
