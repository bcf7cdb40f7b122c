// hostcopy.cpp -- the memcpy the copy threads of the host path run (api.cu: CopyPool), with non-temporal stores.
//
// A chunk of pageable memory is copied once into a pinned bounce buffer that the DMA engine then reads (and chunks of C
// come back the other way): the destination is never read by this core again, so ordinary stores only add a
// read-for-ownership of every destination line and push the source out of the cache.  Streaming stores write whole lines
// straight to memory.  TTV_B200_NT_COPY=0 keeps plain memcpy (A/B measurements).
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ttvb {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_nt_avx2(char* dst, const char* src, size_t bytes)   // dst 32-byte aligned
{
  size_t i = 0;
  for (; i + 128 <= bytes; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < bytes) std::memcpy(dst + i, src + i, bytes - i);
}
#endif

void host_copy(void* dst, const void* src, size_t bytes)
{
#if defined(__x86_64__)
  static const bool use_nt = [] {
    const char* e = std::getenv("TTV_B200_NT_COPY");
    return __builtin_cpu_supports("avx2") && !(e && *e == '0');
  }();
  if (use_nt && bytes >= 16384) {
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
    if (head) { std::memcpy(d, s, head); d += head; s += head; bytes -= head; }
    copy_nt_avx2(d, s, bytes);
    return;
  }
#endif
  std::memcpy(dst, src, bytes);
}

} // namespace ttvb
