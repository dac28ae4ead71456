// rb_hostpack.cpp -- host side of rb_register_host_async: frames as the caller holds them (one
// colour per byte, src/nil.hpp:14-31) -> the packed 4 bit/pixel rows of the device frame store,
// written into pinned staging memory.  This is data marshalling for the PCIe copy (half the bytes
// cross the bus), not a compute path: nothing of kpe / kpm runs on the host.
#include "rb_hostpack.hpp"

#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__AVX2__)
#include <immintrin.h>
#endif

// one row: W pixels (bytes, low nibble significant) -> pitch4 bytes, pixel x in nibble x & 1 of byte x >> 1
template <bool STREAM>
static inline void pack_row(const uint8_t* src, uint32_t W, uint8_t* dst, uint32_t pitch4) {
  uint32_t x = 0;
#if defined(__AVX2__)
  const __m256i lo = _mm256_set1_epi8(0x0F), w = _mm256_set1_epi16(0x1001);
  // 64 pixels -> one 32-byte store (STREAM: non-temporal)
  if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0)
    for (; x + 64 <= W; x += 64) {
      __m256i a = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + x)), lo);
      __m256i b = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + x + 32)), lo);
      a = _mm256_maddubs_epi16(a, w);
      b = _mm256_maddubs_epi16(b, w);
      __m256i v = _mm256_packus_epi16(a, b);               // per 128-bit half: a.half, b.half
      v = _mm256_permute4x64_epi64(v, 0xD8);               // a.lo a.hi b.lo b.hi
      if (STREAM) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + x / 2), v);
      else _mm256_store_si256(reinterpret_cast<__m256i*>(dst + x / 2), v);
    }
  for (; x + 32 <= W; x += 32) {
    __m256i v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + x)), lo);
    v = _mm256_maddubs_epi16(v, w);                       // 16-bit lanes: even pixel + 16 * odd pixel
    v = _mm256_packus_epi16(v, _mm256_setzero_si256());   // bytes, per 128-bit half
    v = _mm256_permute4x64_epi64(v, 0x08);                // halves' low quadwords side by side
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + x / 2), _mm256_castsi256_si128(v));
  }
#endif
  for (; x + 2 <= W; x += 2) dst[x / 2] = (uint8_t)((src[x] & 15) | ((src[x + 1] & 15) << 4));
  if (x < W) { dst[x / 2] = (uint8_t)(src[x] & 15); x += 2; }
  if (x / 2 < pitch4) memset(dst + x / 2, 0, pitch4 - x / 2);
}

void rb_hostpack_frames_cb(const uint8_t* frames, uint32_t W, uint32_t H, size_t n, uint8_t* dst, uint32_t pitch4, int threads,
                           void (*pre)(void*), void* arg) {
  const long long rows = (long long)n * H;
  // an explicit count: launchers such as torchrun export OMP_NUM_THREADS=1, which would serialise the packer
  int nt = threads > 0 ? threads : omp_get_num_procs();
  if (nt < 1) nt = 1;
  // Ordinary stores by default: the staging chunk stays in the last-level cache and the DMA engine reads it from
  // there (measured on the pool's Xeon hosts: 1.32 M frames/s against 0.97 M with non-temporal stores, which send
  // every packed byte to DRAM and back -- the host's DRAM bandwidth is what bounds the packer).  RB_PACK_STREAM=1
  // selects the non-temporal variant (hosts with a small cache).
  static const bool stream = getenv("RB_PACK_STREAM") && atoi(getenv("RB_PACK_STREAM")) != 0;
  // `pre` runs on the CALLING thread (the team's thread 0) while the others already pack: the caller's kernel
  // launches for the previous chunk cost it tens of microseconds per chunk that would otherwise idle the whole team.
  // Rows are handed out dynamically, so thread 0 simply joins late.
  static const bool sched_static = getenv("RB_PACK_STATIC") && atoi(getenv("RB_PACK_STATIC")) != 0;  // experiments
  if (sched_static) {
    if (pre) pre(arg);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long long r = 0; r < rows; ++r) {
      if (stream) pack_row<true>(frames + (size_t)r * W, W, dst + (size_t)r * pitch4, pitch4);
      else pack_row<false>(frames + (size_t)r * W, W, dst + (size_t)r * pitch4, pitch4);
    }
  } else {
#pragma omp parallel num_threads(nt)
    {
      if (pre && omp_get_thread_num() == 0) pre(arg);
      if (stream) {
#pragma omp for schedule(dynamic, 64)
        for (long long r = 0; r < rows; ++r) pack_row<true>(frames + (size_t)r * W, W, dst + (size_t)r * pitch4, pitch4);
      } else {
#pragma omp for schedule(dynamic, 64)
        for (long long r = 0; r < rows; ++r) pack_row<false>(frames + (size_t)r * W, W, dst + (size_t)r * pitch4, pitch4);
      }
    }
  }
#if defined(__AVX2__)
  _mm_sfence();  // streaming stores must be visible before the copy is queued
#endif
}

void rb_hostpack_frames(const uint8_t* frames, uint32_t W, uint32_t H, size_t n, uint8_t* dst, uint32_t pitch4, int threads) {
  rb_hostpack_frames_cb(frames, W, H, n, dst, pitch4, threads, nullptr, nullptr);
}
