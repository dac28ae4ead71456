// rb_common.cuh -- shared definitions for the remap_b200 CUDA kernels (sm_100a).
//
// The kernel bodies are written as RB_HD (host+device) code over plain pointers so that the very
// same source can be compiled for the host by the unit-test harness in tests/emul/ (g++ -x c++);
// that harness is test infrastructure only and is never part of libremap_b200.so.
#pragma once

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#define RB_D __device__ __forceinline__
#define RB_CONSTEXPR_HD __host__ __device__ constexpr
#else
#define RB_HD inline
#define RB_D inline
#define RB_CONSTEXPR_HD constexpr
#endif

#define RB_MAX_REGIONS 32   // grid_w * grid_h limit (reference uses 4 x 2, src/frc.hpp:22-23)
#define RB_STRIP_OUT 28     // output pixels per 32-bit strip word (2-pixel halo on each side)
#define RB_STRIP_HALO 2     // == kpe::kernel_half (src/kpe.hpp:17)

// Result flags of one consecutive-frame pair (rb_offset.flags in include/remap_b200.h)
#define RB_OFFSET_VALID 1u
#define RB_OFFSET_TIE_SENSITIVE 2u

// One region's ballot for one pair: what kpm::details::cast_vote returns (src/kpm.hpp:213-223) plus
// the statistics the tie analysis needs.  Mirrors ro_region_vote in oracle/remap_oracle.c.
struct RbBin {
  int32_t dx, dy;
  uint32_t cnt;
};
struct RbRegionVote {
  uint32_t use_all;
  uint32_t n_prev, n_curr;
  uint32_t w2_prev, w2_curr;
  uint32_t nbins;
  uint32_t nticket;
  RbBin ticket[4];
  uint32_t ngt[4];
  uint32_t nge[4];
  uint32_t hist_hash;  // order-independent digest of the whole offset histogram: sum of rb_bin_hash over its bins
};

// Result of one pair, mirrors ro_match_result.
struct RbPairResult {
  int32_t dx, dy;
  uint32_t valid;
  uint32_t tie_sensitive;
  uint32_t active;
  int32_t top_dx[2], top_dy[2];
  uint32_t top_score[2];
  uint32_t ntop;
};

// Geometry shared by all kernels of one context.
struct RbGeom {
  uint32_t W, H;             // frame size in pixels
  uint32_t pitch;            // bytes between frame rows in HBM (multiple of 16)
  uint64_t frame_stride;     // bytes between frames
  uint32_t NS;               // strips per row: ceil((W - 4) / 28)
  uint32_t mpitch;           // median row pitch; pixel x lives at byte x + 2 (4-byte aligned strips)
  uint64_t median_stride;
  uint32_t grid_w, grid_h, nreg;
  uint32_t weight_switch, region_votes;
  // section s covers columns [col0[s], col1[s]) / rows [row0[s], row1[s])   
  uint32_t col0[8], col1[8], row0[8], row1[8];
};

// Digest of one offset-histogram bin; a region's hist_hash is the wrapping 32-bit SUM over its bins, so it does not
// depend on the order bins are met in.  The full-sequence parity runs compare it with the same sum over the
// reference's totalizator_t (oracle/ref_harness.cpp, digest mode).
RB_HD uint32_t rb_bin_hash(int32_t dx, int32_t dy, uint32_t cnt) {
  uint32_t h = ((uint32_t)dx & 0xFFFFu) | ((uint32_t)dy << 16);
  h = h * 0x9E3779B1u ^ cnt * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

RB_HD uint32_t rb_maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }
RB_HD uint32_t rb_xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }

RB_HD uint32_t rb_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  uint64_t v = (uint64_t)a | ((uint64_t)b << 32);
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}

RB_HD uint32_t rb_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}

RB_HD uint32_t rb_ffs0(uint32_t x) {  // index of lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
  return (uint32_t)(__ffs((int)x) - 1);
#else
  return (uint32_t)__builtin_ctz(x);
#endif
}

// Luminance-ordered colour LUTs of the reference (src/cpl.hpp:163-217; values: SURVEY.md a1).  The
// parity tests check them against the oracle's literal restatement of the generator.
#define RB_N2O_LIST 0, 15, 2, 12, 6, 9, 3, 13, 5, 1, 7, 4, 8, 14, 10, 11
#define RB_O2N_LIST 0, 9, 2, 6, 11, 8, 4, 10, 12, 5, 14, 15, 3, 7, 13, 1

// Truth table of output bit `bit` of a 16-entry LUT as a 16-bit mask over the 4-bit input.
RB_CONSTEXPR_HD uint32_t rb_lut_tt(const int (&lut)[16], int bit) {
  uint32_t m = 0;
  for (int i = 0; i < 16; ++i) m |= (uint32_t)((lut[i] >> bit) & 1) << i;
  return m;
}

// 3-input bitwise function by truth table: result bit = LUT[(a << 2) | (b << 1) | c]  (PTX lop3).
template <uint32_t LUT>
RB_HD uint32_t rb_lop3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return r;
#else
  uint32_t r = 0;
  for (int m = 0; m < 8; ++m)
    if ((LUT >> m) & 1)
      r |= ((m & 4) ? a : ~a) & ((m & 2) ? b : ~b) & ((m & 1) ? c : ~c);
  return r;
#endif
}

// Bit-sliced evaluation of a 4-input boolean function given by truth table TT (bit index
// x3*8 + x2*4 + x1*2 + x0): mux on x3 of two 3-input functions = 3 LOP3.
template <uint32_t TT>
RB_HD uint32_t rb_bool4(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
  const uint32_t g0 = rb_lop3<(TT & 0xFF)>(x2, x1, x0);
  const uint32_t g1 = rb_lop3<((TT >> 8) & 0xFF)>(x2, x1, x0);
  return rb_lop3<0xCA>(x3, g1, g0);  // x3 ? g1 : g0
}

// ---- warp-lockstep loops -------------------------------------------------------------------------
// A data-dependent per-lane loop (one trip per set bit of the lane's word, say) lets the lanes of a warp drift
// apart, and they do not reconverge by themselves: every sub-group then issues its own copy of the loop body
// (measured: 2 to 8 active lanes per issued instruction).  while (RB_WARP_ANY(cond)) { if (cond) { ... } } makes
// all lanes take the same number of trips: a warp vote on the device (all 32 lanes must reach it), plain `cond`
// in the host build.
#if defined(__CUDA_ARCH__)
#define RB_WARP_ANY(c) (__any_sync(0xFFFFFFFFu, (c)) != 0)
#else
#define RB_WARP_ANY(c) (c)
#endif

// ---- block-level execution helpers ------------------------------------------------------------
// Block kernels are written as a sequence of PHASES separated by barriers.  Code inside
// RB_FOR_THREADS runs once per thread; code outside it must be block-uniform (it only reads shared
// memory written in earlier phases).  On the device a phase is the calling thread itself followed
// by __syncthreads(); the host test build (tests/emul) runs the threads of a phase one after
// another, which is one legal interleaving of the same program.
#if defined(__CUDA_ARCH__)
#define RB_FOR_THREADS(tid, NT) for (uint32_t tid = threadIdx.x, rb_once_ = 1; rb_once_; rb_once_ = 0)
#define RB_SYNC() __syncthreads()
RB_D uint32_t rb_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
RB_D uint32_t rb_atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
RB_D void rb_atomic_max64(unsigned long long* p, unsigned long long v) { atomicMax(p, v); }
RB_D uint32_t rb_volatile_load(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
RB_D uint32_t rb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
#else
#define RB_FOR_THREADS(tid, NT) for (uint32_t tid = 0; tid < (NT); ++tid)
#define RB_SYNC() ((void)0)
inline uint32_t rb_atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
inline uint32_t rb_atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
inline void rb_atomic_max64(unsigned long long* p, unsigned long long v) { if (v > *p) *p = v; }
inline uint32_t rb_volatile_load(const uint32_t* p) { return *p; }
inline uint32_t rb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31));
}
#endif
