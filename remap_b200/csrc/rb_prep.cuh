// rb_prep.cuh -- the two small streaming kernels in front of the pipelined matcher (rb_kpm_fast.cuh):
//
//   K0  rb_pack_kernel   frames as uploaded (one colour per byte, src/nil.hpp:14-31) -> the packed
//                        4 bit/pixel frame store that K2 stages through TMA.  Run once per upload.
//   K1c rb_list_kernel   K1's keypoint bit maps -> per (frame, region) POSITION LISTS, i.e. the
//                        membership that kpr::grid::add establishes in the reference
//                        (src/kpr.hpp:189-219; region index = grid_h * colsect + rowsect,
//                        src/kpr.hpp:71-74), with the region's weight counts
//                        (kpr::region::add, src/kpr.hpp:121-124).
//
// List layout: row (frame * nreg + region) of `lists` has `cap` entries x | w2 << 15 | y << 16: the
// weight-2 keypoints in entries [0, n_w2) and the weight-1 keypoints in entries (cap - 1 - m), m = 0 ..
// n_w1 - 1, so that a pair whose weight switch (src/kpm.hpp:219-220) selects weight-2 codes only reads
// a prefix.  counts[frame * nreg + region] = (n_all, n_w2).  When n_all > cap the row is incomplete; the
// matcher sees that in `counts` and hands such a region to the general kernel (rb_kpm.cuh), which works
// from the bit maps.
#pragma once

#include "rb_common.cuh"

#if defined(__CUDACC__)

// 16 pixels per thread: one 128-bit load, one 64-bit store.
__global__ void __launch_bounds__(256) rb_pack_kernel(const uint8_t* __restrict__ src, uint32_t pitch, uint64_t frame_stride,
                                                      uint8_t* __restrict__ dst, uint32_t pitch4, uint64_t frame_stride4,
                                                      uint32_t H, uint32_t nframes) {
  const uint32_t cpr = pitch / 16;  // chunks per row
  const uint64_t total = (uint64_t)nframes * H * cpr;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(i % cpr);
    const uint64_t fy = i / cpr;
    const uint32_t y = (uint32_t)(fy % H);
    const uint64_t f = fy / H;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + f * frame_stride + (uint64_t)y * pitch + 16 * c));
    uint32_t a = v.x & 0x0F0F0F0Fu, b = v.y & 0x0F0F0F0Fu, d = v.z & 0x0F0F0F0Fu, e = v.w & 0x0F0F0F0Fu;
    a = (a | (a >> 4)) & 0x00FF00FFu; a = (a | (a >> 8)) & 0xFFFFu;
    b = (b | (b >> 4)) & 0x00FF00FFu; b = (b | (b >> 8)) & 0xFFFFu;
    d = (d | (d >> 4)) & 0x00FF00FFu; d = (d | (d >> 8)) & 0xFFFFu;
    e = (e | (e >> 4)) & 0x00FF00FFu; e = (e | (e >> 8)) & 0xFFFFu;
    if (8 * c + 8 <= pitch4)
      *reinterpret_cast<uint2*>(dst + f * frame_stride4 + (uint64_t)y * pitch4 + 8 * c) = make_uint2(a | (b << 16), d | (e << 16));
  }
}

namespace rbl {

// bits of strip word j (bit i <-> x = 28 j + i, outputs at bits 2..29) that lie in [X0, X1)
__device__ __forceinline__ uint32_t colmask(uint32_t X0, uint32_t X1, uint32_t j) {
  const int lo = (int)X0 - (int)(RB_STRIP_OUT * j), hi = (int)X1 - (int)(RB_STRIP_OUT * j);
  uint32_t m = 0x3FFFFFFCu;
  if (lo > 2) m &= ~((1u << lo) - 1u);
  if (hi < 30) m &= (1u << (hi < 0 ? 0 : hi)) - 1u;
  return m;
}

}  // namespace rbl

// One warp per (frame, region), one pass over the region's bit-map words.  Lanes are laid out as
// (row within the chunk, strip), floor(32 / nstr) rows per chunk, so a lane's strip and column mask never
// change and its row advances by a constant.  Weight-2 keypoints are written forward from entry 0, weight-1
// keypoints backward from entry cap - 1; with n_all <= cap the two blocks never meet.
__global__ void __launch_bounds__(256) rb_list_kernel(const RbGeom g, const uint32_t* __restrict__ kpbits,
                                                      const uint32_t* __restrict__ w2bits, uint32_t first_frame,
                                                      uint32_t nframes, uint32_t cap, uint32_t* __restrict__ lists,
                                                      uint2* __restrict__ counts) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t nitems = nframes * g.nreg;
  for (uint32_t item = blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < nitems; item += gridDim.x * warps_per_block) {
    const uint32_t frame = first_frame + item / g.nreg, region = item % g.nreg;
    const uint32_t cs = region / g.grid_h, rs = region % g.grid_h;
    const uint32_t X0 = g.col0[cs], X1 = g.col1[cs], Y0 = g.row0[rs], Y1 = g.row1[rs];
    const uint32_t j0 = (X0 - 2) / RB_STRIP_OUT, j1 = (X1 - 1 - 2) / RB_STRIP_OUT, nstr = j1 - j0 + 1;
    // lanes = (row within the chunk, strip): floor(32 / nstr) rows per chunk, fixed for the whole region
    const uint32_t rpc = nstr <= 32 ? 32u / nstr : 0u;
    const uint32_t r0 = nstr <= 32 ? lane / nstr : 0u, k = lane - r0 * nstr;
    const uint32_t nrows = Y1 - Y0;
    const uint32_t* kp = kpbits + ((uint64_t)frame * g.H + Y0) * g.NS + j0 + k;
    const uint32_t* w2 = w2bits + ((uint64_t)frame * g.H + Y0) * g.NS + j0 + k;
    uint32_t* out = lists + ((uint64_t)frame * g.nreg + region) * cap;
    uint32_t b2 = 0, b1 = 0;
    if (rpc != 0) {
      const uint32_t m = r0 < rpc ? rbl::colmask(X0, X1, j0 + k) : 0u;
      const uint32_t xb = RB_STRIP_OUT * (j0 + k);
      uint32_t row = r0;
      uint32_t kw = 0, ww = 0;
      if (m != 0 && row < nrows) { kw = __ldg(kp + (uint64_t)row * g.NS); ww = __ldg(w2 + (uint64_t)row * g.NS); }
      for (uint32_t ra = 0; ra < nrows; ra += rpc) {
        // prefetch the next chunk's words before working on this one
        const uint32_t nrow = row + rpc;
        uint32_t nkw = 0, nww = 0;
        if (m != 0 && nrow < nrows) { nkw = __ldg(kp + (uint64_t)nrow * g.NS); nww = __ldg(w2 + (uint64_t)nrow * g.NS); }
        kw &= m; ww &= m;
        const uint32_t c = __popc(ww) | (__popc(kw & ~ww) << 16);  // weight-2 / weight-1 counts in one word
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if ((int)lane >= o) incl += t;
        }
        const uint32_t excl = incl - c, tot = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t at2 = b2 + (excl & 0xFFFFu), at1 = cap - 1 - (b1 + (excl >> 16));  // weight 1 runs backward
        const uint32_t yv = ((Y0 + row) << 16) + xb;
        while (kw) {
          const uint32_t b = __ffs((int)kw) - 1;
          kw &= kw - 1;
          const bool is2 = (ww >> b) & 1u;
          const uint32_t at = is2 ? at2 : at1;
          if (at < cap) out[at] = (yv + b) | (is2 ? 0x8000u : 0u);  // at1 wraps far above cap when the row is full
          at2 += is2 ? 1u : 0u;
          at1 -= is2 ? 0u : 1u;
        }
        b2 += tot & 0xFFFFu;
        b1 += tot >> 16;
        kw = nkw; ww = nww; row = nrow;
      }
    } else {
      b2 = b1 = cap + 1;  // a region wider than 32 strips: leave it to the general kernel
    }
    if (lane == 0) counts[(uint64_t)frame * g.nreg + region] = make_uint2(b2 + b1, b2);
  }
}

#endif  // __CUDACC__
