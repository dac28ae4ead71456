// rb_blit.cuh -- map assembly: fgm::fragment::blit for all frames of a fragment + fgm::fragment::blend.
//
// The reference keeps, per fragment, a map of dots (16 x uint16 colour histograms per map pixel,
// src/fgm.hpp:12-14) and adds every frame into it one by one: ++dots[pos + xy][colour]
// (src/fgm.hpp:87-97, blit_impl :176-188).  blend() then takes, per map pixel, the first colour with
// the largest count, and marks pixels that were ever covered (src/fgm.hpp:115-135).
//
// Scatter-adding 71,680 increments per frame would be ~10^9 atomics per pass.  Here the sum is GATHERED
// instead: one thread owns one map pixel, one CTA owns a 32 x 8 tile of the map, and the CTA walks
// every frame placed over its tile; a thread reads "its" pixel of that frame (neighbouring threads read
// neighbouring bytes) and bumps its private histogram in shared memory.  No atomics, every frame pixel
// is read exactly once, and the 16-bit counters wrap exactly like the reference's uint16.
#pragma once

#include "rb_common.cuh"

// One frame of the resident store placed on the map: (x, y) = position of its top-left pixel inside
// the map, i.e. fgm::frame::position_ - fragment zero (src/fgm.hpp:33-38,177).  Mirrors rb_placement.
struct RbPlacement {
  uint32_t frame;
  int32_t x, y;
};

#if defined(__CUDACC__)

#define RB_BLIT_TX 32
#define RB_BLIT_TY 8
#define RB_BLIT_NT (RB_BLIT_TX * RB_BLIT_TY)
#define RB_BLIT_CHUNK 1024

// MASKED = true is fdf::filter's blit (src/fdf.hpp:64, src/fgm.hpp:71-85): a frame pixel counts only where the
// frame's foreground mask is 0; fgbits = [placement][H][NW] bit maps written by rb_fg.cuh (bit set = foreground).
template <bool MASKED>
__global__ void __launch_bounds__(RB_BLIT_NT) rb_blit_blend_kernel(const uint8_t* __restrict__ frames, uint32_t pitch,
                                                                   uint64_t frame_stride, uint32_t W, uint32_t H,
                                                                   const RbPlacement* __restrict__ places, uint32_t n,
                                                                   uint32_t mapW, uint32_t mapH, uint16_t* __restrict__ dots,
                                                                   uint8_t* __restrict__ image, uint8_t* __restrict__ mask,
                                                                   const uint32_t* __restrict__ fgbits, uint32_t NW) {
  __shared__ uint16_t hist[16][RB_BLIT_NT];
  __shared__ RbPlacement hit[RB_BLIT_CHUNK];
  __shared__ uint32_t hit_idx[MASKED ? RB_BLIT_CHUNK : 1];
  __shared__ uint32_t nhit;
  const uint32_t tid = threadIdx.x;
  const uint32_t tx0 = blockIdx.x * RB_BLIT_TX, ty0 = blockIdx.y * RB_BLIT_TY;
  const uint32_t mx = tx0 + (tid & (RB_BLIT_TX - 1)), my = ty0 + tid / RB_BLIT_TX;
#pragma unroll
  for (int c = 0; c < 16; ++c) hist[c][tid] = 0;
  for (uint32_t base = 0; base < n; base += RB_BLIT_CHUNK) {
    if (tid == 0) nhit = 0;
    __syncthreads();
    // which of the next placements cover this tile at all
    for (uint32_t i = base + tid; i < n && i < base + RB_BLIT_CHUNK; i += RB_BLIT_NT) {
      const RbPlacement p = places[i];
      const bool over = p.x < (int32_t)(tx0 + RB_BLIT_TX) && p.x + (int32_t)W > (int32_t)tx0 &&
                        p.y < (int32_t)(ty0 + RB_BLIT_TY) && p.y + (int32_t)H > (int32_t)ty0;
      if (over) {  // order does not matter: the increments commute
        const uint32_t at = atomicAdd(&nhit, 1u);
        hit[at] = p;
        if (MASKED) hit_idx[at] = i;
      }
    }
    __syncthreads();
    const uint32_t nh = nhit;
    // One frame pixel (and, masked, one mask word) per hit: four hits' loads are issued before any histogram (eight: more registers, fewer CTAs, no gain)
    // update -- the updates are shared-memory stores the compiler will not move loads across, and a single load
    // in flight per thread left this loop latency-bound (1.3 pixels per cycle and SM).
    auto pixel_of = [&](uint32_t j, bool& take) -> uint32_t {
      const RbPlacement p = hit[j];
      const uint32_t fx = mx - (uint32_t)p.x, fy = my - (uint32_t)p.y;  // unsigned: negative wraps far above W / H
      take = fx < W && fy < H;
      uint32_t c = 0, m = 0;
      if (take) {
        c = frames[(uint64_t)p.frame * frame_stride + (uint64_t)fy * pitch + fx];
        if (MASKED) m = __ldg(fgbits + ((uint64_t)hit_idx[j] * H + fy) * NW + (fx >> 5)) >> (fx & 31);
      }
      if (MASKED && (m & 1u)) take = false;
      return c & 15u;
    };
    uint32_t j = 0;
    for (; j + 4 <= nh; j += 4) {
      bool t0, t1, t2, t3;
      const uint32_t c0 = pixel_of(j, t0), c1 = pixel_of(j + 1, t1), c2 = pixel_of(j + 2, t2), c3 = pixel_of(j + 3, t3);
      if (t0) ++hist[c0][tid];  // uint16: wraps at 65,536 like fgm::dot_type (src/fgm.hpp:14,94)
      if (t1) ++hist[c1][tid];
      if (t2) ++hist[c2][tid];
      if (t3) ++hist[c3][tid];
    }
    for (; j < nh; ++j) {
      bool t0;
      const uint32_t c0 = pixel_of(j, t0);
      if (t0) ++hist[c0][tid];
    }
    __syncthreads();
  }
  if (mx < mapW && my < mapH) {
    const uint64_t at = (uint64_t)my * mapW + mx;
    uint32_t best = 0, bestc = 0, packed[8];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint32_t v = hist[c][tid];
      if (v > best) { best = v; bestc = c; }  // std::max_element: the FIRST largest (src/fgm.hpp:127)
      if (c & 1) packed[c >> 1] |= v << 16; else packed[c >> 1] = v;
    }
    if (dots) {
      uint4* d = reinterpret_cast<uint4*>(dots + at * 16);
      d[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      d[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    }
    if (image) image[at] = (uint8_t)(best ? bestc : 0);  // src/fgm.hpp:128-131
    if (mask) mask[at] = best ? 1 : 0;
  }
}

#endif  // __CUDACC__
