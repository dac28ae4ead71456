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
// K1c list layout: row (frame * nreg + region) of `lists` has `cap` entries x | w2 << 15 | y << 16: the
// weight-2 keypoints in entries [0, n_w2) and the weight-1 keypoints in entries (cap - 1 - m), m = 0 ..
// n_w1 - 1, so that a pair whose weight switch (src/kpm.hpp:219-220) selects weight-2 codes only reads
// a prefix.  counts[frame * nreg + region] = (n_all, n_w2).  When n_all > cap the row is incomplete; the
// matcher sees that in `counts` and hands such a region to the general kernel (rb_kpm.cuh), which works
// from the bit maps.
#pragma once

#include "rb_common.cuh"

// ---- K1c: one warp per (frame, region), one pass over the region's bit-map words ---------------
// Lanes are laid out as (row within the chunk, strip), floor(32 / nstr) rows per chunk, so a lane's
// strip and column mask never change and its row advances by a constant.  Per chunk: every lane
// counts the keypoints of its word, one warp scan turns the counts into write positions, every lane
// emits its word.  Weight-2 keypoints are written forward from entry 0, weight-1 keypoints backward
// from entry cap - 1; with n_all <= cap the two blocks never meet.
// The lane functions are host+device so that tests/emul can run them lane by lane.
struct RbListLane {
  uint32_t Y0, nrows, j, nstr, rows_per_chunk, row0;  // the lane's word of chunk c: row row0 + c * rows_per_chunk
  uint32_t mask;                                      // column mask of its strip (0: idle lane)
};

namespace rbl {

// bits of strip word j (bit i <-> x = 28 j + i, outputs at bits 2..29) that lie in [X0, X1)
RB_HD uint32_t colmask(uint32_t X0, uint32_t X1, uint32_t j) {
  const int lo = (int)X0 - (int)(RB_STRIP_OUT * j), hi = (int)X1 - (int)(RB_STRIP_OUT * j);
  if (lo >= 30 || hi <= 2) return 0u;
  uint32_t m = 0x3FFFFFFCu;
  if (lo > 2) m &= ~((1u << lo) - 1u);
  if (hi < 30) m &= (1u << hi) - 1u;
  return m;
}

RB_HD RbListLane lane_setup(const RbGeom& g, uint32_t region, uint32_t lane) {
  RbListLane L;
  const uint32_t cs = region / g.grid_h, rs = region % g.grid_h;
  const uint32_t X0 = g.col0[cs], X1 = g.col1[cs];
  L.Y0 = g.row0[rs]; L.nrows = g.row1[rs] - g.row0[rs];
  const uint32_t j0 = (X0 - 2) / RB_STRIP_OUT, j1 = (X1 - 1 - 2) / RB_STRIP_OUT;
  L.nstr = j1 - j0 + 1;
  L.mask = 0; L.j = j0; L.row0 = 0; L.rows_per_chunk = 0;
  if (L.nstr > 32) return L;  // wider than a warp: counts are forced above cap, the matcher defers
  L.rows_per_chunk = 32u / L.nstr;
  const uint32_t r = lane / L.nstr, k = lane - r * L.nstr;
  if (r >= L.rows_per_chunk) return L;
  L.j = j0 + k;
  L.row0 = r;
  L.mask = colmask(X0, X1, L.j);
  return L;
}

// the lane's masked words of the chunk that starts at region row `ra`
RB_HD void lane_words(const RbGeom& g, const RbListLane& L, const uint32_t* kpf, const uint32_t* w2f, uint32_t ra,
                      uint32_t& kw, uint32_t& ww) {
  kw = ww = 0;
  const uint32_t row = ra + L.row0;
  if (L.mask && row < L.nrows) {
    const uint64_t at = (uint64_t)(L.Y0 + row) * g.NS + L.j;
    kw = kpf[at] & L.mask;
    ww = w2f[at] & L.mask;
  }
}

// at2 / n1_before: weight-2 / weight-1 keypoints of the region before this lane's word
RB_HD void lane_emit(const RbListLane& L, uint32_t ra, uint32_t kw, uint32_t ww, uint32_t at2, uint32_t n1_before,
                     uint32_t cap, uint32_t* out) {
  uint32_t at1 = cap - 1 - n1_before;  // wraps far above cap when the row is full
  const uint32_t yv = ((L.Y0 + ra + L.row0) << 16) + RB_STRIP_OUT * L.j;
  uint32_t w2 = ww, w1 = kw & ~ww;  // two short loops beat one loop that selects per bit
  // (and beat one warp-lockstep loop that takes a weight-2 and a weight-1 bit per trip: 0.65 vs 0.72 ms, r1o)
  while (w2) {
    const uint32_t b = rb_ffs0(w2);
    w2 &= w2 - 1;
    if (at2 < cap) out[at2] = (yv + b) | 0x8000u;
    ++at2;
  }
  while (w1) {
    const uint32_t b = rb_ffs0(w1);
    w1 &= w1 - 1;
    if (at1 < cap) out[at1] = yv + b;
    --at1;
  }
}

}  // namespace rbl

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

// K0u: the inverse of rb_pack_kernel, for frames that crossed PCIe already packed
// (rb_register_host_async): 16 pixels per thread, one 64-bit load, one 128-bit store.
__global__ void __launch_bounds__(256) rb_unpack_kernel(const uint8_t* __restrict__ src, uint32_t pitch4, uint64_t frame_stride4,
                                                        uint8_t* __restrict__ dst, uint32_t pitch, uint64_t frame_stride,
                                                        uint32_t H, uint32_t nframes) {
  const uint32_t cpr = pitch / 16;  // chunks per row
  const uint64_t total = (uint64_t)nframes * H * cpr;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(i % cpr);
    const uint64_t fy = i / cpr;
    const uint32_t y = (uint32_t)(fy % H);
    const uint64_t f = fy / H;
    uint2 v = make_uint2(0, 0);
    if (8 * c + 8 <= pitch4) v = __ldg(reinterpret_cast<const uint2*>(src + f * frame_stride4 + (uint64_t)y * pitch4 + 8 * c));
    auto spread = [](uint32_t h) {  // 4 nibbles (16 bits) -> 4 bytes
      uint32_t r = (h | (h << 8)) & 0x00FF00FFu;
      return (r | (r << 4)) & 0x0F0F0F0Fu;
    };
    uint4 o;
    o.x = spread(v.x & 0xFFFFu); o.y = spread(v.x >> 16); o.z = spread(v.y & 0xFFFFu); o.w = spread(v.y >> 16);
    *reinterpret_cast<uint4*>(dst + f * frame_stride + (uint64_t)y * pitch + 16 * c) = o;
  }
}

__global__ void __launch_bounds__(256) rb_list_kernel(const RbGeom g, const uint32_t* __restrict__ kpbits,
                                                      const uint32_t* __restrict__ w2bits, uint32_t first_frame,
                                                      uint32_t nframes, uint32_t cap, uint32_t* __restrict__ lists,
                                                      uint2* __restrict__ counts) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t nitems = nframes * g.nreg;
  for (uint32_t item = blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < nitems; item += gridDim.x * warps_per_block) {
    const uint32_t frame = first_frame + item / g.nreg, region = item % g.nreg;
    const RbListLane L = rbl::lane_setup(g, region, lane);
    const uint32_t* kpf = kpbits + (uint64_t)frame * g.H * g.NS;
    const uint32_t* w2f = w2bits + (uint64_t)frame * g.H * g.NS;
    uint32_t* out = lists + ((uint64_t)frame * g.nreg + region) * cap;
    uint32_t b2 = 0, b1 = 0;
    if (L.rows_per_chunk != 0) {
      uint32_t kw, ww, nkw, nww;  // this chunk's words and the next one's; the one after that is requested below
      rbl::lane_words(g, L, kpf, w2f, 0, kw, ww);
      rbl::lane_words(g, L, kpf, w2f, L.rows_per_chunk, nkw, nww);
      for (uint32_t ra = 0; ra < L.nrows; ra += L.rows_per_chunk) {
        uint32_t nnkw, nnww;  // two chunks ahead: the emit loops below are short, one chunk of lead did not cover the loads
        rbl::lane_words(g, L, kpf, w2f, ra + 2 * L.rows_per_chunk, nnkw, nnww);
        const uint32_t c = __popc(ww) | (__popc(kw & ~ww) << 16);  // weight-2 / weight-1 counts in one word
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if ((int)lane >= o) incl += t;
        }
        const uint32_t excl = incl - c, tot = __shfl_sync(0xffffffffu, incl, 31);
        rbl::lane_emit(L, ra, kw, ww, b2 + (excl & 0xFFFFu), b1 + (excl >> 16), cap, out);
        b2 += tot & 0xFFFFu;
        b1 += tot >> 16;
        kw = nkw; ww = nww;
        nkw = nnkw; nww = nnww;
      }
    } else {
      b2 = b1 = cap + 1;  // not listed: forces the matcher to defer this region
    }
    if (lane == 0) counts[(uint64_t)frame * g.nreg + region] = make_uint2(b2 + b1, b2);
  }
}

#endif  // __CUDACC__

// ---- aws::details::compare (src/aws.hpp:37-60) over a run of resident frames ------------------------
// The action-window scan ANDs "this pixel did not change" over consecutive frames into a heat map
// (heat &= prev == curr, one call per frame, src/aws.hpp:121).  Here one pass over frames [0, n): a thread
// owns 16 pixels, streams them through all n frames (each frame byte is read once, 128-bit loads) and
// records the index of the FIRST pair that differs per pixel; heat after k pairs is then
// first_change >= k, so the caller can replay the reference's per-frame states without another pass.
// heat (in/out, 1 byte per pixel like aws::heatmap_type) is cleared where any pair differs;
// first_change[p] = index i of the first pair (i, i + 1) with frame[i][p] != frame[i + 1][p], 0xFFFFFFFF if
// none.  HBM-bound: 1 byte read per pixel and frame.
RB_HD uint32_t rb_diff_bytes(uint32_t a, uint32_t b) {  // 0x80 in every byte that differs
  const uint32_t v = a ^ b;
  return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}

#if defined(__CUDACC__)
// blockIdx.y = segment of the pair range: pairs [seg * seg_len, (seg + 1) * seg_len); segments combine through
// atomicMin on first_change (initialised to 0xFFFFFFFF by the caller), so a long run of frames fills the GPU
// even though one frame has only a few thousand 16-pixel chunks.
__global__ void __launch_bounds__(256) rb_aws_compare_kernel(const uint8_t* __restrict__ frames, uint32_t pitch, uint64_t frame_stride,
                                                             uint32_t W, uint32_t H, uint32_t n, uint32_t seg_len,
                                                             uint8_t* __restrict__ heat, uint32_t* __restrict__ first_change) {
  const uint32_t chunks = pitch / 16, total = chunks * H;
  const uint32_t i0 = blockIdx.y * seg_len;
  const uint32_t i1 = i0 + seg_len < n - 1 ? i0 + seg_len : n - 1;  // pairs [i0, i1)
  if (i0 >= i1) return;
  for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
    const uint32_t y = it / chunks, x = (it - y * chunks) * 16;
    const uint8_t* src = frames + (uint64_t)y * pitch + x;
    uint4 prev = __ldg(reinterpret_cast<const uint4*>(src + (uint64_t)i0 * frame_stride));
    uint32_t fc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) fc[k] = 0xFFFFFFFFu;
    uint32_t seen[4] = {0, 0, 0, 0};  // bytes that have already changed
    auto step = [&](const uint4& cur, uint32_t i) {
      const uint32_t d[4] = {rb_diff_bytes(prev.x, cur.x), rb_diff_bytes(prev.y, cur.y), rb_diff_bytes(prev.z, cur.z),
                             rb_diff_bytes(prev.w, cur.w)};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t fresh = d[q] & ~seen[q];
        if (fresh) {  // rare after the first few frames
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (fresh & (0x80u << (8 * b))) fc[4 * q + b] = i;
          seen[q] |= fresh;
        }
      }
      prev = cur;
    };
    uint32_t i = i0;
    for (; i + 4 <= i1; i += 4) {  // four independent loads in flight
      const uint8_t* q = src + (uint64_t)(i + 1) * frame_stride;
      const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(q));
      const uint4 c1 = __ldg(reinterpret_cast<const uint4*>(q + frame_stride));
      const uint4 c2 = __ldg(reinterpret_cast<const uint4*>(q + 2 * frame_stride));
      const uint4 c3 = __ldg(reinterpret_cast<const uint4*>(q + 3 * frame_stride));
      step(c0, i); step(c1, i + 1); step(c2, i + 2); step(c3, i + 3);
    }
    for (; i < i1; ++i) step(__ldg(reinterpret_cast<const uint4*>(src + (uint64_t)(i + 1) * frame_stride)), i);
    if (seen[0] | seen[1] | seen[2] | seen[3]) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (x + k < W && fc[k] != 0xFFFFFFFFu) {
          const uint64_t at = (uint64_t)y * W + x + k;
          atomicMin(first_change + at, fc[k]);
          if (heat) heat[at] = 0;
        }
      }
    }
  }
}
#endif
