// rb_api.cu -- libremap_b200.so: the C ABI of include/remap_b200.h over the sm_100a kernels.
//
// No host compute path exists in this library: every result comes out of a kernel below.  The only
// host work is geometry set-up, launches and copies.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <new>
#include <time.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/remap_b200.h"
#include "rb_host.hpp"
#include "rb_hostpack.hpp"
#include "rb_kpe.cuh"
#include "rb_kpm.cuh"
#include "rb_kpm_fast.cuh"
#include "rb_kpm_big.cuh"
#include "rb_prep.cuh"
#include "rb_blit.cuh"
#include "rb_fg.cuh"
#include "rb_splice.cuh"

static_assert(sizeof(rb_region_vote) == sizeof(RbRegionVote), "ABI mirror of RbRegionVote");
static_assert(sizeof(rb_bin) == sizeof(RbBin), "ABI mirror of RbBin");
static_assert(sizeof(rb_keypoint) == 24, "rb_keypoint layout");
static_assert(sizeof(rb_offset) == 12, "rb_offset layout");
static_assert(sizeof(rb_placement) == sizeof(RbPlacement), "ABI mirror of RbPlacement");

// ---- small kernels ------------------------------------------------------------------------------

// K3 wrapper: also emits the compact rb_offset the caller fetches.
__global__ void rb_declare_offsets_kernel(const RbGeom g, const RbRegionVote* votes, RbPairResult* results,
                                          rb_offset* offsets, uint32_t npairs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  RbPairResult r;
  rbm::declare_pair(g, votes + (uint64_t)i * g.nreg, &r);
  results[i] = r;
  rb_offset o;
  o.dx = r.dx;
  o.dy = r.dy;
  o.flags = (r.valid ? RB_OFFSET_VALID : 0u) | (r.tie_sensitive ? RB_OFFSET_TIE_SENSITIVE : 0u);
  offsets[i] = o;
}

// K3, cooperative: one GROUP of lanes per pair (one lane per region, group = next power of two), so a
// 20,000-pair launch fills the GPU instead of running four warps per SM over local-memory arrays.  Same
// arithmetic as rbm::declare_pair (rb_kpm.cuh; src/kpm.hpp:172-184,199-211 + the tie analysis of
// DESIGN.md section 2): every lane scores its own <= 3 ticket entries against all entries of the group
// (shuffles), the group reduces to the two best candidates and to the bounds the tie analysis needs.
__global__ void __launch_bounds__(256) rb_declare_group_kernel(const RbGeom g, const RbRegionVote* __restrict__ votes,
                                                               rb_offset* __restrict__ offsets, uint32_t npairs, uint32_t LP) {
  const uint32_t lane = threadIdx.x & 31, r = lane & (LP - 1), gbase = lane - r;
  const uint32_t pair = (blockIdx.x * blockDim.x + threadIdx.x) / LP;
  const bool live = pair < npairs && r < g.nreg;
  const uint32_t rv = g.region_votes;
  // my region's ballot
  uint32_t key[3] = {0, 0, 0}, info[3] = {0, 0, 0};  // info: valid | pts << 1 | hi << 3 | lo << 5 | ambiguous << 7
  uint32_t og = 0, ncurr = 0;
  if (live) {
    const RbRegionVote& v = votes[(uint64_t)pair * g.nreg + r];
    ncurr = v.n_curr;
    const uint32_t nt = v.nticket;
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k)
      if (k < nt && k < rv) {
        key[k] = ((uint32_t)(v.ticket[k].dx + 32768) << 16) | (uint32_t)(v.ticket[k].dy + 32768);
        const uint32_t worst = v.nge[k] - 1;
        const uint32_t hi = v.ngt[k] < rv ? rv - v.ngt[k] : 0u, lo = worst < rv ? rv - worst : 0u;
        info[k] = 1u | ((rv - k) << 1) | (hi << 3) | (lo << 5) | ((v.nge[k] != v.ngt[k] + 1 ? 1u : 0u) << 7);
      }
    if (nt == rv && v.nge[rv - 1] > rv) og = v.ngt[rv - 1] < rv ? rv - v.ngt[rv - 1] : 0u;
  }
  const uint32_t gmask = LP == 32 ? 0xffffffffu : (((1u << LP) - 1u) << gbase);
  const uint32_t active = __popc(__ballot_sync(0xffffffffu, live && ncurr > 0) & gmask);  // current grid only (src/kpm.hpp:400)
  const bool ambiguous = (__ballot_sync(0xffffffffu, ((info[0] | info[1] | info[2]) >> 7) & 1u) & gmask) != 0;
  // score / bounds of my entries against every entry of the group
  uint32_t score[3] = {0, 0, 0}, HI[3] = {0, 0, 0}, LO[3] = {0, 0, 0}, og_all = 0;
  bool rep[3] = {true, true, true};  // first occurrence of its offset in (region, rank) order
  for (uint32_t l = 0; l < LP; ++l) {
    const uint32_t og_l = __shfl_sync(0xffffffffu, og, gbase + l);
    og_all += og_l;
    bool found[3] = {false, false, false};
#pragma unroll
    for (uint32_t m = 0; m < 3; ++m) {
      const uint32_t okey = __shfl_sync(0xffffffffu, key[m], gbase + l), oinf = __shfl_sync(0xffffffffu, info[m], gbase + l);
#pragma unroll
      for (uint32_t k = 0; k < 3; ++k)
        if ((oinf & 1u) && okey == key[k]) {
          score[k] += (oinf >> 1) & 3u;
          HI[k] += (oinf >> 3) & 3u;
          LO[k] += (oinf >> 5) & 3u;
          found[k] = true;
          if (l < r || (l == r && m < k)) rep[k] = false;
        }
    }
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k)
      if (!found[k]) HI[k] += og_l;  // off that region's ticket: it can score there only by displacing a tied entry
  }
  auto gmax64 = [&](unsigned long long v) {
    for (uint32_t o = LP >> 1; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
      v = t > v ? t : v;
    }
    return v;
  };
  auto gmax32 = [&](uint32_t v) {
    for (uint32_t o = LP >> 1; o > 0; o >>= 1) {
      const uint32_t t = __shfl_xor_sync(0xffffffffu, v, o);
      v = t > v ? t : v;
    }
    return v;
  };
  // the two best candidates: score desc, dx asc, dy asc
  unsigned long long c0 = 0;
#pragma unroll
  for (uint32_t k = 0; k < 3; ++k)
    if (info[k] & 1u) {
      const unsigned long long c = ((unsigned long long)score[k] << 32) | (0xFFFFFFFFu - key[k]);
      c0 = c > c0 ? c : c0;
    }
  c0 = gmax64(c0);
  const uint32_t bkey = 0xFFFFFFFFu - (uint32_t)c0, S0 = (uint32_t)(c0 >> 32);
  unsigned long long c1 = 0;
#pragma unroll
  for (uint32_t k = 0; k < 3; ++k)
    if ((info[k] & 1u) && key[k] != bkey) {
      const unsigned long long c = ((unsigned long long)score[k] << 32) | (0xFFFFFFFFu - key[k]);
      c1 = c > c1 ? c : c1;
    }
  c1 = gmax64(c1);
  const uint32_t S1 = (uint32_t)(c1 >> 32), half = active / 2;  // src/kpm.hpp:206
  bool valid = false, tie = false;
  if (active >= g.nreg / 4 && c0 != 0) {  // src/kpm.hpp:401, :202-204
    valid = !(c1 != 0 && S0 < S1 + half);
    // tie sensitivity (see rbm::declare_pair)
    uint32_t hb = 0, ha = 0, lob = 0, l1c = 0;
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k)
      if (info[k] & 1u) {
        ha = HI[k] > ha ? HI[k] : ha;
        if (key[k] != bkey) hb = HI[k] > hb ? HI[k] : hb; else lob = LO[k];
        if (rep[k]) { const uint32_t cc = (LO[k] << 8) | (r * 3 + k + 1); l1c = cc > l1c ? cc : l1c; }
      }
    hb = gmax32(hb); ha = gmax32(ha); lob = gmax32(lob); l1c = gmax32(l1c);
    uint32_t l2 = 0;
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k)
      if ((info[k] & 1u) && rep[k] && ((LO[k] << 8) | (r * 3 + k + 1)) != l1c) l2 = LO[k] > l2 ? LO[k] : l2;
    l2 = gmax32(l2);
    if (ambiguous) {
      if (valid) {
        const uint32_t Hm = hb > og_all ? hb : og_all;
        tie = !(lob >= Hm + (half > 1 ? half : 1));
      } else {
        const uint32_t Hm = ha > og_all ? ha : og_all;
        tie = !(l2 >= 1 && Hm < l2 + half);
      }
    }
  }
  if (live && r == 0) {
    rb_offset o;
    o.dx = valid ? (int32_t)(bkey >> 16) - 32768 : 0;
    o.dy = valid ? (int32_t)(bkey & 0xFFFFu) - 32768 : 0;
    o.flags = (valid ? RB_OFFSET_VALID : 0u) | (tie ? RB_OFFSET_TIE_SENSITIVE : 0u);
    offsets[pair] = o;
  }
}

// Parity tap: expands K1's bit maps of one frame into rb_keypoint records in the reference's
// insertion order (column-major: x outer, y inner; src/kpe.hpp:201-204,289-305), with the 13-byte
// code laid out exactly as kpe::extractor::encode_keypoint does (src/kpe.hpp:342-379).
__global__ void rb_keypoints_kernel(const RbGeom g, const uint8_t* frame, const uint32_t* kpbits,
                                    const uint32_t* w2bits, rb_keypoint* out, uint32_t cap, uint32_t* count) {
  extern __shared__ uint32_t colstart[];  // [W + 1]
  const uint32_t W = g.W, H = g.H;
  for (uint32_t x = threadIdx.x; x < W; x += blockDim.x) {
    uint32_t n = 0;
    if (x >= 2 && x + 2 < W) {
      const uint32_t j = (x - 2) / RB_STRIP_OUT, b = x - RB_STRIP_OUT * j;
      for (uint32_t y = 2; y + 4 < H; ++y) n += (kpbits[(uint64_t)y * g.NS + j] >> b) & 1u;
    }
    colstart[x + 1] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    colstart[0] = 0;
    for (uint32_t x = 0; x < W; ++x) colstart[x + 1] += colstart[x];
    *count = colstart[W];
  }
  __syncthreads();
  for (uint32_t x = threadIdx.x; x < W; x += blockDim.x) {
    if (x < 2 || x + 2 >= W) continue;
    const uint32_t j = (x - 2) / RB_STRIP_OUT, b = x - RB_STRIP_OUT * j;
    uint32_t at = colstart[x];
    uint32_t colbits = 0;
    for (uint32_t s = 0; s < g.grid_w; ++s)
      if (x >= g.col0[s] && x < g.col1[s]) colbits |= 1u << s;
    for (uint32_t y = 2; y + 4 < H; ++y) {
      if (!((kpbits[(uint64_t)y * g.NS + j] >> b) & 1u)) continue;
      if (at < cap) {
        const uint32_t weight = ((w2bits[(uint64_t)y * g.NS + j] >> b) & 1u) ? 2u : 1u;
        uint8_t v[5][5];
        for (int r = 0; r < 5; ++r)
          for (int c = 0; c < 5; ++c) v[r][c] = frame[(uint64_t)(y - 2 + r) * g.pitch + (x - 2 + c)] & 15;
        rb_keypoint k;
        k.code[0] = v[0][0] | (v[0][1] << 4);  k.code[1] = v[0][2] | (v[0][3] << 4);
        k.code[2] = v[1][0] | (v[0][4] << 4);  k.code[3] = v[1][1] | (v[1][2] << 4);
        k.code[4] = v[1][3] | (v[1][4] << 4);  k.code[5] = v[2][0] | (v[2][1] << 4);
        k.code[6] = v[2][2] | (v[2][3] << 4);  k.code[7] = v[3][0] | (v[2][4] << 4);
        k.code[8] = v[3][1] | (v[3][2] << 4);  k.code[9] = v[3][3] | (v[3][4] << 4);
        k.code[10] = v[4][0] | (v[4][1] << 4); k.code[11] = v[4][2] | (v[4][3] << 4);
        k.code[12] = weight | (v[4][4] << 4);
        k.weight = (uint8_t)weight;
        k.x = (uint16_t)x;
        k.y = (uint16_t)y;
        uint32_t mask = 0;  // region idx = grid_h * colsect + rowsect (src/kpr.hpp:71-74)
        for (uint32_t a = 0; a < g.grid_w; ++a)
          if (colbits & (1u << a))
            for (uint32_t bb = 0; bb < g.grid_h; ++bb)
              if (y >= g.row0[bb] && y < g.row1[bb]) mask |= 1u << (g.grid_h * a + bb);
        k.region_mask = mask;
        out[at] = k;
      }
      ++at;
    }
  }
}

// fde::details::generate_mask (src/fde.hpp:19-55): byte-wise equality of the frame with the
// background window at linear offset idx, 0xFF where equal.  16 pixels per thread: one 128-bit
// frame load, five 32-bit background loads funnel-shifted to the frame's alignment, one 128-bit
// store; scalar tail.  HBM-bound: 2 bytes read + 1 byte written per pixel.
__global__ void rb_fgmask_kernel(const uint8_t* __restrict__ bg, long long idx, uint32_t bgW,
                                 const uint8_t* __restrict__ frame, uint32_t fpitch, uint8_t* __restrict__ mask,
                                 uint32_t mpitch, uint32_t W, uint32_t H) {
  const uint32_t chunks = (W + 15) / 16;
  const uint32_t total = chunks * H;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t y = i / chunks, x = (i % chunks) * 16;
    const uint8_t* f = frame + (uint64_t)y * fpitch + x;
    const uint8_t* b = bg + idx + (long long)y * bgW + x;
    uint8_t* m = mask + (uint64_t)y * mpitch + x;
    const bool vec = x + 16 <= W && ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(m)) & 15) == 0;
    if (vec) {
      const uint4 fv = __ldg(reinterpret_cast<const uint4*>(f));
      const uintptr_t ba = reinterpret_cast<uintptr_t>(b);
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(ba & ~(uintptr_t)3);
      const uint32_t sh = (uint32_t)(ba & 3) * 8;
      const uint32_t w0 = __ldg(bw), w1 = __ldg(bw + 1), w2 = __ldg(bw + 2), w3 = __ldg(bw + 3);
      const uint32_t w4 = sh ? __ldg(bw + 4) : 0u;
      uint4 o;
      o.x = __vcmpeq4(__funnelshift_r(w0, w1, sh), fv.x);
      o.y = __vcmpeq4(__funnelshift_r(w1, w2, sh), fv.y);
      o.z = __vcmpeq4(__funnelshift_r(w2, w3, sh), fv.z);
      o.w = __vcmpeq4(__funnelshift_r(w3, w4, sh), fv.w);
      *reinterpret_cast<uint4*>(m) = o;
    } else {
      for (uint32_t k = 0; k < 16 && x + k < W; ++k) m[k] = b[k] == f[k] ? 0xFF : 0x00;
    }
  }
}


// Parity tap for full-sequence comparisons (tests/digest_check.py against the compiled reference in digest mode): per frame,
// order-independent 64-bit digests of kpe's two outputs -- the median image and every (region, point, code)
// insertion kpe::extractor makes into the grid (src/kpe.hpp:225-229,301-303; code layout src/kpe.hpp:342-379).
//   median_hash = sum over pixels with value v != 0 at index i = y * W + x of splitmix(i << 8 | v)
//   kp_hash     = sum over insertions of splitmix((x | y << 16 | region << 32) ^ splitmix(code[0..7] ^ splitmix(code[8..12])))
__device__ __forceinline__ unsigned long long rb_splitmix(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) rb_digest_kernel(const RbGeom g, const uint8_t* __restrict__ frames, const uint8_t* __restrict__ median,
                                                        const uint32_t* __restrict__ kpbits, const uint32_t* __restrict__ w2bits,
                                                        uint32_t first, rb_frame_digest* __restrict__ out) {
  const uint32_t f = first + blockIdx.x, tid = threadIdx.x, W = g.W, H = g.H;
  unsigned long long mh = 0, kh = 0;
  uint32_t nk = 0, ni = 0;
  if (median) {
    const uint8_t* m = median + (uint64_t)f * g.median_stride + 2;  // pixel x at byte x + 2 of a row
    for (uint32_t i = tid; i < W * H; i += blockDim.x) {
      const uint32_t y = i / W, x = i - y * W;
      const uint32_t v = m[(uint64_t)y * g.mpitch + x];
      if (v) mh += rb_splitmix(((unsigned long long)i << 8) | v);
    }
  }
  const uint8_t* fr = frames + (uint64_t)f * g.frame_stride;
  const uint32_t* kpf = kpbits + (uint64_t)f * H * g.NS;
  const uint32_t* w2f = w2bits + (uint64_t)f * H * g.NS;
  for (uint32_t wi = tid; wi < H * g.NS; wi += blockDim.x) {
    const uint32_t y = wi / g.NS, j = wi - y * g.NS;
    uint32_t w = kpf[wi];
    const uint32_t w2 = w2f[wi];
    while (w) {
      const uint32_t b = (uint32_t)__ffs((int)w) - 1;
      w &= w - 1;
      const uint32_t x = RB_STRIP_OUT * j + b, weight = ((w2 >> b) & 1u) ? 2u : 1u;
      uint8_t v[25];
      for (int r = 0; r < 5; ++r)
        for (int cc = 0; cc < 5; ++cc) v[5 * r + cc] = fr[(uint64_t)(y - 2 + r) * g.pitch + (x - 2 + cc)] & 15;
      auto at = [&](int r, int cc) { return (unsigned long long)v[5 * r + cc]; };
      auto byte = [&](int lo_r, int lo_c, int hi_r, int hi_c) { return at(lo_r, lo_c) | (at(hi_r, hi_c) << 4); };
      const unsigned long long lo = byte(0, 0, 0, 1) | (byte(0, 2, 0, 3) << 8) | (byte(1, 0, 0, 4) << 16) | (byte(1, 1, 1, 2) << 24) |
                                    (byte(1, 3, 1, 4) << 32) | (byte(2, 0, 2, 1) << 40) | (byte(2, 2, 2, 3) << 48) | (byte(3, 0, 2, 4) << 56);
      const unsigned long long hi = byte(3, 1, 3, 2) | (byte(3, 3, 3, 4) << 8) | (byte(4, 0, 4, 1) << 16) | (byte(4, 2, 4, 3) << 24) |
                                    (((unsigned long long)weight | (at(4, 4) << 4)) << 32);
      const unsigned long long ch = rb_splitmix(lo ^ rb_splitmix(hi));
      ++nk;
      for (uint32_t a = 0; a < g.grid_w; ++a)
        if (x >= g.col0[a] && x < g.col1[a])
          for (uint32_t bb = 0; bb < g.grid_h; ++bb)
            if (y >= g.row0[bb] && y < g.row1[bb]) {
              const unsigned long long r = g.grid_h * a + bb;  // src/kpr.hpp:71-74
              kh += rb_splitmix(((unsigned long long)x | ((unsigned long long)y << 16) | (r << 32)) ^ ch);
              ++ni;
            }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mh += __shfl_down_sync(0xffffffffu, mh, o);
    kh += __shfl_down_sync(0xffffffffu, kh, o);
    nk += __shfl_down_sync(0xffffffffu, nk, o);
    ni += __shfl_down_sync(0xffffffffu, ni, o);
  }
  if ((tid & 31) == 0) {
    rb_frame_digest* d = out + blockIdx.x;
    if (mh) atomicAdd(reinterpret_cast<unsigned long long*>(&d->median_hash), mh);
    if (kh) atomicAdd(reinterpret_cast<unsigned long long*>(&d->kp_hash), kh);
    if (nk) atomicAdd(&d->keypoints, nk);
    if (ni) atomicAdd(&d->insertions, ni);
  }
}

// popcount of K1's keypoint bit maps (statistics for the roofline's K term)
__global__ void rb_count_kernel(const uint32_t* __restrict__ bits, size_t nwords, unsigned long long* total) {
  unsigned long long loc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
    loc += __popc(__ldg(bits + i));
  for (int o = 16; o > 0; o >>= 1) loc += __shfl_down_sync(0xffffffffu, loc, o);
  if ((threadIdx.x & 31) == 0 && loc) atomicAdd(total, loc);
}

// The general matcher (rb_kpm.cuh) over the (pair, region) list the pipelined matcher deferred.
__global__ void __launch_bounds__(1024) rb_kpm_deferred_kernel(const RbKpmParams p, const uint2* __restrict__ list,
                                                              const uint32_t* __restrict__ count, uint32_t cap,
                                                              uint32_t* __restrict__ total) {
  extern __shared__ __align__(16) uint32_t rb_kpm_smem[];
  uint32_t n = *count;
  if (n > cap) n = cap;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(total, n);  // statistics (rb_deferred_count)
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint2 it = list[i];
    rbm::kpm_block(p, it.x, it.y, rb_kpm_smem, blockDim.x);
    __syncthreads();
  }
}

// ---- context ------------------------------------------------------------------------------------
#define RB_MAX_BATCHES 16

// profile marks of one batch: K1 start / end on the extraction stream; matcher start, after K1c, after
// K2, after the deferred / general kernel, after K3 on the main stream
struct RbBatchEvents { cudaEvent_t e[7]; };


struct rb_ctx {
  rb_config cfg;
  RbGeom g;
  int device;
  int sm_count;
  cudaStream_t stream;
  bool own_stream;
  uint8_t* d_frames;
  uint8_t* d_frames4;    // packed 4 bit/pixel copy of the frame store (K0), read by the pipelined matcher
  uint32_t pitch4;
  uint64_t frame_stride4;
  uint32_t* d_lists;     // [frame][region][list_cap] keypoint positions (K1c)
  uint2* d_counts;       // [frame][region] (n_all, n_w2)
  uint2* d_deferred;     // [pair][region] worst case: what the first matcher pass deferred
  uint2* d_deferred2;    // ... and what the second, large-list pass deferred (general kernel)
  RbKpmFastParams big;   // rb_kpm_big_kernel: one CTA per SM, lists of thousands of entries.  Second pass over what the
  size_t big_smem;       //   first-pass kernel deferred (one pair per work item), or -- large regions -- THE matcher
  bool big_primary;
  uint32_t* d_work;      // [0] work-item counter, [1] deferred count
  CUtensorMap tmap;      // 3-D tensor map of d_frames4: (row bytes, rows, frames)
  RbKpmFastParams fast;  // geometry-dependent constants of the pipelined matcher
  size_t fast_smem;
  int fast_ctas_per_sm;
  uint8_t* d_median;
  uint32_t* d_kp;
  uint32_t* d_w2;
  RbRegionVote* d_votes;
  RbPairResult* d_results;
  rb_offset* d_offsets;
  RbBin* d_tap_bins;
  uint32_t* d_tap_count;
  rb_keypoint* d_kps;
  RbPlacement* d_places;  // rb_blit_blend scratch (grown on demand)
  size_t places_cap;
  uint8_t* d_map;         // dots (32 B / map pixel) + image + mask
  size_t map_cap;
  uint32_t map_w, map_h;  // geometry of the map the scratch holds (last rb_blit_blend / rb_filter_fragment)
  std::vector<std::pair<std::string, void*>> ipc_open;  // peers' map scratch mapped through CUDA IPC (handle bytes -> address)
  uint8_t* d_bg;       // scratch for rb_foreground_mask
  size_t bg_cap;
  uint8_t* d_fgframe;  // dense frame scratch
  // pass-2 filter (rb_filter_fragment)
  bool fg_ready, fg_fast_ok;
  uint32_t fg_NW, fg_rcap, fg_scap, fg_grid, fg_gen_rcap, fg_gen_grid, fg_gen_nt, fg_nt;
  size_t fg_smem, fg_gen_smem, fg_slab;
  uint32_t* d_fgbits;      // [placement][H][NW]
  uint32_t* d_fg_nkept;    // [placement]
  uint32_t* d_fg_deferred; // [placement]
  uint32_t* d_fg_count;    // [0] frames deferred by the shared-memory variant
  size_t fg_cap;           // placements the three arrays above hold
  uint8_t* d_fg_scratch;   // slabs of the general variant
  uint8_t* d_fg_bytes;     // parity tap: masks as bytes, one chunk
  cudaEvent_t fg_ev[4];
  float fg_ms[3];
  uint32_t fg_last_deferred;
  size_t med_lo, med_hi;   // frames whose medians K1 has written
  uint8_t* d_mask;     // dense mask [H][W]
  size_t bytes;
  uint32_t code_slots, off_slots, tile_pitch, tile_rows;
  size_t kpm_smem;
  uint32_t kpm_nt;       // threads per CTA of the general matcher: 256, or more when shared memory allows few CTAs per SM
  size_t uploaded;       // frames [0, uploaded) hold data
  size_t reg_first, reg_n;
  RbBatchEvents bev[RB_MAX_BATCHES];  // profile marks per batch of the last call
  size_t nbatch_used;
  cudaStream_t kpe_stream;   // K1 runs here, concurrently with the matcher of the previous batch
  cudaEvent_t ev_kpe[RB_MAX_BATCHES];
  cudaStream_t copy_stream;  // host -> device copies of rb_register_host_async
  cudaEvent_t ev_copy[2], ev_entry;
  uint8_t* h_stage[3];     // pinned staging: chunks of packed 4 bit/pixel frames (RB_STAGE_BUFS of them in use, default 3)
  uint32_t* h_flags;       // pinned: copy of d_work[0..3] fetched with the offsets ([2] = matcher error word)
  // rb_register_host_async: one "landed" event per chunk (timing enabled: the link rate is estimated from them)
  std::vector<cudaEvent_t> ev_chunk, ev_chunk_start;
  double link_Bps;         // estimated host -> device rate of this context's link (bytes / s)
  double pack_fps;         // estimated rate of the host packer (frames / s)
  uint64_t lane_raw, lane_packed;  // chunks of the last rb_register_host_async call that went raw / packed
  uint64_t lane_bytes;             // bytes that call put on the link
  int host_threads;        // packer threads of this context
  size_t stage_frames;
  uint64_t launches;
  bool debug_sync;
  bool k3_reference;
  std::string err;
};

#define RB_CUDA(ctx, call)                                                                      \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
      return RB_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// RB_DEBUG_SYNC=1 in the environment: wait after every launch so that a device fault names its kernel.
#define RB_LAUNCHED(ctx, name)                                                                  \
  do {                                                                                          \
    ++(ctx)->launches;                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                        \
    if (e_ == cudaSuccess && (ctx)->debug_sync) e_ = cudaStreamSynchronize((ctx)->stream);      \
    if (e_ != cudaSuccess) {                                                                    \
      (ctx)->err = std::string(name) + ": " + cudaGetErrorString(e_);                           \
      return RB_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

static uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

static std::atomic<int> g_live_contexts{0};  // contexts alive in this process (host packer thread share)

template <typename T>
static cudaError_t dmalloc(rb_ctx* c, T** p, size_t bytes) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
  if (e == cudaSuccess) c->bytes += bytes;
  return e;
}

extern "C" {

uint32_t rb_abi_version(void) { return RB_ABI_VERSION; }

void rb_default_config(rb_config* cfg, uint32_t width, uint32_t height, uint32_t max_frames) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->width = width;
  cfg->height = height;
  cfg->grid_w = 4;          // src/frc.hpp:22
  cfg->grid_h = 2;          // src/frc.hpp:23
  cfg->overlap = 16;        // src/frc.hpp:24
  cfg->weight_switch = 10;  // src/frc.hpp:32
  cfg->region_votes = 3;    // src/frc.hpp:33
  cfg->device = 0;
  cfg->max_frames = max_frames;
  cfg->compute_median = 1;
}

int rb_create(const rb_config* cfg, rb_ctx** out) {
  if (!cfg || !out) return RB_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return RB_ERR_NO_DEVICE;  // there is deliberately no CPU fallback
  }
  rb_ctx* c = new (std::nothrow) rb_ctx();
  if (!c) return RB_ERR_INVALID;
  c->cfg = *cfg;
  c->device = cfg->device;
  c->debug_sync = getenv("RB_DEBUG_SYNC") != nullptr;
  c->k3_reference = getenv("RB_K3_REFERENCE") != nullptr;
  ++g_live_contexts;
  *out = c;  // returned even on failure so that rb_last_error can be read; caller rb_destroy()s it
  if (cfg->max_frames < 2) { c->err = "max_frames must be >= 2"; return RB_ERR_INVALID; }
  if (rb_make_geom(cfg->width, cfg->height, cfg->grid_w, cfg->grid_h, cfg->overlap, cfg->weight_switch,
                   cfg->region_votes, &c->g) != 0) {
    c->err = "unsupported frame size / grid (need W/grid_w > overlap/2, H/grid_h > overlap/2, grid <= 8x8, W,H < 32768)";
    return RB_ERR_INVALID;
  }
  if (cfg->width >= 32768 || cfg->height >= 32768) { c->err = "frame too large"; return RB_ERR_INVALID; }
  const RbGeom& g = c->g;
  RB_CUDA(c, cudaSetDevice(c->device));
  RB_CUDA(c, cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
  if (cfg->stream) { c->stream = static_cast<cudaStream_t>(cfg->stream); c->own_stream = false; }
  else { RB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  for (int b = 0; b < RB_MAX_BATCHES; ++b) {
    for (int i = 0; i < 7; ++i) RB_CUDA(c, cudaEventCreate(&c->bev[b].e[i]));
    RB_CUDA(c, cudaEventCreateWithFlags(&c->ev_kpe[b], cudaEventDisableTiming));
  }
  {
    int lo = 0, hi = 0;  // numerically larger = lower priority: the matcher's CTAs get SM resources first
    RB_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    RB_CUDA(c, cudaStreamCreateWithPriority(&c->kpe_stream, cudaStreamNonBlocking, lo));
  }
  RB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) RB_CUDA(c, cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
  RB_CUDA(c, cudaEventCreateWithFlags(&c->ev_entry, cudaEventDisableTiming));
  RB_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_flags), 64, cudaHostAllocDefault));
  memset(c->h_flags, 0, 64);

  // K2 shared-memory budget
  uint32_t maxw = 0, maxh = 0, maxcols = 0;
  for (uint32_t s = 0; s < g.grid_w; ++s) {
    const uint32_t tx0 = (g.col0[s] - 2) & ~7u, tw = (g.col1[s] + 2 - tx0 + 7) / 8;
    if (tw > maxw) maxw = tw;
    if (g.col1[s] - g.col0[s] > maxcols) maxcols = g.col1[s] - g.col0[s];
  }
  for (uint32_t s = 0; s < g.grid_h; ++s)
    if (g.row1[s] - g.row0[s] > maxh) maxh = g.row1[s] - g.row0[s];
  c->tile_pitch = maxw + 1;
  c->tile_rows = maxh + 4;
  const uint32_t area = maxcols * maxh;
  c->code_slots = cfg->code_slots ? cfg->code_slots : next_pow2(area / 4 < 1024 ? 1024 : area / 4);
  c->off_slots = cfg->offset_slots ? cfg->offset_slots : 1024;
  if (c->code_slots > 8192 && !cfg->code_slots) c->code_slots = 8192;  // 12-bit list index in the code table
  if ((c->code_slots & (c->code_slots - 1)) || (c->off_slots & (c->off_slots - 1)) || c->code_slots < 64 ||
      c->code_slots > 8192 || c->code_slots / 2 < maxcols || c->off_slots <= 2 * 256) {
    c->err = "code_slots / offset_slots must be powers of two, 64 <= code_slots <= 8192, code_slots/2 >= region width, offset_slots > 512";
    return RB_ERR_INVALID;
  }
  int smem_max = 0;
  RB_CUDA(c, cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  for (;;) {
    RbKpmParams p;
    memset(&p, 0, sizeof(p));
    p.code_slots = c->code_slots; p.off_slots = c->off_slots; p.tile_pitch = c->tile_pitch; p.tile_rows = c->tile_rows;
    c->kpm_smem = rbm::smem_words(p, 256) * sizeof(uint32_t);
    if (c->kpm_smem <= (size_t)smem_max) break;
    if (!cfg->code_slots && c->code_slots / 2 >= 2 * maxcols && c->code_slots > 1024) { c->code_slots /= 2; continue; }
    c->err = "region tiles + hash tables exceed the shared memory of one CTA";
    return RB_ERR_INVALID;
  }
  // Large regions (640x480: 125 KB of tiles and tables) leave room for one CTA per SM; with 256 threads that is 8
  // warps per SM and the kernel crawls.  The body is written for any block size: give such a CTA the whole SM.
  c->kpm_nt = 256;
  {
    int smem_sm = 0;
    RB_CUDA(c, cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device));
    const size_t per_sm = (size_t)smem_sm / (c->kpm_smem + 1024);
    const uint32_t nt = per_sm <= 1 ? 1024u : per_sm == 2 ? 512u : 256u;
    if (nt != 256) {
      RbKpmParams q;
      memset(&q, 0, sizeof(q));
      q.code_slots = c->code_slots; q.off_slots = c->off_slots; q.tile_pitch = c->tile_pitch; q.tile_rows = c->tile_rows;
      const size_t need = rbm::smem_words(q, nt) * sizeof(uint32_t);
      if (need <= (size_t)smem_max) { c->kpm_nt = nt; c->kpm_smem = need; }
    }
  }
  RB_CUDA(c, cudaFuncSetAttribute(rb_kpm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->kpm_smem));
  RB_CUDA(c, cudaFuncSetAttribute(rb_kpm_deferred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->kpm_smem));

  const size_t N = cfg->max_frames;
  // ---- pipelined matcher: packed frame store, lists, TMA tensor map -------------------------------
  {
    c->pitch4 = rb_align_up((g.W + 1) / 2, 16);
    c->frame_stride4 = (uint64_t)c->pitch4 * g.H;
    RbKpmFastParams& f = c->fast;
    memset(&f, 0, sizeof(f));
    f.g = g;
    uint32_t bx = 0, by = 0;
    for (uint32_t s = 0; s < g.grid_w; ++s) {
      const uint32_t tx0 = (g.col0[s] - 2) & ~31u, bytes = (g.col1[s] + 2 - tx0 + 1) / 2;
      if (rb_align_up(bytes, 16) > bx) bx = rb_align_up(bytes, 16);
    }
    for (uint32_t s = 0; s < g.grid_h; ++s)
      if (g.row1[s] - g.row0[s] + 4 > by) by = g.row1[s] - g.row0[s] + 4;
    f.box_x = bx;
    f.nbox_y = (by + 255) / 256;
    f.box_y = (by + f.nbox_y - 1) / f.nbox_y;
    if (f.nbox_y > 1) f.box_y = rb_align_up(f.box_y, 8);  // every box must land 128-byte aligned in shared memory
    uint32_t dxbits = 0, dybits = 0;
    while ((1u << dxbits) <= 2 * g.W) ++dxbits;  // dx + W < 2 W, never all ones
    while ((1u << dybits) <= 2 * g.H) ++dybits;
    f.dybits = dybits;
    f.offbits = dxbits + dybits;
    const uint32_t cntmax = f.offbits < 32 ? (1u << (32 - f.offbits)) - 1u : 0u;
    uint32_t cap = cfg->list_cap ? cfg->list_cap : 1280;
    if (cap > 2047) cap = 2047;
    if (cap > cntmax) cap = cntmax;  // a bin's count never exceeds the shorter list
    cap &= ~3u;                      // 16-byte list rows
    f.cap = cap;
    f.lcap = 2 * cap;  // list rows hold all keypoints of a region even when only the weight-2 ones take part
    f.tslots = next_pow2(cap < 2048 ? (cap < 64 ? 64 : (cap > 1024 ? 2048 : 2 * cap)) : cap);  // chained buckets: load <= 1 is fine
    f.oslots = 1024;
    f.run = cfg->run_pairs ? cfg->run_pairs : 40;  // auto: 8 ... 40 pairs per work item, chosen per launch (pick_run); shared memory sized for 40
    int smem_max = 0;
    RB_CUDA(c, cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    c->fast_smem = rbf::smem_bytes(f);
    const bool tma_ok = bx <= 256 && f.box_y <= 256 && g.W < 16384 && g.H < 4096 && f.run <= 1024;
    bool fast_ok = tma_ok && cap >= 16 && c->fast_smem <= (size_t)smem_max && f.offbits <= 24 && maxcols * maxh < 65536 &&
                   (g.W << dybits) < (1u << 28);
    // ---- the large-region matcher (rb_kpm_big.cuh): region-relative offsets, 6 bytes per keypoint and frame ------
    RbKpmFastParams& B = c->big;
    B = f;
    bool big_ok = tma_ok;
    {
      uint32_t lxmax = 0;
      for (uint32_t s = 0; s < g.grid_w; ++s) {
        const uint32_t tx0 = (g.col0[s] - 2) & ~31u;
        if (g.col1[s] - 3 - tx0 > lxmax) lxmax = g.col1[s] - 3 - tx0;
      }
      uint32_t dxb = 0, dyb = 0;
      while ((1u << dxb) <= 2 * maxcols) ++dxb;
      while ((1u << dyb) <= 2 * maxh) ++dyb;
      B.dybits = dyb;
      B.offbits = dxb + dyb;
      B.bias_x = maxcols;
      B.bias_y = maxh;
      B.epoch0 = getenv("RB_BIG_EPOCH0") ? (uint32_t)atoi(getenv("RB_BIG_EPOCH0")) : 1u;  // tests: start near the 16-bit wrap
      big_ok = big_ok && lxmax <= 255 && maxh <= 256 && B.offbits <= 19;  // 16-bit positions; >= 13 bits of count per bin
      const uint32_t cntmax_b = (1u << (32 - B.offbits)) - 1u;
      B.oslots = 1024;  // 2 x 6 KB; a region whose pair has more distinct offsets (scene cuts, heavy repetition) is deferred
      B.cap = 0;
      // the largest bucket table whose lists still hold what a dense frame puts into a region (~ a keypoint per 8 pixels;
      // BASELINE configs[3] has one per 9): short chains matter more than the last thousand entries of capacity --
      // the bucket walk's trip count is the longest chain among a warp's 32 keypoints
      const uint32_t want = cfg->list_cap ? cfg->list_cap : maxcols * maxh / 8;
      for (uint32_t ts = 8192; big_ok && ts >= 1024; ts >>= 1) {
        B.tslots = ts;
        B.cap = 0;
        const size_t fixed = rbb::smem_bytes(B) + 16;
        if (fixed >= (size_t)smem_max) continue;
        uint32_t fit = (uint32_t)(((size_t)smem_max - fixed) / 12);
        if (fit > 65532) fit = 65532;
        if (fit > cntmax_b) fit = cntmax_b;
        if (fit > 8 * ts) fit = 8 * ts;
        fit &= ~3u;
        B.cap = fit;
        if (fit >= want || ts == 1024) break;
      }
      if (B.cap < 64) big_ok = false;
    }
    // Which kernel takes the regions first: rb_kpm_fast_kernel (two CTAs per SM, full codes in shared memory) where
    // typical region lists fit its 1,280 entries; rb_kpm_big_kernel where a sparse frame already overflows them.
    // kpm_mode: 0 = by region size, 1 = general kernel only, 2 = large-region kernel first, 3 = fast kernel first.
    c->big_primary = big_ok && (cfg->kpm_mode == 2 || (cfg->kpm_mode == 0 && maxcols * maxh > 24000) || !fast_ok);
    if (cfg->kpm_mode == 3 && fast_ok) c->big_primary = false;
    if (!fast_ok && !big_ok) c->cfg.kpm_mode = 1;  // geometry outside both pipelined matchers' limits
    if (c->cfg.kpm_mode != 1) {
      c->cfg.kpm_mode = 0;
      if (c->big_primary) {
        B.lcap = B.cap;  // one row length for K1c and the matcher
        f.lcap = B.lcap;
      } else {
        if (B.cap > f.lcap) B.cap = f.lcap & ~3u;  // second pass: bounded by the list rows K1c writes
        B.lcap = f.lcap;
        if (B.cap <= f.cap) big_ok = false;        // nothing to gain
      }
      c->big_smem = big_ok ? rbb::smem_bytes(B) : 0;
      if (c->big_smem)
        RB_CUDA(c, cudaFuncSetAttribute(rb_kpm_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->big_smem));
      if (!c->big_primary) {
        RB_CUDA(c, cudaFuncSetAttribute(rb_kpm_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->fast_smem));
        int occ = 0;
        RB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rb_kpm_fast_kernel, RB_FAST_NT, c->fast_smem));
        c->fast_ctas_per_sm = occ < 1 ? 1 : occ;
      }
      RB_CUDA(c, dmalloc(c, &c->d_frames4, c->frame_stride4 * N + 256));
      RB_CUDA(c, dmalloc(c, &c->d_lists, (size_t)N * g.nreg * f.lcap * 4 + 256));
      RB_CUDA(c, dmalloc(c, &c->d_counts, (size_t)N * g.nreg * sizeof(uint2)));
      RB_CUDA(c, dmalloc(c, &c->d_deferred, (size_t)N * g.nreg * sizeof(uint2)));
      RB_CUDA(c, dmalloc(c, &c->d_deferred2, (size_t)N * g.nreg * sizeof(uint2)));
      RB_CUDA(c, dmalloc(c, &c->d_work, 256));
      RB_CUDA(c, cudaMemsetAsync(c->d_frames4, 0, c->frame_stride4 * N + 256, c->stream));
      RB_CUDA(c, cudaMemsetAsync(c->d_work, 0, 256, c->stream));
      typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      RB_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess) { c->err = "cuTensorMapEncodeTiled unavailable"; return RB_ERR_CUDA; }
      const cuuint64_t dims[3] = {c->pitch4, g.H, N};
      const cuuint64_t strides[2] = {c->pitch4, c->frame_stride4};
      const cuuint32_t box[3] = {f.box_x, f.box_y, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      const CUresult r = reinterpret_cast<encode_fn>(fn)(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->d_frames4, dims, strides,
                                                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { c->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return RB_ERR_CUDA; }
    }
  }
  RB_CUDA(c, dmalloc(c, &c->d_frames, g.frame_stride * N + 256));
  RB_CUDA(c, dmalloc(c, &c->d_kp, (size_t)N * g.H * g.NS * 4));
  RB_CUDA(c, dmalloc(c, &c->d_w2, (size_t)N * g.H * g.NS * 4));
  RB_CUDA(c, dmalloc(c, &c->d_votes, (size_t)N * g.nreg * sizeof(RbRegionVote)));
  RB_CUDA(c, dmalloc(c, &c->d_results, (size_t)N * sizeof(RbPairResult)));
  RB_CUDA(c, dmalloc(c, &c->d_offsets, (size_t)N * sizeof(rb_offset)));
  RB_CUDA(c, dmalloc(c, &c->d_tap_bins, ((size_t)1 << 20) * sizeof(RbBin)));
  RB_CUDA(c, dmalloc(c, &c->d_tap_count, 256));
  RB_CUDA(c, dmalloc(c, &c->d_kps, (size_t)g.W * g.H * sizeof(rb_keypoint)));
  RB_CUDA(c, dmalloc(c, &c->d_fgframe, (size_t)g.W * g.H + 256));
  RB_CUDA(c, dmalloc(c, &c->d_mask, (size_t)g.W * g.H + 256));
  RB_CUDA(c, cudaMemsetAsync(c->d_frames, 0, g.frame_stride * N + 256, c->stream));
  RB_CUDA(c, cudaMemsetAsync(c->d_kp, 0, (size_t)N * g.H * g.NS * 4, c->stream));
  RB_CUDA(c, cudaMemsetAsync(c->d_w2, 0, (size_t)N * g.H * g.NS * 4, c->stream));
  if (cfg->compute_median) {
    RB_CUDA(c, dmalloc(c, &c->d_median, g.median_stride * N + 256));
    // rows/columns outside the keypoint domain are never written and must read 0 (src/frc.hpp:104)
    RB_CUDA(c, cudaMemsetAsync(c->d_median, 0, g.median_stride * N + 256, c->stream));
  }
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

void rb_destroy(rb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_frames4); cudaFree(c->d_lists); cudaFree(c->d_counts); cudaFree(c->d_deferred); cudaFree(c->d_deferred2); cudaFree(c->d_work);
  cudaFree(c->d_frames); cudaFree(c->d_median); cudaFree(c->d_kp); cudaFree(c->d_w2); cudaFree(c->d_votes);
  cudaFree(c->d_results); cudaFree(c->d_offsets); cudaFree(c->d_tap_bins); cudaFree(c->d_tap_count);
  cudaFree(c->d_kps); cudaFree(c->d_places); cudaFree(c->d_map); cudaFree(c->d_bg); cudaFree(c->d_fgframe); cudaFree(c->d_mask);
  for (auto& kv : c->ipc_open) cudaIpcCloseMemHandle(kv.second);
  cudaFree(c->d_fgbits); cudaFree(c->d_fg_nkept); cudaFree(c->d_fg_deferred); cudaFree(c->d_fg_count); cudaFree(c->d_fg_scratch);
  cudaFree(c->d_fg_bytes);
  for (int i = 0; i < 4; ++i)
    if (c->fg_ev[i]) cudaEventDestroy(c->fg_ev[i]);
  for (int b = 0; b < RB_MAX_BATCHES; ++b) {
    for (int i = 0; i < 7; ++i)
      if (c->bev[b].e[i]) cudaEventDestroy(c->bev[b].e[i]);
    if (c->ev_kpe[b]) cudaEventDestroy(c->ev_kpe[b]);
  }
  if (c->kpe_stream) { cudaStreamSynchronize(c->kpe_stream); cudaStreamDestroy(c->kpe_stream); }
  for (int i = 0; i < 2; ++i)
    if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
  if (c->ev_entry) cudaEventDestroy(c->ev_entry);
  for (int i = 0; i < 3; ++i)
    if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  for (auto e : c->ev_chunk) cudaEventDestroy(e);
  for (auto e : c->ev_chunk_start) cudaEventDestroy(e);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  --g_live_contexts;
  delete c;
}

const char* rb_last_error(rb_ctx* c) { return c ? c->err.c_str() : "null context"; }
const char* rb_matcher_kernel(rb_ctx* c) {
  if (!c) return "";
  return c->cfg.kpm_mode == 1 ? "rb_kpm_kernel" : c->big_primary ? "rb_kpm_big_kernel" : "rb_kpm_fast_kernel";
}
const rb_offset* rb_offsets_device(rb_ctx* c) { return c ? c->d_offsets + c->reg_first : nullptr; }

int rb_count_keypoints(rb_ctx* c, size_t first, size_t n, uint64_t* total) {
  if (!c || !total) return RB_ERR_INVALID;
  if (first + n > c->cfg.max_frames) return RB_ERR_CAPACITY;
  RB_CUDA(c, cudaSetDevice(c->device));
  unsigned long long* d = reinterpret_cast<unsigned long long*>(c->d_tap_count + 2);  // 8-byte aligned scratch
  RB_CUDA(c, cudaMemsetAsync(d, 0, 8, c->stream));
  const size_t nwords = n * c->g.H * c->g.NS;
  rb_count_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->d_kp + first * c->g.H * c->g.NS, nwords, d);
  RB_LAUNCHED(c, "rb_count_kernel");
  unsigned long long h = 0;
  RB_CUDA(c, cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  *total = h;
  return RB_OK;
}
void* rb_stream(rb_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }
uint64_t rb_kernel_launches(rb_ctx* c) { return c ? c->launches : 0; }
size_t rb_device_bytes(rb_ctx* c) { return c ? c->bytes : 0; }

int rb_deferred_count(rb_ctx* c, uint32_t* count) {
  if (!c || !count) return RB_ERR_INVALID;
  *count = 0;
  if (!c->d_work) return RB_OK;
  RB_CUDA(c, cudaSetDevice(c->device));
  uint32_t w[4] = {0, 0, 0, 0};
  RB_CUDA(c, cudaMemcpyAsync(w, c->d_work, 16, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  *count = w[1];
  if (w[2]) { c->err = "pipelined matcher: a TMA transaction timed out"; return RB_ERR_CUDA; }
  return RB_OK;
}

void* rb_alloc_host(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void rb_free_host(void* p) {
  if (p) cudaFreeHost(p);
}

int rb_synchronize(rb_ctx* c) {
  if (!c) return RB_ERR_INVALID;
  RB_CUDA(c, cudaSetDevice(c->device));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

// host -> frame store, on `stream`
static int copy_frames(rb_ctx* c, const uint8_t* frames, size_t first, size_t n, cudaStream_t stream) {
  const RbGeom& g = c->g;
  uint8_t* dst = c->d_frames + g.frame_stride * first;
  if (g.pitch == g.W)
    RB_CUDA(c, cudaMemcpyAsync(dst, frames, (size_t)g.W * g.H * n, cudaMemcpyHostToDevice, stream));
  else
    RB_CUDA(c, cudaMemcpy2DAsync(dst, g.pitch, frames, g.W, g.W, (size_t)g.H * n, cudaMemcpyHostToDevice, stream));
  return RB_OK;
}

// K0: the packed 4 bit/pixel copy of frames [first, first + n) that the pipelined matcher reads
static int pack_frames(rb_ctx* c, size_t first, size_t n) {
  if (!c->d_frames4 || n == 0) return RB_OK;
  const RbGeom& g = c->g;
  const uint64_t chunks = (uint64_t)n * g.H * (g.pitch / 16);
  uint64_t blocks = (chunks + 255) / 256;
  const uint64_t maxb = (uint64_t)c->sm_count * 16;
  if (blocks > maxb) blocks = maxb;
  rb_pack_kernel<<<(uint32_t)blocks, 256, 0, c->stream>>>(c->d_frames + g.frame_stride * first, g.pitch, g.frame_stride,
                                                         c->d_frames4 + c->frame_stride4 * first, c->pitch4, c->frame_stride4,
                                                         g.H, (uint32_t)n);
  RB_LAUNCHED(c, "rb_pack_kernel");
  return RB_OK;
}

int rb_upload(rb_ctx* c, const uint8_t* frames, size_t first, size_t n) {
  if (!c || !frames) return RB_ERR_INVALID;
  if (first + n > c->cfg.max_frames) { c->err = "rb_upload: beyond max_frames"; return RB_ERR_CAPACITY; }
  if (n == 0) return RB_OK;
  RB_CUDA(c, cudaSetDevice(c->device));
  int rc = copy_frames(c, frames, first, n, c->stream);
  if (rc != RB_OK) return rc;
  rc = pack_frames(c, first, n);
  if (rc != RB_OK) return rc;
  if (first + n > c->uploaded) c->uploaded = first + n;
  return RB_OK;
}

int rb_upload_medians(rb_ctx* c, const uint8_t* medians, size_t first, size_t n) {
  if (!c || !medians) return RB_ERR_INVALID;
  if (!c->d_median) { c->err = "rb_upload_medians: context created with compute_median = 0"; return RB_ERR_STATE; }
  if (first + n > c->cfg.max_frames) { c->err = "rb_upload_medians: beyond max_frames"; return RB_ERR_CAPACITY; }
  if (n == 0) return RB_OK;
  const RbGeom& g = c->g;
  RB_CUDA(c, cudaSetDevice(c->device));
  // pixel x lives at byte x + 2 of a median row (the inverse of rb_fetch_medians)
  RB_CUDA(c, cudaMemcpy2DAsync(c->d_median + g.median_stride * first + 2, g.mpitch, medians, g.W, g.W, (size_t)g.H * n,
                               cudaMemcpyHostToDevice, c->stream));
  if (c->med_hi == c->med_lo) { c->med_lo = first; c->med_hi = first + n; }
  else {
    if (first < c->med_lo) c->med_lo = first;
    if (first + n > c->med_hi) c->med_hi = first + n;
  }
  return RB_OK;
}

// Picks the number of row segments per strip: enough CTAs for a few waves on every SM without
// paying too many 4-row warm-ups.
static uint32_t pick_segments(const rb_ctx* c, size_t n) {
  const RbGeom& g = c->g;
  const uint32_t rows = g.H - 6;
  const double slots = (double)c->sm_count * 4 /*CTAs of 128 threads per SM at 128 regs*/ * 128;
  uint32_t best = 1;
  double best_eff = 0;
  for (uint32_t s = 1; s <= 32 && rows / s >= 8; ++s) {
    const double items = (double)n * g.NS * s;
    const double waves = items / slots;
    const double eff = waves / (double)(size_t)(waves + 0.999999) * (double)rows / (rows + 4.0 * s);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

// K1 on frames [f0, f0 + kn), on stream `st`
static int launch_kpe(rb_ctx* c, size_t f0, size_t kn, cudaStream_t st) {
  const RbGeom& g = c->g;
  RbKpeParams p;
  p.g = g;
  p.frames = c->d_frames + g.frame_stride * f0;
  p.median = c->d_median ? c->d_median + g.median_stride * f0 : nullptr;
  p.kpbits = c->d_kp + (size_t)f0 * g.H * g.NS;
  p.w2bits = c->d_w2 + (size_t)f0 * g.H * g.NS;
  p.nframes = (uint32_t)kn;
  p.nseg = pick_segments(c, kn);
  p.seg_rows = (g.H - 6 + p.nseg - 1) / p.nseg;
  const size_t items = kn * p.nseg * g.NS;
  const uint32_t blocks = (uint32_t)((items + 127) / 128);
  rb_kpe_kernel<<<blocks, 128, 0, st>>>(p);
  ++c->launches;
  RB_CUDA(c, cudaGetLastError());
  return RB_OK;
}

// The matcher + K3 for the n - 1 pairs of frames [first, first + n), on the main stream.  Frames
// [list_first, first + n) still need their region lists (K1c).  Ballots, results and offsets are stored at
// the pair's absolute index (= index of its first frame).
// Pairs per work item of a pipelined matcher launch.  The CTAs take (region, run) items off one queue; the launch ends
// when the last item does, so what counts is whole WAVES of items: ceil(items / CTAs) waves of run + 1 steps each (a run
// of r pairs walks r + 1 frames).  Among 8 ... 40 pairs per run (longer runs would leave too few items to even out the
// regions' different sizes), the cheapest in those terms: 4,999 pairs of 8 regions on 148 CTAs: 34 pairs -> 148 runs ->
// exactly 8 waves of 35 steps, against 9 waves of 33 with 32; a 512-frame chunk of the host path: 10 pairs -> 416 items
// on 296 CTAs instead of 128.
static uint32_t pick_run(const rb_ctx* c, uint32_t npairs, uint32_t nreg, uint32_t ctas) {
  if (c->cfg.run_pairs) return c->cfg.run_pairs;
  uint32_t best = 32;
  double best_cost = 1e300;
  for (uint32_t r = 8; r <= 40; ++r) {
    const uint32_t runs = (npairs + r - 1) / r, items = runs * nreg, waves = (items + ctas - 1) / ctas;
    const double cost = (double)waves * (r + 1);
    if (cost < best_cost * 0.999) { best_cost = cost; best = r; }
  }
  return best;
}

static int launch_match(rb_ctx* c, size_t first, size_t n, size_t list_first, cudaEvent_t* ev) {
  const RbGeom& g = c->g;
  if (ev) RB_CUDA(c, cudaEventRecord(ev[2], c->stream));
  if (n >= 2) {
    RbKpmParams p;
    memset(&p, 0, sizeof(p));
    p.g = g;
    p.frames = c->d_frames;
    p.kpbits = c->d_kp;
    p.w2bits = c->d_w2;
    p.votes = c->d_votes + first * g.nreg;
    p.first_frame = (uint32_t)first;
    p.npairs = (uint32_t)(n - 1);
    p.code_slots = c->code_slots; p.off_slots = c->off_slots; p.tile_pitch = c->tile_pitch; p.tile_rows = c->tile_rows;
    if (c->cfg.kpm_mode == 0) {
      // K1c: per-(frame, region) keypoint lists; K2: pipelined matcher; then the general kernel over
      // whatever K2 deferred (normally nothing: the grid exits on an empty list)
      if (list_first < first + n) {
        const uint32_t items = (uint32_t)(first + n - list_first) * g.nreg;
        rb_list_kernel<<<(items + 7) / 8, 256, 0, c->stream>>>(g, c->d_kp, c->d_w2, (uint32_t)list_first,
                                                              (uint32_t)(first + n - list_first), c->fast.lcap, c->d_lists,
                                                              c->d_counts);
        RB_LAUNCHED(c, "rb_list_kernel");
      }
      if (ev) RB_CUDA(c, cudaEventRecord(ev[3], c->stream));
      RB_CUDA(c, cudaMemsetAsync(c->d_work, 0, 4, c->stream));  // work counter; deferred count / error word accumulate
      const uint32_t dcap = (uint32_t)((n - 1) * g.nreg);
      auto fill = [&](RbKpmFastParams& q) {
        q.lists = c->d_lists; q.counts = c->d_counts; q.votes = p.votes;
        q.first_frame = (uint32_t)first; q.npairs = (uint32_t)(n - 1);
        q.work_counter = c->d_work;
        q.deferred_cap = dcap;
      };
      RB_CUDA(c, cudaMemsetAsync(c->d_work + 4, 0, 4, c->stream));
      const uint2* glist = c->d_deferred;
      const uint32_t* gcount = c->d_work + 4;
      if (c->big_primary) {  // large regions: K2b over runs of pairs, one CTA per SM
        RbKpmFastParams f = c->big;
        fill(f);
        f.deferred_count = c->d_work + 4;  // per-launch list position; d_work[1] keeps the running total
        f.deferred = c->d_deferred;
        f.items = nullptr; f.nitems = nullptr;
        f.run = pick_run(c, f.npairs, g.nreg, (uint32_t)c->sm_count);
        const uint32_t witems = ((f.npairs + f.run - 1) / f.run) * g.nreg;
        uint32_t grid = (uint32_t)c->sm_count;
        if (grid > witems) grid = witems;
        rb_kpm_big_kernel<<<grid, RB_BIG_NT, c->big_smem, c->stream>>>(c->tmap, f);
        RB_LAUNCHED(c, "rb_kpm_big_kernel");
        if (ev) RB_CUDA(c, cudaEventRecord(ev[4], c->stream));
      } else {
        RbKpmFastParams f = c->fast;
        fill(f);
        f.deferred_count = c->d_work + 4;
        f.deferred = c->d_deferred;
        f.items = nullptr; f.nitems = nullptr;
        f.run = pick_run(c, f.npairs, g.nreg, (uint32_t)(c->sm_count * c->fast_ctas_per_sm));
        const uint32_t witems = ((f.npairs + f.run - 1) / f.run) * g.nreg;
        uint32_t grid = (uint32_t)(c->sm_count * c->fast_ctas_per_sm);
        if (grid > witems) grid = witems;
        rb_kpm_fast_kernel<<<grid, RB_FAST_NT, c->fast_smem, c->stream>>>(c->tmap, f);
        RB_LAUNCHED(c, "rb_kpm_fast_kernel");
        if (ev) RB_CUDA(c, cudaEventRecord(ev[4], c->stream));
        if (c->big_smem) {  // second pass over what the first deferred (exits at once on an empty list)
          RB_CUDA(c, cudaMemsetAsync(c->d_work, 0, 4, c->stream));
          RB_CUDA(c, cudaMemsetAsync(c->d_work + 5, 0, 4, c->stream));
          RbKpmFastParams f2 = c->big;
          fill(f2);
          f2.items = c->d_deferred; f2.nitems = c->d_work + 4;
          f2.deferred_count = c->d_work + 5;
          f2.deferred = c->d_deferred2;
          uint32_t grid2 = (uint32_t)c->sm_count;
          if (grid2 > dcap) grid2 = dcap;
          rb_kpm_big_kernel<<<grid2, RB_BIG_NT, c->big_smem, c->stream>>>(c->tmap, f2);
          RB_LAUNCHED(c, "rb_kpm_big_kernel (second pass)");
          glist = c->d_deferred2;
          gcount = c->d_work + 5;
        }
      }
      const struct { uint32_t deferred_cap; } f = {dcap};
      uint32_t dgrid = (uint32_t)c->sm_count * 2;
      if (dgrid > f.deferred_cap) dgrid = f.deferred_cap;
      rb_kpm_deferred_kernel<<<dgrid, c->kpm_nt, c->kpm_smem, c->stream>>>(p, glist, gcount, f.deferred_cap, c->d_work + 1);
      RB_LAUNCHED(c, "rb_kpm_deferred_kernel");
    } else {
      if (ev) { RB_CUDA(c, cudaEventRecord(ev[3], c->stream)); RB_CUDA(c, cudaEventRecord(ev[4], c->stream)); }
      rb_kpm_kernel<<<(uint32_t)((n - 1) * g.nreg), c->kpm_nt, c->kpm_smem, c->stream>>>(p);
      RB_LAUNCHED(c, "rb_kpm_kernel");
    }
    if (ev) RB_CUDA(c, cudaEventRecord(ev[5], c->stream));
    if (c->k3_reference) {  // RB_K3_REFERENCE=1: one thread per pair, rbm::declare_pair as the tests' host build runs it
      rb_declare_offsets_kernel<<<(uint32_t)((n - 1 + 127) / 128), 128, 0, c->stream>>>(g, p.votes, c->d_results + first,
                                                                                     c->d_offsets + first, (uint32_t)(n - 1));
      RB_LAUNCHED(c, "rb_declare_offsets_kernel");
    } else {
      uint32_t LP = 1;
      while (LP < g.nreg) LP <<= 1;
      const uint64_t threads = (uint64_t)(n - 1) * LP;
      rb_declare_group_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, c->stream>>>(g, p.votes, c->d_offsets + first,
                                                                                      (uint32_t)(n - 1), LP);
      RB_LAUNCHED(c, "rb_declare_group_kernel");
    }
  } else if (ev) {
    for (int i = 3; i <= 5; ++i) RB_CUDA(c, cudaEventRecord(ev[i], c->stream));
  }
  if (ev) RB_CUDA(c, cudaEventRecord(ev[6], c->stream));
  return RB_OK;
}

// Enqueues the registration of frames [first, first + n): K1 on frames [kpe_first, first + n) (earlier
// ones were extracted by a previous call), the matcher on the n - 1 pairs, K3.
// K1 is bound by the ALU pipe and uses no shared memory; the matcher is bound by instruction issue and
// its barrier and lives in shared memory.  With overlap_batches > 1 they run CONCURRENTLY: the frames are
// cut into batches, K1 of batch b + 1 goes to a second, lower-priority stream while the matcher of batch
// b runs on the main stream.  On B200 this measured SLOWER than one after the other (K1's CTAs hold the
// whole register file of an SM, the matcher's CTAs queue behind them), so the default is 1.
static int enqueue_range(rb_ctx* c, size_t first, size_t n, size_t kpe_first) {
  const bool prof = c->cfg.profile != 0;
  const size_t kn = kpe_first < first + n ? first + n - kpe_first : 0;
  size_t nb = c->cfg.overlap_batches ? c->cfg.overlap_batches : 1;  // measured: overlapping does not pay (profiles/README.md)
  if (nb > RB_MAX_BATCHES) nb = RB_MAX_BATCHES;
  if (kn < nb * 512) nb = kn / 512 ? kn / 512 : 1;  // batches too small to fill the GPU gain nothing
  c->nbatch_used = 0;
  if (kn == 0) return launch_match(c, first, n, first + n, prof ? c->bev[0].e : nullptr);
  // K1 must come after whatever put the frames in place on the main stream (and after earlier readers of
  // the bit maps it will overwrite)
  RB_CUDA(c, cudaEventRecord(c->ev_entry, c->stream));
  RB_CUDA(c, cudaStreamWaitEvent(c->kpe_stream, c->ev_entry, 0));
  for (size_t b = 0; b < nb; ++b) {
    const size_t b0 = kpe_first + kn * b / nb, b1 = kpe_first + kn * (b + 1) / nb;
    cudaEvent_t* ev = prof ? c->bev[b].e : nullptr;
    if (ev) RB_CUDA(c, cudaEventRecord(ev[0], c->kpe_stream));
    int rc = launch_kpe(c, b0, b1 - b0, c->kpe_stream);
    if (rc != RB_OK) return rc;
    if (ev) RB_CUDA(c, cudaEventRecord(ev[1], c->kpe_stream));
    RB_CUDA(c, cudaEventRecord(c->ev_kpe[b], c->kpe_stream));
    RB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_kpe[b], 0));
    // the pairs whose second frame lies in this batch
    const size_t m0 = b == 0 ? first : b0 - 1;
    rc = launch_match(c, m0, b1 - m0, b0, ev);
    if (rc != RB_OK) return rc;
    c->nbatch_used = b + 1;
  }
  return RB_OK;
}

int rb_register_async(rb_ctx* c, size_t first, size_t n) {
  if (!c) return RB_ERR_INVALID;
  if (n < 1 || first + n > c->cfg.max_frames) { c->err = "rb_register: frame range"; return RB_ERR_CAPACITY; }
  if (first + n > c->uploaded) { c->err = "rb_register: frames not uploaded"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  if (c->d_work) RB_CUDA(c, cudaMemsetAsync(c->d_work, 0, 16, c->stream));  // work counter, deferred total, error word
  const int rc = enqueue_range(c, first, n, first);
  if (rc != RB_OK) return rc;
  c->reg_first = first;
  c->reg_n = n;
  if (c->med_hi == c->med_lo) { c->med_lo = first; c->med_hi = first + n; }
  else {  // the frames in between keep whatever an earlier call wrote; callers register contiguous ranges
    if (first < c->med_lo) c->med_lo = first;
    if (first + n > c->med_hi) c->med_hi = first + n;
  }
  return RB_OK;
}

// Threads the host packer of one context may use: an explicit rb_config.host_threads / RB_HOST_THREADS, else this
// context's fair share of the host: processors / max(ranks of this node (LOCAL_WORLD_SIZE, set by torchrun and
// most launchers), contexts alive in this process).  NOT processors / visible GPUs: one rank on an 8-GPU node owns
// the whole host.
// GPUs that share this host with the calling context: ranks of this node (LOCAL_WORLD_SIZE, set by torchrun and most
// launchers) or contexts alive in this process, whichever is larger
static int host_share() {
  int share = g_live_contexts.load();
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v > share) share = v; }
  return share < 1 ? 1 : share;
}
static int packer_threads(const rb_ctx* c) {
  if (const char* e = getenv("RB_HOST_THREADS")) { const int v = atoi(e); if (v > 0) return v; }
  if (c->cfg.host_threads) return (int)c->cfg.host_threads;
  long procs = sysconf(_SC_NPROCESSORS_ONLN);
  if (procs < 1) procs = 1;
  const long nt = procs / host_share();
  return (int)(nt > 0 ? nt : 1);
}

static void note_registered(rb_ctx* c, size_t first, size_t n) {
  c->reg_first = first;
  c->reg_n = n;
  if (c->med_hi == c->med_lo) { c->med_lo = first; c->med_hi = first + n; }
  else {  // the frames in between keep whatever an earlier call wrote; callers register contiguous ranges
    if (first < c->med_lo) c->med_lo = first;
    if (first + n > c->med_hi) c->med_hi = first + n;
  }
}

static int ensure_chunk_events(rb_ctx* c, size_t nchunks) {
  while (c->ev_chunk.size() < nchunks) {
    cudaEvent_t a = nullptr, b = nullptr;
    RB_CUDA(c, cudaEventCreate(&a));
    RB_CUDA(c, cudaEventCreate(&b));
    c->ev_chunk_start.push_back(a);
    c->ev_chunk.push_back(b);
  }
  return RB_OK;
}

// device side of one landed chunk: the store the chunk did NOT arrive in is derived on the device (K0 / K0u), then
// the chunk is registered with the last frame of its predecessor as its first `previous`
static int chunk_landed(rb_ctx* c, size_t first, size_t at, size_t m, bool arrived_packed, cudaEvent_t landed) {
  const RbGeom& g = c->g;
  RB_CUDA(c, cudaStreamWaitEvent(c->stream, landed, 0));
  if (arrived_packed) {  // K0u: the one-colour-per-byte store K1 reads
    const uint64_t chunks16 = (uint64_t)m * g.H * (g.pitch / 16);
    uint64_t blocks = (chunks16 + 255) / 256;
    const uint64_t maxb = (uint64_t)c->sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    rb_unpack_kernel<<<(uint32_t)blocks, 256, 0, c->stream>>>(c->d_frames4 + c->frame_stride4 * (first + at), c->pitch4,
                                                             c->frame_stride4, c->d_frames + g.frame_stride * (first + at),
                                                             g.pitch, g.frame_stride, g.H, (uint32_t)m);
    RB_LAUNCHED(c, "rb_unpack_kernel");
  } else {
    const int rc = pack_frames(c, first + at, m);  // K0 (no-op without the 4 bit/pixel store)
    if (rc != RB_OK) return rc;
  }
  if (first + at + m > c->uploaded) c->uploaded = first + at + m;
  return at == 0 ? enqueue_range(c, first, m, first) : enqueue_range(c, first + at - 1, m + 1, first + at);
}

// rb_upload + rb_register_async in one call for frames in HOST memory, pipelined.  A chunk of frames reaches the
// device through one of two lanes that work AT THE SAME TIME:
//   packed  host threads pack the chunk to 4 bit/pixel in pinned staging (half the bytes cross PCIe), one
//           cudaMemcpyAsync, K0u derives the byte store on the device;
//   raw     one cudaMemcpyAsync of the caller's bytes as they are (no host work at all; needs page-locked memory),
//           K0 derives the packed store on the device.
// Which lane a chunk takes is decided when its turn comes, from measured rates: the calling thread packs
// continuously; a chunk goes raw when the link would otherwise run dry before the packer could deliver it
// (bytes still queued on the link / measured link rate < chunk frames / measured packer rate).  With many host threads per GPU nearly
// everything goes packed (the link is the limit and packed halves its load); with few threads per GPU (one rank of
// eight on a node) nearly everything goes raw.  Copies of later chunks run under the kernels of earlier ones.
struct RbPendingChunk {  // a landed chunk whose kernels have not been launched yet
  rb_ctx* c;
  size_t first, at, m;
  bool packed, live;
  cudaEvent_t landed;
  int rc;
};
static void flush_pending(void* arg) {
  RbPendingChunk* p = static_cast<RbPendingChunk*>(arg);
  if (!p->live) return;
  p->live = false;
  const int rc = chunk_landed(p->c, p->first, p->at, p->m, p->packed, p->landed);
  if (rc != RB_OK) p->rc = rc;
}

int rb_register_host_async(rb_ctx* c, const uint8_t* frames, size_t first, size_t n) {
  if (!c || !frames) return RB_ERR_INVALID;
  if (n < 1 || first + n > c->cfg.max_frames) { c->err = "rb_register_host: frame range"; return RB_ERR_CAPACITY; }
  RB_CUDA(c, cudaSetDevice(c->device));
  const RbGeom& g = c->g;
  // frames per chunk: what keeps a staging buffer at ~18 MB (see the staging buffers below), 128 ... 512
  size_t chunk = c->cfg.upload_chunk;
  if (!chunk) {
    chunk = (size_t)(18.4e6 / (double)(c->frame_stride4 ? c->frame_stride4 : g.frame_stride / 2 + 1)) / 64 * 64;
    chunk = chunk < 128 ? 128 : chunk > 512 ? 512 : chunk;
  }
  // Chunk sizes: `chunk` frames each, but a long call ramps up (128, 256, ... frames) and down again, so that the
  // link starts after packing 128 frames, not 512, and the last copy + kernels that nothing overlaps are short.
  std::vector<size_t> sizes;
  {
    std::vector<size_t> ramp;
    size_t ramp_sum = 0;
    if (chunk >= 256 && n >= 6 * chunk)
      for (size_t r = 128; r < chunk; r *= 2) { ramp.push_back(r); ramp_sum += r; }
    for (size_t r : ramp) sizes.push_back(r);
    for (size_t left = n - 2 * ramp_sum; left > 0;) { const size_t m = left < chunk ? left : chunk; sizes.push_back(m); left -= m; }
    for (size_t i = ramp.size(); i-- > 0;) sizes.push_back(ramp[i]);
  }
  const size_t nchunks = sizes.size();
  const bool have4 = c->d_frames4 != nullptr;  // the 4 bit/pixel store exists: frames may travel packed
  bool pinned = false;
  {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, frames) == cudaSuccess) pinned = at.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  int lane_force = 0;  // RB_HOST_LANE=raw|packed|auto: experiments and tests
  // Four or more GPUs on one host: the host's DRAM, not a GPU's link, bounds the transfer (measured on the 8-GPU
  // node: 217 GB/s of DMA reads in all, 27 GB/s per link), and a packed chunk costs the DRAM two to three times the
  // traffic of a raw one (read 1 + write 1/2 + DMA read 1/2 against DMA read 1): everything goes raw there
  // (2.59 M frames/s on 8 GPUs against 1.22 M all packed and 2.13 M mixed).
  if (pinned && host_share() >= 4) lane_force = 1;
  if (const char* e = getenv("RB_HOST_LANE")) lane_force = !strcmp(e, "raw") ? 1 : !strcmp(e, "packed") ? 2 : !strcmp(e, "auto") ? 0 : lane_force;
  // Two staging buffers of 512 frames (2 x 18 MB at 320x224) stay in a 60 MB last-level cache, from where the DMA engine
  // reads them; measured: a third buffer (RB_STAGE_BUFS=3) or 1,024-frame chunks cost 10 % (1.13 -> 1.01 / 0.98 M frames/s).
  int nstage = 2;
  if (const char* e = getenv("RB_STAGE_BUFS")) { const int v = atoi(e); if (v == 2 || v == 3) nstage = v; }
  if (have4 && c->stage_frames < chunk) {  // (no chunk is larger than `chunk`)
    for (int i = 0; i < 3; ++i) {
      if (c->h_stage[i]) { cudaFreeHost(c->h_stage[i]); c->h_stage[i] = nullptr; }
      RB_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_stage[i]), c->frame_stride4 * chunk, cudaHostAllocDefault));
    }
    c->stage_frames = chunk;
  }
  int rc = ensure_chunk_events(c, nchunks);
  if (rc != RB_OK) return rc;
  // a previous call's packed copies may still be reading the staging buffers (callers need not synchronise in between)
  RB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  const int pack_threads = packer_threads(c);
  // the copies must not overtake earlier work on the main stream that still reads these slots
  RB_CUDA(c, cudaEventRecord(c->ev_entry, c->stream));
  RB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_entry, 0));
  if (c->d_work) RB_CUDA(c, cudaMemsetAsync(c->d_work, 0, 16, c->stream));
  c->lane_raw = c->lane_packed = 0;
  c->lane_bytes = 0;
  const double raw_bytes_pf = (double)g.W * g.H, packed_bytes_pf = (double)c->frame_stride4;
  size_t oldest = 0;          // first chunk whose copy may still be in flight
  double inflight_bytes = 0;  // bytes queued on the link and not yet known to have landed
  std::vector<double> chunk_bytes(nchunks, 0.0);
  RbPendingChunk pending;
  pending.c = c; pending.first = first; pending.at = pending.m = 0; pending.packed = pending.live = false; pending.landed = nullptr;
  pending.rc = RB_OK;
  int stage_owner[3] = {-1, -1, -1};  // chunk whose packed copy last read staging buffer i
  int stage_next = 0;
  size_t at = 0;
  for (size_t k = 0; k < nchunks; at += sizes[k], ++k) {
    const size_t m = sizes[k];
    // retire landed copies; every retired copy refines the link estimate
    while (oldest < k && cudaEventQuery(c->ev_chunk[oldest]) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, c->ev_chunk_start[oldest], c->ev_chunk[oldest]) == cudaSuccess && ms > 0.02f) {
        const double bps = chunk_bytes[oldest] / (ms * 1e-3);
        c->link_Bps = c->link_Bps > 0 ? 0.7 * c->link_Bps + 0.3 * bps : bps;
      }
      inflight_bytes -= chunk_bytes[oldest];
      ++oldest;
    }
    cudaGetLastError();  // cudaErrorNotReady from the query is not an error
    bool raw;
    if (!have4) raw = true;
    else if (!pinned) raw = false;  // pageable memory: the driver would stage it through its own bounce buffer
    else if (lane_force) raw = lane_force == 1;
    else {
      const double link = c->link_Bps > 0 ? c->link_Bps : 50e9, pack = c->pack_fps > 0 ? c->pack_fps : 1e6;
      // the packer delivers its next chunk to the link in t_pack; if the link's backlog runs out before that, the
      // link would idle: give it this chunk as it is and let the packer start on the next one right away
      // -- unless packed traffic alone already fills the link: then a raw chunk (twice the bytes) only sets it back
      const double t_left = inflight_bytes / link, t_pack = m / pack;
      raw = t_left < t_pack && pack * packed_bytes_pf < 0.85 * link;
    }
    cudaEvent_t landed = c->ev_chunk[k];
    if (raw) {
      RB_CUDA(c, cudaEventRecord(c->ev_chunk_start[k], c->copy_stream));
      rc = copy_frames(c, frames + (size_t)g.W * g.H * at, first + at, m, c->copy_stream);
      if (rc != RB_OK) return rc;
      chunk_bytes[k] = m * raw_bytes_pf;
      ++c->lane_raw;
    } else {
      const int sb = stage_next;
      stage_next = (stage_next + 1) % nstage;
      if (stage_owner[sb] >= 0) RB_CUDA(c, cudaEventSynchronize(c->ev_chunk[stage_owner[sb]]));  // staging buffer free again
      timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      // the previous chunk's kernel launches ride along: issued by this thread while the other packer threads already work
      rb_hostpack_frames_cb(frames + (size_t)g.W * g.H * at, g.W, g.H, m, c->h_stage[sb], c->pitch4, pack_threads,
                            pending.live ? flush_pending : nullptr, &pending);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (pending.rc != RB_OK) return pending.rc;
      const double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
      if (dt > 0) { const double fps = m / dt; c->pack_fps = c->pack_fps > 0 ? 0.7 * c->pack_fps + 0.3 * fps : fps; }
      RB_CUDA(c, cudaEventRecord(c->ev_chunk_start[k], c->copy_stream));
      RB_CUDA(c, cudaMemcpyAsync(c->d_frames4 + c->frame_stride4 * (first + at), c->h_stage[sb], c->frame_stride4 * m,
                                 cudaMemcpyHostToDevice, c->copy_stream));
      stage_owner[sb] = (int)k;
      chunk_bytes[k] = m * packed_bytes_pf;
      ++c->lane_packed;
    }
    RB_CUDA(c, cudaEventRecord(landed, c->copy_stream));
    inflight_bytes += chunk_bytes[k];
    c->lane_bytes += (uint64_t)chunk_bytes[k];
    // the device side of this chunk (stream-ordered behind its copy) is launched with the NEXT chunk's packing, or at once
    // when nothing will be packed next (raw chunks, the last chunk)
    flush_pending(&pending);
    if (pending.rc != RB_OK) return pending.rc;
    pending.live = true; pending.at = at; pending.m = m; pending.packed = !raw; pending.landed = landed;
    if (raw || k + 1 == nchunks) {
      flush_pending(&pending);
      if (pending.rc != RB_OK) return pending.rc;
    }
  }
  note_registered(c, first, n);
  return RB_OK;
}

// The same for callers that already hold 4 bit/pixel frames (pixel x in nibble x & 1 of byte x >> 1, rows of
// `row_bytes` >= ceil(W / 2) bytes, frames back to back): one copy per chunk, no host work.
int rb_register_host_packed4(rb_ctx* c, const uint8_t* packed, size_t row_bytes, size_t first, size_t n) {
  if (!c || !packed) return RB_ERR_INVALID;
  if (n < 1 || first + n > c->cfg.max_frames) { c->err = "rb_register_host_packed4: frame range"; return RB_ERR_CAPACITY; }
  if (!c->d_frames4) { c->err = "rb_register_host_packed4: this geometry has no 4 bit/pixel store (kpm_mode = 1)"; return RB_ERR_STATE; }
  const RbGeom& g = c->g;
  if (row_bytes < (g.W + 1) / 2) { c->err = "rb_register_host_packed4: row_bytes < ceil(W / 2)"; return RB_ERR_INVALID; }
  RB_CUDA(c, cudaSetDevice(c->device));
  const size_t chunk = c->cfg.upload_chunk ? c->cfg.upload_chunk : 1024;
  int rc = ensure_chunk_events(c, (n + chunk - 1) / chunk);
  if (rc != RB_OK) return rc;
  RB_CUDA(c, cudaEventRecord(c->ev_entry, c->stream));
  RB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_entry, 0));
  if (c->d_work) RB_CUDA(c, cudaMemsetAsync(c->d_work, 0, 16, c->stream));
  size_t k = 0;
  for (size_t at = 0; at < n; at += chunk, ++k) {
    const size_t m = at + chunk < n ? chunk : n - at;
    RB_CUDA(c, cudaEventRecord(c->ev_chunk_start[k], c->copy_stream));
    if (row_bytes == c->pitch4)
      RB_CUDA(c, cudaMemcpyAsync(c->d_frames4 + c->frame_stride4 * (first + at), packed + row_bytes * g.H * at,
                                 c->frame_stride4 * m, cudaMemcpyHostToDevice, c->copy_stream));
    else  // the pad bytes of the device rows stay zero (rb_create clears the store)
      RB_CUDA(c, cudaMemcpy2DAsync(c->d_frames4 + c->frame_stride4 * (first + at), c->pitch4, packed + row_bytes * g.H * at,
                                   row_bytes, (g.W + 1) / 2, (size_t)g.H * m, cudaMemcpyHostToDevice, c->copy_stream));
    RB_CUDA(c, cudaEventRecord(c->ev_chunk[k], c->copy_stream));
    rc = chunk_landed(c, first, at, m, true, c->ev_chunk[k]);
    if (rc != RB_OK) return rc;
  }
  note_registered(c, first, n);
  return RB_OK;
}

// Last rb_register_host_async: chunks that went raw / packed, and the rate estimates the choice was made from.
int rb_host_lane_stats(rb_ctx* c, uint64_t* raw_chunks, uint64_t* packed_chunks, double* link_GBps, double* pack_fps, int* threads,
                       uint64_t* h2d_bytes) {
  if (!c) return RB_ERR_INVALID;
  if (raw_chunks) *raw_chunks = c->lane_raw;
  if (packed_chunks) *packed_chunks = c->lane_packed;
  if (link_GBps) *link_GBps = c->link_Bps / 1e9;
  if (pack_fps) *pack_fps = c->pack_fps;
  if (threads) *threads = packer_threads(c);
  if (h2d_bytes) *h2d_bytes = c->lane_bytes;
  return RB_OK;
}

int rb_kernel_times(rb_ctx* c, float* ms, size_t n) {
  if (!c || !ms || n < 3) return RB_ERR_INVALID;
  if (!c->cfg.profile) { c->err = "context created without profile=1"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  // each kernel's own elapsed time, summed over the batches of the last call (K1 of one batch and the
  // matcher of the previous one overlap in wall time; these are durations, not a timeline)
  float d[5] = {0, 0, 0, 0, 0};
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  const size_t nbu = c->nbatch_used ? c->nbatch_used : 1;
  for (size_t b = 0; b < nbu; ++b) {
    float t;
    if (c->nbatch_used) { RB_CUDA(c, cudaEventElapsedTime(&t, c->bev[b].e[0], c->bev[b].e[1])); d[0] += t; }
    for (int i = 1; i < 5; ++i) { RB_CUDA(c, cudaEventElapsedTime(&t, c->bev[b].e[i + 1], c->bev[b].e[i + 2])); d[i] += t; }
  }
  const float all[6] = {d[0], d[1] + d[2] + d[3], d[4], d[1], d[2], d[3]};  // kpe, matcher total, declare, lists, match, deferred
  for (size_t i = 0; i < n && i < 6; ++i) ms[i] = all[i];
  return RB_OK;
}

int rb_fetch_offsets(rb_ctx* c, rb_offset* out, size_t n_pairs) {
  if (!c || (!out && n_pairs)) return RB_ERR_INVALID;
  if (c->reg_n == 0 || n_pairs > c->reg_n - 1) { c->err = "rb_fetch_offsets: nothing registered"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  if (n_pairs)
    RB_CUDA(c, cudaMemcpyAsync(out, c->d_offsets + c->reg_first, n_pairs * sizeof(rb_offset), cudaMemcpyDeviceToHost, c->stream));
  // the pipelined matcher's error word (a TMA transaction that never completed) travels with the offsets: results
  // computed from an incomplete tile must not be reported as RB_OK
  if (c->d_work && c->h_flags) RB_CUDA(c, cudaMemcpyAsync(c->h_flags, c->d_work, 16, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->d_work && c->h_flags && c->h_flags[2]) { c->err = "pipelined matcher: a TMA transaction timed out"; return RB_ERR_CUDA; }
  return RB_OK;
}

int rb_fetch_medians(rb_ctx* c, size_t first, size_t n, uint8_t* out) {
  if (!c || !out) return RB_ERR_INVALID;
  if (!c->d_median) { c->err = "context created with compute_median = 0"; return RB_ERR_STATE; }
  if (first + n > c->cfg.max_frames) return RB_ERR_CAPACITY;
  const RbGeom& g = c->g;
  RB_CUDA(c, cudaSetDevice(c->device));
  // pixel x lives at byte x + 2 of a median row
  RB_CUDA(c, cudaMemcpy2DAsync(out, g.W, c->d_median + g.median_stride * first + 2, g.mpitch, g.W, (size_t)g.H * n,
                               cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

int rb_register(rb_ctx* c, size_t first, size_t n, rb_offset* out, uint8_t* out_median) {
  int rc = rb_register_async(c, first, n);
  if (rc != RB_OK) return rc;
  rc = rb_fetch_offsets(c, out, n - 1);
  if (rc != RB_OK) return rc;
  if (out_median) rc = rb_fetch_medians(c, first, n, out_median);
  return rc;
}

int rb_keypoints(rb_ctx* c, size_t frame, rb_keypoint* out, size_t cap, size_t* count) {
  if (!c || !count) return RB_ERR_INVALID;
  if (frame < c->reg_first || frame >= c->reg_first + c->reg_n) {
    c->err = "rb_keypoints: frame not in the last registered range";
    return RB_ERR_STATE;
  }
  const RbGeom& g = c->g;
  RB_CUDA(c, cudaSetDevice(c->device));
  const uint32_t dcap = g.W * g.H;
  rb_keypoints_kernel<<<1, 1024, (g.W + 1) * sizeof(uint32_t), c->stream>>>(
      g, c->d_frames + g.frame_stride * frame, c->d_kp + frame * g.H * g.NS, c->d_w2 + frame * g.H * g.NS, c->d_kps,
      dcap, c->d_tap_count);
  RB_LAUNCHED(c, "rb_keypoints_kernel");
  uint32_t n = 0;
  RB_CUDA(c, cudaMemcpyAsync(&n, c->d_tap_count, 4, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  *count = n;
  const size_t m = n < cap ? n : cap;
  if (m && out) {
    RB_CUDA(c, cudaMemcpyAsync(out, c->d_kps, m * sizeof(rb_keypoint), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return RB_OK;
}

int rb_region_ballots(rb_ctx* c, size_t pair, rb_region_vote* out) {
  if (!c || !out) return RB_ERR_INVALID;
  if (c->reg_n < 2 || pair >= c->reg_n - 1) { c->err = "rb_region_ballots: pair not registered"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  RB_CUDA(c, cudaMemcpyAsync(out, c->d_votes + (c->reg_first + pair) * c->g.nreg, c->g.nreg * sizeof(RbRegionVote),
                             cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}


int rb_frame_digests(rb_ctx* c, size_t first, size_t n, rb_frame_digest* out) {
  if (!c || (!out && n)) return RB_ERR_INVALID;
  if (first + n > c->uploaded || first < c->med_lo || first + n > c->med_hi) {
    c->err = "rb_frame_digests: frames not registered";
    return RB_ERR_STATE;
  }
  if (n == 0) return RB_OK;
  RB_CUDA(c, cudaSetDevice(c->device));
  rb_frame_digest* d = nullptr;
  RB_CUDA(c, cudaMalloc(&d, n * sizeof(rb_frame_digest)));
  cudaError_t e = cudaMemsetAsync(d, 0, n * sizeof(rb_frame_digest), c->stream);
  if (e == cudaSuccess) {
    rb_digest_kernel<<<(uint32_t)n, 256, 0, c->stream>>>(c->g, c->d_frames, c->d_median, c->d_kp, c->d_w2, (uint32_t)first, d);
    ++c->launches;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, n * sizeof(rb_frame_digest), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) { c->err = std::string("rb_frame_digests: ") + cudaGetErrorString(e); return RB_ERR_CUDA; }
  return RB_OK;
}

int rb_fetch_ballots(rb_ctx* c, size_t pair, size_t n_pairs, rb_region_vote* out) {
  if (!c || (!out && n_pairs)) return RB_ERR_INVALID;
  if (c->reg_n < 2 || pair + n_pairs > c->reg_n - 1) { c->err = "rb_fetch_ballots: pairs not registered"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  if (n_pairs)
    RB_CUDA(c, cudaMemcpyAsync(out, c->d_votes + (c->reg_first + pair) * c->g.nreg, n_pairs * c->g.nreg * sizeof(RbRegionVote),
                               cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

// Re-runs the production K2 kernel on one (pair, region) with the bin dump enabled.
int rb_region_votes(rb_ctx* c, size_t pair, uint32_t region, rb_bin* out, size_t cap, size_t* count) {
  if (!c || !count) return RB_ERR_INVALID;
  if (c->reg_n < 2 || pair >= c->reg_n - 1 || region >= c->g.nreg) { c->err = "rb_region_votes: pair/region"; return RB_ERR_STATE; }
  const RbGeom& g = c->g;
  RB_CUDA(c, cudaSetDevice(c->device));
  RB_CUDA(c, cudaMemsetAsync(c->d_tap_count, 0, 4, c->stream));
  RbKpmParams p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.frames = c->d_frames; p.kpbits = c->d_kp; p.w2bits = c->d_w2;
  p.votes = c->d_votes + (size_t)c->cfg.max_frames * g.nreg - g.nreg;  // scratch slot: last pair row is never used
  p.first_frame = (uint32_t)(c->reg_first + pair);
  p.npairs = 1;
  p.code_slots = c->code_slots; p.off_slots = c->off_slots; p.tile_pitch = c->tile_pitch; p.tile_rows = c->tile_rows;
  p.tap_bins = c->d_tap_bins; p.tap_cap = 1u << 20; p.tap_count = c->d_tap_count;
  p.tap_pair = 0; p.tap_region = region;
  // launch all regions of the pair (blockIdx -> region), only `region` dumps
  rb_kpm_kernel<<<g.nreg, c->kpm_nt, c->kpm_smem, c->stream>>>(p);
  RB_LAUNCHED(c, "rb_kpm_kernel");
  uint32_t n = 0;
  RB_CUDA(c, cudaMemcpyAsync(&n, c->d_tap_count, 4, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  *count = n;
  const size_t m = n < cap ? n : cap;
  if (m && out) {
    RB_CUDA(c, cudaMemcpyAsync(out, c->d_tap_bins, m * sizeof(rb_bin), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return RB_OK;
}

int rb_blit_blend(rb_ctx* c, const rb_placement* placements, size_t n, uint32_t mapW, uint32_t mapH, uint16_t* out_dots,
                  uint8_t* out_image, uint8_t* out_mask) {
  if (!c || (!placements && n) || mapW == 0 || mapH == 0) return RB_ERR_INVALID;
  const RbGeom& g = c->g;
  for (size_t i = 0; i < n; ++i) {
    const rb_placement& p = placements[i];
    if (p.frame >= c->uploaded) { c->err = "rb_blit_blend: frame not uploaded"; return RB_ERR_STATE; }
    if (p.x < 0 || p.y < 0 || (uint64_t)p.x + g.W > mapW || (uint64_t)p.y + g.H > mapH) {
      c->err = "rb_blit_blend: a frame lies outside the map";
      return RB_ERR_INVALID;
    }
  }
  RB_CUDA(c, cudaSetDevice(c->device));
  if (n > c->places_cap) {
    if (c->d_places) { cudaFree(c->d_places); c->bytes -= c->places_cap * sizeof(RbPlacement); c->d_places = nullptr; }
    c->places_cap = 0;
    RB_CUDA(c, dmalloc(c, &c->d_places, n * sizeof(RbPlacement)));
    c->places_cap = n;
  }
  const size_t px = (size_t)mapW * mapH, need = px * 34 + 256;
  if (need > c->map_cap) {
    if (c->d_map) { cudaFree(c->d_map); c->bytes -= c->map_cap; c->d_map = nullptr; }
    c->map_cap = 0;
    RB_CUDA(c, dmalloc(c, &c->d_map, need));
    c->map_cap = need;
  }
  uint16_t* d_dots = reinterpret_cast<uint16_t*>(c->d_map);
  uint8_t* d_img = c->d_map + px * 32;
  uint8_t* d_msk = d_img + px;
  c->map_w = mapW; c->map_h = mapH;
  if (n) RB_CUDA(c, cudaMemcpyAsync(c->d_places, placements, n * sizeof(RbPlacement), cudaMemcpyHostToDevice, c->stream));
  const dim3 grid((mapW + RB_BLIT_TX - 1) / RB_BLIT_TX, (mapH + RB_BLIT_TY - 1) / RB_BLIT_TY);
  rb_blit_blend_kernel<false><<<grid, RB_BLIT_NT, 0, c->stream>>>(c->d_frames, g.pitch, g.frame_stride, g.W, g.H, c->d_places,
                                                                  (uint32_t)n, mapW, mapH, d_dots, d_img, d_msk, nullptr, 0);
  RB_LAUNCHED(c, "rb_blit_blend_kernel");
  if (out_dots) RB_CUDA(c, cudaMemcpyAsync(out_dots, d_dots, px * 32, cudaMemcpyDeviceToHost, c->stream));
  if (out_image) RB_CUDA(c, cudaMemcpyAsync(out_image, d_img, px, cudaMemcpyDeviceToHost, c->stream));
  if (out_mask) RB_CUDA(c, cudaMemcpyAsync(out_mask, d_msk, px, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

static int fgmask_common(rb_ctx* c, const uint8_t* bg, uint32_t bgW, uint32_t bgH, int32_t px, int32_t py,
                         const uint8_t* dframe, uint32_t fpitch, uint8_t* out_mask) {
  const RbGeom& g = c->g;
  if (px < 0 || py < 0 || (uint64_t)px + g.W > bgW || (uint64_t)py + g.H > bgH) {
    c->err = "rb_foreground_mask: frame window outside the background";
    return RB_ERR_INVALID;
  }
  const size_t need = (size_t)bgW * bgH + 64;
  if (need > c->bg_cap) {
    if (c->d_bg) { cudaFree(c->d_bg); c->bytes -= c->bg_cap; c->d_bg = nullptr; c->bg_cap = 0; }
    RB_CUDA(c, dmalloc(c, &c->d_bg, need));
    c->bg_cap = need;
  }
  RB_CUDA(c, cudaMemcpyAsync(c->d_bg, bg, (size_t)bgW * bgH, cudaMemcpyHostToDevice, c->stream));
  const long long idx = (long long)bgW * py + px;  // cdt::to_index (src/cdt.hpp:173-177; src/fde.hpp:87)
  const uint32_t total = ((g.W + 15) / 16) * g.H;
  uint32_t blocks = (total + 255) / 256;
  const uint32_t cap = (uint32_t)c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  rb_fgmask_kernel<<<blocks, 256, 0, c->stream>>>(c->d_bg, idx, bgW, dframe, fpitch, c->d_mask, g.W, g.W, g.H);
  RB_LAUNCHED(c, "rb_fgmask_kernel");
  if (out_mask)
    RB_CUDA(c, cudaMemcpyAsync(out_mask, c->d_mask, (size_t)g.W * g.H, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

int rb_foreground_mask(rb_ctx* c, const uint8_t* bg, uint32_t bgW, uint32_t bgH, int32_t px, int32_t py,
                       const uint8_t* frame, uint8_t* out_mask) {
  if (!c || !bg || !frame || !out_mask) return RB_ERR_INVALID;
  RB_CUDA(c, cudaSetDevice(c->device));
  RB_CUDA(c, cudaMemcpyAsync(c->d_fgframe, frame, (size_t)c->g.W * c->g.H, cudaMemcpyHostToDevice, c->stream));
  return fgmask_common(c, bg, bgW, bgH, px, py, c->d_fgframe, c->g.W, out_mask);
}

int rb_foreground_mask_resident(rb_ctx* c, const uint8_t* bg, uint32_t bgW, uint32_t bgH, int32_t px, int32_t py,
                                size_t frame, uint8_t* out_mask) {
  if (!c || !bg) return RB_ERR_INVALID;
  if (frame >= c->uploaded) { c->err = "rb_foreground_mask_resident: frame not uploaded"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  return fgmask_common(c, bg, bgW, bgH, px, py, c->d_frames + c->g.frame_stride * frame, c->g.pitch, out_mask);
}

// ---- pass 2: fdf::filter for one fragment (src/fdf.hpp:40-75) ---------------------------------------
static int fg_setup(rb_ctx* c) {
  if (c->fg_ready) return RB_OK;
  const RbGeom& g = c->g;
  if (g.H < 4 || g.W < 3) { c->err = "rb_filter_fragment: frame too small"; return RB_ERR_INVALID; }
  c->fg_NW = (g.W + 31) / 32;
  int smem_max = 0, smem_sm = 0, smem_res = 0;
  RB_CUDA(c, cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  RB_CUDA(c, cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device));
  RB_CUDA(c, cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, c->device));
  const size_t fixed = (rbg::fixed_bytes(g.H, c->fg_NW) + 15) & ~(size_t)15;
  const uint32_t rmax = (g.W - 2) * (g.H - 3);  // every interior pixel its own run
  int want = getenv("RB_FG_CTAS") ? atoi(getenv("RB_FG_CTAS")) : 2;
  if (want < 1) want = 1;
  if (want > 4) want = 4;
  // shared-memory variant: `want` CTAs per SM share the SM's shared memory; a quarter of the table space goes
  // to the statistics slots, the rest to the labels
  size_t budget = (size_t)smem_sm / want - smem_res;
  if (budget > (size_t)smem_max) budget = smem_max;
  c->fg_fast_ok = false;
  if (budget > fixed + 8192) {
    const size_t avail = budget - fixed - 64;
    uint32_t scap = (uint32_t)(avail / 4 / 16);
    if (scap > 4096) scap = 4096;
    size_t r = (avail - (size_t)scap * 16) * 4 / 9;  // 2 bytes label + 2 bits per run
    if (r > rmax) r = rmax;
    if (r + scap > 65535) r = 65535 - scap;
    r &= ~(size_t)31;
    if (getenv("RB_FG_RCAP")) r = (size_t)atoi(getenv("RB_FG_RCAP"));  // tests: force deferrals
    if (getenv("RB_FG_SCAP")) scap = (uint32_t)atoi(getenv("RB_FG_SCAP"));
    c->fg_rcap = (uint32_t)r;
    c->fg_scap = scap;
    c->fg_smem = fixed + rbg::table_bytes<uint16_t>(c->fg_rcap, c->fg_scap);
    if (c->fg_rcap >= 64 && scap >= 1 && c->fg_smem <= (size_t)smem_max && c->fg_rcap + scap <= 65535) {
      RB_CUDA(c, cudaFuncSetAttribute(rb_fg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->fg_smem));
      // 1,024 threads per CTA, two CTAs per SM (32 registers): 2.02 ms per 4,000 frames against 2.12 ms with 512
      c->fg_nt = getenv("RB_FG_NT") ? (uint32_t)atoi(getenv("RB_FG_NT")) : 2 * RB_FG_NT;
      if (c->fg_nt != RB_FG_NT && c->fg_nt != 2 * RB_FG_NT) c->fg_nt = 2 * RB_FG_NT;
      int occ = 0;
      RB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rb_fg_kernel, (int)c->fg_nt, c->fg_smem));
      if (occ >= 1) { c->fg_fast_ok = true; c->fg_grid = (uint32_t)(occ * c->sm_count); }
    }
  }
  if (getenv("RB_FG_GENERAL_ONLY")) c->fg_fast_ok = false;
  // general variant: bit maps in shared memory, worst-case tables in a global slab per CTA
  c->fg_gen_rcap = rmax;
  c->fg_gen_smem = fixed;
  if (c->fg_gen_smem > (size_t)smem_max) { c->err = "rb_filter_fragment: frame bit maps exceed shared memory"; return RB_ERR_INVALID; }
  RB_CUDA(c, cudaFuncSetAttribute(rb_fg_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->fg_gen_smem));
  int occ = 0;
  RB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rb_fg_general_kernel, RB_FG_NT, c->fg_gen_smem));
  if (occ < 1) occ = 1;
  if (occ > 2) occ = 2;
  c->fg_gen_nt = occ == 1 ? 2 * RB_FG_NT : RB_FG_NT;  // a CTA that has the SM to itself gets twice the threads
  c->fg_gen_grid = (uint32_t)(occ * c->sm_count);
  c->fg_slab = (rbg::table_bytes<uint32_t>(rmax, rmax) + 255) & ~(size_t)255;
  RB_CUDA(c, dmalloc(c, &c->d_fg_scratch, c->fg_slab * c->fg_gen_grid));
  RB_CUDA(c, dmalloc(c, &c->d_fg_count, 256));
  for (int i = 0; i < 4; ++i) RB_CUDA(c, cudaEventCreate(&c->fg_ev[i]));
  c->fg_ready = true;
  return RB_OK;
}

int rb_filter_fragment(rb_ctx* c, const rb_placement* placements, size_t n, uint32_t mapW, uint32_t mapH,
                       const uint8_t* background, uint16_t* out_dots, uint8_t* out_image, uint8_t* out_mask,
                       uint8_t* out_fgmasks, uint32_t* out_ncontours) {
  if (!c || (!placements && n) || mapW == 0 || mapH == 0 || n > 0xFFFFFFFFu) return RB_ERR_INVALID;
  if (!c->d_median) { c->err = "rb_filter_fragment: context created with compute_median = 0"; return RB_ERR_STATE; }
  const RbGeom& g = c->g;
  for (size_t i = 0; i < n; ++i) {
    const rb_placement& p = placements[i];
    if (p.frame >= c->uploaded || p.frame < c->med_lo || p.frame >= c->med_hi) {
      c->err = "rb_filter_fragment: frame not registered (its median image is needed, src/fdf.hpp:60)";
      return RB_ERR_STATE;
    }
    if (p.x < 0 || p.y < 0 || (uint64_t)p.x + g.W > mapW || (uint64_t)p.y + g.H > mapH) {
      c->err = "rb_filter_fragment: a frame lies outside the map";
      return RB_ERR_INVALID;
    }
  }
  RB_CUDA(c, cudaSetDevice(c->device));
  int rc = fg_setup(c);
  if (rc != RB_OK) return rc;
  if (n > c->places_cap) {
    if (c->d_places) { cudaFree(c->d_places); c->bytes -= c->places_cap * sizeof(RbPlacement); c->d_places = nullptr; }
    c->places_cap = 0;
    RB_CUDA(c, dmalloc(c, &c->d_places, n * sizeof(RbPlacement)));
    c->places_cap = n;
  }
  const size_t px = (size_t)mapW * mapH, need = px * 34 + 256;
  if (need > c->map_cap) {
    if (c->d_map) { cudaFree(c->d_map); c->bytes -= c->map_cap; c->d_map = nullptr; }
    c->map_cap = 0;
    RB_CUDA(c, dmalloc(c, &c->d_map, need));
    c->map_cap = need;
  }
  if (px + 64 > c->bg_cap) {
    if (c->d_bg) { cudaFree(c->d_bg); c->bytes -= c->bg_cap; c->d_bg = nullptr; c->bg_cap = 0; }
    RB_CUDA(c, dmalloc(c, &c->d_bg, px + 64));
    c->bg_cap = px + 64;
  }
  const size_t fwords = (size_t)g.H * c->fg_NW;
  if (n > c->fg_cap) {
    cudaFree(c->d_fgbits); cudaFree(c->d_fg_nkept); cudaFree(c->d_fg_deferred);
    c->bytes -= c->fg_cap * (fwords * 4 + 8);
    c->d_fgbits = c->d_fg_nkept = c->d_fg_deferred = nullptr;
    c->fg_cap = 0;
    RB_CUDA(c, dmalloc(c, &c->d_fgbits, n * fwords * 4));
    RB_CUDA(c, dmalloc(c, &c->d_fg_nkept, n * 4));
    RB_CUDA(c, dmalloc(c, &c->d_fg_deferred, n * 4));
    c->fg_cap = n;
  }
  uint16_t* d_dots = reinterpret_cast<uint16_t*>(c->d_map);
  uint8_t* d_img = c->d_map + px * 32;
  uint8_t* d_msk = d_img + px;
  c->map_w = mapW; c->map_h = mapH;
  const dim3 grid((mapW + RB_BLIT_TX - 1) / RB_BLIT_TX, (mapH + RB_BLIT_TY - 1) / RB_BLIT_TY);
  if (n) RB_CUDA(c, cudaMemcpyAsync(c->d_places, placements, n * sizeof(RbPlacement), cudaMemcpyHostToDevice, c->stream));
  RB_CUDA(c, cudaEventRecord(c->fg_ev[0], c->stream));
  // 1. the background: given, or fragment.blend() of the plain blit (fdf::details::get_background, src/fdf.hpp:21-34)
  if (background) {
    RB_CUDA(c, cudaMemcpyAsync(c->d_bg, background, px, cudaMemcpyHostToDevice, c->stream));
  } else {
    rb_blit_blend_kernel<false><<<grid, RB_BLIT_NT, 0, c->stream>>>(c->d_frames, g.pitch, g.frame_stride, g.W, g.H, c->d_places,
                                                                    (uint32_t)n, mapW, mapH, nullptr, d_img, nullptr, nullptr, 0);
    RB_LAUNCHED(c, "rb_blit_blend_kernel");
    RB_CUDA(c, cudaMemcpyAsync(c->d_bg, d_img, px, cudaMemcpyDeviceToDevice, c->stream));
  }
  RB_CUDA(c, cudaEventRecord(c->fg_ev[1], c->stream));
  // 2. every frame's foreground bit map (fde::extractor::extract + fde::mask, src/fdf.hpp:62-63)
  RB_CUDA(c, cudaMemsetAsync(c->d_fg_count, 0, 8, c->stream));
  if (n) {
    RbFgParams p;
    memset(&p, 0, sizeof(p));
    p.g = g;
    p.frames = c->d_frames; p.median = c->d_median; p.bg = c->d_bg; p.bgW = mapW; p.bgH = mapH;
    p.places = c->d_places; p.n = (uint32_t)n; p.NW = c->fg_NW;
    p.area_limit = (uint32_t)(((uint64_t)g.W * g.H) / 5);
    p.fgbits = c->d_fgbits; p.nkept = c->d_fg_nkept; p.deferred = c->d_fg_deferred; p.ndeferred = c->d_fg_count;
    if (c->fg_fast_ok) {
      p.rcap = c->fg_rcap; p.scap = c->fg_scap;
      const uint32_t blocks = n < c->fg_grid ? (uint32_t)n : c->fg_grid;
      rb_fg_kernel<<<blocks, c->fg_nt, c->fg_smem, c->stream>>>(p);
      RB_LAUNCHED(c, "rb_fg_kernel");
      p.todo = c->d_fg_deferred; p.ntodo = c->d_fg_count;
    }
    p.rcap = c->fg_gen_rcap; p.scap = c->fg_gen_rcap;
    p.scratch = c->d_fg_scratch; p.scratch_stride = c->fg_slab;
    const uint32_t blocks = n < c->fg_gen_grid ? (uint32_t)n : c->fg_gen_grid;
    rb_fg_general_kernel<<<blocks, c->fg_gen_nt, c->fg_gen_smem, c->stream>>>(p);
    RB_LAUNCHED(c, "rb_fg_general_kernel");
  }
  RB_CUDA(c, cudaEventRecord(c->fg_ev[2], c->stream));
  // 3. the masked blit of every frame + blend of the result (src/fdf.hpp:64, src/fgm.hpp:71-85,115-135)
  rb_blit_blend_kernel<true><<<grid, RB_BLIT_NT, 0, c->stream>>>(c->d_frames, g.pitch, g.frame_stride, g.W, g.H, c->d_places,
                                                                 (uint32_t)n, mapW, mapH, d_dots, d_img, d_msk, c->d_fgbits, c->fg_NW);
  RB_LAUNCHED(c, "rb_blit_blend_kernel");
  RB_CUDA(c, cudaEventRecord(c->fg_ev[3], c->stream));
  if (out_dots) RB_CUDA(c, cudaMemcpyAsync(out_dots, d_dots, px * 32, cudaMemcpyDeviceToHost, c->stream));
  if (out_image) RB_CUDA(c, cudaMemcpyAsync(out_image, d_img, px, cudaMemcpyDeviceToHost, c->stream));
  if (out_mask) RB_CUDA(c, cudaMemcpyAsync(out_mask, d_msk, px, cudaMemcpyDeviceToHost, c->stream));
  if (out_ncontours && n) RB_CUDA(c, cudaMemcpyAsync(out_ncontours, c->d_fg_nkept, n * 4, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaMemcpyAsync(&c->fg_last_deferred, c->d_fg_count, 4, cudaMemcpyDeviceToHost, c->stream));
  if (out_fgmasks && n) {  // parity tap: bit maps -> bytes, in chunks through a small scratch buffer
    const size_t chunk = 256, fbytes = (size_t)g.W * g.H;
    if (!c->d_fg_bytes) RB_CUDA(c, dmalloc(c, &c->d_fg_bytes, chunk * fbytes));
    for (size_t f0 = 0; f0 < n; f0 += chunk) {
      const size_t m = n - f0 < chunk ? n - f0 : chunk;
      rb_fgbits_bytes_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_fgbits + f0 * fwords, (uint32_t)m, g.W, g.H, c->fg_NW,
                                                                      c->d_fg_bytes);
      RB_LAUNCHED(c, "rb_fgbits_bytes_kernel");
      RB_CUDA(c, cudaMemcpyAsync(out_fgmasks + f0 * fbytes, c->d_fg_bytes, m * fbytes, cudaMemcpyDeviceToHost, c->stream));
    }
  }
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 3; ++i) RB_CUDA(c, cudaEventElapsedTime(&c->fg_ms[i], c->fg_ev[i], c->fg_ev[i + 1]));
  return RB_OK;
}

int rb_filter_times(rb_ctx* c, float* ms, size_t n, uint32_t* frames_deferred) {
  if (!c || !c->fg_ready) return RB_ERR_STATE;
  for (size_t i = 0; i < n && i < 3; ++i) ms[i] = c->fg_ms[i];
  if (frames_deferred) *frames_deferred = c->fg_last_deferred;
  return RB_OK;
}

// ---- fragment splicing: fgs::details::extract_single + the cellular kpm::match ----------
struct rb_snippet {
  int device;
  int sm_count;
  cudaStream_t stream;
  RbGeom g;
  uint16_t* d_dots;   // the fragment's dot map (16 x uint16 per map pixel), resident: rb_snippet_merge adds maps in place
  uint8_t* d_image;   // blend image, g.pitch bytes per row
  uint8_t* d_mask;    // blend mask, W bytes per row
  uint32_t* d_kp;
  uint32_t* d_w2;
  RbSnipKp* d_kps;
  uint32_t nkp;
  unsigned long long* d_count;
  uint8_t* d_scratch;   // rb_snippet_match work area of this snippet as `prev` (grown on demand, kept)
  size_t scratch_cap;
  std::string err;
};

#define RS_CUDA(s, call)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (s)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      return RB_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// Snippet buffers come from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync, with the pool told to keep
// what is freed): a splice creates and destroys a snippet per merge, and plain cudaMalloc / cudaFree (milliseconds each
// for 100 MB maps, plus a device-wide synchronisation per free) cost more than the merge's kernels.
static cudaError_t snip_alloc(rb_snippet* s, void** p, size_t bytes) { return cudaMallocAsync(p, bytes, s->stream); }
static void snip_free(rb_snippet* s, void* p) { if (p) cudaFreeAsync(p, s->stream); }
static void snip_pool_setup(int device) {
  static std::atomic<unsigned long long> done{0};
  if (device < 0 || device >= 64 || (done.load() >> device) & 1ull) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done.fetch_or(1ull << device);
}

void rb_snippet_destroy(rb_snippet* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) {
    snip_free(s, s->d_dots); snip_free(s, s->d_image); snip_free(s, s->d_mask); snip_free(s, s->d_kp); snip_free(s, s->d_w2);
    snip_free(s, s->d_kps); snip_free(s, s->d_count); snip_free(s, s->d_scratch);
    cudaStreamSynchronize(s->stream);
    cudaStreamDestroy(s->stream);
  }
  delete s;
}

const char* rb_snippet_last_error(rb_snippet* s) { return s ? s->err.c_str() : "null snippet"; }

// a snippet object for a W x H map on `device`, dots allocated (not filled)
static int snippet_new(int device, uint32_t W, uint32_t H, rb_snippet** out) {
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return RB_ERR_NO_DEVICE;  // no CPU fallback
  }
  rb_snippet* s = new (std::nothrow) rb_snippet();
  if (!s) return RB_ERR_INVALID;
  *out = s;
  s->device = device;
  // kpe::extractor<kpr::grid<1, 1>, 0> (src/fgs.hpp:16,85): one region, no overlap
  if (rb_make_geom(W, H, 1, 1, 0, 10, 3, &s->g) != 0 || W >= 32768 || H >= 32768) { s->err = "unsupported map size"; return RB_ERR_INVALID; }
  RS_CUDA(s, cudaSetDevice(device));
  RS_CUDA(s, cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device));
  RS_CUDA(s, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  snip_pool_setup(device);
  RS_CUDA(s, snip_alloc(s, reinterpret_cast<void**>(&s->d_dots), (size_t)W * H * 32));
  return RB_OK;
}

// fgs::details::extract_single (src/fgs.hpp:80-89) over the resident dots: blend, K1 with a 1 x 1 grid, keypoint records
static int snippet_extract(rb_snippet* s) {
  const RbGeom& g = s->g;
  const uint32_t W = g.W, H = g.H;
  const size_t px = (size_t)W * H, words = (size_t)H * g.NS;
  RS_CUDA(s, snip_alloc(s, reinterpret_cast<void**>(&s->d_image), g.frame_stride + 256));
  RS_CUDA(s, snip_alloc(s, reinterpret_cast<void**>(&s->d_mask), px + 256));
  RS_CUDA(s, snip_alloc(s, reinterpret_cast<void**>(&s->d_kp), words * 4));
  RS_CUDA(s, snip_alloc(s, reinterpret_cast<void**>(&s->d_w2), words * 4));
  RS_CUDA(s, snip_alloc(s, reinterpret_cast<void**>(&s->d_count), 64));
  RS_CUDA(s, cudaMemsetAsync(s->d_image, 0, g.frame_stride + 256, s->stream));
  RS_CUDA(s, cudaMemsetAsync(s->d_kp, 0, words * 4, s->stream));
  RS_CUDA(s, cudaMemsetAsync(s->d_w2, 0, words * 4, s->stream));
  RS_CUDA(s, cudaMemsetAsync(s->d_count, 0, 64, s->stream));
  // fragment.blend() (src/fgs.hpp:81, src/fgm.hpp:115-135)
  rb_blend_kernel<<<s->sm_count * 8, 256, 0, s->stream>>>(s->d_dots, W, H, s->d_image, g.pitch, s->d_mask);
  RS_CUDA(s, cudaGetLastError());
  // extractor.extract(image, median, ...) (src/fgs.hpp:85-88): K1 on the one map image, no median output
  RbKpeParams p;
  p.g = g;
  p.frames = s->d_image; p.median = nullptr; p.kpbits = s->d_kp; p.w2bits = s->d_w2; p.nframes = 1;
  uint32_t nseg = (g.H - 6) / 8;  // a single image: as many row segments as the 4-row warm-up allows
  if (nseg > 64) nseg = 64;
  if (nseg < 1) nseg = 1;
  p.nseg = nseg;
  p.seg_rows = (g.H - 6 + nseg - 1) / nseg;
  const size_t items = (size_t)nseg * g.NS;
  rb_kpe_kernel<<<(uint32_t)((items + 127) / 128), 128, 0, s->stream>>>(p);
  RS_CUDA(s, cudaGetLastError());
  rb_count_kernel<<<s->sm_count * 4, 256, 0, s->stream>>>(s->d_kp, words, s->d_count);
  RS_CUDA(s, cudaGetLastError());
  unsigned long long n = 0;
  RS_CUDA(s, cudaMemcpyAsync(&n, s->d_count, 8, cudaMemcpyDeviceToHost, s->stream));
  RS_CUDA(s, cudaStreamSynchronize(s->stream));
  s->nkp = (uint32_t)n;
  RS_CUDA(s, snip_alloc(s, reinterpret_cast<void**>(&s->d_kps), ((size_t)n + 1) * sizeof(RbSnipKp)));
  uint32_t* cnt = reinterpret_cast<uint32_t*>(s->d_count + 1);
  rb_snip_emit_kernel<<<s->sm_count * 4, 256, 0, s->stream>>>(g, s->d_image, s->d_kp, s->d_w2, s->d_kps, s->nkp, cnt);
  RS_CUDA(s, cudaGetLastError());
  RS_CUDA(s, cudaStreamSynchronize(s->stream));
  return RB_OK;
}

int rb_snippet_create(int device, const uint16_t* dots, uint32_t W, uint32_t H, rb_snippet** out) {
  if (!dots || !out) return RB_ERR_INVALID;
  const int rc = snippet_new(device, W, H, out);
  if (rc != RB_OK) return rc;
  rb_snippet* s = *out;
  RS_CUDA(s, cudaMemcpyAsync(s->d_dots, dots, (size_t)W * H * 32, cudaMemcpyHostToDevice, s->stream));
  return snippet_extract(s);
}

// fgm::fragment::blit(pos, fragment&&) (src/fgm.hpp:99-113) on the device: out's W x H map is zero except for a's map
// at (ax, ay) plus b's map at (bx, by), uint16 counters adding with wrap-around like the reference's `+=` on
// std::uint16_t; then extract_single over the merged map (src/fgs.hpp:146-152).  The caller computes the geometry
// (fragment::ensure / extend, src/fgm.hpp:190-233; include/fgs_b200.hpp does).  a and b stay valid.
__global__ void __launch_bounds__(256) rb_dots_place_kernel(const uint4* __restrict__ src, uint32_t sW, uint32_t sH, uint4* __restrict__ dst,
                                                            uint32_t dW, uint32_t ox, uint32_t oy, int add) {
  const size_t total = (size_t)sW * sH * 2;  // two 16-byte halves per dot
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pxl = i >> 1;
    const uint32_t y = (uint32_t)(pxl / sW), x = (uint32_t)(pxl - (size_t)y * sW);
    const size_t at = (((size_t)(y + oy) * dW + (x + ox)) << 1) | (i & 1);
    uint4 v = __ldg(src + i);
    if (add) {
      const uint4 d = dst[at];
      v.x = __vadd2(v.x, d.x); v.y = __vadd2(v.y, d.y); v.z = __vadd2(v.z, d.z); v.w = __vadd2(v.w, d.w);  // per-halfword, wrapping
    }
    dst[at] = v;
  }
}

int rb_snippet_merge(rb_snippet* a, uint32_t ax, uint32_t ay, rb_snippet* b, uint32_t bx, uint32_t by, uint32_t W, uint32_t H,
                     rb_snippet** out) {
  if (!a || !b || !out) return RB_ERR_INVALID;
  *out = nullptr;
  if (a->device != b->device) { a->err = "rb_snippet_merge: snippets on different devices"; return RB_ERR_INVALID; }
  if ((uint64_t)ax + a->g.W > W || (uint64_t)ay + a->g.H > H || (uint64_t)bx + b->g.W > W || (uint64_t)by + b->g.H > H) {
    a->err = "rb_snippet_merge: a map lies outside the merged map";
    return RB_ERR_INVALID;
  }
  const int rc = snippet_new(a->device, W, H, out);
  if (rc != RB_OK) return rc;
  rb_snippet* s = *out;
  RS_CUDA(s, cudaStreamSynchronize(a->stream));
  RS_CUDA(s, cudaStreamSynchronize(b->stream));
  RS_CUDA(s, cudaMemsetAsync(s->d_dots, 0, (size_t)W * H * 32, s->stream));
  const uint32_t grid = (uint32_t)s->sm_count * 8;
  rb_dots_place_kernel<<<grid, 256, 0, s->stream>>>(reinterpret_cast<const uint4*>(a->d_dots), a->g.W, a->g.H,
                                                    reinterpret_cast<uint4*>(s->d_dots), W, ax, ay, 0);
  rb_dots_place_kernel<<<grid, 256, 0, s->stream>>>(reinterpret_cast<const uint4*>(b->d_dots), b->g.W, b->g.H,
                                                    reinterpret_cast<uint4*>(s->d_dots), W, bx, by, 1);
  RS_CUDA(s, cudaGetLastError());
  return snippet_extract(s);
}

// The snippet's dot map (H * W * 16 uint16): what fgm::fragment::dots() holds after the merges.
int rb_snippet_fetch_dots(rb_snippet* s, uint16_t* out) {
  if (!s || !out) return RB_ERR_INVALID;
  RS_CUDA(s, cudaSetDevice(s->device));
  RS_CUDA(s, cudaMemcpyAsync(out, s->d_dots, (size_t)s->g.W * s->g.H * 32, cudaMemcpyDeviceToHost, s->stream));
  RS_CUDA(s, cudaStreamSynchronize(s->stream));
  return RB_OK;
}

int rb_snippet_fetch(rb_snippet* s, uint32_t* nkeypoints, uint8_t* out_image, uint8_t* out_mask, rb_keypoint* out_kps, size_t cap) {
  if (!s) return RB_ERR_INVALID;
  const RbGeom& g = s->g;
  RS_CUDA(s, cudaSetDevice(s->device));
  if (nkeypoints) *nkeypoints = s->nkp;
  if (out_image) RS_CUDA(s, cudaMemcpy2DAsync(out_image, g.W, s->d_image, g.pitch, g.W, g.H, cudaMemcpyDeviceToHost, s->stream));
  if (out_mask) RS_CUDA(s, cudaMemcpyAsync(out_mask, s->d_mask, (size_t)g.W * g.H, cudaMemcpyDeviceToHost, s->stream));
  std::vector<RbSnipKp> h;
  if (out_kps && s->nkp) {
    h.resize(s->nkp);
    RS_CUDA(s, cudaMemcpyAsync(h.data(), s->d_kps, (size_t)s->nkp * sizeof(RbSnipKp), cudaMemcpyDeviceToHost, s->stream));
  }
  RS_CUDA(s, cudaStreamSynchronize(s->stream));
  for (size_t i = 0; i < h.size() && i < cap; ++i) {  // parity tap: the reference's 13-byte layout (src/kpe.hpp:342-379)
    uint8_t v[25];
    for (int n = 0; n < 25; ++n) v[n] = (h[i].c[n >> 3] >> (4 * (n & 7))) & 15;
    const uint32_t weight = (h[i].c[3] >> 4) & 3;
    rb_keypoint& k = out_kps[i];
    memset(&k, 0, sizeof(k));
    auto at = [&](int r, int c) { return v[5 * r + c]; };
    k.code[0] = at(0, 0) | (at(0, 1) << 4);  k.code[1] = at(0, 2) | (at(0, 3) << 4);
    k.code[2] = at(1, 0) | (at(0, 4) << 4);  k.code[3] = at(1, 1) | (at(1, 2) << 4);
    k.code[4] = at(1, 3) | (at(1, 4) << 4);  k.code[5] = at(2, 0) | (at(2, 1) << 4);
    k.code[6] = at(2, 2) | (at(2, 3) << 4);  k.code[7] = at(3, 0) | (at(2, 4) << 4);
    k.code[8] = at(3, 1) | (at(3, 2) << 4);  k.code[9] = at(3, 3) | (at(3, 4) << 4);
    k.code[10] = at(4, 0) | (at(4, 1) << 4); k.code[11] = at(4, 2) | (at(4, 3) << 4);
    k.code[12] = (uint8_t)(weight | (at(4, 4) << 4));
    k.weight = (uint8_t)weight;
    k.x = (uint16_t)(h[i].xy & 0xFFFFu);
    k.y = (uint16_t)(h[i].xy >> 16);
    k.region_mask = 1;
  }
  return RB_OK;
}

// kpm::details::get_limits (src/kpm.hpp:301-315), size_t arithmetic included
static void cell_limits(int32_t delta, uint64_t previous, uint64_t current, uint64_t* clo, uint64_t* chi) {
  if (delta < 0) {
    const uint64_t d = (uint64_t)(-(int64_t)delta);
    *clo = d;
    *chi = current < previous + d ? current : previous + d;
  } else {
    const uint64_t d = (uint64_t)delta;
    *clo = 0;
    *chi = current < previous - d ? current : previous - d;  // wraps like the reference when d > previous
  }
}

int rb_snippet_match(rb_snippet* a, rb_snippet* b, uint32_t cell_w, uint32_t cell_h, rb_cell_match* out) {
  if (!a || !b || !out || cell_w == 0 || cell_h == 0 || cell_w > 255 || cell_h > 255) return RB_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  if (a->device != b->device) { a->err = "rb_snippet_match: snippets on different devices"; return RB_ERR_INVALID; }
  if (a->nkp == 0 || b->nkp == 0) return RB_OK;  // count_offsets finds nothing: no vote (src/kpm.hpp:379-381)
  RS_CUDA(a, cudaSetDevice(a->device));
  cudaStream_t st = a->stream;
  RS_CUDA(a, cudaStreamSynchronize(b->stream));
  RbCellParams p;
  memset(&p, 0, sizeof(p));
  p.prev = a->d_kps; p.np = a->nkp; p.curr = b->d_kps; p.nc = b->nkp;
  p.pW = a->g.W; p.pH = a->g.H; p.cW = b->g.W; p.cH = b->g.H;
  p.nbuckets = next_pow2(2 * p.np < 1024 ? 1024 : 2 * p.np);
  p.OW = p.pW + p.cW - 1; p.OH = p.pH + p.cH - 1;
  p.cell_w = cell_w; p.cell_h = cell_h;
  const size_t nbins = (size_t)p.OW * p.OH;
  p.CW = (p.pW > p.cW ? p.pW : p.cW) / cell_w + 1;
  const uint32_t CH = (p.pH > p.cH ? p.pH : p.cH) / cell_h + 1;
  const size_t cellwords = ((size_t)p.CW * CH + 31) / 32;
  p.AW = p.cW / cell_w + 1;
  const size_t actwords = ((size_t)p.AW * (p.cH / cell_h + 1) + 31) / 32;
  const size_t bytes = ((size_t)p.nbuckets + p.np + nbins + cellwords + actwords) * 4 + 64;
  if (bytes > a->scratch_cap) {  // cudaMalloc / cudaFree per match cost more than the match itself
    snip_free(a, a->d_scratch);
    a->d_scratch = nullptr;
    a->scratch_cap = 0;
    RS_CUDA(a, snip_alloc(a, reinterpret_cast<void**>(&a->d_scratch), bytes));
    a->scratch_cap = bytes;
  }
  uint8_t* mem = a->d_scratch;
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(mem);
  p.head = reinterpret_cast<uint32_t*>(mem + 64);
  p.next = p.head + p.nbuckets;
  p.hist = p.next + p.np;
  p.cellbits = p.hist + nbins;
  p.actbits = p.cellbits + cellwords;
  RS_CUDA(a, cudaMemsetAsync(mem, 0, bytes, st));
  RS_CUDA(a, cudaMemsetAsync(p.head, 0xFF, (size_t)p.nbuckets * 4, st));
  const uint32_t grid = (uint32_t)a->sm_count * 8;
  rb_cell_build_kernel<<<grid, 256, 0, st>>>(p);
  rb_cell_vote_kernel<0><<<grid, 256, 0, st>>>(p);
  rb_cell_best_kernel<<<grid, 256, 0, st>>>(p.hist, nbins, d_out);
  RS_CUDA(a, cudaGetLastError());
  unsigned long long h[8] = {0};
  RS_CUDA(a, cudaMemcpyAsync(h, d_out, 24, cudaMemcpyDeviceToHost, st));
  RS_CUDA(a, cudaStreamSynchronize(st));
  out->offsets = (uint32_t)h[1];
  out->pairs = h[2];
  if (h[1] == 0) return RB_OK;  // no code in common
  const uint32_t votes = (uint32_t)(h[0] >> 32), bin = 0xFFFFFFFFu - (uint32_t)(h[0] & 0xFFFFFFFFu);
  p.best_dx = (int32_t)(bin % p.OW) - (int32_t)(p.cW - 1);
  p.best_dy = (int32_t)(bin / p.OW) - (int32_t)(p.cH - 1);
  // count_active_cells (src/kpm.hpp:349-369): the part of curr that prev covers at this offset
  uint64_t l, r, t, bt;
  cell_limits(p.best_dx, p.pW, p.cW, &l, &r);
  cell_limits(p.best_dy, p.pH, p.cH, &t, &bt);
  p.lim_l = (uint32_t)l; p.lim_r = (uint32_t)(r > 0xFFFFFFFFull ? 0xFFFFFFFFull : r);
  p.lim_t = (uint32_t)t; p.lim_b = (uint32_t)(bt > 0xFFFFFFFFull ? 0xFFFFFFFFull : bt);
  p.pmask = a->d_mask;
  rb_cell_ties_kernel<<<grid, 256, 0, st>>>(p.hist, nbins, votes, d_out);
  rb_cell_vote_kernel<1><<<grid, 256, 0, st>>>(p);
  rb_cell_active_kernel<<<grid, 256, 0, st>>>(p);
  rb_count_kernel<<<grid, 256, 0, st>>>(p.cellbits, cellwords, d_out + 4);
  rb_count_kernel<<<grid, 256, 0, st>>>(p.actbits, actwords, d_out + 5);
  RS_CUDA(a, cudaGetLastError());
  RS_CUDA(a, cudaMemcpyAsync(h, d_out, 48, cudaMemcpyDeviceToHost, st));
  RS_CUDA(a, cudaStreamSynchronize(st));
  out->dx = p.best_dx; out->dy = p.best_dy;
  out->matched_keypoints = votes;
  out->ties = (uint32_t)h[3];
  out->matched_cells = (uint32_t)h[4];
  out->active_cells = (uint32_t)h[5];
  out->valid = !((float)out->matched_cells < (float)out->active_cells * 0.66f);  // src/kpm.hpp:387-389
  return RB_OK;
}

// aws::details::compare (src/aws.hpp:37-60) for every consecutive pair of frames [first, first + n)
int rb_aws_compare(rb_ctx* c, size_t first, size_t n, uint8_t* heat, uint32_t* first_change) {
  if (!c || (!heat && !first_change)) return RB_ERR_INVALID;
  if (n < 1 || first + n > c->uploaded) { c->err = "rb_aws_compare: frames not uploaded"; return RB_ERR_STATE; }
  const RbGeom& g = c->g;
  RB_CUDA(c, cudaSetDevice(c->device));
  const size_t px = (size_t)g.W * g.H, need = px * 5 + 256;
  if (need > c->map_cap) {
    if (c->d_map) { cudaFree(c->d_map); c->bytes -= c->map_cap; c->d_map = nullptr; }
    c->map_cap = 0;
    RB_CUDA(c, dmalloc(c, &c->d_map, need));
    c->map_cap = need;
  }
  c->map_w = c->map_h = 0;  // the scratch no longer holds a fragment map (rb_map_device / rb_blend_map must not serve it)
  uint32_t* d_fc = reinterpret_cast<uint32_t*>(c->d_map);
  uint8_t* d_heat = c->d_map + px * 4;
  if (heat) RB_CUDA(c, cudaMemcpyAsync(d_heat, heat, px, cudaMemcpyHostToDevice, c->stream));
  RB_CUDA(c, cudaMemsetAsync(d_fc, 0xFF, px * 4, c->stream));
  if (n >= 2) {
    const uint32_t total = (g.pitch / 16) * g.H, blocks = (total + 255) / 256;
    // enough segments of the pair range to put ~4 waves of threads on the GPU, at least 8 pairs each
    const uint64_t want = (uint64_t)c->sm_count * 2048 * 4;
    uint32_t segs = (uint32_t)((want + total - 1) / total);
    const uint32_t pairs = (uint32_t)n - 1;
    if (segs > (pairs + 7) / 8) segs = (pairs + 7) / 8;
    if (segs < 1) segs = 1;
    const uint32_t seg_len = (pairs + segs - 1) / segs;
    segs = (pairs + seg_len - 1) / seg_len;
    rb_aws_compare_kernel<<<dim3(blocks, segs), 256, 0, c->stream>>>(c->d_frames + g.frame_stride * first, g.pitch, g.frame_stride,
                                                                      g.W, g.H, (uint32_t)n, seg_len, heat ? d_heat : nullptr, d_fc);
  }
  RB_LAUNCHED(c, "rb_aws_compare_kernel");
  if (heat) RB_CUDA(c, cudaMemcpyAsync(heat, d_heat, px, cudaMemcpyDeviceToHost, c->stream));
  if (first_change) RB_CUDA(c, cudaMemcpyAsync(first_change, d_fc, px * 4, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

// Multi-GPU map assembly: the device addresses of the map the scratch holds, and blend over dots that the
// caller has reduced across ranks in place.
int rb_map_device(rb_ctx* c, uint16_t** dots, uint8_t** image, uint8_t** mask, uint32_t* mapW, uint32_t* mapH) {
  if (!c) return RB_ERR_INVALID;
  if (!c->d_map || c->map_w == 0) { c->err = "rb_map_device: no map assembled yet"; return RB_ERR_STATE; }
  const size_t px = (size_t)c->map_w * c->map_h;
  if (dots) *dots = reinterpret_cast<uint16_t*>(c->d_map);
  if (image) *image = c->d_map + px * 32;
  if (mask) *mask = c->d_map + px * 33;
  if (mapW) *mapW = c->map_w;
  if (mapH) *mapH = c->map_h;
  return RB_OK;
}

int rb_blend_map(rb_ctx* c, uint16_t* out_dots, uint8_t* out_image, uint8_t* out_mask) {
  if (!c) return RB_ERR_INVALID;
  if (!c->d_map || c->map_w == 0) { c->err = "rb_blend_map: no map assembled yet"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  const size_t px = (size_t)c->map_w * c->map_h;
  uint16_t* d_dots = reinterpret_cast<uint16_t*>(c->d_map);
  uint8_t* d_img = c->d_map + px * 32;
  uint8_t* d_msk = d_img + px;
  rb_blend_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_dots, c->map_w, c->map_h, d_img, c->map_w, d_msk);
  RB_LAUNCHED(c, "rb_blend_kernel");
  if (out_dots) RB_CUDA(c, cudaMemcpyAsync(out_dots, d_dots, px * 32, cudaMemcpyDeviceToHost, c->stream));
  if (out_image) RB_CUDA(c, cudaMemcpyAsync(out_image, d_img, px, cudaMemcpyDeviceToHost, c->stream));
  if (out_mask) RB_CUDA(c, cudaMemcpyAsync(out_mask, d_msk, px, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

// Fused multi-GPU map assembly: export this rank's partial map, and (on the destination rank) sum + blend over
// the peers' maps through CUDA IPC.
int rb_map_export(rb_ctx* c, rb_map_handle* out) {
  if (!c || !out) return RB_ERR_INVALID;
  if (!c->d_map || c->map_w == 0) { c->err = "rb_map_export: no map assembled yet"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  memset(out, 0, sizeof(*out));
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(out->opaque), "rb_map_handle too small");
  cudaIpcMemHandle_t h;
  RB_CUDA(c, cudaIpcGetMemHandle(&h, c->d_map));
  memcpy(out->opaque, &h, sizeof(h));
  out->map_w = c->map_w; out->map_h = c->map_h; out->device = (uint32_t)c->device;
  return RB_OK;
}

// a peer's scratch keeps its address (and handle) until it grows: mappings are opened once and kept
static int peer_address(rb_ctx* c, const rb_map_handle& h, const uint16_t** out) {
  if (h.map_w != c->map_w || h.map_h != c->map_h) { c->err = "peer map geometry differs"; return RB_ERR_INVALID; }
  const std::string key(reinterpret_cast<const char*>(h.opaque), sizeof(cudaIpcMemHandle_t));
  for (auto& kv : c->ipc_open)
    if (kv.first == key) { *out = static_cast<const uint16_t*>(kv.second); return RB_OK; }
  cudaIpcMemHandle_t ih;
  memcpy(&ih, h.opaque, sizeof(ih));
  void* addr = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&addr, ih, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { c->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); return RB_ERR_CUDA; }
  c->ipc_open.emplace_back(key, addr);
  *out = static_cast<const uint16_t*>(addr);
  return RB_OK;
}

// reduce-scatter step: this rank sums slice `self` of `world` over all other ranks' maps into its own scratch
int rb_sum_map_slice(rb_ctx* c, const rb_map_handle* ranks, size_t world, size_t self) {
  if (!c || !ranks || world < 1 || world > RB_MAX_PEERS || self >= world) return RB_ERR_INVALID;
  if (!c->d_map || c->map_w == 0) { c->err = "rb_sum_map_slice: no map assembled yet"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  RbPeerMaps pm;
  memset(&pm, 0, sizeof(pm));
  for (size_t r = 0; r < world; ++r) {
    if (r == self) continue;
    const int rc = peer_address(c, ranks[r], &pm.dots[pm.n]);
    if (rc != RB_OK) return rc;
    ++pm.n;
  }
  const size_t px = (size_t)c->map_w * c->map_h, len = (px + world - 1) / world;
  const size_t i0 = self * len < px ? self * len : px, i1 = (self + 1) * len < px ? (self + 1) * len : px;
  if (i1 > i0) {
    rb_sum_slice_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(reinterpret_cast<uint16_t*>(c->d_map), pm, i0, i1);
    RB_LAUNCHED(c, "rb_sum_slice_kernel");
  }
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

// gather step on the destination rank: pull every slice from the rank that reduced it, blend
int rb_blend_map_slices(rb_ctx* c, const rb_map_handle* ranks, size_t world, size_t self, uint16_t* out_dots, uint8_t* out_image,
                        uint8_t* out_mask) {
  if (!c || !ranks || world < 1 || world > RB_MAX_PEERS || self >= world) return RB_ERR_INVALID;
  if (!c->d_map || c->map_w == 0) { c->err = "rb_blend_map_slices: no map assembled yet"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  RbPeerMaps pm;
  memset(&pm, 0, sizeof(pm));
  pm.n = (uint32_t)world;
  for (size_t r = 0; r < world; ++r) {
    if (r == self) continue;
    const int rc = peer_address(c, ranks[r], &pm.dots[r]);
    if (rc != RB_OK) return rc;
  }
  const size_t px = (size_t)c->map_w * c->map_h, len = (px + world - 1) / world;
  uint16_t* d_dots = reinterpret_cast<uint16_t*>(c->d_map);
  uint8_t* d_img = c->d_map + px * 32;
  uint8_t* d_msk = d_img + px;
  rb_gather_blend_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_dots, pm, len, c->map_w, c->map_h, d_img, c->map_w, d_msk);
  RB_LAUNCHED(c, "rb_gather_blend_kernel");
  if (out_dots) RB_CUDA(c, cudaMemcpyAsync(out_dots, d_dots, px * 32, cudaMemcpyDeviceToHost, c->stream));
  if (out_image) RB_CUDA(c, cudaMemcpyAsync(out_image, d_img, px, cudaMemcpyDeviceToHost, c->stream));
  if (out_mask) RB_CUDA(c, cudaMemcpyAsync(out_mask, d_msk, px, cudaMemcpyDeviceToHost, c->stream));
  RB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RB_OK;
}

int rb_blend_map_peers(rb_ctx* c, const rb_map_handle* peers, size_t npeers, uint16_t* out_dots, uint8_t* out_image,
                       uint8_t* out_mask) {
  if (!c || (!peers && npeers) || npeers > RB_MAX_PEERS) return RB_ERR_INVALID;
  if (!c->d_map || c->map_w == 0) { c->err = "rb_blend_map_peers: no map assembled yet"; return RB_ERR_STATE; }
  RB_CUDA(c, cudaSetDevice(c->device));
  RbPeerMaps pm;
  memset(&pm, 0, sizeof(pm));
  for (size_t i = 0; i < npeers; ++i) {
    const int rc = peer_address(c, peers[i], &pm.dots[i]);
    if (rc != RB_OK) return rc;
  }
  pm.n = (uint32_t)npeers;
  const size_t px = (size_t)c->map_w * c->map_h;
  uint16_t* d_dots = reinterpret_cast<uint16_t*>(c->d_map);
  uint8_t* d_img = c->d_map + px * 32;
  uint8_t* d_msk = d_img + px;
  rb_blend_peers_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_dots, pm, c->map_w, c->map_h, d_img, c->map_w, d_msk);
  ++c->launches;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && out_dots) e = cudaMemcpyAsync(out_dots, d_dots, px * 32, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && out_image) e = cudaMemcpyAsync(out_image, d_img, px, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && out_mask) e = cudaMemcpyAsync(out_mask, d_msk, px, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { c->err = std::string("rb_blend_map_peers: ") + cudaGetErrorString(e); return RB_ERR_CUDA; }
  return RB_OK;
}

}  // extern "C"
