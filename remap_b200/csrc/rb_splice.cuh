// rb_splice.cuh -- fragment splicing: the device side of fgs::splice (src/fgs.hpp:187-213).
//
//   fgs::details::extract_single (src/fgs.hpp:80-89)   fragment.blend() -> image + mask, then kpe with a 1 x 1
//                                                      grid and no overlap over the whole map image;
//   kpm::match, cellular variant (src/kpm.hpp:371-393) count_offsets (all pairs of equal codes vote
//                                                      prev - curr, each vote filed under the cell of
//                                                      min(prev, curr), :231-262), find_best (offset with the
//                                                      most votes, :280-299), count_active_cells (:301-369).
//
// The reference keeps unordered_map<offset, unordered_map<cell, count>> and only ever uses the entry of the
// winning offset.  Here: a DENSE offset histogram (one word per possible offset, (Wp + Wc - 1) x (Hp + Hc - 1)
// words -- every offset of a pair lies in that box) filled by a chained hash join over the 101-bit codes,
// an arg-max pass, then a second walk of the join that files only the winner's votes into a cell bit map;
// the active cells are a third bit map.  No per-offset containers exist.
//
// Keypoint records hold the 5 x 5 patch as 25 nibbles (row-major, nibble n in word n / 8) + the weight in
// bits 4..5 of word 3: equal records <=> equal kpr::code (src/kpr.hpp:20-27, layout src/kpe.hpp:342-379 is a
// permutation of the same 25 nibbles + weight).
#pragma once

#include "rb_common.cuh"

struct RbSnipKp {
  uint32_t c[4];
  uint32_t xy;  // x | y << 16
};

struct RbCellParams {
  const RbSnipKp* prev; uint32_t np;
  const RbSnipKp* curr; uint32_t nc;
  uint32_t pW, pH, cW, cH;   // map sizes of the two snippets
  uint32_t* head;            // [nbuckets] chain heads over prev (0xFFFFFFFF = empty)
  uint32_t* next;            // [np]
  uint32_t nbuckets;         // power of two
  uint32_t* hist;            // [(pH + cH - 1)][(pW + cW - 1)] votes per offset; bin (dx + cW - 1, dy + cH - 1)
  uint32_t OW, OH;
  uint32_t cell_w, cell_h;   // kpm::cell_size_t (src/fgs.hpp:121: 15 x 15)
  int32_t best_dx, best_dy;  // second walk: the winning offset
  uint32_t* cellbits;        // matched cells of the winner: bit (cy * CW + cx)
  uint32_t CW;
  // active cells (src/kpm.hpp:321-347)
  const uint8_t* pmask;      // prev's blend mask, pW bytes per row
  uint32_t lim_l, lim_t, lim_r, lim_b;  // clim
  uint32_t* actbits;
  uint32_t AW;
};

namespace rbs {

RB_HD uint32_t code_hash(const RbSnipKp& k) {
  uint32_t h = k.c[0] * 0x9E3779B1u;
  h = (h ^ (h >> 15)) + k.c[1] * 0x85EBCA77u;
  h = (h ^ (h >> 13)) + k.c[2] * 0xC2B2AE3Du;
  h = (h ^ (h >> 16)) + k.c[3] * 0x27D4EB2Fu;
  return h ^ (h >> 15);
}
RB_HD bool same_code(const RbSnipKp& a, const RbSnipKp& b) {
  return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3];
}

// fgm::fragment::blend (src/fgm.hpp:115-135) of one map pixel: first colour with the largest count.
RB_HD void blend_pixel(const uint16_t* dot, uint8_t* image, uint8_t* mask) {
  uint32_t best = 0, bestc = 0;
  for (uint32_t c = 0; c < 16; ++c) {
    const uint32_t v = dot[c];
    if (v > best) { best = v; bestc = c; }
  }
  *image = (uint8_t)(best ? bestc : 0);
  *mask = best ? 1 : 0;
}

// The keypoints of strip word (y, j) of K1's bit maps -> records (any order; the join does not care).
RB_HD void emit_word(const RbGeom& g, const uint8_t* image, const uint32_t* kpbits, const uint32_t* w2bits, uint32_t y,
                     uint32_t j, RbSnipKp* out, uint32_t cap, uint32_t* count) {
  uint32_t w = kpbits[(uint64_t)y * g.NS + j];
  const uint32_t w2 = w2bits[(uint64_t)y * g.NS + j];
  while (w) {
    const uint32_t b = rb_ffs0(w);
    w &= w - 1;
    const uint32_t x = RB_STRIP_OUT * j + b;
    RbSnipKp k;
    k.c[0] = k.c[1] = k.c[2] = k.c[3] = 0;
    for (uint32_t r = 0; r < 5; ++r)
      for (uint32_t c = 0; c < 5; ++c) {
        const uint32_t n = 5 * r + c;
        k.c[n >> 3] |= (uint32_t)(image[(uint64_t)(y - 2 + r) * g.pitch + (x - 2 + c)] & 15) << (4 * (n & 7));
      }
    k.c[3] |= (((w2 >> b) & 1u) ? 2u : 1u) << 4;  // kpr::weight, low nibble of code byte 12 (src/kpr.hpp:25-27)
    k.xy = x | (y << 16);
    const uint32_t at = rb_atomic_add(count, 1u);
    if (at < cap) out[at] = k;
  }
}

// chain insert of prev keypoint i
RB_HD void build_item(const RbCellParams& p, uint32_t i) {
  const uint32_t b = code_hash(p.prev[i]) & (p.nbuckets - 1);
#if defined(__CUDA_ARCH__)
  p.next[i] = atomicExch(p.head + b, i);
#else
  p.next[i] = p.head[b];
  p.head[b] = i;
#endif
}

// kpm::details::get_offsets for curr keypoint j against every prev keypoint of the same code
// (src/kpm.hpp:231-247).  WALK 0: ++hist[offset].  WALK 1: votes of the winning offset mark their cell
// (to_cell = min(prev, curr) / cell, src/kpm.hpp:225-229).
template <int WALK>
RB_HD void vote_item(const RbCellParams& p, uint32_t j) {
  const RbSnipKp c = p.curr[j];
  const int32_t cx = (int32_t)(c.xy & 0xFFFFu), cy = (int32_t)(c.xy >> 16);
  for (uint32_t i = p.head[code_hash(c) & (p.nbuckets - 1)]; i != 0xFFFFFFFFu; i = p.next[i]) {
    const RbSnipKp a = p.prev[i];
    if (!same_code(a, c)) continue;
    const int32_t px = (int32_t)(a.xy & 0xFFFFu), py = (int32_t)(a.xy >> 16);
    const int32_t ox = px - cx, oy = py - cy;
    if (WALK == 0) {
      rb_atomic_add(p.hist + (uint64_t)(oy + (int32_t)p.cH - 1) * p.OW + (uint32_t)(ox + (int32_t)p.cW - 1), 1u);
    } else if (ox == p.best_dx && oy == p.best_dy) {
      const uint32_t cellx = (uint32_t)(px < cx ? px : cx) / p.cell_w, celly = (uint32_t)(py < cy ? py : cy) / p.cell_h;
      const uint32_t bit = celly * p.CW + cellx;
#if defined(__CUDA_ARCH__)
      atomicOr(p.cellbits + (bit >> 5), 1u << (bit & 31));
#else
      p.cellbits[bit >> 5] |= 1u << (bit & 31);
#endif
    }
  }
}

// kpm::details::filter_keypoints (src/kpm.hpp:321-347) for curr keypoint j: inside clim, and prev's mask is
// set at point + delta -> its cell, relative to clim's corner, is active.
RB_HD void active_item(const RbCellParams& p, uint32_t j) {
  const uint32_t xy = p.curr[j].xy;
  const uint32_t x = xy & 0xFFFFu, y = xy >> 16;
  if (x < p.lim_l || x >= p.lim_r || y < p.lim_t || y >= p.lim_b) return;  // region::contains (src/cdt.hpp:268-272)
  const int64_t idx = (int64_t)p.pW * ((int32_t)y + p.best_dy) + ((int32_t)x + p.best_dx);  // cdt::to_index
  if (idx < 0 || idx >= (int64_t)p.pW * p.pH) return;  // cannot happen for an offset of a real pair
  if (p.pmask[idx] == 0) return;
  const uint32_t bit = ((y - p.lim_t) / p.cell_h) * p.AW + (x - p.lim_l) / p.cell_w;
#if defined(__CUDA_ARCH__)
  atomicOr(p.actbits + (bit >> 5), 1u << (bit & 31));
#else
  p.actbits[bit >> 5] |= 1u << (bit & 31);
#endif
}

}  // namespace rbs

#if defined(__CUDACC__)

__global__ void rb_blend_kernel(const uint16_t* __restrict__ dots, uint32_t W, uint32_t H, uint8_t* __restrict__ image,
                                uint32_t pitch, uint8_t* __restrict__ mask) {
  const size_t total = (size_t)W * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i - (size_t)y * W);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(dots + i * 16)), b = __ldg(reinterpret_cast<const uint4*>(dots + i * 16) + 1);
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint16_t d[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[2 * k] = (uint16_t)(v[k] & 0xFFFFu); d[2 * k + 1] = (uint16_t)(v[k] >> 16); }
    rbs::blend_pixel(d, image + (size_t)y * pitch + x, mask + i);
  }
}

// Multi-GPU map assembly as ONE kernel: rank 0 sums its own partial dot map and the peers' partial maps -- read
// in place over NVLink through CUDA-IPC mappings of their map scratch -- and blends in the same pass
// (fgm::fragment::blend, src/fgm.hpp:115-135).  The uint16 counters add per 16-bit lane and wrap like the
// reference's (src/fgm.hpp:12-14,94).  32 B per map pixel and peer cross the link, once.
#define RB_MAX_PEERS 15
struct RbPeerMaps {
  const uint16_t* dots[RB_MAX_PEERS];
  uint32_t n;
};
__device__ __forceinline__ uint32_t rb_add16x2(uint32_t a, uint32_t b) {  // two independent uint16 sums, each mod 65,536
  return ((a & 0x7FFF7FFFu) + (b & 0x7FFF7FFFu)) ^ ((a ^ b) & 0x80008000u);
}
__global__ void rb_blend_peers_kernel(uint16_t* __restrict__ dots, const RbPeerMaps peers, uint32_t W, uint32_t H,
                                      uint8_t* __restrict__ image, uint32_t pitch, uint8_t* __restrict__ mask) {
  const size_t total = (size_t)W * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    uint4* own = reinterpret_cast<uint4*>(dots + i * 16);
    uint4 a = own[0], b = own[1];
    for (uint32_t r = 0; r < peers.n; ++r) {
      const uint4* q = reinterpret_cast<const uint4*>(peers.dots[r] + i * 16);
      const uint4 pa = q[0], pb = q[1];  // peer memory: plain loads over NVLink
      a.x = rb_add16x2(a.x, pa.x); a.y = rb_add16x2(a.y, pa.y); a.z = rb_add16x2(a.z, pa.z); a.w = rb_add16x2(a.w, pa.w);
      b.x = rb_add16x2(b.x, pb.x); b.y = rb_add16x2(b.y, pb.y); b.z = rb_add16x2(b.z, pb.z); b.w = rb_add16x2(b.w, pb.w);
    }
    own[0] = a; own[1] = b;
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint16_t d[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[2 * k] = (uint16_t)(v[k] & 0xFFFFu); d[2 * k + 1] = (uint16_t)(v[k] >> 16); }
    const uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i - (size_t)y * W);
    rbs::blend_pixel(d, image + (size_t)y * pitch + x, mask + i);
  }
}

// The same reduction spread over all links (reduce-scatter, then gather): every rank sums ONE slice of the map --
// pixels [i0, i1) -- over all peers into its own scratch (rb_sum_slice_kernel, all NVLinks busy at once), then the
// destination rank pulls each slice from the rank that reduced it and blends (rb_gather_blend_kernel).  With the
// single-destination kernel above, 7 x 275 MB enter one GPU; here 7/8 of one map does, twice.
__global__ void rb_sum_slice_kernel(uint16_t* __restrict__ dots, const RbPeerMaps peers, size_t i0, size_t i1) {
  for (size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (size_t)gridDim.x * blockDim.x) {
    uint4* own = reinterpret_cast<uint4*>(dots + i * 16);
    uint4 a = own[0], b = own[1];
    for (uint32_t r = 0; r < peers.n; ++r) {
      const uint4* q = reinterpret_cast<const uint4*>(peers.dots[r] + i * 16);
      const uint4 pa = q[0], pb = q[1];
      a.x = rb_add16x2(a.x, pa.x); a.y = rb_add16x2(a.y, pa.y); a.z = rb_add16x2(a.z, pa.z); a.w = rb_add16x2(a.w, pa.w);
      b.x = rb_add16x2(b.x, pb.x); b.y = rb_add16x2(b.y, pb.y); b.z = rb_add16x2(b.z, pb.z); b.w = rb_add16x2(b.w, pb.w);
    }
    own[0] = a; own[1] = b;
  }
}
// owners.dots[k] = scratch of the rank that reduced slice k (nullptr: this rank); slice k = pixels [k * len, (k + 1) * len)
__global__ void rb_gather_blend_kernel(uint16_t* __restrict__ dots, const RbPeerMaps owners, size_t slice_len, uint32_t W,
                                       uint32_t H, uint8_t* __restrict__ image, uint32_t pitch, uint8_t* __restrict__ mask) {
  const size_t total = (size_t)W * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint16_t* src = owners.dots[i / slice_len];
    uint4* own = reinterpret_cast<uint4*>(dots + i * 16);
    uint4 a, b;
    if (src) {
      a = reinterpret_cast<const uint4*>(src + i * 16)[0];
      b = reinterpret_cast<const uint4*>(src + i * 16)[1];
      own[0] = a; own[1] = b;
    } else {
      a = own[0]; b = own[1];
    }
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint16_t d[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[2 * k] = (uint16_t)(v[k] & 0xFFFFu); d[2 * k + 1] = (uint16_t)(v[k] >> 16); }
    const uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i - (size_t)y * W);
    rbs::blend_pixel(d, image + (size_t)y * pitch + x, mask + i);
  }
}

__global__ void rb_snip_emit_kernel(const RbGeom g, const uint8_t* __restrict__ image, const uint32_t* __restrict__ kpbits,
                                    const uint32_t* __restrict__ w2bits, RbSnipKp* out, uint32_t cap, uint32_t* count) {
  const uint32_t total = g.H * g.NS;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    rbs::emit_word(g, image, kpbits, w2bits, i / g.NS, i % g.NS, out, cap, count);
}

__global__ void rb_cell_build_kernel(const RbCellParams p) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.np; i += gridDim.x * blockDim.x) rbs::build_item(p, i);
}

template <int WALK>
__global__ void rb_cell_vote_kernel(const RbCellParams p) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < p.nc; j += gridDim.x * blockDim.x) rbs::vote_item<WALK>(p, j);
}

__global__ void rb_cell_active_kernel(const RbCellParams p) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < p.nc; j += gridDim.x * blockDim.x) rbs::active_item(p, j);
}

// kpm::details::find_best (src/kpm.hpp:280-299) over the dense histogram: out[0] = max over bins of
// (votes << 32 | ~bin) -- the most votes, ties to the smallest bin (smallest dy, then dx; the reference takes
// the first maximum in unordered_map order, so a tie is reported, see out[3]) --, out[1] = non-empty bins,
// out[2] = all votes.
__global__ void rb_cell_best_kernel(const uint32_t* __restrict__ hist, size_t nbins, unsigned long long* out) {
  unsigned long long best = 0, nz = 0, sum = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t v = __ldg(hist + i);
    if (v) {
      ++nz;
      sum += v;
      const unsigned long long key = ((unsigned long long)v << 32) | (0xFFFFFFFFu - (uint32_t)i);
      if (key > best) best = key;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ob = __shfl_down_sync(0xffffffffu, best, o);
    if (ob > best) best = ob;
    nz += __shfl_down_sync(0xffffffffu, nz, o);
    sum += __shfl_down_sync(0xffffffffu, sum, o);
  }
  if ((threadIdx.x & 31) == 0 && nz) {
    atomicMax(out, best);
    atomicAdd(out + 1, nz);
    atomicAdd(out + 2, sum);
  }
}
// out[3] = bins holding exactly `votes` votes
__global__ void rb_cell_ties_kernel(const uint32_t* __restrict__ hist, size_t nbins, uint32_t votes, unsigned long long* out) {
  unsigned long long n = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (size_t)gridDim.x * blockDim.x)
    n += __ldg(hist + i) == votes;
  for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(out + 3, n);
}

#endif  // __CUDACC__
