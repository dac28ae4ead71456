// rb_kpm_fast.cuh -- K2 (pipelined): per-region matching + offset voting for RUNS of consecutive
// frames.  Replaces, like rb_kpm.cuh, kpm::details::cast_vote / count_offsets / get_offsets /
// top_offsets (src/kpm.hpp:91-159,213-223) under kpm::match (src/kpm.hpp:395-415).
//
// The reference rebuilds, per frame, an unordered_map<code, points> per region (kpr::region,
// src/kpr.hpp:93-156) and joins the current frame's map against the previous frame's.  A frame is
// therefore "current" once and "previous" once.  Here a persistent CTA takes one region and a run of
// consecutive frames and walks the frames in order; per frame it
//   1. receives the region's tile of the packed 4 bit/pixel frame (TMA, cp.async.bulk.tensor.3d) and
//      the region's keypoint list (cp.async.bulk) in shared memory -- both were requested two steps
//      earlier and signal one mbarrier, so no thread waits on HBM in steady state;
//   2. computes every keypoint's code (the 5x5 patch as 100 bits: src/kpe.hpp:342-379 is a bijective
//      packing of the 25 nibbles, the weight nibble is a function of the patch) ONCE, stores it,
//      inserts it into this frame's hash table and, in the same pass, probes the previous frame's
//      table, verifies the full code on a tag hit and votes prev - curr (src/kpm.hpp:96-98) into an
//      offset table (one word per bin: offset id | count);
//   3. selects the ticket (top region_votes bins by count desc, dx asc, dy asc -- a DEFINED order
//      where the reference has unordered_map iteration order, src/kpm.hpp:134-138) with warp-wide
//      REDUX reductions, counts the bins tied with each ticket entry, writes the ballot;
//   4. clears the tables it no longer needs and requests the tile + list of frame f + 2.
// The frame's table and code list then serve as "previous" in the next step.
//
// Bounded shared memory: a list longer than `cap`, or more distinct offsets than the offset table
// holds, DEFERS that (pair, region) to the general kernel (rb_kpm.cuh, exact for any input) through a
// device-side work list; nothing is approximated.
//
// Device-only (TMA, mbarrier, REDUX): its results are checked on the GPU against the oracle and
// against the general kernel, which also has a host build (tests/emul).
#pragma once

#include "rb_common.cuh"

#if defined(__CUDACC__)

#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#ifndef RB_FAST_NWP
#define RB_FAST_NWP 23  // warps that build / probe / vote (768 threads with the ballot warp: 40 registers at 2 CTAs per SM)
#endif
#define RB_FAST_NTP (32 * RB_FAST_NWP)
#define RB_FAST_NT (RB_FAST_NTP + 32)  // + one warp that turns finished offset tables into ballots

struct RbKpmFastParams {
  RbGeom g;
  const uint32_t* lists;   // [frame][region][lcap]  (rb_prep.cuh)
  const uint2* counts;     // [frame][region] (n_all, n_w2)
  RbRegionVote* votes;     // [npairs][nreg]
  uint32_t first_frame, npairs;
  uint32_t cap;            // entries of one frame's region that can take part (shared-memory lists, <= 2047)
  uint32_t lcap;           // entries per row of `lists` (>= cap, multiple of 4); a row is complete iff n_all <= lcap
  uint32_t tslots;         // code table slots, power of two >= 2 * cap
  uint32_t oslots;         // offset table slots, power of two
  uint32_t run;            // pairs per work item
  uint32_t box_x, box_y;   // TMA box: bytes per tile row, rows per box
  uint32_t nbox_y;         // boxes stacked vertically per tile (tile rows = nbox_y * box_y)
  uint32_t dybits, offbits;  // offset id = (dx + W) << dybits | (dy + H), offbits bits in all
  uint32_t bias_x, bias_y;   // rb_kpm_big_kernel only: offset id = (dx + bias_x) << dybits | (dy + bias_y)
  uint32_t epoch0;           // rb_kpm_big_kernel only: first bucket-head epoch of a launch (1; tests start near the wrap)
  const uint2* items;      // nullptr: work item i = (region i % nreg, run i / nreg).  Otherwise a second pass
  const uint32_t* nitems;  //   over (pair, region) entries another launch deferred: item i = items[i], one pair each
  uint32_t* work_counter;  // [0] work items, [2] error word; zeroed before launch
  uint32_t* deferred_count;
  uint2* deferred;         // (pair, region)
  uint32_t deferred_cap;
};

namespace rbf {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t NONE = 0x80000000u;  // "no offset": never a valid offset id (offbits <= 24)
constexpr uint32_t MAXPROBE = 96;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded: a lost transaction must not hang the GPU.  On time-out the error word is set, the CTA
// carries on with whatever is in shared memory, and the host reports RB_ERR_CUDA for the call.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t* error_word) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 22)) { *error_word = 1u; break; }
}
// NOTE (measured on B200): the innermost TMA coordinate times the element size must be a multiple
// of 16 bytes; an unaligned start raises "illegal instruction".  Tiles therefore start at 32-pixel
// boundaries of the 4 bit/pixel store.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint32_t x, uint32_t y, uint32_t z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Smem {           // stage k of a per-frame array lives at base + k * stride
  uint8_t* tile;        // [2][tile_rows][box_x] packed 4 bit/pixel, + 16 bytes of slack
  uint32_t* plist;      // [2][cap] positions as delivered by the bulk copy
  uint4* ents;          // [2][cap] (c0, c1, c2, x << dybits | y | c3 << 28): the frame's codes, kept for its turn as "previous"
  uint16_t* next;       // [2][cap] chain links: next entry of the same frame in the same bucket, or NIL16
  uint32_t* head;       // [3][tslots] first entry of each bucket's chain, or NIL
  uint32_t* otab;       // [2][oslots] offset id << cntbits | count, or EMPTY
  uint16_t* touched;    // [2][oslots] slots claimed in otab
  uint32_t* plan;       // [run + 3] per-step control words of the current work item (see PLAN_*)
  uint32_t* ctl;        // [0..1] touched counts, [2..3] overflow flags, [4] work item
  uint64_t* mbar;       // [2]
  uint32_t tile_stride, plist_stride;  // bytes / words between the stages
};

// plan word of step t (frame fa + t): entries taking part | flags | weight-2 entries of the frame
constexpr uint32_t PLAN_LMASK = 0x7FFu, PLAN_FITS = 1u << 11, PLAN_PAIR_OK = 1u << 12, PLAN_USE_ALL = 1u << 13;
constexpr uint32_t PLAN_NW2_SHIFT = 16;

__host__ __device__ inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline size_t tile_bytes(const RbKpmFastParams& p) {
  return align_up_sz((size_t)p.box_x * p.box_y * p.nbox_y + 16, 128);
}

__host__ __device__ inline size_t smem_bytes(const RbKpmFastParams& p) {
  size_t b = 0;
  b += 2 * tile_bytes(p);
  b += 2 * align_up_sz((size_t)p.cap * 4, 128);
  b += 2 * (size_t)p.cap * 16;
  b += align_up_sz(2 * (size_t)p.cap * 2, 16);
  b += 3 * (size_t)p.tslots * 4;
  b += 2 * (size_t)p.oslots * 4;
  b += 2 * (size_t)p.oslots * 2;
  b += align_up_sz(((size_t)p.run + 3) * 4, 16);
  b += 8 * 4 + 2 * 8;
  return b;
}

// Pointers are formed as (shared array + offset) so that the compiler keeps them in the shared
// state space (LDS / STS / ATOMS instead of generic accesses).  The dynamic shared array is declared
// 128-byte aligned, which the TMA destination needs.
__device__ __forceinline__ void carve(const RbKpmFastParams& p, uint8_t* base, Smem& s) {
  size_t o = 0;
  s.tile_stride = (uint32_t)tile_bytes(p);
  s.plist_stride = (uint32_t)(align_up_sz((size_t)p.cap * 4, 128) / 4);
  s.tile = base + o; o += 2 * tile_bytes(p);
  s.plist = reinterpret_cast<uint32_t*>(base + o); o += 2 * align_up_sz((size_t)p.cap * 4, 128);
  s.ents = reinterpret_cast<uint4*>(base + o); o += 2 * (size_t)p.cap * 16;
  s.next = reinterpret_cast<uint16_t*>(base + o); o += align_up_sz(2 * (size_t)p.cap * 2, 16);
  s.head = reinterpret_cast<uint32_t*>(base + o); o += 3 * (size_t)p.tslots * 4;
  s.otab = reinterpret_cast<uint32_t*>(base + o); o += 2 * (size_t)p.oslots * 4;
  s.touched = reinterpret_cast<uint16_t*>(base + o); o += 2 * (size_t)p.oslots * 2;
  s.plan = reinterpret_cast<uint32_t*>(base + o); o += align_up_sz(((size_t)p.run + 3) * 4, 16);
  s.ctl = reinterpret_cast<uint32_t*>(base + o); o += 8 * 4;
  s.mbar = reinterpret_cast<uint64_t*>(base + o);
}

struct Code { uint32_t c0, c1, c2, c3; };

// The 5x5 patch whose top-left pixel is (lx, ly) in tile coordinates, as 100 bits (row r, column c
// of the patch = nibble 5r + c).  wpr = words per tile row.
__device__ __forceinline__ Code code_at(const uint32_t* tile, uint32_t wpr, uint32_t lx, uint32_t ly) {
  uint32_t r[5];
  const uint32_t sh = (lx & 7) * 4;
  const uint32_t* row = tile + ly * wpr + (lx >> 3);
#pragma unroll
  for (int k = 0; k < 5; ++k) r[k] = __funnelshift_r(row[k * wpr], row[k * wpr + 1], sh) & 0xFFFFFu;
  Code c;
  c.c0 = r[0] | (r[1] << 20);
  c.c1 = (r[1] >> 12) | (r[2] << 8) | (r[3] << 28);
  c.c2 = (r[3] >> 4) | (r[4] << 16);
  c.c3 = r[4] >> 16;
  return c;
}

__device__ __forceinline__ uint32_t code_hash(const Code& c) {
  uint32_t h = c.c0 * 0x9E3779B1u ^ c.c1 * 0x85EBCA77u ^ c.c2 * 0xC2B2AE3Du ^ c.c3 * 0x27D4EB2Fu;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
  return h;
}

__device__ __forceinline__ uint32_t off_hash(uint32_t key) {
  uint32_t h = key * 0x9E3779B1u;
  h ^= h >> 15;
  return h;
}

// sorted insertion of k into (t0 >= t1 >= t2)
__device__ __forceinline__ void top3_insert(uint32_t k, uint32_t& t0, uint32_t& t1, uint32_t& t2) {
  if (k > t2) {
    if (k > t1) {
      t2 = t1;
      if (k > t0) { t1 = t0; t0 = k; } else t1 = k;
    } else t2 = k;
  }
}

// three rounds of "warp maximum, owner pops": the warp's three largest keys (keys are distinct)
__device__ __forceinline__ void warp_top3(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t& m0, uint32_t& m1, uint32_t& m2) {
  m0 = __reduce_max_sync(0xffffffffu, t0);
  if (t0 == m0) { t0 = t1; t1 = t2; t2 = 0; }
  m1 = __reduce_max_sync(0xffffffffu, t0);
  if (t0 == m1) { t0 = t1; t1 = t2; }
  m2 = __reduce_max_sync(0xffffffffu, t0);
}

constexpr uint32_t NIL = 0xFFFFFFFFu;

}  // namespace rbf

__global__ void __launch_bounds__(RB_FAST_NT, 2) rb_kpm_fast_kernel(const __grid_constant__ CUtensorMap tmap, const RbKpmFastParams p) {
  using namespace rbf;
  extern __shared__ __align__(128) uint8_t rb_fast_smem[];
  Smem s;
  carve(p, rb_fast_smem, s);
  const RbGeom& g = p.g;
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  constexpr uint32_t NT = RB_FAST_NT, NTP = RB_FAST_NTP;
  const bool ballot_warp = tid >= NTP;
  const uint32_t runs = (p.npairs + p.run - 1) / p.run;
  uint32_t nitems = runs * g.nreg;
  if (p.items) { nitems = *p.nitems; if (nitems > p.deferred_cap) nitems = p.deferred_cap; }

  // ---- one-off set-up: barriers, clean tables -------------------------------------------------
  if (tid == 0) {
    mbar_init(&s.mbar[0], 1);
    mbar_init(&s.mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (uint32_t i = tid; i < 3 * p.tslots; i += NT) s.head[i] = NIL;
  for (uint32_t i = tid; i < 2 * p.oslots; i += NT) s.otab[i] = EMPTY;
  if (tid < 8) s.ctl[tid] = 0;
  __syncthreads();
  uint32_t ph0 = 0, ph1 = 0;  // mbarrier phase parities (persist across work items)

  // ---- the ballot warp's job: one finished offset table -> one RbRegionVote ----------------------
  // t = step whose votes are in the table (pair = frames fa + t - 1, fa + t)
  auto make_ballot = [&](uint32_t t, uint32_t fa, uint32_t region) {
    const uint32_t par = t & 1, w = s.plan[t];
    const uint32_t cntbits = 32 - p.offbits, cntmask = (1u << cntbits) - 1u, offmask = (1u << p.offbits) - 1u;
    uint32_t* otab = s.otab + par * p.oslots;
    const uint16_t* touched = s.touched + par * p.oslots;
    const uint32_t nt = s.ctl[par];
    const bool overflow = s.ctl[2 + par] != 0;
    const uint32_t pairidx = fa + t - 1 - p.first_frame;
    if ((w & PLAN_PAIR_OK) && !overflow) {
      uint32_t t0 = 0, t1 = 0, t2 = 0;
      for (uint32_t j = lane; j < nt; j += 32) {
        const uint32_t v = otab[touched[j]];
        // larger key = earlier in the ticket: count desc, then offset id asc (dx asc, dy asc)
        top3_insert(((v & cntmask) << p.offbits) | (offmask - (v >> cntbits)), t0, t1, t2);
      }
      uint32_t g0, g1, g2;
      warp_top3(t0, t1, t2, g0, g1, g2);
      const uint32_t c0 = g0 >> p.offbits, c1 = g1 >> p.offbits, c2 = g2 >> p.offbits;
      uint32_t e01 = 0, e2 = 0;  // bins tied with ticket 0 / 1 (16 bits each) and 2
      uint32_t hh = 0;           // digest of the histogram (rb_bin_hash summed over the bins)
      for (uint32_t j = lane; j < nt; j += 32) {
        const uint32_t sl = touched[j];
        const uint32_t v = otab[sl], c = v & cntmask, oid = v >> cntbits;
        e01 += (c == c0 ? 1u : 0u) + (c == c1 ? 0x10000u : 0u);
        e2 += c == c2;
        hh += rb_bin_hash((int32_t)(oid >> p.dybits) - (int32_t)g.W, (int32_t)(oid & ((1u << p.dybits) - 1u)) - (int32_t)g.H, c);
        otab[sl] = EMPTY;
      }
      e01 = __reduce_add_sync(0xffffffffu, e01);
      e2 = __reduce_add_sync(0xffffffffu, e2);
      hh = __reduce_add_sync(0xffffffffu, hh);
      if (lane == 0) {
        const uint32_t rv = g.region_votes;
        const uint2 cp = __ldg(p.counts + (uint64_t)(fa + t - 1) * g.nreg + region);
        const uint2 cc = __ldg(p.counts + (uint64_t)(fa + t) * g.nreg + region);
        const uint32_t E0 = e01 & 0xFFFFu, E1 = e01 >> 16, E2 = e2;
        RbRegionVote vt;
        vt.use_all = (w & PLAN_USE_ALL) ? 1u : 0u;
        vt.n_prev = cp.x; vt.n_curr = cc.x; vt.w2_prev = cp.y; vt.w2_curr = cc.y;
        vt.nbins = nt;
        vt.hist_hash = hh;
        vt.nticket = nt < rv ? nt : rv;
        const uint32_t gk[3] = {g0, g1, g2}, ck[3] = {c0, c1, c2};
        // bins with a larger / larger-or-equal count than ticket k (counts are sorted c0 >= c1 >= c2)
        uint32_t ngt[3], nge[3];
        ngt[0] = 0; nge[0] = E0;
        ngt[1] = c1 == c0 ? 0u : nge[0]; nge[1] = ngt[1] + E1;
        ngt[2] = c2 == c1 ? ngt[1] : nge[1]; nge[2] = ngt[2] + E2;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
          RbBin b; b.dx = 0; b.dy = 0; b.cnt = 0;
          vt.ticket[k] = b; vt.ngt[k] = 0; vt.nge[k] = 0;
          if (k < 3 && k < vt.nticket) {
            const uint32_t oid = offmask - (gk[k] & offmask);
            vt.ticket[k].dx = (int32_t)(oid >> p.dybits) - (int32_t)g.W;
            vt.ticket[k].dy = (int32_t)(oid & ((1u << p.dybits) - 1u)) - (int32_t)g.H;
            vt.ticket[k].cnt = ck[k];
            vt.ngt[k] = ngt[k]; vt.nge[k] = nge[k];
          }
        }
        p.votes[(uint64_t)pairidx * g.nreg + region] = vt;
      }
    } else {
      // the pair cannot be finished here (a list or the offset table does not fit): defer it
      if (lane == 0) {
        const uint32_t at = atomicAdd(p.deferred_count, 1u);
        if (at < p.deferred_cap) p.deferred[at] = make_uint2(pairidx, region);
      }
      if (overflow) { for (uint32_t i = lane; i < p.oslots; i += 32) otab[i] = EMPTY; }
      else { for (uint32_t j = lane; j < nt; j += 32) otab[touched[j]] = EMPTY; }
    }
    __syncwarp();
    if (lane == 0) { s.ctl[par] = 0; s.ctl[2 + par] = 0; }
  };

  for (;;) {
    if (tid == 0) s.ctl[4] = atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const uint32_t item = s.ctl[4];
    if (item >= nitems) break;
    // regions vary fastest so that the CTAs running at the same time share frames in L2
    uint32_t region = item % g.nreg, pa = (item / g.nreg) * p.run;
    uint32_t pb = pa + p.run < p.npairs ? pa + p.run : p.npairs;  // pairs [pa, pb)
    if (p.items) { const uint2 it = p.items[item]; pa = it.x; pb = pa + 1; region = it.y; }
    const uint32_t fa = p.first_frame + pa;                             // frames fa .. fa + nsteps - 1
    const uint32_t nsteps = pb - pa + 1;
    const uint32_t cs = region / g.grid_h, rs = region % g.grid_h;     // idx = grid_h*col + row (src/kpr.hpp:71-74)
    const uint32_t X0 = g.col0[cs], Y0 = g.row0[rs];
    const uint32_t tx0 = (X0 - 2) & ~31u;  // tile origin (pixels), see tma_load_3d

    // ---- the run's plan: one control word per step, computed once, by one thread per step ----------
    if (tid < nsteps) {
      const uint2* cnts = p.counts + (uint64_t)fa * g.nreg + region;
      const uint32_t t = tid, ws = g.weight_switch;
      auto load = [&](int k) { return k >= 0 && k < (int)nsteps ? __ldg(cnts + (uint64_t)k * g.nreg) : make_uint2(0, 0); };
      const uint2 cm1 = load((int)t - 1), c0 = load((int)t), cp1 = load((int)t + 1);
      // weight switch of a pair (a, b): src/kpm.hpp:219-220 ('<' on previous, '<=' on current)
      auto sw = [&](const uint2& a, const uint2& b) { return a.y < ws || b.y <= ws; };
      const bool ua_0 = t >= 1 && sw(cm1, c0);                 // pair (t - 1, t)
      const bool ua_p1 = t + 1 < nsteps && sw(c0, cp1);        // pair (t, t + 1)
      const uint32_t L = (ua_0 || ua_p1) ? c0.x : c0.y;        // entries of this frame that take part
      const uint32_t Lm1 = ((t >= 2 && sw(load((int)t - 2), cm1)) || ua_0) ? cm1.x : cm1.y;  // ... of the previous frame
      // a list row is complete only if all of the region's keypoints fitted into it
      const bool fits = c0.x <= p.lcap && L <= p.cap, fits_m1 = cm1.x <= p.lcap && Lm1 <= p.cap;
      uint32_t w = (fits ? (L | PLAN_FITS | (c0.y << PLAN_NW2_SHIFT)) : 0u) | (ua_0 ? PLAN_USE_ALL : 0u);
      if (t >= 1 && fits && fits_m1) w |= PLAN_PAIR_OK;
      s.plan[t] = w;
    }
    __syncthreads();

    auto request = [&](uint32_t t) {  // thread 0 only: tile + list of step t into stage t & 1
      const uint32_t stage = t & 1, frame = fa + t, w = s.plan[t];
      // weight-2 entries sit at the front of the list row, weight-1 entries (if this frame needs them) at its end
      const uint32_t L = w & PLAN_LMASK, nw2 = (w >> PLAN_NW2_SHIFT) & PLAN_LMASK;
      const uint32_t bytes2 = (w & PLAN_FITS) ? (nw2 * 4 + 15) & ~15u : 0u;
      const uint32_t a1 = (p.cap - (L - nw2)) & ~3u;  // shared-memory index; the row's index is a1 + lcap - cap
      const uint32_t bytes1 = ((w & PLAN_FITS) && L > nw2) ? (p.cap - a1) * 4 : 0u;
      mbar_expect_tx(&s.mbar[stage], p.box_x * p.box_y * p.nbox_y + bytes2 + bytes1);
      for (uint32_t b = 0; b < p.nbox_y; ++b)
        tma_load_3d(s.tile + stage * s.tile_stride + b * p.box_x * p.box_y, &tmap, tx0 / 2, Y0 - 2 + b * p.box_y, frame,
                    &s.mbar[stage]);
      const uint32_t* row = p.lists + ((uint64_t)frame * g.nreg + region) * p.lcap;
      if (bytes2) bulk_load(s.plist + stage * s.plist_stride, row, bytes2, &s.mbar[stage]);
      if (bytes1) bulk_load(s.plist + stage * s.plist_stride + a1, row + a1 + (p.lcap - p.cap), bytes1, &s.mbar[stage]);
    };
    if (tid == 0) {
      request(0);
      if (nsteps > 1) request(1);
    }
    uint32_t tb = 0;  // bucket table built in this step; probed table = tb - 1, table to clear = tb + 1 (mod 3)

    for (uint32_t t = 0; t < nsteps; ++t) {
      const uint32_t st = t & 1;
      if (!ballot_warp) {
        const uint32_t w = s.plan[t];
        const uint32_t tp = tb >= 1 ? tb - 1 : 2, tc = tb == 2 ? 0 : tb + 1;
        // the table that will be built next step still holds the chains of two steps ago
        {
          uint4* hc = reinterpret_cast<uint4*>(s.head + tc * p.tslots);
          const uint32_t nq = p.tslots / 4;
          if (nq <= NTP) { if (tid < nq) hc[tid] = make_uint4(NIL, NIL, NIL, NIL); }
          else for (uint32_t i = tid; i < nq; i += NTP) hc[i] = make_uint4(NIL, NIL, NIL, NIL);
        }
        if (st == 0) { mbar_wait(&s.mbar[0], ph0, p.work_counter + 2); ph0 ^= 1; }
        else { mbar_wait(&s.mbar[1], ph1, p.work_counter + 2); ph1 ^= 1; }

        // ---- codes, build this frame's buckets, probe the previous frame's, vote -------------------
        if (w & PLAN_FITS) {
          const uint32_t L = w & PLAN_LMASK, nw2 = (w >> PLAN_NW2_SHIFT) & PLAN_LMASK;
          const bool pair_ok = (w & PLAN_PAIR_OK) != 0, use_all = (w & PLAN_USE_ALL) != 0;
          const uint32_t tmask = p.tslots - 1, omask = p.oslots - 1, cntbits = 32 - p.offbits, wpr = p.box_x / 4;
          const uint32_t* tile = reinterpret_cast<const uint32_t*>(s.tile + st * s.tile_stride);
          const uint32_t* plist = s.plist + st * s.plist_stride;
          uint4* ents = s.ents + st * p.cap;
          uint16_t* next = s.next + st * p.cap;
          const uint4* pents = s.ents + (st ^ 1) * p.cap;
          const uint16_t* pnext = s.next + (st ^ 1) * p.cap;
          uint32_t* head = s.head + tb * p.tslots;
          const uint32_t* phead = s.head + tp * p.tslots;
          uint32_t* otab = s.otab + st * p.oslots;
          uint16_t* touched = s.touched + st * p.oslots;
          // One bin usually takes nearly every vote of a region (the true camera offset), so votes are
          // aggregated per warp first: lanes with the same offset elect a leader that adds their number.
          auto vote = [&](uint32_t oid, uint32_t k) {
            uint32_t os = off_hash(oid) & omask;
            uint32_t probes = 0;
            for (;;) {
              uint32_t v = otab[os];
              if (v == EMPTY) {
                v = atomicCAS(&otab[os], EMPTY, (oid << cntbits) | k);
                if (v == EMPTY) { touched[atomicAdd(&s.ctl[st], 1u)] = (uint16_t)os; break; }
              }
              if ((v >> cntbits) == oid) { atomicAdd(&otab[os], k); break; }
              os = (os + 1) & omask;
              if (++probes > MAXPROBE) { s.ctl[2 + st] = 1; break; }  // table (nearly) full: defer
            }
          };
          const uint32_t obias = (g.W << p.dybits) | g.H;  // (dx + W) << dybits | (dy + H), both fields stay positive
          const uint32_t Lw = (L + 31) & ~31u;  // whole warps iterate together
          for (uint32_t i = tid; i < Lw; i += NTP) {
            uint32_t oid0 = NONE | lane, oid1 = NONE | lane;  // unique per lane: groups of one in match_any
            if (i < L) {
              const uint32_t pos = plist[i < nw2 ? i : p.cap - 1 - (i - nw2)];
              const uint32_t x = pos & 0x7FFFu, y = pos >> 16;
              const Code c = code_at(tile, wpr, x - 2 - tx0, y - Y0);
              const uint32_t slot = code_hash(c) & tmask;
              // position as x << dybits | y (< 2^24): prev - curr + bias is then the offset id in one subtraction;
              // the top nibble carries the code's last 4 bits
              const uint32_t key = (x << p.dybits) | y;
              const uint32_t w3 = key | (c.c3 << 28);
              ents[i] = make_uint4(c.c0, c.c1, c.c2, w3);
              next[i] = (uint16_t)atomicExch(&head[slot], i);  // push onto the bucket's chain (NIL -> 0xFFFF)
              if (pair_ok && (use_all || (pos & 0x8000u))) {   // !use_all: weight-2 codes only (src/kpm.hpp:113-117)
                uint32_t j = phead[slot];
                while (j != NIL) {
                  const uint4 pe = pents[j];
                  if (pe.x == c.c0 && pe.y == c.c1 && pe.z == c.c2 && ((pe.w ^ w3) >> 28) == 0) {
                    // equal codes: vote prev - curr (src/kpm.hpp:96-98)
                    const uint32_t oid = (pe.w & 0x0FFFFFFFu) - key + obias;
                    if (oid0 & NONE) oid0 = oid;
                    else if (oid1 & NONE) oid1 = oid;
                    else vote(oid, 1u);  // third and later matches of one keypoint: rare, one by one
                  }
                  const uint32_t nx = pnext[j];
                  j = nx == 0xFFFFu ? NIL : nx;
                }
              }
            }
            if (pair_ok) {
              const uint32_t grp = __match_any_sync(0xffffffffu, oid0);
              if (!(oid0 & NONE) && lane == (uint32_t)(__ffs((int)grp) - 1)) vote(oid0, (uint32_t)__popc(grp));
              if (__any_sync(0xffffffffu, !(oid1 & NONE))) {
                const uint32_t grp1 = __match_any_sync(0xffffffffu, oid1);
                if (!(oid1 & NONE) && lane == (uint32_t)(__ffs((int)grp1) - 1)) vote(oid1, (uint32_t)__popc(grp1));
              }
            }
          }
        }
        tb = tc;
      } else if (t >= 2) {
        make_ballot(t - 1, fa, region);  // the pair voted on during the previous step
      }
      __syncthreads();  // votes of this step complete; tile[st] / plist[st] consumed; last step's ballot written
      if (tid == 0 && t + 2 < nsteps) request(t + 2);
    }
    // run finished: the ballot warp still owes the last pair; the next work item's first barrier waits for it
    if (ballot_warp && nsteps >= 2) make_ballot(nsteps - 1, fa, region);
    // bucket tables: the two built last are still filled
    if (!ballot_warp) {
      const uint32_t t1 = tb >= 1 ? tb - 1 : 2, t2 = t1 >= 1 ? t1 - 1 : 2;
      uint4* h1 = reinterpret_cast<uint4*>(s.head + t1 * p.tslots);
      uint4* h2 = reinterpret_cast<uint4*>(s.head + t2 * p.tslots);
      for (uint32_t i = tid; i < p.tslots / 4; i += NTP) { h1[i] = make_uint4(NIL, NIL, NIL, NIL); h2[i] = make_uint4(NIL, NIL, NIL, NIL); }
    }
  }
}

#endif  // __CUDACC__
