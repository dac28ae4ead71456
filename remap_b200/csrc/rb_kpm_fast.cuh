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

#ifndef RB_FAST_NT
#define RB_FAST_NT 512
#endif
#define RB_MAX_OSLOTS_PER_THREAD 8  // offset table <= 2048 slots

struct RbKpmFastParams {
  RbGeom g;
  const uint32_t* lists;   // [frame][region][cap]  (rb_prep.cuh)
  const uint2* counts;     // [frame][region] (n_all, n_w2)
  RbRegionVote* votes;     // [npairs][nreg]
  uint32_t first_frame, npairs;
  uint32_t cap;            // list capacity (<= 2048)
  uint32_t tslots;         // code table slots, power of two >= 2 * cap
  uint32_t oslots;         // offset table slots, power of two, multiple of RB_FAST_NT
  uint32_t run;            // pairs per work item
  uint32_t box_x, box_y;   // TMA box: bytes per tile row, rows per box
  uint32_t nbox_y;         // boxes stacked vertically per tile (tile rows = nbox_y * box_y)
  uint32_t dybits, offbits;  // offset id = (dx + W) << dybits | (dy + H), offbits bits in all
  uint32_t* work_counter;  // zeroed before launch
  uint32_t* deferred_count;
  uint2* deferred;         // (pair, region)
  uint32_t deferred_cap;
};

namespace rbf {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
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

struct Smem {           // two stages of everything per-frame: stage k lives at base + k * stride
  uint8_t* tile;        // [2][tile_rows][box_x] packed 4 bit/pixel, + 16 bytes of slack
  uint32_t* plist;      // [2][cap] positions as delivered by the bulk copy
  uint4* ents;          // [2][cap] (c0, c1, c2, pos): the frame's codes, kept for its turn as "previous"
  uint32_t* ctab;       // [2][tslots] idx | c3 << 11 | 16 hash bits << 15 (bit 31 clear), or EMPTY
  uint32_t tile_stride, plist_stride;  // bytes / words between the stages
  uint32_t* otab;       // [oslots] offset id << cntbits | count, or EMPTY
  uint32_t* cand;       // [NW * 3 + NW] per-warp ticket candidates and bin counts (NW <= 16)
  uint32_t* stat;       // [2][4]: bins tied with ticket 0/1/2, overflow flag
  uint32_t* misc;       // [4]: work item
  uint64_t* mbar;       // [2]
};

__host__ __device__ inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline size_t tile_bytes(const RbKpmFastParams& p) {
  return align_up_sz((size_t)p.box_x * p.box_y * p.nbox_y + 16, 128);
}

__host__ __device__ inline size_t smem_bytes(const RbKpmFastParams& p) {
  size_t b = 0;
  b += 2 * tile_bytes(p);
  b += 2 * align_up_sz((size_t)p.cap * 4, 128);
  b += 2 * (size_t)p.cap * 16;
  b += 2 * (size_t)p.tslots * 4;
  b += (size_t)p.oslots * 4;
  b += 64 * 4 + 8 * 4 + 4 * 4 + 2 * 8;
  return b + 128;  // base alignment slack
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
  s.ctab = reinterpret_cast<uint32_t*>(base + o); o += 2 * (size_t)p.tslots * 4;
  s.otab = reinterpret_cast<uint32_t*>(base + o); o += (size_t)p.oslots * 4;
  s.cand = reinterpret_cast<uint32_t*>(base + o); o += 64 * 4;
  s.stat = reinterpret_cast<uint32_t*>(base + o); o += 8 * 4;
  s.misc = reinterpret_cast<uint32_t*>(base + o); o += 4 * 4;
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

}  // namespace rbf

__global__ void __launch_bounds__(RB_FAST_NT) rb_kpm_fast_kernel(const __grid_constant__ CUtensorMap tmap, const RbKpmFastParams p) {
  using namespace rbf;
  extern __shared__ __align__(128) uint8_t rb_fast_smem[];
  Smem s;
  carve(p, rb_fast_smem, s);
  const RbGeom& g = p.g;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t NT = RB_FAST_NT, NW = RB_FAST_NT / 32;
  const uint32_t tmask = p.tslots - 1, omask = p.oslots - 1;
  const uint32_t cntbits = 32 - p.offbits, cntmask = (1u << cntbits) - 1u, offmask = (1u << p.offbits) - 1u;
  const uint32_t wpr = p.box_x / 4;
  const uint32_t tile_tx_bytes = p.box_x * p.box_y * p.nbox_y;
  const uint32_t rv = g.region_votes;
  const uint32_t runs = (p.npairs + p.run - 1) / p.run;
  const uint32_t nitems = runs * g.nreg;

  // ---- one-off set-up: barriers, clean tables -------------------------------------------------
  if (tid == 0) {
    mbar_init(&s.mbar[0], 1);
    mbar_init(&s.mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (uint32_t i = tid; i < 2 * p.tslots; i += NT) s.ctab[i] = EMPTY;
  for (uint32_t i = tid; i < p.oslots; i += NT) s.otab[i] = EMPTY;
  if (tid < 8) s.stat[tid] = 0;
  __syncthreads();
  uint32_t ph0 = 0, ph1 = 0;  // mbarrier phase parities (persist across work items)
  uint32_t spar = 0;          // stat[] double buffer

  for (;;) {
    if (tid == 0) s.misc[0] = atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const uint32_t item = s.misc[0];
    if (item >= nitems) break;
    // regions vary fastest so that the CTAs running at the same time share frames in L2
    const uint32_t region = item % g.nreg, runidx = item / g.nreg;
    const uint32_t pa = runidx * p.run;
    const uint32_t pb = pa + p.run < p.npairs ? pa + p.run : p.npairs;  // pairs [pa, pb)
    const uint32_t fa = p.first_frame + pa, fb = p.first_frame + pb;   // frames fa .. fb inclusive
    const uint32_t cs = region / g.grid_h, rs = region % g.grid_h;     // idx = grid_h*col + row (src/kpr.hpp:71-74)
    const uint32_t X0 = g.col0[cs], Y0 = g.row0[rs];
    const uint32_t tx0 = (X0 - 2) & ~31u;  // tile origin: TMA wants the inner coordinate 16-byte aligned (32 px)
    const uint2* cnts = p.counts + region;

    auto request = [&](uint32_t frame, uint32_t stage, uint32_t n_all) {  // thread 0 only
      const uint32_t n = n_all < p.cap ? n_all : p.cap;
      const uint32_t lbytes = (n * 4 + 15) & ~15u;
      mbar_expect_tx(&s.mbar[stage], tile_tx_bytes + lbytes);
      for (uint32_t b = 0; b < p.nbox_y; ++b)
        tma_load_3d(s.tile + stage * s.tile_stride + b * p.box_x * p.box_y, &tmap, tx0 / 2, Y0 - 2 + b * p.box_y, frame,
                    &s.mbar[stage]);
      if (lbytes) bulk_load(s.plist + stage * s.plist_stride, p.lists + ((uint64_t)frame * g.nreg + region) * p.cap, lbytes, &s.mbar[stage]);
    };

    // counts of frames f, f + 1, f + 2 (rolling registers, block-uniform)
    uint2 cA = __ldg(cnts + (uint64_t)fa * g.nreg);
    uint2 cB = fa + 1 <= fb ? __ldg(cnts + (uint64_t)(fa + 1) * g.nreg) : make_uint2(0, 0);
    uint2 cC = make_uint2(0, 0);
    if (tid == 0) {
      request(fa, 0, cA.x);
      if (fa + 1 <= fb) request(fa + 1, 1, cB.x);
    }
    bool prev_valid = false;
    uint32_t prev_n = 0, prev_w2 = 0;
    bool use_all = false;  // weight switch of the pair (f - 1, f)

    for (uint32_t f = fa, t = 0; f <= fb; ++f, ++t) {
      const uint32_t st = t & 1;
      if (f + 2 <= fb) cC = __ldg(cnts + (uint64_t)(f + 2) * g.nreg);
      // src/kpm.hpp:219-220 ('<' on previous, '<=' on current), for the pair (f, f + 1)
      const bool use_all_next = f < fb && ((cA.y < g.weight_switch) || (cB.y <= g.weight_switch));
      const bool need_w1 = (t > 0 && use_all) || use_all_next;
      const uint32_t L = need_w1 ? cA.x : cA.y;  // entries of this frame that take part
      const bool fits = L <= p.cap;
      const bool pair = t > 0;
      const bool pair_ok = pair && prev_valid && fits;

      if (st == 0) { mbar_wait(&s.mbar[0], ph0, p.work_counter + 2); ph0 ^= 1; }
      else { mbar_wait(&s.mbar[1], ph1, p.work_counter + 2); ph1 ^= 1; }

      // ---- P: codes, build this frame's table, probe the previous one, vote --------------------
      if (fits) {
        const uint32_t* tile = reinterpret_cast<const uint32_t*>(s.tile + st * s.tile_stride);
        const uint32_t* plist = s.plist + st * s.plist_stride;
        uint4* ents = s.ents + st * p.cap;
        const uint4* pents = s.ents + (st ^ 1) * p.cap;
        uint32_t* ctab = s.ctab + st * p.tslots;
        const uint32_t* ptab = s.ctab + (st ^ 1) * p.tslots;
        // One bin usually takes nearly every vote of a region (the true camera offset), so votes are
        // aggregated per warp first: lanes with the same offset elect a leader that adds their number.
        auto vote = [&](uint32_t oid, uint32_t k) {
          uint32_t os = off_hash(oid) & omask;
          uint32_t probes = 0;
          for (;;) {
            uint32_t v = s.otab[os];
            if (v == EMPTY) {
              v = atomicCAS(&s.otab[os], EMPTY, (oid << cntbits) | k);
              if (v == EMPTY) break;
            }
            if ((v >> cntbits) == oid) { atomicAdd(&s.otab[os], k); break; }
            os = (os + 1) & omask;
            if (++probes > MAXPROBE) { s.stat[spar * 4 + 3] = 1; break; }  // table (nearly) full: defer
          }
        };
        const uint32_t Lw = (L + 31) & ~31u;  // whole warps iterate together
        for (uint32_t i = tid; i < Lw; i += NT) {
          uint32_t first_oid = 0x80000000u | lane;  // "no vote": unique per lane
          if (i < L) {
            const uint32_t pos = plist[i];
            const uint32_t x = pos & 0x7FFFu, y = pos >> 16;
            const Code c = code_at(tile, wpr, x - 2 - tx0, y - Y0);
            const uint32_t h = code_hash(c);
            ents[i] = make_uint4(c.c0, c.c1, c.c2, pos);
            const uint32_t tag = (c.c3 << 11) | ((h >> 16) << 15);  // bit 31 stays clear: h >> 16 has 16 bits
            {
              uint32_t slot = h & tmask;
              while (atomicCAS(&ctab[slot], EMPTY, tag | i) != EMPTY) slot = (slot + 1) & tmask;
            }
            if (pair_ok && (use_all || (pos & 0x8000u))) {  // !use_all: weight-2 codes only (src/kpm.hpp:113-117)
              uint32_t slot = h & tmask, e;
              while ((e = ptab[slot]) != EMPTY) {
                if (((e ^ tag) >> 11) == 0) {
                  const uint4 pe = pents[e & 0x7FFu];
                  if (pe.x == c.c0 && pe.y == c.c1 && pe.z == c.c2) {
                    // offset = prev - curr (src/kpm.hpp:96-98)
                    const uint32_t dxb = (pe.w & 0x7FFFu) - x + g.W, dyb = (pe.w >> 16) - y + g.H;
                    const uint32_t oid = (dxb << p.dybits) | dyb;
                    if (first_oid & 0x80000000u) first_oid = oid; else vote(oid, 1u);
                  }
                }
                slot = (slot + 1) & tmask;
              }
            }
          }
          const uint32_t grp = __match_any_sync(0xffffffffu, first_oid);
          if (!(first_oid & 0x80000000u) && lane == (uint32_t)(__ffs((int)grp) - 1)) vote(first_oid, (uint32_t)__popc(grp));
        }
      }
      __syncthreads();  // B2: votes complete; tile[st] / plist[st] consumed

      if (tid == 0 && f + 2 <= fb) request(f + 2, st, cC.x);

      if (pair) {
        const uint32_t pairidx = f - 1 - p.first_frame;
        const bool overflow = s.stat[spar * 4 + 3] != 0;
        if (pair_ok && !overflow) {
          // ---- S: ticket, tie statistics ----------------------------------------------------
          uint32_t v[RB_MAX_OSLOTS_PER_THREAD];
          uint32_t t0 = 0, t1 = 0, t2 = 0, nb = 0;
          const uint32_t per = p.oslots / NT;
#pragma unroll
          for (uint32_t k = 0; k < RB_MAX_OSLOTS_PER_THREAD; ++k) {
            v[k] = EMPTY;
            if (k < per) {
              v[k] = s.otab[tid + k * NT];
              if (v[k] != EMPTY) {
                ++nb;
                // larger key = earlier in the ticket: count desc, then offset id asc (dx asc, dy asc)
                top3_insert(((v[k] & cntmask) << p.offbits) | (offmask - (v[k] >> cntbits)), t0, t1, t2);
              }
            }
          }
          uint32_t m0, m1, m2;
          warp_top3(t0, t1, t2, m0, m1, m2);
          nb = __reduce_add_sync(0xffffffffu, nb);
          if (lane == 0) {
            s.cand[warp * 3] = m0; s.cand[warp * 3 + 1] = m1; s.cand[warp * 3 + 2] = m2;
            s.cand[NW * 3 + warp] = nb;
          }
          __syncthreads();  // B3
          uint32_t cv0 = lane < NW * 3 ? s.cand[lane] : 0u, cv1 = 0, cv2 = 0;
          if (lane + 32 < NW * 3) top3_insert(s.cand[lane + 32], cv0, cv1, cv2);
          uint32_t g0, g1, g2;
          warp_top3(cv0, cv1, cv2, g0, g1, g2);
          const uint32_t nbins = __reduce_add_sync(0xffffffffu, lane < NW ? s.cand[NW * 3 + lane] : 0u);
          const uint32_t c0 = g0 >> p.offbits, c1 = g1 >> p.offbits, c2 = g2 >> p.offbits;
          uint32_t e0 = 0, e1 = 0, e2 = 0;
#pragma unroll
          for (uint32_t k = 0; k < RB_MAX_OSLOTS_PER_THREAD; ++k) {
            if (k < per && v[k] != EMPTY) {
              const uint32_t c = v[k] & cntmask;
              e0 += c == c0; e1 += c == c1; e2 += c == c2;
              s.otab[tid + k * NT] = EMPTY;
            }
          }
          e0 = __reduce_add_sync(0xffffffffu, e0 | (e1 << 16));  // e0, e1 <= oslots < 65536
          e2 = __reduce_add_sync(0xffffffffu, e2);
          if (lane == 0) {
            atomicAdd(&s.stat[spar * 4 + 0], e0 & 0xFFFFu);
            atomicAdd(&s.stat[spar * 4 + 1], e0 >> 16);
            atomicAdd(&s.stat[spar * 4 + 2], e2);
          }
          for (uint32_t i = tid; i < p.tslots; i += NT) s.ctab[(st ^ 1) * p.tslots + i] = EMPTY;
          if (tid < 4) s.stat[(spar ^ 1) * 4 + tid] = 0;
          __syncthreads();  // B4
          if (tid == 0) {
            const uint32_t E0 = s.stat[spar * 4 + 0], E1 = s.stat[spar * 4 + 1], E2 = s.stat[spar * 4 + 2];
            RbRegionVote vt;
            vt.use_all = use_all ? 1u : 0u;
            vt.n_prev = prev_n; vt.n_curr = cA.x; vt.w2_prev = prev_w2; vt.w2_curr = cA.y;
            vt.nbins = nbins;
            vt.nticket = nbins < rv ? nbins : rv;
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
          spar ^= 1;
        } else {
          // the pair cannot be finished here (a list or the offset table does not fit): defer it
          if (tid == 0) {
            const uint32_t at = atomicAdd(p.deferred_count, 1u);
            if (at < p.deferred_cap) p.deferred[at] = make_uint2(pairidx, region);
          }
          for (uint32_t i = tid; i < p.oslots; i += NT) s.otab[i] = EMPTY;
          for (uint32_t i = tid; i < p.tslots; i += NT) s.ctab[(st ^ 1) * p.tslots + i] = EMPTY;
          if (tid < 4) s.stat[(spar ^ 1) * 4 + tid] = 0;
          __syncthreads();
          if (tid < 4) s.stat[spar * 4 + tid] = 0;
          spar ^= 1;
        }
      }
      prev_valid = fits;
      prev_n = cA.x; prev_w2 = cA.y;
      use_all = use_all_next;
      cA = cB; cB = cC;
    }
    // leave both code tables clean for the next work item (the last frame's table is still filled)
    __syncthreads();
    for (uint32_t i = tid; i < 2 * p.tslots; i += NT) s.ctab[i] = EMPTY;
    __syncthreads();
  }
}

#endif  // __CUDACC__
