// rb_kpm_big.cuh -- K2b: the pipelined matcher for LARGE regions (640x480: 4-5 k keypoints per region and
// frame; any region list the first-pass kernel of rb_kpm_fast.cuh cannot hold).  Same job, same results:
// kpm::details::cast_vote / count_offsets / get_offsets / top_offsets (src/kpm.hpp:91-159,213-223) for runs of
// consecutive frames; count_offsets has no size limit (src/kpm.hpp:105-125), so neither may the fast path.
//
// rb_kpm_fast_kernel keeps every keypoint's full 100-bit code (16 bytes) for two frames in shared memory; that caps
// a region at ~2 k keypoints.  Here one CTA has the SM to itself (1,024 threads, up to 227 KB) and keeps, per
// keypoint and frame, only 6 bytes:
//   pos   16 bit   position inside the region's tile (x | y << 8; tiles are at most 256 x 256 pixels)
//   link  32 bit   next entry of the same hash bucket (16 bit) | 16 more bits of the code's hash (tag)
// and THREE packed 4 bit/pixel tiles (previous frame, current frame, the next one arriving by TMA).  A probe walks
// the previous frame's bucket chain, skips entries whose tag differs, and verifies a tag hit by re-forming the
// previous keypoint's code from the previous tile -- the tile is the most compact exact store of the codes
// (28 KB against 16 B x 5 k = 80 KB per frame at 640x480).  Keypoint positions are read straight from the list
// rows in HBM (coalesced, each entry once, the next trip's entry requested before this trip's work), so no
// shared memory goes to staging them.  Everything else follows rb_kpm_fast_kernel: persistent CTAs, work items =
// (region, run of consecutive pairs), a frame's table is built once and serves as "current" and then as
// "previous", warp-aggregated votes into a one-word-per-bin offset table, a ballot warp that turns the finished
// table into the region's ballot while the other warps are on the next frame, and a device-side list that DEFERS
// whatever does not fit (to the general kernel, rb_kpm.cuh) -- nothing is approximated.
//
// Offsets are encoded relative to the region: both keypoints of a vote lie in the same region, so
// |dx| < tile width and |dy| < tile rows; that leaves >= 13 bits of a bin word for the count at 640x480
// (the frame-relative encoding of the fast kernel would leave 11: a 640x480 region's winning bin holds ~3,000 votes).
#pragma once

#include "rb_kpm_fast.cuh"

#if defined(__CUDACC__)

#define RB_BIG_NWP 31
#define RB_BIG_NTP (32 * RB_BIG_NWP)
#define RB_BIG_NT (RB_BIG_NTP + 32)  // + the ballot warp

namespace rbb {

using namespace rbf;

constexpr uint32_t F_FITS = 1u, F_PAIR_OK = 2u, F_USE_ALL = 4u;
constexpr uint32_t NIL16 = 0xFFFFu;

struct Smem {
  uint8_t* tile;      // [3][tile_bytes] packed 4 bit/pixel tiles; frame t of a run lives in stage t % 3
  uint16_t* pos;      // [2][cap] tile-relative position of entry i: (x - 2 - tx0) | (y - Y0) << 8
  uint32_t* link;     // [2][cap] next entry of the bucket (NIL16: none) | hash tag << 16
  uint32_t* head;     // [2][tslots] epoch << 16 | first entry of the bucket's chain (valid only for the frame of that epoch)
  uint32_t* otab;     // [2][oslots] offset id << cntbits | count, or EMPTY
  uint16_t* touched;  // [2][oslots]
  uint32_t* planL;    // [run + 3] entries taking part | weight-2 entries << 16
  uint32_t* planF;    // [run + 3] F_* flags
  uint32_t* ctl;      // [0..1] touched counts, [2..3] overflow flags, [4] work item
  uint64_t* mbar;     // [3]
  uint32_t tile_stride;
};

__host__ __device__ inline size_t smem_bytes(const RbKpmFastParams& p) {
  size_t b = 0;
  b += 3 * tile_bytes(p);
  b += align_up_sz(2 * (size_t)p.cap * 2, 16);
  b += 2 * (size_t)p.cap * 4;
  b += 2 * (size_t)p.tslots * 4;
  b += 2 * (size_t)p.oslots * 4;
  b += 2 * (size_t)p.oslots * 2;
  b += 2 * align_up_sz(((size_t)p.run + 3) * 4, 16);
  b += 8 * 4 + 4 * 8;
  return b;
}

__device__ __forceinline__ void carve(const RbKpmFastParams& p, uint8_t* base, Smem& s) {
  size_t o = 0;
  s.tile_stride = (uint32_t)tile_bytes(p);
  s.tile = base + o; o += 3 * tile_bytes(p);
  s.pos = reinterpret_cast<uint16_t*>(base + o); o += align_up_sz(2 * (size_t)p.cap * 2, 16);
  s.link = reinterpret_cast<uint32_t*>(base + o); o += 2 * (size_t)p.cap * 4;
  s.head = reinterpret_cast<uint32_t*>(base + o); o += 2 * (size_t)p.tslots * 4;
  s.otab = reinterpret_cast<uint32_t*>(base + o); o += 2 * (size_t)p.oslots * 4;
  s.touched = reinterpret_cast<uint16_t*>(base + o); o += 2 * (size_t)p.oslots * 2;
  s.planL = reinterpret_cast<uint32_t*>(base + o); o += align_up_sz(((size_t)p.run + 3) * 4, 16);
  s.planF = reinterpret_cast<uint32_t*>(base + o); o += align_up_sz(((size_t)p.run + 3) * 4, 16);
  s.ctl = reinterpret_cast<uint32_t*>(base + o); o += 8 * 4;
  s.mbar = reinterpret_cast<uint64_t*>(base + o);
}

__device__ __forceinline__ void workers_sync() {  // the 31 worker warps only; the ballot warp is elsewhere
  asm volatile("bar.sync 1, %0;" ::"n"(RB_BIG_NTP) : "memory");
}

}  // namespace rbb

__global__ void __launch_bounds__(RB_BIG_NT, 1) rb_kpm_big_kernel(const __grid_constant__ CUtensorMap tmap, const RbKpmFastParams p) {
  using namespace rbb;
  extern __shared__ __align__(128) uint8_t rb_big_smem[];
  rbb::Smem s;
  rbb::carve(p, rb_big_smem, s);
  const RbGeom& g = p.g;
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  constexpr uint32_t NT = RB_BIG_NT, NTP = RB_BIG_NTP;
  const bool ballot_warp = tid >= NTP;
  const uint32_t runs = (p.npairs + p.run - 1) / p.run;
  uint32_t nitems = runs * g.nreg;
  if (p.items) { nitems = *p.nitems; if (nitems > p.deferred_cap) nitems = p.deferred_cap; }
  const uint32_t cntbits = 32 - p.offbits, cntmask = (1u << cntbits) - 1u, offmask = (1u << p.offbits) - 1u;
  const uint32_t dymask = (1u << p.dybits) - 1u;

  if (tid == 0) {
    mbar_init(&s.mbar[0], 1);
    mbar_init(&s.mbar[1], 1);
    mbar_init(&s.mbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (uint32_t i = tid; i < 2 * p.tslots; i += NT) s.head[i] = 0;  // epoch 0: never a frame's epoch
  for (uint32_t i = tid; i < 2 * p.oslots; i += NT) s.otab[i] = EMPTY;
  if (tid < 8) s.ctl[tid] = 0;
  __syncthreads();
  uint32_t phbits = 0;  // mbarrier phase parities, bit per stage (persist across work items)
  // Bucket heads are never cleared: a head word is (epoch << 16 | entry), every frame a CTA builds gets the next
  // epoch, and a head whose epoch is not the frame's is an empty bucket.  (Clearing cost a pass over the table and a
  // second barrier per step.)  The 16-bit epoch wraps after 65,535 frames: then both tables are zeroed once.
  uint32_t ebase = p.epoch0 >= 1 && p.epoch0 < 0xFFFFu ? p.epoch0 : 1u;

  // ---- the ballot warp's job: one finished offset table -> one RbRegionVote (as in rb_kpm_fast_kernel) --------
  auto make_ballot = [&](uint32_t t, uint32_t fa, uint32_t region) {
    const uint32_t par = t & 1, fl = s.planF[t];
    uint32_t* otab = s.otab + par * p.oslots;
    const uint16_t* touched = s.touched + par * p.oslots;
    const uint32_t nt = s.ctl[par];
    const bool overflow = s.ctl[2 + par] != 0;
    const uint32_t pairidx = fa + t - 1 - p.first_frame;
    if ((fl & F_PAIR_OK) && !overflow) {
      uint32_t t0 = 0, t1 = 0, t2 = 0;
      for (uint32_t j = lane; j < nt; j += 32) {
        const uint32_t v = otab[touched[j]];
        // larger key = earlier in the ticket: count desc, then offset id asc (dx asc, dy asc)
        top3_insert(((v & cntmask) << p.offbits) | (offmask - (v >> cntbits)), t0, t1, t2);
      }
      uint32_t g0, g1, g2;
      warp_top3(t0, t1, t2, g0, g1, g2);
      const uint32_t c0 = g0 >> p.offbits, c1 = g1 >> p.offbits, c2 = g2 >> p.offbits;
      uint32_t e01 = 0, e2 = 0, hh = 0;  // bins tied with ticket 0 / 1 (16 bits each) and 2; histogram digest
      for (uint32_t j = lane; j < nt; j += 32) {
        const uint32_t sl = touched[j];
        const uint32_t v = otab[sl], c = v & cntmask, oid = v >> cntbits;
        e01 += (c == c0 ? 1u : 0u) + (c == c1 ? 0x10000u : 0u);
        e2 += c == c2;
        hh += rb_bin_hash((int32_t)(oid >> p.dybits) - (int32_t)p.bias_x, (int32_t)(oid & dymask) - (int32_t)p.bias_y, c);
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
        vt.use_all = (fl & F_USE_ALL) ? 1u : 0u;
        vt.n_prev = cp.x; vt.n_curr = cc.x; vt.w2_prev = cp.y; vt.w2_curr = cc.y;
        vt.nbins = nt;
        vt.hist_hash = hh;
        vt.nticket = nt < rv ? nt : rv;
        const uint32_t gk[3] = {g0, g1, g2}, ck[3] = {c0, c1, c2};
        uint32_t ngt[3], nge[3];  // counts are sorted c0 >= c1 >= c2
        ngt[0] = 0; nge[0] = E0;
        ngt[1] = c1 == c0 ? 0u : nge[0]; nge[1] = ngt[1] + E1;
        ngt[2] = c2 == c1 ? ngt[1] : nge[1]; nge[2] = ngt[2] + E2;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
          RbBin b; b.dx = 0; b.dy = 0; b.cnt = 0;
          vt.ticket[k] = b; vt.ngt[k] = 0; vt.nge[k] = 0;
          if (k < 3 && k < vt.nticket) {
            const uint32_t oid = offmask - (gk[k] & offmask);
            vt.ticket[k].dx = (int32_t)(oid >> p.dybits) - (int32_t)p.bias_x;
            vt.ticket[k].dy = (int32_t)(oid & dymask) - (int32_t)p.bias_y;
            vt.ticket[k].cnt = ck[k];
            vt.ngt[k] = ngt[k]; vt.nge[k] = nge[k];
          }
        }
        p.votes[(uint64_t)pairidx * g.nreg + region] = vt;
      }
    } else {
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
    uint32_t region = item % g.nreg, pa = (item / g.nreg) * p.run;
    uint32_t pb = pa + p.run < p.npairs ? pa + p.run : p.npairs;  // pairs [pa, pb)
    if (p.items) { const uint2 it = p.items[item]; pa = it.x; pb = pa + 1; region = it.y; }
    const uint32_t fa = p.first_frame + pa;
    const uint32_t nsteps = pb - pa + 1;
    const uint32_t cs = region / g.grid_h, rs = region % g.grid_h;  // idx = grid_h*col + row (src/kpr.hpp:71-74)
    const uint32_t X0 = g.col0[cs], Y0 = g.row0[rs];
    const uint32_t tx0 = (X0 - 2) & ~31u;  // tile origin (pixels): TMA needs a 16-byte aligned start

    // ---- the run's plan, one thread per step (the same rules as rb_kpm_fast_kernel) -----------------------------
    if (tid < nsteps) {
      const uint2* cnts = p.counts + (uint64_t)fa * g.nreg + region;
      const uint32_t t = tid, ws = g.weight_switch;
      auto load = [&](int k) { return k >= 0 && k < (int)nsteps ? __ldg(cnts + (uint64_t)k * g.nreg) : make_uint2(0, 0); };
      const uint2 cm1 = load((int)t - 1), c0 = load((int)t), cp1 = load((int)t + 1);
      auto sw = [&](const uint2& a, const uint2& b) { return a.y < ws || b.y <= ws; };  // src/kpm.hpp:219-220
      const bool ua_0 = t >= 1 && sw(cm1, c0), ua_p1 = t + 1 < nsteps && sw(c0, cp1);
      const uint32_t L = (ua_0 || ua_p1) ? c0.x : c0.y;
      const uint32_t Lm1 = ((t >= 2 && sw(load((int)t - 2), cm1)) || ua_0) ? cm1.x : cm1.y;
      const bool fits = c0.x <= p.lcap && L <= p.cap, fits_m1 = cm1.x <= p.lcap && Lm1 <= p.cap;
      s.planL[t] = fits ? (L | (c0.y << 16)) : 0u;
      s.planF[t] = (fits ? F_FITS : 0u) | (ua_0 ? F_USE_ALL : 0u) | ((t >= 1 && fits && fits_m1) ? F_PAIR_OK : 0u);
    }
    __syncthreads();

    if (ebase + nsteps > 0xFFFFu) {  // block-uniform
      for (uint32_t i = tid; i < 2 * p.tslots; i += NT) s.head[i] = 0;
      ebase = 1;
      __syncthreads();
    }
    auto request = [&](uint32_t t) {  // thread 0 only: the tile of step t into stage t % 3
      const uint32_t stage = t % 3;
      mbar_expect_tx(&s.mbar[stage], p.box_x * p.box_y * p.nbox_y);
      for (uint32_t b = 0; b < p.nbox_y; ++b)
        tma_load_3d(s.tile + stage * s.tile_stride + b * p.box_x * p.box_y, &tmap, tx0 / 2, Y0 - 2 + b * p.box_y, fa + t,
                    &s.mbar[stage]);
    };
    if (tid == 0) {
      request(0);
      if (nsteps > 1) request(1);
    }
    // entry i of step t's list: weight-2 entries sit at the front of the list row, weight-1 entries at its end
    // (rb_prep.cuh).  Every worker's FIRST entry of a step is requested during the step before, so that no step
    // opens with a round trip to L2 / HBM; later entries one trip ahead.
    auto list_entry = [&](uint32_t t, uint32_t i) -> uint32_t {
      const uint32_t wl = s.planL[t], L = wl & 0xFFFFu, nw2 = wl >> 16;
      if (i >= L) return 0u;
      const uint32_t* row = p.lists + ((uint64_t)(fa + t) * g.nreg + region) * p.lcap;
      return __ldg(row + (i < nw2 ? i : p.lcap - 1 - (i - nw2)));
    };
    uint32_t pfirst = ballot_warp ? 0u : list_entry(0, tid);

    for (uint32_t t = 0; t < nsteps; ++t) {
      const uint32_t st = t & 1, stage = t % 3;
      if (!ballot_warp) {
        const uint32_t pcur = pfirst;                          // this worker's first entry of step t
        if (t + 1 < nsteps) pfirst = list_entry(t + 1, tid);  // ... and of step t + 1: lands while this step works
        mbar_wait(&s.mbar[stage], (phbits >> stage) & 1u, p.work_counter + 2);
        phbits ^= 1u << stage;
        const uint32_t fl = s.planF[t];
        if (fl & F_FITS) {
          const uint32_t wl = s.planL[t], L = wl & 0xFFFFu, nw2 = wl >> 16;
          const bool pair_ok = (fl & F_PAIR_OK) != 0, use_all = (fl & F_USE_ALL) != 0;
          const uint32_t tmask = p.tslots - 1, omask = p.oslots - 1, wpr = p.box_x / 4;
          const uint32_t* tile = reinterpret_cast<const uint32_t*>(s.tile + stage * s.tile_stride);
          const uint32_t* ptile = reinterpret_cast<const uint32_t*>(s.tile + ((t + 2) % 3) * s.tile_stride);
          uint16_t* pos = s.pos + st * p.cap;
          uint32_t* link = s.link + st * p.cap;
          const uint16_t* ppos = s.pos + (st ^ 1) * p.cap;
          const uint32_t* plink = s.link + (st ^ 1) * p.cap;
          uint32_t* head = s.head + st * p.tslots;
          const uint32_t* phead = s.head + (st ^ 1) * p.tslots;
          uint32_t* otab = s.otab + st * p.oslots;
          uint16_t* touched = s.touched + st * p.oslots;
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
          const uint32_t obias = (p.bias_x << p.dybits) | p.bias_y;
          const uint32_t* row = p.lists + ((uint64_t)(fa + t) * g.nreg + region) * p.lcap;
          auto list_at = [&](uint32_t i) { return i < L ? __ldg(row + (i < nw2 ? i : p.lcap - 1 - (i - nw2))) : 0u; };
          const uint32_t Lw = (L + 31) & ~31u;  // whole warps iterate together
          const uint32_t ecur = (ebase + t) << 16, eprev = (ebase + t - 1) << 16;
          // re-forms the code of previous-frame entry j from the previous tile and compares: the offset id, or NONE
          auto verify = [&](uint32_t j, const Code& c, uint32_t key) -> uint32_t {
            const uint32_t pp = ppos[j], plx = pp & 0xFFu, ply = pp >> 8;
            const Code d = code_at(ptile, wpr, plx, ply);
            if (d.c0 == c.c0 && d.c1 == c.c1 && d.c2 == c.c2 && d.c3 == c.c3)
              return ((plx << p.dybits) | ply) - key + obias;  // equal codes: prev - curr (src/kpm.hpp:96-98)
            return NONE;
          };
          uint32_t pnext = pcur;
          for (uint32_t i = tid; i < Lw; i += NTP) {
            const uint32_t pw = pnext;
            pnext = list_at(i + NTP);
            uint32_t oid0 = NONE | lane, oid1 = NONE | lane;  // unique per lane: groups of one in match_any
            // The bucket walk only COLLECTS the entries whose 16-bit tag matches (lanes leave the walk at different
            // trips); the expensive part -- re-forming the previous keypoint's code from the tile -- runs after the
            // walk with the warp converged again.  A third tag hit (rare: heavy repetition) takes the slow path.
            uint32_t cand0 = NIL16, cand1 = NIL16, key = 0, jhead = NIL16, tag = 0;
            bool more = false;
            Code c;
            c.c0 = c.c1 = c.c2 = c.c3 = 0;
            if (i < L) {
              const uint32_t lx = (pw & 0x7FFFu) - 2 - tx0, ly = (pw >> 16) - Y0;
              c = code_at(tile, wpr, lx, ly);
              const uint32_t h = code_hash(c), slot = h & tmask;
              tag = h & 0xFFFF0000u;
              key = (lx << p.dybits) | ly;
              pos[i] = (uint16_t)(lx | (ly << 8));
              const uint32_t old = atomicExch(&head[slot], ecur | i);  // push onto the bucket's chain
              link[i] = ((old & 0xFFFF0000u) == ecur ? (old & 0xFFFFu) : NIL16) | tag;
              if (pair_ok && (use_all || (pw & 0x8000u))) {  // !use_all: weight-2 codes only (src/kpm.hpp:113-117)
                const uint32_t ph = phead[slot];
                jhead = (ph & 0xFFFF0000u) == eprev ? (ph & 0xFFFFu) : NIL16;
                uint32_t j = jhead;
                while (j != NIL16) {
                  const uint32_t e = plink[j];
                  if ((e & 0xFFFF0000u) == tag) {
                    if (cand0 == NIL16) cand0 = j;
                    else if (cand1 == NIL16) cand1 = j;
                    else more = true;
                  }
                  j = e & 0xFFFFu;
                }
              }
            }
            if (cand0 != NIL16) oid0 = verify(cand0, c, key);
            if (oid0 & NONE) oid0 = NONE | lane;
            if (__any_sync(0xffffffffu, cand1 != NIL16)) {
              if (cand1 != NIL16) {
                const uint32_t o = verify(cand1, c, key);
                if (!(o & NONE)) { if (oid0 & NONE) oid0 = o; else oid1 = o; }
              }
              if (more) {  // third and later tag hits of one keypoint: one by one
                uint32_t seen = 0;
                for (uint32_t j = jhead; j != NIL16;) {
                  const uint32_t e = plink[j];
                  if ((e & 0xFFFF0000u) == tag && ++seen > 2) {
                    const uint32_t o = verify(j, c, key);
                    if (!(o & NONE)) vote(o, 1u);
                  }
                  j = e & 0xFFFFu;
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
      } else if (t >= 2) {
        make_ballot(t - 1, fa, region);  // the pair voted on during the previous step
      }
      __syncthreads();  // votes of this step complete; the tile of frame t - 1 is no longer needed
      if (tid == 0 && t + 2 < nsteps) request(t + 2);  // into the stage frame t - 1 occupied
    }
    if (ballot_warp && nsteps >= 2) make_ballot(nsteps - 1, fa, region);
    ebase += nsteps;  // the next work item's frames get fresh epochs: its tables start out empty without a clear
  }
}

#endif  // __CUDACC__
