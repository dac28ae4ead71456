// rb_fg.cuh -- pass-2 foreground extraction: per frame, the mask that fdf::filter
// builds before its masked blit (src/fdf.hpp:58-66):
//
//   fde::extractor::extract   (src/fde.hpp:83-103)  generate_mask (frame vs background window), then
//                             cte::extractor::extract on the MEDIAN image seeded where they differ
//                             (src/cte.hpp:60-166: breadth-first fill of 4-connected equal-colour pixels,
//                             one contour per fill), then contours larger than area / 5 are dropped;
//   fde::mask                 (src/fde.hpp:122-146)  every kept contour's pixels + its enclosure rectangle.
//
// The reference walks a queue pixel by pixel.  Here one CTA takes one frame and works on BIT MAPS and
// RUNS, all of it in shared memory:
//   1. three bit maps of the frame interior, one bit per pixel, built 32 pixels per thread with byte-wise
//      SWAR compares: run starts (median differs from its left neighbour), vertical links (median equals
//      the pixel above) and seeds (frame differs from the background window, i.e. generate_mask == 0);
//   2. horizontal runs get ids by a prefix count of the run-start bits (row-major, so the smallest id of a
//      component is the run the reference's scan meets first);
//   3. union-find over runs: one union per (run, overlapping run above) pair, found as the bits of
//      links & (a run starts here, above or the link chain starts here); min-root linking with CAS;
//   4. seeds mark their component's root; seeded roots get statistics slots; every run adds its length,
//      its extent and its row to its component's slot (shared-memory atomics);
//   5. kept components (area <= limit) paint their runs, then their enclosures, into the output bit map.
// The interior is what cte's horizon leaves: columns 1..W-2, rows 1..H-3 (clear_outline marks row 0, the
// side columns AND the last two rows, src/cte.hpp:159-177).  The enclosure follows the reference to the
// letter, including cdt::limits::update's `else if` (src/cdt.hpp:183-190): its left bound is the smallest
// column among the contour's rows BELOW its first row (none: empty rectangle), right/bottom are exclusive
// (src/fde.hpp:133-143).  Derivation in DESIGN.md; equality with the reference is what the tests check.
//
// Frames with more runs or more seeded components than the shared-memory tables hold are DEFERRED to the
// general variant of the same code (labels and statistics in a global-memory scratch slab per CTA, 32-bit
// labels), nothing is approximated.  Contour ids in the reference are uint16 and wrap after 65,534
// contours per frame; such frames are outside the contract.
//
// Output per frame: H x NW words, bit x & 31 of word x >> 5 set <=> fde::mask != 0 (foreground); consumed
// by the masked blit (rb_blit.cuh) so that the byte mask never exists in HBM.
#pragma once

#include "rb_blit.cuh"
#include "rb_common.cuh"

#define RB_FG_NT 512
#define RB_FG_NONE 0xFFFFFFFFu

struct RbFgParams {
  RbGeom g;
  const uint8_t* frames;   // 8 bit/pixel frame store
  const uint8_t* median;   // median store (pixel x at byte x + 2 of a row)
  const uint8_t* bg;       // background map, bgW bytes per row (fdf::background::image_, src/fdf.hpp:13-16)
  uint32_t bgW, bgH;
  const RbPlacement* places;  // frame slot + position of the frame inside the background map
  uint32_t n;
  uint32_t NW;             // words per bit-map row: ceil(W / 32)
  uint32_t rcap, scap;     // run / slot capacity of this variant
  uint32_t area_limit;     // frame area / 5 (src/fde.hpp:94)
  uint32_t* fgbits;        // out [n][H][NW]
  uint32_t* nkept;         // out [n]: kept contours (parity tap)
  uint32_t* deferred;      // frames (placement indices) this variant could not hold
  uint32_t* ndeferred;
  const uint32_t* todo;    // general variant: list of placement indices to process (NULL: all)
  const uint32_t* ntodo;
  uint8_t* scratch;        // general variant: per-CTA slab
  uint64_t scratch_stride;
};

// Working set of one CTA.  L = label type of the union-find (uint16_t in shared memory, uint32_t in the
// general variant).
template <typename L>
struct RbFgWork {
  uint32_t* start;     // [H][NW] run-start bits
  uint32_t* ev;        // [H][NW] merge events, later the output bit map
  uint32_t* seed;      // [H][NW] seed bits
  uint16_t* base16;    // [H][NW] run starts in the row before this word
  uint32_t* rowbase;   // [H]     run starts in the rows before this row
  uint32_t* misc;      // [0] R, [1] slots, [2] kept, [8..40) scan scratch
  L* parent;           // [rcap]
  uint32_t* seedbits;  // [rcap / 32] roots that hold a seed
  uint32_t* selbits;   // [rcap / 32] runs whose component holds a seed
  uint32_t* area;      // [scap]
  uint32_t* yl;        // [scap] first row << 16 | smallest column below the first row (0xFFFF: none)
  uint32_t* maxx;      // [scap]
  uint32_t* maxy;      // [scap]
};

namespace rbg {

RB_HD size_t fixed_bytes(uint32_t H, uint32_t NW) {  // bit maps + prefix tables + misc
  return (size_t)H * NW * (3 * 4 + 2) + (size_t)H * 4 + 64 * 4 + 16;
}
template <typename L>
RB_HD size_t table_bytes(uint32_t rcap, uint32_t scap) {
  return (size_t)rcap * sizeof(L) + ((size_t)rcap + 31) / 32 * 8 + (size_t)scap * 16 + 32;
}

// carve the fixed part out of `fix` and the tables out of `tab` (both 16-byte aligned)
template <typename L>
RB_HD RbFgWork<L> carve(uint8_t* fix, uint8_t* tab, uint32_t H, uint32_t NW, uint32_t rcap, uint32_t scap) {
  RbFgWork<L> s;
  const size_t nw = (size_t)H * NW;
  s.start = reinterpret_cast<uint32_t*>(fix);
  s.ev = s.start + nw;
  s.seed = s.ev + nw;
  s.rowbase = s.seed + nw;
  s.misc = s.rowbase + H;
  s.base16 = reinterpret_cast<uint16_t*>(s.misc + 64);
  s.seedbits = reinterpret_cast<uint32_t*>(tab);
  s.selbits = s.seedbits + (rcap + 31) / 32;
  s.area = s.selbits + (rcap + 31) / 32;
  s.yl = s.area + scap;
  s.maxx = s.yl + scap;
  s.maxy = s.maxx + scap;
  s.parent = reinterpret_cast<L*>(s.maxy + scap);
  return s;
}

#if defined(__CUDA_ARCH__)
RB_D uint32_t cas_label(uint16_t* p, uint32_t cmp, uint32_t v) {
  return atomicCAS(reinterpret_cast<unsigned short*>(p), (unsigned short)cmp, (unsigned short)v);
}
RB_D uint32_t cas_label(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
RB_D uint32_t ld_label(const uint16_t* p) { return *reinterpret_cast<const volatile uint16_t*>(p); }
RB_D uint32_t ld_label(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
RB_D void a_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
RB_D void a_min(uint32_t* p, uint32_t v) { atomicMin(p, v); }
RB_D void a_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
RB_D uint32_t ldg32(const uint32_t* p) { return __ldg(p); }
#else
template <typename L>
inline uint32_t cas_label(L* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = (L)v; return o; }
template <typename L>
inline uint32_t ld_label(const L* p) { return *p; }
inline void a_or(uint32_t* p, uint32_t v) { *p |= v; }
inline void a_min(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
inline void a_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
inline uint32_t ldg32(const uint32_t* p) { return *p; }
#endif

// 4 bytes with values 0..15 each -> 4 bits, bit i set <=> byte i != 0
RB_HD uint32_t nz4(uint32_t v) {
  const uint32_t t = ((v + 0x0F0F0F0Fu) & 0x10101010u) >> 4;  // bits 0, 8, 16, 24
  return ((t * 0x00204081u) >> 21) & 15u;                      // gathered into bits 0..3 (no carries)
}

// 9 aligned words starting at byte pointer p rounded down to 4
RB_HD void load9(const uint8_t* p, uint32_t* w) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = ldg32(q + i);
}

RB_HD uint32_t low_mask(uint32_t b) { return 0xFFFFFFFFu >> (31u - b); }  // bits 0..b, b <= 31

// columns of word w that belong to the interior (1 .. W - 2)
RB_HD uint32_t interior_cols(uint32_t W, uint32_t w) {
  const uint32_t x0 = 32u * w;
  uint32_t m = 0xFFFFFFFFu;
  if (x0 == 0) m &= ~1u;
  if (W - 1 <= x0) return 0u;
  if (W - 1 < x0 + 32u) m &= (1u << (W - 1 - x0)) - 1u;  // bits for x < W - 1
  return m;
}

// statistics slot of run k's component, RB_FG_NONE when the component holds no seed.  After the slot
// phase a seeded root holds R + slot, every other run holds its root's id (< R).
template <typename L>
RB_HD uint32_t slot_of(const RbFgWork<L>& s, uint32_t R, uint32_t k) {
  uint32_t r = s.parent[k];
  if (r < R) {
    r = s.parent[r];
    if (r < R) return RB_FG_NONE;
  }
  return r - R;
}

// ---- phase bodies (one work item each; the kernel and tests/emul loop them over the threads) ----------

// Phase A: the three bit maps of word (y, w).
template <typename L>
RB_HD void build_word(const RbFgParams& p, const RbFgWork<L>& s, const RbPlacement& pl, uint32_t y, uint32_t w) {
  const RbGeom& g = p.g;
  const uint32_t at = y * p.NW + w;
  uint32_t st = 0, ev = 0, sd = 0;
  const uint32_t cols = interior_cols(g.W, w);
  if (y >= 1 && y + 3 <= g.H && cols) {  // rows 1 .. H - 3
    uint32_t m[9], u[9];
    const uint8_t* mrow = p.median + (uint64_t)pl.frame * g.median_stride + (uint64_t)y * g.mpitch + 32u * w;
    load9(mrow, m);            // bytes 32w .. 32w+35 of the row = pixels 32w-2 .. 32w+33
    load9(mrow - g.mpitch, u);  // the row above (row 0 exists)
    uint32_t ne_l = 0, ne_u = 0, ne_ul = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t cur = rb_funnel_r(m[j], m[j + 1], 16), lft = rb_funnel_r(m[j], m[j + 1], 8);
      const uint32_t up = rb_funnel_r(u[j], u[j + 1], 16), ulf = rb_funnel_r(u[j], u[j + 1], 8);
      ne_l |= nz4(cur ^ lft) << (4 * j);
      ne_u |= nz4(cur ^ up) << (4 * j);
      ne_ul |= nz4(up ^ ulf) << (4 * j);
    }
    const uint32_t first = w == 0 ? 2u : 0u;  // x == 1 always starts a run (its left neighbour is horizon)
    st = (ne_l | first) & cols;
    if (y >= 2) {
      const uint32_t eq = ~ne_u & cols;                // same colour as the interior pixel above
      const uint32_t st_up = (ne_ul | first) & cols;   // run starts of the row above
      ev = eq & (st | st_up | ~(eq << 1));             // one event per (run, run above) overlap (+ word starts)
    }
    // seeds: fde::details::generate_mask == 0 (src/fde.hpp:19-55,87), background window at (pl.x, pl.y)
    uint32_t b[9];
    const uint8_t* brow = p.bg + (uint64_t)(pl.y + (int32_t)y) * p.bgW + (uint32_t)pl.x + 32u * w;
    load9(brow, b);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(brow) & 3) * 8;
    const uint8_t* frow = p.frames + (uint64_t)pl.frame * g.frame_stride + (uint64_t)y * g.pitch + 32u * w;  // 16-byte aligned
    uint32_t f[8];
#if defined(__CUDA_ARCH__)
    {
      const uint4 f0 = __ldg(reinterpret_cast<const uint4*>(frow)), f1 = __ldg(reinterpret_cast<const uint4*>(frow) + 1);
      f[0] = f0.x; f[1] = f0.y; f[2] = f0.z; f[3] = f0.w; f[4] = f1.x; f[5] = f1.y; f[6] = f1.z; f[7] = f1.w;
    }
#else
    for (int j = 0; j < 8; ++j) f[j] = reinterpret_cast<const uint32_t*>(frow)[j];
#endif
    uint32_t ne_b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) ne_b |= nz4(f[j] ^ rb_funnel_r(b[j], b[j + 1], sh)) << (4 * j);
    sd = ne_b & cols;
  }
  s.start[at] = st;
  s.ev[at] = ev;
  s.seed[at] = sd;
}

// ---- warp-lockstep item loops ---------------------------------------------------------------------
// Every phase below is "for each word, for each set bit / run segment of the word".  The counts differ
// from word to word, and a plain per-thread loop lets the lanes of a warp drift apart for good (measured:
// 2 to 5 active lanes per issued instruction).  The loops are therefore written so that all lanes of a
// warp take the same number of trips: RB_WARP_ANY(c) (rb_common.cuh) is a warp vote on the device (all 32 lanes
// must reach it) and plain `c` in the host build.

// find with path halving: any ancestor is a valid parent, so the plain store races harmlessly with other
// finds; a CAS can only succeed on a root, and a root is never written here.
template <typename L>
RB_HD uint32_t find_halve(L* parent, uint32_t a) {
  for (;;) {
    const uint32_t p = ld_label(parent + a);
    if (p == a) return a;
    const uint32_t gp = ld_label(parent + p);
    if (gp == p) return p;
    parent[a] = (L)gp;
    a = gp;
  }
}

template <typename L>
RB_HD void unite_halve(L* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = find_halve(parent, a);
    b = find_halve(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }
    if (cas_label(parent + a, a, b) == a) return;  // the larger root now points at the smaller one
  }
}

// Phase D: the unions of word `it` (on = this lane has a word).
template <typename L>
RB_HD void unite_word(const RbFgParams& p, const RbFgWork<L>& s, uint32_t it, bool on) {
  const uint32_t y = on ? it / p.NW : 0;
  on = on && y >= 2 && y + 3 <= p.g.H;
  uint32_t e = on ? s.ev[it] : 0u;
  uint32_t stc = 0, stu = 0, bc = 0, bu = 0;
  if (e) {
    stc = s.start[it]; stu = s.start[it - p.NW];
    bc = s.rowbase[y] + s.base16[it] - 1u; bu = s.rowbase[y - 1] + s.base16[it - p.NW] - 1u;
  }
  while (RB_WARP_ANY(e != 0)) {
    if (e) {
      const uint32_t lm = low_mask(rb_ffs0(e));
      e &= e - 1;
      unite_halve(s.parent, bc + rb_popc(stc & lm), bu + rb_popc(stu & lm));
    }
  }
}

// Phase E: label k -> its root.
template <typename L>
RB_HD void flatten_label(const RbFgWork<L>& s, uint32_t k, bool on) {
  uint32_t r = on ? k : 0;
  bool go = on;
  while (RB_WARP_ANY(go)) {
    if (go) {
      const uint32_t q = ld_label(s.parent + r);
      if (q == r) go = false; else r = q;
    }
  }
  if (on) s.parent[k] = (L)r;
}

// Phase F: the seeds of word `it` mark their roots (parent[] is flat by now).
template <typename L>
RB_HD void seed_word(const RbFgParams& p, const RbFgWork<L>& s, uint32_t it, bool on) {
  uint32_t e = 0, st = 0, base = 0;
  if (on) {
    const uint32_t sd = s.seed[it];
    st = s.start[it];
    e = sd & ~((sd << 1) & ~st);  // a seed right of a seed of the same run adds nothing
    if (e) base = s.rowbase[it / p.NW] + s.base16[it] - 1u;
  }
  while (RB_WARP_ANY(e != 0)) {
    if (e) {
      const uint32_t lm = low_mask(rb_ffs0(e));
      e &= e - 1;
      const uint32_t r = s.parent[base + rb_popc(st & lm)];
      a_or(s.seedbits + (r >> 5), 1u << (r & 31));
    }
  }
}

// Phase G: seeded roots that START in word `it` take a statistics slot.
template <typename L>
RB_HD void slot_word(const RbFgParams& p, const RbFgWork<L>& s, uint32_t R, uint32_t it, bool on) {
  if (!on) return;
  const uint32_t n = rb_popc(s.start[it]);
  if (!n) return;
  const uint32_t y = it / p.NW, k0 = s.rowbase[y] + s.base16[it];
  // the seed marks of runs k0 .. k0 + n - 1 (n <= 32) as one word
  const uint32_t lo = s.seedbits[k0 >> 5] >> (k0 & 31);
  const uint32_t hi = (k0 & 31) && ((k0 + n - 1) >> 5) != (k0 >> 5) ? s.seedbits[(k0 >> 5) + 1] << (32 - (k0 & 31)) : 0u;
  uint32_t m = (lo | hi) & low_mask(n - 1);
  while (m) {  // only roots are ever marked; few per word
    const uint32_t k = k0 + rb_ffs0(m);
    m &= m - 1;
    const uint32_t slot = rb_atomic_add(s.misc + 1, 1u);
    if (slot < p.scap) {
      s.parent[k] = (L)(R + slot);
      s.area[slot] = 0;
      s.yl[slot] = (y << 16) | 0xFFFFu;  // the root is the component's first run in row-major order: y = top row
      s.maxx[slot] = 0;
      s.maxy[slot] = 0;
    }
  }
}

// Phase G2: bit k of selbits <=> run k belongs to a component that holds a seed.  One bit per thread, 32 runs
// per warp and trip; the warp's bits leave as one word (ballot), no atomics.
template <typename L>
RB_HD void select_run(const RbFgWork<L>& s, uint32_t R, uint32_t k, bool on) {
  bool bit = false;
  if (on) {
    const uint32_t r = s.parent[k];
    bit = r >= R || ((s.seedbits[r >> 5] >> (r & 31)) & 1u);  // a seeded root already holds R + slot
  }
#if defined(__CUDA_ARCH__)
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
  if ((threadIdx.x & 31) == 0 && (k >> 5) < (R + 31) / 32) s.selbits[k >> 5] = m;
#else
  if (bit) s.selbits[k >> 5] |= 1u << (k & 31);  // zeroed in phase C3
#endif
}

// bits [k0, k0 + n) of selbits, n <= 33 (bit 32 is dropped: callers pass n <= 32 and treat run k0 + 32 separately)
template <typename L>
RB_HD uint32_t sel_range(const RbFgWork<L>& s, uint32_t k0, uint32_t n) {
  const uint32_t lo = s.selbits[k0 >> 5] >> (k0 & 31);
  const uint32_t hi = (k0 & 31) && ((k0 + n - 1) >> 5) != (k0 >> 5) ? s.selbits[(k0 >> 5) + 1] << (32 - (k0 & 31)) : 0u;
  return (lo | hi) & low_mask(n - 1);
}

// Iterator over the run segments of one word: (run id k, first bit b, last bit e).
struct RbFgSeg {
  uint32_t st, top, b, k;
  uint32_t sel;  // bit j: the j-th segment from here on belongs to a seeded component
  bool more;
};
template <typename L>
RB_HD RbFgSeg seg_begin(const RbFgParams& p, const RbFgWork<L>& s, uint32_t it, bool on) {
  RbFgSeg g;
  g.st = g.top = g.b = g.k = g.sel = 0;
  g.more = false;
  if (!on) return g;
  const uint32_t y = it / p.NW, w = it - y * p.NW;
  const uint32_t cols = interior_cols(p.g.W, w);
  if (!cols || y < 1 || y + 3 > p.g.H) return g;
  g.st = s.start[it];
#if defined(__CUDA_ARCH__)
  g.top = 31u - (uint32_t)__clz((int)cols);
#else
  g.top = 31u - (uint32_t)__builtin_clz(cols);
#endif
  g.b = rb_ffs0(cols);
  g.k = s.rowbase[y] + s.base16[it] + rb_popc(g.st & low_mask(g.b)) - 1u;
  // the word's segments are runs g.k .. g.k + nseg - 1 (nseg <= 32: at most 31 starts after the first interior bit);
  // words without a seeded run -- most of them -- are skipped whole
  const uint32_t nseg = 1u + rb_popc(g.st & ~low_mask(g.b));
  g.sel = sel_range(s, g.k, nseg);
  g.more = g.sel != 0;
  return g;
}
RB_HD uint32_t seg_end(const RbFgSeg& g) {  // last bit of the current segment
  const uint32_t higher = g.b < 31 ? (g.st >> (g.b + 1)) << (g.b + 1) : 0u;
  return higher ? rb_ffs0(higher) - 1u : g.top;
}
RB_HD void seg_next(RbFgSeg& g, uint32_t e) {
  g.sel >>= 1;
  if (e >= g.top || g.sel == 0) { g.more = false; return; }  // nothing seeded further right
  g.b = e + 1;
  ++g.k;
}

// Phase H: statistics of the seeded components.
template <typename L>
RB_HD void stats_word(const RbFgParams& p, const RbFgWork<L>& s, uint32_t R, uint32_t it, bool on) {
  RbFgSeg g = seg_begin(p, s, it, on);
  const uint32_t y = it / p.NW, x32 = 32u * (it - y * p.NW);
  while (RB_WARP_ANY(g.more)) {
    if (g.more) {
      const uint32_t e = seg_end(g);
      if (g.sel & 1u) {
        const uint32_t slot = slot_of(s, R, g.k);
        const uint32_t x0 = x32 + g.b, x1 = x32 + e;
        rb_atomic_add(s.area + slot, e - g.b + 1u);
        if (x1 > s.maxx[slot]) a_max(s.maxx + slot, x1);
        if (y > s.maxy[slot]) a_max(s.maxy + slot, y);
        const uint32_t yl = s.yl[slot];
        if (y > (yl >> 16) && x0 < (yl & 0xFFFFu)) a_min(s.yl + slot, (yl & 0xFFFF0000u) | x0);
      }
      seg_next(g, e);
    }
  }
}

// Phase I: kept components paint their runs (every word is written by exactly one item).
template <typename L>
RB_HD void paint_word(const RbFgParams& p, const RbFgWork<L>& s, uint32_t R, uint32_t it, bool on) {
  RbFgSeg g = seg_begin(p, s, it, on);
  uint32_t out = 0;
  while (RB_WARP_ANY(g.more)) {
    if (g.more) {
      const uint32_t e = seg_end(g);
      if ((g.sel & 1u) && s.area[slot_of(s, R, g.k)] <= p.area_limit) out |= low_mask(e) & ~(low_mask(g.b) >> 1);
      seg_next(g, e);
    }
  }
  if (on) s.ev[it] = out;
}

// Phase J: the enclosures of the kept components (src/fde.hpp:133-143).  Most contours are a few pixels, so
// J1 gives every slot ONE lane that paints its rectangle if it has at most RB_FG_BOX_SMALL (row, word) items;
// larger rectangles are queued (in the seed bit map's memory, dead since phase F) and painted by a whole warp
// each in J2.  A full queue only costs time: the lane then paints its large rectangle itself.
#define RB_FG_BOX_SMALL 16u

struct RbFgBox {
  uint32_t top, w0, nww, total, left, right;
};
template <typename L>
RB_HD RbFgBox box_of(const RbFgParams& p, const RbFgWork<L>& s, uint32_t slot) {
  RbFgBox b;
  b.top = s.yl[slot] >> 16;
  b.left = s.yl[slot] & 0xFFFFu;  // 0xFFFF: single-row contour, nothing below the first row
  b.right = s.maxx[slot];
  const uint32_t bottom = s.maxy[slot];
  b.w0 = b.left >> 5;
  b.nww = 0; b.total = 0;
  if (b.left < b.right && b.top < bottom) {
    b.nww = ((b.right - 1) >> 5) - b.w0 + 1;
    b.total = (bottom - b.top) * b.nww;
  }
  return b;
}
template <typename L>
RB_HD void box_item(const RbFgParams& p, const RbFgWork<L>& s, const RbFgBox& b, uint32_t j) {
  const uint32_t y = b.top + j / b.nww, w = b.w0 + j % b.nww;
  uint32_t m = 0xFFFFFFFFu;
  if (w == b.w0) m &= ~(low_mask(b.left & 31) >> 1);                              // bits >= left
  if (w == b.w0 + b.nww - 1 && (b.right & 31)) m &= low_mask((b.right & 31) - 1);  // bits < right
  a_or(s.ev + y * p.NW + w, m);
}

template <typename L>
RB_HD void box_small(const RbFgParams& p, const RbFgWork<L>& s, uint32_t slot, bool on, uint32_t qcap) {
  RbFgBox b;
  b.total = 0; b.nww = 1; b.top = b.w0 = b.left = b.right = 0;
  const bool kept = on && s.area[slot] <= p.area_limit;
#if defined(__CUDA_ARCH__)
  const uint32_t nk = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, kept));
  if ((threadIdx.x & 31) == 0 && nk) atomicAdd(s.misc + 2, nk);
#else
  if (kept) ++s.misc[2];
#endif
  if (kept) {
    b = box_of(p, s, slot);
    if (b.total > RB_FG_BOX_SMALL) {
      const uint32_t at = rb_atomic_add(s.misc + 3, 1u);
      if (at < qcap) { s.seed[at] = slot; b.total = 0; }
    }
  }
  uint32_t j = 0;
  while (RB_WARP_ANY(j < b.total)) {
    if (j < b.total) { box_item(p, s, b, j); ++j; }
  }
}

template <typename L>
RB_HD void box_large_lane(const RbFgParams& p, const RbFgWork<L>& s, uint32_t slot, uint32_t lane) {
  const RbFgBox b = box_of(p, s, slot);
  for (uint32_t j = lane; j < b.total; j += 32) box_item(p, s, b, j);
}

// One frame through all phases.  Returns false (block-uniformly) when the frame does not fit the tables.
template <typename L>
RB_HD bool frame_body(const RbFgParams& p, const RbFgWork<L>& s, uint32_t i, uint32_t NT) {
  const RbGeom& g = p.g;
  const uint32_t H = g.H, NW = p.NW, nwords = H * NW;
  const RbPlacement pl = p.places[i];
  RB_FOR_THREADS(tid, NT) {  // A: bit maps
    for (uint32_t it = tid; it < nwords; it += NT) build_word(p, s, pl, it / NW, it % NW);
    if (tid < 8) s.misc[tid] = 0;
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // B: run starts before each word of a row; row totals
    for (uint32_t y = tid; y < H; y += NT) {
      uint32_t acc = 0;
      for (uint32_t w = 0; w < NW; ++w) {
        s.base16[y * NW + w] = (uint16_t)acc;
        acc += rb_popc(s.start[y * NW + w]);
      }
      s.rowbase[y] = acc;
    }
  }
  RB_SYNC();
  const uint32_t chunk = (H + 31) / 32;
  RB_FOR_THREADS(tid, NT) {  // C1: 32 chunk sums
    if (tid < 32) {
      uint32_t acc = 0;
      for (uint32_t y = tid * chunk; y < (tid + 1) * chunk && y < H; ++y) acc += s.rowbase[y];
      s.misc[8 + tid] = acc;
    }
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // C2: scan of the chunk sums
    if (tid == 0) {
      uint32_t acc = 0;
      for (uint32_t c = 0; c < 32; ++c) {
        const uint32_t v = s.misc[8 + c];
        s.misc[8 + c] = acc;
        acc += v;
      }
      s.misc[0] = acc;
    }
  }
  RB_SYNC();
  const uint32_t R = s.misc[0];
  if (R > p.rcap) return false;
  RB_FOR_THREADS(tid, NT) {  // C3: row bases; labels; seed marks
    if (tid < 32) {
      uint32_t acc = s.misc[8 + tid];
      for (uint32_t y = tid * chunk; y < (tid + 1) * chunk && y < H; ++y) {
        const uint32_t v = s.rowbase[y];
        s.rowbase[y] = acc;
        acc += v;
      }
    }
    for (uint32_t k = tid; k < R; k += NT) s.parent[k] = (L)k;
    for (uint32_t k = tid; k < (R + 31) / 32; k += NT) { s.seedbits[k] = 0; s.selbits[k] = 0; }
  }
  RB_SYNC();
  // the item loops below run the same number of trips in every lane (see RB_WARP_ANY)
  RB_FOR_THREADS(tid, NT) {  // D: unions
    for (uint32_t base = 0; base < nwords; base += NT) unite_word(p, s, base + tid, base + tid < nwords);
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // E: flatten
    for (uint32_t base = 0; base < R; base += NT) flatten_label(s, base + tid, base + tid < R);
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // F: seeds
    for (uint32_t base = 0; base < nwords; base += NT) seed_word(p, s, base + tid, base + tid < nwords);
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // G: slots
    for (uint32_t base = 0; base < nwords; base += NT) slot_word(p, s, R, base + tid, base + tid < nwords);
  }
  RB_SYNC();
  const uint32_t nslots = s.misc[1];
  if (nslots > p.scap) return false;
  RB_FOR_THREADS(tid, NT) {  // G2: runs of seeded components
    for (uint32_t base = 0; base < R; base += NT) select_run(s, R, base + tid, base + tid < R);
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // H: statistics
    for (uint32_t base = 0; base < nwords; base += NT) stats_word(p, s, R, base + tid, base + tid < nwords);
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // I: runs of the kept components
    for (uint32_t base = 0; base < nwords; base += NT) paint_word(p, s, R, base + tid, base + tid < nwords);
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // J1: small enclosures, one lane each
    for (uint32_t base = 0; base < nslots; base += NT) box_small(p, s, base + tid, base + tid < nslots, nwords);
  }
  RB_SYNC();
  const uint32_t nlarge = s.misc[3] < nwords ? s.misc[3] : nwords;
  RB_FOR_THREADS(tid, NT) {  // J2: large enclosures, one warp each
    for (uint32_t q = tid / 32; q < nlarge; q += NT / 32) box_large_lane(p, s, s.seed[q], tid & 31);
  }
  RB_SYNC();
  RB_FOR_THREADS(tid, NT) {  // K: out
    uint32_t* out = p.fgbits + (uint64_t)i * nwords;
    for (uint32_t it = tid; it < nwords; it += NT) out[it] = s.ev[it];
    if (tid == 0 && p.nkept) p.nkept[i] = s.misc[2];
  }
  RB_SYNC();
  return true;
}

}  // namespace rbg

#if defined(__CUDACC__)

// Fast variant: everything in shared memory, 16-bit labels.  Frame i -> CTA i mod gridDim.
__global__ void __launch_bounds__(2 * RB_FG_NT) rb_fg_kernel(const RbFgParams p) {
  extern __shared__ __align__(16) uint8_t rb_fg_smem[];
  const size_t fixed = (rbg::fixed_bytes(p.g.H, p.NW) + 15) & ~(size_t)15;
  const RbFgWork<uint16_t> s = rbg::carve<uint16_t>(rb_fg_smem, rb_fg_smem + fixed, p.g.H, p.NW, p.rcap, p.scap);
  for (uint32_t i = blockIdx.x; i < p.n; i += gridDim.x) {
    if (!rbg::frame_body(p, s, i, blockDim.x)) {
      if (threadIdx.x == 0) p.deferred[atomicAdd(p.ndeferred, 1u)] = i;
      __syncthreads();
    }
  }
}

// General variant: bit maps in shared memory, labels (32 bit) and statistics in this CTA's global slab.  Launched
// with RB_FG_NT threads, or with twice as many when the bit maps leave room for only one CTA per SM (640x480).
__global__ void __launch_bounds__(2 * RB_FG_NT) rb_fg_general_kernel(const RbFgParams p) {
  extern __shared__ __align__(16) uint8_t rb_fg_smem[];
  const RbFgWork<uint32_t> s =
      rbg::carve<uint32_t>(rb_fg_smem, p.scratch + p.scratch_stride * blockIdx.x, p.g.H, p.NW, p.rcap, p.scap);
  const uint32_t n = p.todo ? *p.ntodo : p.n;
  for (uint32_t j = blockIdx.x; j < n; j += gridDim.x) {
    const uint32_t i = p.todo ? p.todo[j] : j;
    if (!rbg::frame_body(p, s, i, blockDim.x)) {  // cannot happen: the slab holds the worst case
      if (threadIdx.x == 0) p.nkept[i] = 0xFFFFFFFFu;
      __syncthreads();
    }
  }
}

// Parity tap: the bit map of frames [0, n) as bytes (1 = foreground), H*W per frame.
__global__ void rb_fgbits_bytes_kernel(const uint32_t* __restrict__ bits, uint32_t n, uint32_t W, uint32_t H, uint32_t NW,
                                       uint8_t* __restrict__ out) {
  const size_t total = (size_t)n * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(i % W);
    const size_t row = i / W;
    out[i] = (bits[row * NW + (x >> 5)] >> (x & 31)) & 1u;
  }
}

#endif  // __CUDACC__
