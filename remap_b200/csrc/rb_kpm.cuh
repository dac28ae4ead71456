// rb_kpm.cuh -- K2: per-region matching + offset voting, K3: Borda count + declaration.
// Replaces kpm::match(cfg, prev_grid, curr_grid) (src/kpm.hpp:395-415) and everything under it:
// get_active :186-197, cast_vote :213-223, count_offsets :105-125, get_offsets :91-103,
// top_offsets :127-159, count :172-184, declare :199-211.
//
// The reference keeps, per region, an unordered_map<13-byte code, vector<point>> and joins the
// current map against the previous one, pushing every (prev, curr) pair of an equal code into an
// unordered_map<offset, count>.  Here one CTA owns one (frame pair, region):
//   1. both frames' region tiles (+2 px halo) are packed to 4 bit/pixel in shared memory;
//      the 5x5 patch IS the code (src/kpe.hpp:342-379 is a bijective packing of the 25 nibbles; the
//      weight nibble is a function of the patch), so equality of codes == equality of patches;
//   2. keypoints are enumerated straight from K1's bit maps (no keypoint records in HBM);
//   3. previous keypoints go into an open-addressing hash table in shared memory (tag = 32-bit
//      hash of the 100-bit patch, payload = position); current keypoints probe it, verify the full
//      patch on a tag hit, and vote prev - curr (src/kpm.hpp:96-98) into a second shared-memory
//      hash table keyed by offset (atomicCAS to claim a bin, atomicAdd to count);
//   4. the ticket = top region_votes bins by (count desc, dx asc, dy asc) -- a DEFINED order where
//      the reference has unordered_map iteration order (src/kpm.hpp:134-138) -- plus the tie
//      statistics (#bins > and >= each ticket count) that K3 needs to flag tie-sensitive pairs.
// Bounded shared memory, exact for any input: previous keypoints are processed in row BANDS that
// fit the code table, and if the offset table would overflow the CTA restarts with the offset
// space hash-PARTITIONED into Q parts handled one after another (votes are additive).
#pragma once

#include "rb_common.cuh"

struct RbKpmParams {
  RbGeom g;
  const uint8_t* frames;
  const uint32_t* kpbits;
  const uint32_t* w2bits;
  RbRegionVote* votes;   // [npairs][nreg]; pair i = frames (first_frame + i, first_frame + i + 1)
  uint32_t first_frame;
  uint32_t npairs;
  uint32_t code_slots;   // power of two; a band holds at most code_slots / 2 previous keypoints
  uint32_t off_slots;    // power of two; a partition holds at most off_slots / 2 distinct offsets
  uint32_t tile_pitch;   // words per packed tile row (8 px per word) incl. one pad word
  uint32_t tile_rows;    // max region height + 4
  // parity tap (rb_region_votes): dump every bin of (tap_pair, tap_region); otherwise tap_bins == nullptr
  RbBin* tap_bins;
  uint32_t tap_cap;
  uint32_t* tap_count;
  uint32_t tap_pair, tap_region;
};

namespace rbm {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;

struct Smem {  // carved from dynamic shared memory, all uint32_t-aligned
  uint32_t* tile[2];     // [tile_rows][tile_pitch] packed 4 bit/pixel tiles of prev / curr frame
  uint32_t* rowcnt[4];   // per region row: prev all, prev w2, curr all, curr w2
  uint32_t* segend[2];   // slow path only: row ends of the prev bands / curr chunks
  uint32_t* ctab;        // [code_slots] 19-bit hash tag (bits 12..30) | 12-bit index into plist; never == EMPTY
  uint32_t* plist;       // [code_slots / 2] previous keypoints of the band, (y << 16) | x
  uint32_t* clist;       // [code_slots / 2] current keypoints of the chunk
  uint32_t* okey;        // [off_slots]
  uint32_t* ocnt;        // [off_slots]
  uint32_t* touched;     // [off_slots / 2 + NT + 1] slots claimed in this partition
  uint32_t* scal;        // scalars, see S_*
  unsigned long long* best;  // [4]: per-partition selection rounds
};
enum {
  S_NPREV = 0, S_W2PREV, S_NCURR, S_W2CURR, S_NTOUCHED, S_OVERFLOW, S_NGT0, S_NGE0 = S_NGT0 + 3,
  S_PFILL = S_NGE0 + 3, S_CFILL, S_NBAND, S_NCHUNK, S_HASH, S_COUNT
};

RB_HD size_t smem_words(const RbKpmParams& p, uint32_t NT) {
  return (size_t)2 * p.tile_rows * p.tile_pitch + 6 * (size_t)p.tile_rows + 2 * (size_t)p.code_slots +
         2 * (size_t)p.off_slots + (p.off_slots / 2 + NT + 1) + S_COUNT + 2 /*align*/ + 8 /*best*/;
}

RB_HD void carve(const RbKpmParams& p, uint32_t NT, uint32_t* base, Smem& s) {
  uint32_t* q = base;
  s.tile[0] = q; q += (size_t)p.tile_rows * p.tile_pitch;
  s.tile[1] = q; q += (size_t)p.tile_rows * p.tile_pitch;
  for (int k = 0; k < 4; ++k) { s.rowcnt[k] = q; q += p.tile_rows; }
  for (int k = 0; k < 2; ++k) { s.segend[k] = q; q += p.tile_rows; }
  s.ctab = q; q += p.code_slots;
  s.plist = q; q += p.code_slots / 2;
  s.clist = q; q += p.code_slots / 2;
  s.okey = q; q += p.off_slots;
  s.ocnt = q; q += p.off_slots;
  s.touched = q; q += p.off_slots / 2 + NT + 1;
  s.scal = q; q += S_COUNT;
  if ((reinterpret_cast<uintptr_t>(q) & 7) != 0) ++q;
  s.best = reinterpret_cast<unsigned long long*>(q);
}

// 8 byte-pixels (two words) -> 8 nibbles in one word, pixel i in bits [4i, 4i+4)
RB_HD uint32_t pack8(uint32_t a, uint32_t b) {
  a &= 0x0F0F0F0Fu; b &= 0x0F0F0F0Fu;
  a = (a | (a >> 4)) & 0x00FF00FFu; a = (a | (a >> 8)) & 0xFFFFu;
  b = (b | (b >> 4)) & 0x00FF00FFu; b = (b | (b >> 8)) & 0xFFFFu;
  return a | (b << 16);
}

struct Code { uint32_t c0, c1, c2, c3; };

// The 5x5 patch around (lx + 2, ly + 2) in tile coordinates as 100 bits: row r, column c of the
// patch = nibble 5r + c.  lx, ly = tile coordinates of the patch's top-left pixel.
RB_HD Code code_at(const uint32_t* tile, uint32_t pitch, uint32_t lx, uint32_t ly) {
  uint32_t r[5];
  const uint32_t wi = lx >> 3, sh = (lx & 7) * 4;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint32_t* row = tile + (size_t)(ly + k) * pitch + wi;
    r[k] = rb_funnel_r(row[0], row[1], sh) & 0xFFFFFu;
  }
  Code c;
  c.c0 = r[0] | (r[1] << 20);
  c.c1 = (r[1] >> 12) | (r[2] << 8) | (r[3] << 28);
  c.c2 = (r[3] >> 4) | (r[4] << 16);
  c.c3 = r[4] >> 16;
  return c;
}

RB_HD uint32_t code_hash(const Code& c) {
  uint32_t h = c.c0 * 0x9E3779B1u ^ c.c1 * 0x85EBCA77u ^ c.c2 * 0xC2B2AE3Du ^ c.c3 * 0x27D4EB2Fu;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
  return h;
}

RB_HD uint32_t off_hash(uint32_t key) {
  uint32_t h = key * 0x9E3779B1u;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15;
  return h;
}

// bits of strip word j (bit i <-> x = 28 j + i, outputs at bits 2..29) that lie in [X0, X1)
RB_HD uint32_t colmask(uint32_t X0, uint32_t X1, uint32_t j) {
  const int lo = (int)X0 - (int)(RB_STRIP_OUT * j), hi = (int)X1 - (int)(RB_STRIP_OUT * j);
  uint32_t m = 0x3FFFFFFCu;
  if (lo > 2) m &= ~((1u << lo) - 1u);
  if (hi < 30) m &= (1u << (hi < 0 ? 0 : hi)) - 1u;
  return m;
}

// selection key: larger = earlier in the ticket (count desc, dx asc, dy asc)
RB_HD unsigned long long sel_key(uint32_t okey, uint32_t cnt) {
  const uint32_t dxb = okey & 0xFFFFu, dyb = okey >> 16;  // biased by 32768
  const uint32_t ord = (dxb << 16) | dyb;
  return ((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - ord);
}
RB_HD RbBin sel_decode(unsigned long long k) {
  const uint32_t ord = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
  RbBin b;
  b.dx = (int32_t)(ord >> 16) - 32768;
  b.dy = (int32_t)(ord & 0xFFFFu) - 32768;
  b.cnt = (uint32_t)(k >> 32);
  return b;
}

// Appends the keypoints of one region row (strip words [j0, j0+nstr) masked to [X0, X1)) to a list.
RB_HD void emit_row(const uint32_t* bits, uint64_t rowbase, uint32_t j0, uint32_t nstr, uint32_t X0, uint32_t X1,
                    uint32_t y, uint32_t* list, uint32_t at) {
  for (uint32_t k = 0; k < nstr; ++k) {
    const uint32_t j = j0 + k;
#if defined(__CUDA_ARCH__)
    uint32_t w = __ldg(bits + rowbase + j) & colmask(X0, X1, j);
#else
    uint32_t w = bits[rowbase + j] & colmask(X0, X1, j);
#endif
    while (w) {
      const uint32_t b = rb_ffs0(w);
      w &= w - 1;
      list[at++] = (y << 16) | (RB_STRIP_OUT * j + b);
    }
  }
}

// One (pair, region).  smem_base: dynamic shared memory (device) or a heap buffer (host test).
RB_HD void kpm_block(const RbKpmParams& p, uint32_t pair, uint32_t region, uint32_t* smem_base, uint32_t NT) {
  const RbGeom& g = p.g;
  Smem s;
  carve(p, NT, smem_base, s);
  const uint32_t cs = region / g.grid_h, rs = region % g.grid_h;  // idx = grid_h*col + row (src/kpr.hpp:71-74)
  const uint32_t X0 = g.col0[cs], X1 = g.col1[cs], Y0 = g.row0[rs], Y1 = g.row1[rs];
  const uint32_t tx0 = (X0 - 2) & ~7u;                     // tile origin (pixels), 8-aligned
  const uint32_t tw = (X1 + 2 - tx0 + 7) / 8;              // packed words per tile row
  const uint32_t th = Y1 - Y0 + 4;                         // tile rows: Y0-2 .. Y1+1
  const uint32_t nrows = Y1 - Y0;
  const uint32_t j0 = (X0 - 2) / RB_STRIP_OUT, j1 = (X1 - 1 - 2) / RB_STRIP_OUT;  // strips that hold [X0, X1)
  const uint32_t nstr = j1 - j0 + 1;
  const uint32_t fprev = p.first_frame + pair, fcurr = fprev + 1;
  const bool tap = p.tap_bins != nullptr && pair == p.tap_pair && region == p.tap_region;
  const uint64_t kprow_prev = ((uint64_t)fprev * g.H + Y0) * g.NS, kprow_curr = ((uint64_t)fcurr * g.H + Y0) * g.NS;

  // ---- phase Z: zero the scalars ----------------------------------------------------------------
  RB_FOR_THREADS(tid, NT) {
    for (uint32_t i = tid; i < S_COUNT; i += NT) s.scal[i] = 0;
    if (tid < 4) s.best[tid] = 0;
  }
  RB_SYNC();

  // ---- phase 0: load + pack both tiles; per-row keypoint counts (all / weight 2) ------------------
  RB_FOR_THREADS(tid, NT) {
    // tiles: KW = power of two >= tw + 1 columns per row so that row/column come from shifts
    uint32_t lkw = 4;
    while ((1u << lkw) < tw + 1) ++lkw;
    const uint32_t KW = 1u << lkw;
    const uint32_t items = th << lkw;
    for (uint32_t fr = 0; fr < 2; ++fr) {
      const uint8_t* base = p.frames + (uint64_t)(fr ? fcurr : fprev) * g.frame_stride + (uint64_t)(Y0 - 2) * g.pitch + tx0;
      uint32_t* dst = s.tile[fr];
      for (uint32_t i0 = tid; i0 < items; i0 += 4 * NT) {
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // issue the loads of four rows before packing any
          const uint32_t i = i0 + u * NT, row = i >> lkw, k = i & (KW - 1);
          lo[u] = hi[u] = 0;
          if (i < items && k < tw) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(base + (uint64_t)row * g.pitch + 8 * k);
#if defined(__CUDA_ARCH__)
            const uint2 ab = __ldg(reinterpret_cast<const uint2*>(src));
            lo[u] = ab.x; hi[u] = ab.y;
#else
            lo[u] = src[0]; hi[u] = src[1];
#endif
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t i = i0 + u * NT, row = i >> lkw, k = i & (KW - 1);
          if (i < items && k <= tw) dst[(size_t)row * p.tile_pitch + k] = pack8(lo[u], hi[u]);  // k == tw: pad word 0
        }
      }
    }
    // counts: one thread per (frame, row)
    for (uint32_t i = tid; i < 2 * nrows; i += NT) {
      const uint32_t fr = i >= nrows ? 1u : 0u, row = fr ? i - nrows : i;
      const uint64_t rb = (fr ? kprow_curr : kprow_prev) + (uint64_t)row * g.NS;
      uint32_t ca = 0, cb = 0;
      for (uint32_t k = 0; k < nstr; ++k) {
        const uint32_t m = colmask(X0, X1, j0 + k);
        ca += rb_popc(p.kpbits[rb + j0 + k] & m);
        cb += rb_popc(p.w2bits[rb + j0 + k] & m);
      }
      s.rowcnt[2 * fr][row] = ca;
      s.rowcnt[2 * fr + 1][row] = cb;
      if (ca) rb_atomic_add(&s.scal[S_NPREV + 2 * fr], ca);
      if (cb) rb_atomic_add(&s.scal[S_W2PREV + 2 * fr], cb);
    }
  }
  RB_SYNC();

  const uint32_t n_prev = s.scal[S_NPREV], w2_prev = s.scal[S_W2PREV];
  const uint32_t n_curr = s.scal[S_NCURR], w2_curr = s.scal[S_W2CURR];
  // src/kpm.hpp:219-220 ('<' on previous, '<=' on current)
  const bool use_all = (w2_prev < g.weight_switch) || (w2_curr <= g.weight_switch);
  const uint32_t* prow = use_all ? s.rowcnt[0] : s.rowcnt[1];  // effective keypoints per row
  const uint32_t* crow = use_all ? s.rowcnt[2] : s.rowcnt[3];
  const uint32_t eff_prev = use_all ? n_prev : w2_prev;
  const uint32_t eff_curr = use_all ? n_curr : w2_curr;
  const uint32_t* kp_src = use_all ? p.kpbits : p.w2bits;      // !use_all: weight-2 codes only (src/kpm.hpp:113-117)

  const uint32_t rv = g.region_votes;
  const uint32_t lcap = p.code_slots / 2, off_cap = p.off_slots / 2;
  unsigned long long gtop[3] = {0, 0, 0};
  uint32_t nbins = 0;
  uint32_t Q = 1;

  if (eff_prev != 0 && eff_curr != 0) {
    // Row bands of the previous frame / row chunks of the current frame that fit the lists.  The
    // common case is one band and one chunk; otherwise two threads cut the rows serially.
    const bool fast = eff_prev <= lcap && eff_curr <= lcap;
    if (!fast) {
      RB_FOR_THREADS(tid, NT) {
        if (tid < 2) {
          const uint32_t* cnt = tid ? crow : prow;
          uint32_t nseg = 0, acc = 0;
          for (uint32_t r = 0; r < nrows; ++r) {
            if (acc + cnt[r] > lcap) { s.segend[tid][nseg++] = r; acc = 0; }  // a single row always fits (host check)
            acc += cnt[r];
          }
          s.segend[tid][nseg++] = nrows;
          s.scal[S_NBAND + tid] = nseg;
        }
      }
      RB_SYNC();
    }
    const uint32_t nband = fast ? 1u : s.scal[S_NBAND], nchunk = fast ? 1u : s.scal[S_NCHUNK];

    bool restart = true;
    while (restart) {
      restart = false;
      gtop[0] = gtop[1] = gtop[2] = 0;
      nbins = 0;
      for (uint32_t sweep = 0; sweep < 2 && !restart; ++sweep) {
        for (uint32_t q = 0; q < Q && !restart; ++q) {
          if (sweep == 0 || Q > 1) {
            // ---- build the offset histogram of partition q ---------------------------------
            for (uint32_t band = 0; band < nband; ++band) {
              const uint32_t ra = (fast || band == 0) ? 0u : s.segend[0][band - 1];
              const uint32_t rb = fast ? nrows : s.segend[0][band];
              uint32_t bcnt = eff_prev;
              if (!fast) { bcnt = 0; for (uint32_t r = ra; r < rb; ++r) bcnt += prow[r]; }
              uint32_t cslots = 64;
              while (cslots < 2 * bcnt) cslots <<= 1;
              const uint32_t cmask = cslots - 1;
              // phase A: clear tables and fill counters
              RB_FOR_THREADS(tid, NT) {
                for (uint32_t i = tid; i < cslots; i += NT) s.ctab[i] = EMPTY;
                if (band == 0) {
                  for (uint32_t i = tid; i < p.off_slots; i += NT) { s.okey[i] = EMPTY; s.ocnt[i] = 0; }
                  if (tid == 0) { s.scal[S_NTOUCHED] = 0; s.scal[S_OVERFLOW] = 0; }
                  if (tid < 6) s.scal[S_NGT0 + tid] = (sweep == 0) ? 0u : s.scal[S_NGT0 + tid];
                  if (tid < 3 && sweep == 0) s.best[tid] = 0;
                  if (tap && tid == 0 && sweep == 0 && q == 0) *p.tap_count = 0;
                  if (tid == 0 && sweep == 0 && q == 0) s.scal[S_HASH] = 0;  // also after a restart
                }
                if (tid == 0) { s.scal[S_PFILL] = 0; s.scal[S_CFILL] = 0; }
              }
              RB_SYNC();
              // phase B: compact the band's previous keypoints (and, with a single chunk, the current ones)
              RB_FOR_THREADS(tid, NT) {
                const uint32_t nprow = rb - ra;
                const uint32_t extra = (nchunk == 1 && band == 0) ? nrows : 0u;
                for (uint32_t i = tid; i < nprow + extra; i += NT) {
                  if (i < nprow) {
                    const uint32_t row = ra + i, c = prow[row];
                    if (c) emit_row(kp_src, kprow_prev + (uint64_t)row * g.NS, j0, nstr, X0, X1, Y0 + row, s.plist,
                                    rb_atomic_add(&s.scal[S_PFILL], c));
                  } else {
                    const uint32_t row = i - nprow, c = crow[row];
                    if (c) emit_row(kp_src, kprow_curr + (uint64_t)row * g.NS, j0, nstr, X0, X1, Y0 + row, s.clist,
                                    rb_atomic_add(&s.scal[S_CFILL], c));
                  }
                }
              }
              RB_SYNC();
              // phase C: insert the previous keypoints into the code table
              RB_FOR_THREADS(tid, NT) {
                for (uint32_t i = tid; i < bcnt; i += NT) {
                  const uint32_t pp = s.plist[i];
                  const Code c = code_at(s.tile[0], p.tile_pitch, (pp & 0xFFFFu) - 2 - tx0, (pp >> 16) - Y0);
                  const uint32_t h = code_hash(c);
                  uint32_t slot = h & cmask;
                  while (rb_atomic_cas(&s.ctab[slot], EMPTY, (h & 0x7FFFF000u) | i) != EMPTY) slot = (slot + 1) & cmask;
                }
              }
              RB_SYNC();
              for (uint32_t chunk = 0; chunk < nchunk; ++chunk) {
                uint32_t ccnt = eff_curr;
                if (nchunk > 1) {
                  const uint32_t ca = chunk == 0 ? 0u : s.segend[1][chunk - 1], cb = s.segend[1][chunk];
                  ccnt = 0;
                  for (uint32_t r = ca; r < cb; ++r) ccnt += crow[r];
                  RB_FOR_THREADS(tid, NT) {
                    if (tid == 0) s.scal[S_CFILL] = 0;
                  }
                  RB_SYNC();
                  RB_FOR_THREADS(tid, NT) {
                    for (uint32_t i = tid; i < cb - ca; i += NT) {
                      const uint32_t row = ca + i, c = crow[row];
                      if (c) emit_row(kp_src, kprow_curr + (uint64_t)row * g.NS, j0, nstr, X0, X1, Y0 + row, s.clist,
                                      rb_atomic_add(&s.scal[S_CFILL], c));
                    }
                  }
                  RB_SYNC();
                }
                // phase D: probe with every current keypoint of the chunk, vote prev - curr
                RB_FOR_THREADS(tid, NT) {
                  for (uint32_t i = tid; i < ccnt; i += NT) {
                    const uint32_t cp = s.clist[i];
                    const uint32_t x = cp & 0xFFFFu, y = cp >> 16;
                    const Code c = code_at(s.tile[1], p.tile_pitch, x - 2 - tx0, y - Y0);
                    const uint32_t h = code_hash(c);
                    uint32_t slot = h & cmask;
                    uint32_t e;
                    while ((e = s.ctab[slot]) != EMPTY) {
                      if (((e ^ h) & 0x7FFFF000u) == 0) {
                        const uint32_t pp = s.plist[e & 0xFFFu];
                        const uint32_t px = pp & 0xFFFFu, py = pp >> 16;
                        const Code d = code_at(s.tile[0], p.tile_pitch, px - 2 - tx0, py - Y0);
                        if (d.c0 == c.c0 && d.c1 == c.c1 && d.c2 == c.c2 && d.c3 == c.c3) {
                          // offset = prev - curr (src/kpm.hpp:96-98), biased into 16 + 16 bits
                          const uint32_t key = ((py - y + 32768u) << 16) | ((px - x + 32768u) & 0xFFFFu);
                          const uint32_t oh = off_hash(key);
                          if (Q == 1 || (oh >> 12) % Q == q) {
                            uint32_t os = oh & (p.off_slots - 1);
                            while (true) {
                              uint32_t k = s.okey[os];
                              if (k == EMPTY) {
                                if (rb_volatile_load(&s.scal[S_NTOUCHED]) >= off_cap) { s.scal[S_OVERFLOW] = 1; break; }
                                k = rb_atomic_cas(&s.okey[os], EMPTY, key);
                                if (k == EMPTY) {
                                  const uint32_t t = rb_atomic_add(&s.scal[S_NTOUCHED], 1);
                                  s.touched[t] = os;
                                  k = key;
                                }
                              }
                              if (k == key) { rb_atomic_add(&s.ocnt[os], 1); break; }
                              os = (os + 1) & (p.off_slots - 1);
                            }
                          }
                        }
                      }
                      slot = (slot + 1) & cmask;
                    }
                  }
                }
                RB_SYNC();
              }
            }
            if (s.scal[S_OVERFLOW] != 0) {  // block-uniform: read after the barrier
              Q *= 4;
              restart = true;
              RB_SYNC();
              break;
            }
          }
          const uint32_t nt = s.scal[S_NTOUCHED];
          if (sweep == 0) {
            nbins += nt;
            RB_FOR_THREADS(tid, NT) {  // digest of the histogram: every bin lives in exactly one partition
              uint32_t hh = 0;
              for (uint32_t i = tid; i < nt; i += NT) {
                const uint32_t os = s.touched[i];
                const RbBin b = sel_decode(sel_key(s.okey[os], s.ocnt[os]));
                hh += rb_bin_hash(b.dx, b.dy, b.cnt);
              }
              if (hh) rb_atomic_add(&s.scal[S_HASH], hh);
            }
            if (tap) {
              RB_FOR_THREADS(tid, NT) {
                for (uint32_t i = tid; i < nt; i += NT) {
                  const uint32_t os = s.touched[i];
                  const uint32_t at = rb_atomic_add(p.tap_count, 1);
                  if (at < p.tap_cap) p.tap_bins[at] = sel_decode(sel_key(s.okey[os], s.ocnt[os]));
                }
              }
            }
            // ---- ticket of this partition: rv rounds of "largest key below the previous one" --
            unsigned long long below = ~0ull;
            unsigned long long ptop[3] = {0, 0, 0};
            for (uint32_t round = 0; round < rv; ++round) {
              RB_FOR_THREADS(tid, NT) {
                unsigned long long loc = 0;
                for (uint32_t i = tid; i < nt; i += NT) {
                  const uint32_t os = s.touched[i];
                  const unsigned long long k = sel_key(s.okey[os], s.ocnt[os]);
                  if (k < below && k > loc) loc = k;
                }
                if (loc) rb_atomic_max64(&s.best[round], loc);
              }
              RB_SYNC();
              ptop[round] = s.best[round];
              below = ptop[round];
              if (below == 0) break;
            }
            if (Q > 1) RB_SYNC();  // best[] is reset by the next partition's phase A
            // merge into the running top (uniform, registers only)
            for (uint32_t a = 0; a < rv; ++a) {
              unsigned long long v = ptop[a];
              if (v == 0) break;
              for (uint32_t b2 = 0; b2 < rv; ++b2)
                if (v > gtop[b2]) { const unsigned long long t = gtop[b2]; gtop[b2] = v; v = t; }
            }
          } else {
            // ---- tie statistics against the final ticket ------------------------------------
            RB_FOR_THREADS(tid, NT) {
              uint32_t gt[3] = {0, 0, 0}, ge[3] = {0, 0, 0};
              for (uint32_t i = tid; i < nt; i += NT) {
                const uint32_t c = s.ocnt[s.touched[i]];
                for (uint32_t k = 0; k < rv; ++k) {
                  const uint32_t ck = (uint32_t)(gtop[k] >> 32);
                  if (ck != 0) { gt[k] += c > ck; ge[k] += c >= ck; }
                }
              }
              for (uint32_t k = 0; k < rv; ++k) {
                if (gt[k]) rb_atomic_add(&s.scal[S_NGT0 + k], gt[k]);
                if (ge[k]) rb_atomic_add(&s.scal[S_NGE0 + k], ge[k]);
              }
            }
            RB_SYNC();
          }
        }
      }
    }
  }

  // ---- write the region's ballot ---------------------------------------------------------------
  RB_FOR_THREADS(tid, NT) {
    if (tid == 0) {
      RbRegionVote v;
      v.use_all = use_all ? 1u : 0u;
      v.n_prev = n_prev; v.n_curr = n_curr; v.w2_prev = w2_prev; v.w2_curr = w2_curr;
      v.nbins = nbins;
      v.hist_hash = s.scal[S_HASH];
      v.nticket = nbins < rv ? nbins : rv;
      for (uint32_t k = 0; k < 4; ++k) {
        RbBin b; b.dx = 0; b.dy = 0; b.cnt = 0;
        v.ticket[k] = b; v.ngt[k] = 0; v.nge[k] = 0;
        if (k < v.nticket) {
          v.ticket[k] = sel_decode(gtop[k]);
          v.ngt[k] = s.scal[S_NGT0 + k];
          v.nge[k] = s.scal[S_NGE0 + k];
        }
      }
      p.votes[(uint64_t)pair * g.nreg + region] = v;
    }
  }
}

// ---- K3: Borda count + declare + tie sensitivity for one pair (src/kpm.hpp:172-184,199-211) ----
// Same analysis as ro_declare in oracle/remap_oracle.c (documented there and in DESIGN.md).
RB_HD void declare_pair(const RbGeom& g, const RbRegionVote* votes, RbPairResult* res) {
  const uint32_t nreg = g.nreg, rv = g.region_votes;
  RbPairResult r;
  r.dx = r.dy = 0; r.valid = 0; r.tie_sensitive = 0; r.active = 0; r.ntop = 0;
  r.top_dx[0] = r.top_dx[1] = r.top_dy[0] = r.top_dy[1] = 0; r.top_score[0] = r.top_score[1] = 0;
  uint32_t active = 0;
  for (uint32_t i = 0; i < nreg; ++i) active += votes[i].n_curr > 0;  // current grid only (src/kpm.hpp:400)
  r.active = active;
  if (active >= nreg / 4) {  // src/kpm.hpp:401
    int32_t cdx[RB_MAX_REGIONS * 3], cdy[RB_MAX_REGIONS * 3];
    uint32_t csc[RB_MAX_REGIONS * 3];
    uint32_t nc = 0;
    for (uint32_t i = 0; i < nreg; ++i)
      for (uint32_t k = 0; k < votes[i].nticket; ++k) {  // total[off] += rank-- (src/kpm.hpp:176-181)
        const int32_t dx = votes[i].ticket[k].dx, dy = votes[i].ticket[k].dy;
        uint32_t j = 0;
        while (j < nc && !(cdx[j] == dx && cdy[j] == dy)) ++j;
        if (j == nc) { cdx[nc] = dx; cdy[nc] = dy; csc[nc] = 0; ++nc; }
        csc[j] += rv - k;
      }
    if (nc != 0) {
      int b0 = -1, b1 = -1;
      for (uint32_t j = 0; j < nc; ++j) {
        auto before = [&](uint32_t a, int b) {
          if (csc[a] != csc[b]) return csc[a] > csc[b];
          if (cdx[a] != cdx[b]) return cdx[a] < cdx[b];
          return cdy[a] < cdy[b];
        };
        if (b0 < 0 || before(j, b0)) { b1 = b0; b0 = (int)j; }
        else if (b1 < 0 || before(j, b1)) { b1 = (int)j; }
      }
      r.ntop = b1 >= 0 ? 2 : 1;
      r.top_dx[0] = cdx[b0]; r.top_dy[0] = cdy[b0]; r.top_score[0] = csc[b0];
      if (b1 >= 0) { r.top_dx[1] = cdx[b1]; r.top_dy[1] = cdy[b1]; r.top_score[1] = csc[b1]; }
      const uint32_t half = active / 2;  // src/kpm.hpp:206
      const uint32_t S0 = csc[b0], S1 = b1 >= 0 ? csc[b1] : 0;
      if (b1 >= 0 && S0 < S1 + half) r.valid = 0;
      else { r.valid = 1; r.dx = cdx[b0]; r.dy = cdy[b0]; }
      // Tie sensitivity: bounds lo(o) <= score(o) <= hi(o) under any order of count-tied bins; see
      // ro_declare in the oracle and DESIGN.md for the derivation.
      uint32_t lo[RB_MAX_REGIONS * 3], hi[RB_MAX_REGIONS * 3], hi_out = 0;
      bool ambiguous = false;
      for (uint32_t j = 0; j < nc; ++j) lo[j] = hi[j] = 0;
      for (uint32_t i = 0; i < nreg; ++i) {
        const RbRegionVote& v = votes[i];
        uint32_t og = 0;
        if (v.nticket == rv && v.nge[rv - 1] > rv) og = v.ngt[rv - 1] < rv ? rv - v.ngt[rv - 1] : 0;
        hi_out += og;
        for (uint32_t j = 0; j < nc; ++j) {
          uint32_t k = 0;
          while (k < v.nticket && !(v.ticket[k].dx == cdx[j] && v.ticket[k].dy == cdy[j])) ++k;
          if (k < v.nticket) {
            const uint32_t worst = v.nge[k] - 1;
            hi[j] += v.ngt[k] < rv ? rv - v.ngt[k] : 0;
            lo[j] += worst < rv ? rv - worst : 0;
            if (v.nge[k] != v.ngt[k] + 1) ambiguous = true;
          } else {
            hi[j] += og;
          }
        }
      }
      if (!ambiguous) r.tie_sensitive = 0;
      else if (r.valid) {
        uint32_t Hm = hi_out;
        for (uint32_t j = 0; j < nc; ++j)
          if ((int)j != b0 && hi[j] > Hm) Hm = hi[j];
        r.tie_sensitive = !(lo[b0] >= Hm + (half > 1 ? half : 1));
      } else {
        uint32_t Hm = hi_out, l1 = 0, l2 = 0;
        for (uint32_t j = 0; j < nc; ++j) {
          if (hi[j] > Hm) Hm = hi[j];
          if (lo[j] > l1) { l2 = l1; l1 = lo[j]; } else if (lo[j] > l2) l2 = lo[j];
        }
        r.tie_sensitive = !(l2 >= 1 && Hm < l2 + half);
      }
    }
  }
  *res = r;
}

}  // namespace rbm

#if defined(__CUDACC__)
__global__ void __launch_bounds__(1024) rb_kpm_kernel(const RbKpmParams p) {
  extern __shared__ __align__(16) uint32_t rb_kpm_smem[];
  const uint32_t pair = blockIdx.x / p.g.nreg, region = blockIdx.x % p.g.nreg;
  rbm::kpm_block(p, pair, region, rb_kpm_smem, blockDim.x);
}
#endif
