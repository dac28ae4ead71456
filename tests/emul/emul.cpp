// tests/emul/emul.cpp -- HOST build of the kernel bodies for CPU unit tests.
//
// TEST INFRASTRUCTURE ONLY.  The CUDA kernels in remap_b200/csrc/*.cuh are written as plain
// host+device code; this file compiles the very same source with g++ and runs every work item (or
// every thread of a block, phase by phase) in a serial loop, so kernel logic can be diffed against
// the oracle in the GPU-less build container.  It is never linked into libremap_b200.so and no
// product path can reach it (the product library has no host compute path at all).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define RB_EMUL 1
#include "../../remap_b200/csrc/rb_host.hpp"
#include "../../remap_b200/csrc/rb_kpe.cuh"
#include "../../remap_b200/csrc/rb_prep.cuh"
#include "../../remap_b200/csrc/rb_fg.cuh"
#include "../../remap_b200/csrc/rb_splice.cuh"

extern "C" {

// frames: n*H*W bytes (dense).  median: n*H*W (dense, may be NULL).  kp/w2: n*H*NS words.
int emul_kpe(const uint8_t* frames, uint32_t n, uint32_t W, uint32_t H, uint32_t nseg, uint8_t* median,
             uint32_t* kpbits, uint32_t* w2bits) {
  RbKpeParams p;
  if (rb_make_geom(W, H, 4, 2, 16, 10, 3, &p.g) != 0) return -1;
  const RbGeom& g = p.g;
  std::vector<uint8_t> dfr((size_t)g.frame_stride * n + 64, 0xEE);  // garbage beyond the rows
  for (uint32_t f = 0; f < n; ++f)
    for (uint32_t y = 0; y < H; ++y) memcpy(&dfr[f * g.frame_stride + (size_t)y * g.pitch], frames + ((size_t)f * H + y) * W, W);
  std::vector<uint8_t> dmed((size_t)g.median_stride * n + 64, 0);
  p.frames = dfr.data();
  p.median = median ? dmed.data() : nullptr;
  p.kpbits = kpbits;
  p.w2bits = w2bits;
  p.nframes = n;
  p.nseg = nseg;
  const uint32_t rows = H - 6;
  p.seg_rows = (rows + nseg - 1) / nseg;
  memset(kpbits, 0, (size_t)n * H * g.NS * 4);
  memset(w2bits, 0, (size_t)n * H * g.NS * 4);
  for (uint32_t f = 0; f < n; ++f)
    for (uint32_t s = 0; s < nseg; ++s)
      for (uint32_t j = 0; j < g.NS; ++j) rbk::kpe_strip(p, f, s, j);
  if (median)
    for (uint32_t f = 0; f < n; ++f)
      for (uint32_t y = 0; y < H; ++y)
        memcpy(median + ((size_t)f * H + y) * W, &dmed[f * g.median_stride + (size_t)y * g.mpitch + 2], W);
  return (int)g.NS;
}

// K1c (rb_prep.cuh): the lane functions of rb_list_kernel run lane by lane, with the warp scan done
// serially here.  kp / w2: n*H*NS words (from emul_kpe).  lists: n*8*cap words, counts: n*8*2 words.
int emul_lists(const uint32_t* kp, const uint32_t* w2, uint32_t n, uint32_t W, uint32_t H, uint32_t cap, uint32_t* lists,
               uint32_t* counts) {
  RbGeom g;
  if (rb_make_geom(W, H, 4, 2, 16, 10, 3, &g) != 0) return -1;
  memset(lists, 0xFF, (size_t)n * g.nreg * cap * 4);
  for (uint32_t f = 0; f < n; ++f)
    for (uint32_t r = 0; r < g.nreg; ++r) {
      const uint32_t* kpf = kp + (size_t)f * H * g.NS;
      const uint32_t* w2f = w2 + (size_t)f * H * g.NS;
      RbListLane L[32];
      uint32_t a2 = 0, a1 = 0;
      for (uint32_t lane = 0; lane < 32; ++lane) L[lane] = rbl::lane_setup(g, r, lane);
      if (L[0].rows_per_chunk == 0) { a2 = a1 = cap + 1; }
      else
        for (uint32_t ra = 0; ra < L[0].nrows; ra += L[0].rows_per_chunk)
          for (uint32_t lane = 0; lane < 32; ++lane) {  // lanes in order == the warp's exclusive scan
            uint32_t kw, ww;
            rbl::lane_words(g, L[lane], kpf, w2f, ra, kw, ww);
            rbl::lane_emit(L[lane], ra, kw, ww, a2, a1, cap, lists + ((size_t)f * g.nreg + r) * cap);
            a2 += (uint32_t)__builtin_popcount(ww); a1 += (uint32_t)__builtin_popcount(kw & ~ww);
          }
      counts[((size_t)f * g.nreg + r) * 2] = a2 + a1;
      counts[((size_t)f * g.nreg + r) * 2 + 1] = a2;
    }
  return (int)g.nreg;
}

uint32_t emul_strips(uint32_t W) { return (W - 4 + RB_STRIP_OUT - 1) / RB_STRIP_OUT; }

}  // extern "C"

#include "../../remap_b200/csrc/rb_kpm.cuh"

extern "C" {

size_t emul_sizeof_vote() { return sizeof(RbRegionVote); }
size_t emul_sizeof_result() { return sizeof(RbPairResult); }

// Runs K1 (host build) on all frames, then K2 on every (pair, region) and K3 on every pair.
// votes: (n-1)*8 RbRegionVote, results: (n-1) RbPairResult.  tap: bins of (tap_pair, tap_region).
int emul_register(const uint8_t* frames, uint32_t n, uint32_t W, uint32_t H, uint32_t code_slots, uint32_t off_slots,
                  uint32_t NT, RbRegionVote* votes, RbPairResult* results, int32_t tap_pair, int32_t tap_region,
                  RbBin* tap_bins, uint32_t tap_cap, uint32_t* tap_count) {
  RbKpmParams p;
  memset(&p, 0, sizeof(p));
  if (rb_make_geom(W, H, 4, 2, 16, 10, 3, &p.g) != 0) return -1;
  const RbGeom& g = p.g;
  std::vector<uint8_t> dfr((size_t)g.frame_stride * n + 64, 0xEE);
  for (uint32_t f = 0; f < n; ++f)
    for (uint32_t y = 0; y < H; ++y) memcpy(&dfr[f * g.frame_stride + (size_t)y * g.pitch], frames + ((size_t)f * H + y) * W, W);
  std::vector<uint32_t> kp((size_t)n * H * g.NS, 0), w2((size_t)n * H * g.NS, 0);
  {
    RbKpeParams k;
    k.g = g; k.frames = dfr.data(); k.median = nullptr; k.kpbits = kp.data(); k.w2bits = w2.data();
    k.nframes = n; k.nseg = 2; k.seg_rows = (H - 6 + 1) / 2;
    for (uint32_t f = 0; f < n; ++f)
      for (uint32_t s = 0; s < k.nseg; ++s)
        for (uint32_t j = 0; j < g.NS; ++j) rbk::kpe_strip(k, f, s, j);
  }
  uint32_t maxw = 0, maxh = 0;
  for (uint32_t s = 0; s < g.grid_w; ++s) {
    const uint32_t tx0 = (g.col0[s] - 2) & ~7u;
    const uint32_t tw = (g.col1[s] + 2 - tx0 + 7) / 8;
    if (tw > maxw) maxw = tw;
  }
  for (uint32_t s = 0; s < g.grid_h; ++s)
    if (g.row1[s] - g.row0[s] > maxh) maxh = g.row1[s] - g.row0[s];
  p.frames = dfr.data(); p.kpbits = kp.data(); p.w2bits = w2.data(); p.votes = votes;
  p.first_frame = 0; p.npairs = n - 1; p.code_slots = code_slots; p.off_slots = off_slots;
  p.tile_pitch = maxw + 1; p.tile_rows = maxh + 4;
  p.tap_bins = tap_pair >= 0 ? tap_bins : nullptr; p.tap_cap = tap_cap; p.tap_count = tap_count;
  p.tap_pair = (uint32_t)tap_pair; p.tap_region = (uint32_t)tap_region;
  std::vector<uint32_t> smem(rbm::smem_words(p, NT) + 4);
  for (uint32_t pair = 0; pair + 1 < n; ++pair) {
    for (uint32_t r = 0; r < g.nreg; ++r) {
      std::fill(smem.begin(), smem.end(), 0xDEADBEEFu);  // shared memory is not zero-initialised
      rbm::kpm_block(p, pair, r, smem.data(), NT);
    }
    rbm::declare_pair(g, votes + (size_t)pair * g.nreg, results + pair);
  }
  return 0;
}

}  // extern "C"

extern "C" {

// Pass-2 foreground (rb_fg.cuh): the phases of rbg::frame_body run thread by thread.  frames / medians:
// n*H*W bytes (dense), bg: bgH*bgW, places: n x (frame, x, y).  general = 0: 16-bit labels, tables of
// nplaces x (frame, x, y) placements; (rcap, scap) entries, returns the number of frames that did not fit (their bits are left untouched);
// general = 1: 32-bit labels, worst-case tables.  bits: n*H*NW words, nkept: n.
int emul_fg(const uint8_t* frames, const uint8_t* medians, uint32_t n, uint32_t W, uint32_t H, const uint8_t* bg, uint32_t bgW,
            uint32_t bgH, const int32_t* places, uint32_t nplaces, int general, uint32_t rcap, uint32_t scap, uint32_t* bits, uint32_t* nkept) {
  RbFgParams p;
  memset(&p, 0, sizeof(p));
  if (rb_make_geom(W, H, 4, 2, 16, 10, 3, &p.g) != 0) return -1;
  const RbGeom& g = p.g;
  std::vector<uint8_t> dfr((size_t)g.frame_stride * n + 256, 0x0E), dmed((size_t)g.median_stride * n + 256, 0x0D);
  for (uint32_t f = 0; f < n; ++f)
    for (uint32_t y = 0; y < H; ++y) {
      memcpy(&dfr[f * g.frame_stride + (size_t)y * g.pitch], frames + ((size_t)f * H + y) * W, W);
      memcpy(&dmed[f * g.median_stride + (size_t)y * g.mpitch + 2], medians + ((size_t)f * H + y) * W, W);
    }
  std::vector<uint8_t> dbg((size_t)bgW * bgH + 64, 0x0C);
  memcpy(dbg.data(), bg, (size_t)bgW * bgH);
  std::vector<RbPlacement> pl(nplaces);
  for (uint32_t i = 0; i < nplaces; ++i) { pl[i].frame = (uint32_t)places[3 * i]; pl[i].x = places[3 * i + 1]; pl[i].y = places[3 * i + 2]; }
  p.frames = dfr.data(); p.median = dmed.data(); p.bg = dbg.data(); p.bgW = bgW; p.bgH = bgH;
  p.places = pl.data(); p.n = nplaces; p.NW = (W + 31) / 32;
  p.area_limit = (uint32_t)(((uint64_t)W * H) / 5);
  p.fgbits = bits; p.nkept = nkept;
  const uint32_t rmax = (W - 2) * (H - 3);
  p.rcap = general ? rmax : rcap;
  p.scap = general ? rmax : scap;
  const size_t fixed = (rbg::fixed_bytes(H, p.NW) + 15) & ~(size_t)15;
  int deferred = 0;
  if (general) {
    std::vector<uint8_t> mem(fixed + rbg::table_bytes<uint32_t>(p.rcap, p.scap) + 64, 0xA5);
    const RbFgWork<uint32_t> s = rbg::carve<uint32_t>(mem.data(), mem.data() + fixed, H, p.NW, p.rcap, p.scap);
    for (uint32_t i = 0; i < nplaces; ++i)
      if (!rbg::frame_body(p, s, i, RB_FG_NT)) ++deferred;
  } else {
    if (p.rcap + p.scap > 65535) return -2;
    std::vector<uint8_t> mem(fixed + rbg::table_bytes<uint16_t>(p.rcap, p.scap) + 64, 0xA5);
    const RbFgWork<uint16_t> s = rbg::carve<uint16_t>(mem.data(), mem.data() + fixed, H, p.NW, p.rcap, p.scap);
    for (uint32_t i = 0; i < nplaces; ++i)
      if (!rbg::frame_body(p, s, i, RB_FG_NT)) ++deferred;
  }
  return deferred;
}

}  // extern "C"

extern "C" {

// Fragment splicing (rb_splice.cuh), item by item.  emul_snippet: dots (H*W*16 u16) -> blend image / mask and the
// keypoint records of K1 with a 1 x 1 grid; returns the keypoint count.  recs: cap x 5 words (c[4], xy).
int emul_snippet(const uint16_t* dots, uint32_t W, uint32_t H, uint8_t* image, uint8_t* mask, uint32_t* recs, uint32_t cap) {
  RbKpeParams k;
  if (rb_make_geom(W, H, 1, 1, 0, 10, 3, &k.g) != 0) return -1;
  const RbGeom& g = k.g;
  std::vector<uint8_t> dimg((size_t)g.frame_stride + 256, 0);
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t x = 0; x < W; ++x)
      rbs::blend_pixel(dots + ((size_t)y * W + x) * 16, &dimg[(size_t)y * g.pitch + x], mask + (size_t)y * W + x);
  for (uint32_t y = 0; y < H; ++y) memcpy(image + (size_t)y * W, &dimg[(size_t)y * g.pitch], W);
  std::vector<uint32_t> kp((size_t)H * g.NS, 0), w2((size_t)H * g.NS, 0);
  k.frames = dimg.data(); k.median = nullptr; k.kpbits = kp.data(); k.w2bits = w2.data(); k.nframes = 1;
  uint32_t nseg = (H - 6) / 8;
  if (nseg > 64) nseg = 64;
  if (nseg < 1) nseg = 1;
  k.nseg = nseg; k.seg_rows = (H - 6 + nseg - 1) / nseg;
  for (uint32_t s = 0; s < nseg; ++s)
    for (uint32_t j = 0; j < g.NS; ++j) rbk::kpe_strip(k, 0, s, j);
  uint32_t count = 0;
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t j = 0; j < g.NS; ++j)
      rbs::emit_word(g, dimg.data(), kp.data(), w2.data(), y, j, reinterpret_cast<RbSnipKp*>(recs), cap, &count);
  return (int)count;
}

// The cellular match of two record lists; out = {valid, dx, dy, matched_keypoints, matched_cells, active_cells,
// offsets, ties, pairs_lo, pairs_hi}.  The host steps between the kernels mirror rb_snippet_match.
int emul_cell_match(const uint32_t* prev, uint32_t np, const uint8_t* pmask, uint32_t pW, uint32_t pH, const uint32_t* curr,
                    uint32_t nc, uint32_t cW, uint32_t cH, uint32_t cell_w, uint32_t cell_h, uint32_t* out) {
  memset(out, 0, 10 * 4);
  if (np == 0 || nc == 0) return 0;
  RbCellParams p;
  memset(&p, 0, sizeof(p));
  p.prev = reinterpret_cast<const RbSnipKp*>(prev); p.np = np;
  p.curr = reinterpret_cast<const RbSnipKp*>(curr); p.nc = nc;
  p.pW = pW; p.pH = pH; p.cW = cW; p.cH = cH;
  p.nbuckets = 1024;
  while (p.nbuckets < 2 * np) p.nbuckets <<= 1;
  p.OW = pW + cW - 1; p.OH = pH + cH - 1;
  p.cell_w = cell_w; p.cell_h = cell_h;
  p.CW = (pW > cW ? pW : cW) / cell_w + 1;
  const uint32_t CH = (pH > cH ? pH : cH) / cell_h + 1;
  p.AW = cW / cell_w + 1;
  std::vector<uint32_t> head(p.nbuckets, 0xFFFFFFFFu), next(np), hist((size_t)p.OW * p.OH, 0), cells(((size_t)p.CW * CH + 31) / 32, 0),
      act(((size_t)p.AW * (cH / cell_h + 1) + 31) / 32, 0);
  p.head = head.data(); p.next = next.data(); p.hist = hist.data(); p.cellbits = cells.data(); p.actbits = act.data();
  for (uint32_t i = 0; i < np; ++i) rbs::build_item(p, i);
  for (uint32_t j = 0; j < nc; ++j) rbs::vote_item<0>(p, j);
  uint64_t best = 0, nz = 0, sum = 0;
  for (size_t i = 0; i < hist.size(); ++i)
    if (hist[i]) {
      ++nz; sum += hist[i];
      const uint64_t key = ((uint64_t)hist[i] << 32) | (0xFFFFFFFFu - (uint32_t)i);
      if (key > best) best = key;
    }
  out[6] = (uint32_t)nz; out[8] = (uint32_t)sum; out[9] = (uint32_t)(sum >> 32);
  if (!nz) return 0;
  const uint32_t votes = (uint32_t)(best >> 32), bin = 0xFFFFFFFFu - (uint32_t)best;
  p.best_dx = (int32_t)(bin % p.OW) - (int32_t)(cW - 1);
  p.best_dy = (int32_t)(bin / p.OW) - (int32_t)(cH - 1);
  auto limits = [](int32_t delta, uint64_t previous, uint64_t current, uint64_t* lo, uint64_t* hi) {
    if (delta < 0) { const uint64_t d = (uint64_t)(-(int64_t)delta); *lo = d; *hi = current < previous + d ? current : previous + d; }
    else { *lo = 0; *hi = current < previous - (uint64_t)delta ? current : previous - (uint64_t)delta; }
  };
  uint64_t l, r, t, b;
  limits(p.best_dx, pW, cW, &l, &r);
  limits(p.best_dy, pH, cH, &t, &b);
  p.lim_l = (uint32_t)l; p.lim_r = (uint32_t)r; p.lim_t = (uint32_t)t; p.lim_b = (uint32_t)b;
  p.pmask = pmask;
  uint32_t ties = 0;
  for (size_t i = 0; i < hist.size(); ++i) ties += hist[i] == votes;
  for (uint32_t j = 0; j < nc; ++j) rbs::vote_item<1>(p, j);
  for (uint32_t j = 0; j < nc; ++j) rbs::active_item(p, j);
  uint32_t mc = 0, ac = 0;
  for (uint32_t w : cells) mc += __builtin_popcount(w);
  for (uint32_t w : act) ac += __builtin_popcount(w);
  out[1] = (uint32_t)p.best_dx; out[2] = (uint32_t)p.best_dy; out[3] = votes; out[4] = mc; out[5] = ac; out[7] = ties;
  out[0] = !((float)mc < (float)ac * 0.66f);
  return 0;
}

}  // extern "C"
