/* oracle/remap_oracle.c -- CPU restatement of kataklinger/remap's per-frame registration path.
 *
 * TEST INFRASTRUCTURE ONLY.  Plain scalar C, no SIMD, no hash maps.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product path
 * (remap_b200/, include/) never does.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures, so
 * this restatement is pinned against outputs of the reference ITSELF, compiled in the build
 * container by oracle/build_ref.py (oracle/_ref/ref_harness): tests/test_oracle_vs_ref.py diffs
 * every intermediate on seeded sequences, and tests/golden/ holds committed dumps of the real
 * reference (made by tests/golden/make_golden.py) that the CPU test-suite checks on every run.
 *
 * One deliberate difference, flagged not hidden: kpm::details::top_offsets breaks count ties by
 * std::unordered_map iteration order (src/kpm.hpp:134-138), which is implementation-defined
 * (MSVC != libstdc++).  This file uses a DEFINED order (count desc, dx asc, dy asc) and computes a
 * conservative tie_sensitive flag per pair; on unflagged pairs the declared offset provably does
 * not depend on the tie order, so it must equal the reference's.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RO_MAX_REGIONS 64
#define RO_CODE_LEN 13

typedef struct {
  uint8_t code[RO_CODE_LEN]; /* src/kpr.hpp:20-23, layout src/kpe.hpp:342-379            */
  uint8_t weight;            /* 1 or 2 == code[12] & 0xf (src/kpr.hpp:25-27)              */
  uint16_t x, y;
  uint32_t region_mask;      /* bit i set <=> inserted into region i (src/kpr.hpp:189-219) */
} ro_keypoint;

typedef struct {
  uint32_t width, height, grid_w, grid_h, overlap, weight_switch, region_votes;
} ro_config;

typedef struct { int32_t dx, dy; uint32_t cnt; } ro_bin;

typedef struct {
  uint32_t use_all;        /* the weight switch of kpm::details::cast_vote (src/kpm.hpp:217-222) */
  uint32_t n_prev, n_curr; /* insertions into this region (w1 + w2)                                */
  uint32_t w2_prev, w2_curr;
  uint32_t nbins;          /* distinct offsets                                                     */
  uint32_t nticket;        /* min(region_votes, nbins)                                             */
  ro_bin ticket[4];        /* defined order: cnt desc, dx asc, dy asc                              */
  uint32_t ngt[4];         /* #bins with cnt >  ticket[k].cnt                                      */
  uint32_t nge[4];         /* #bins with cnt >= ticket[k].cnt (includes ticket[k])                 */
  uint32_t hist_hash;      /* wrapping sum over all bins of ro_bin_hash (digest of the whole histogram)   */
} ro_region_vote;

/* Digest of one bin of kpm's totalizator_t (src/kpm.hpp:70-76); defined in include/remap_b200.h. */
static uint32_t ro_bin_hash(int32_t dx, int32_t dy, uint32_t cnt) {
  uint32_t h = ((uint32_t)dx & 0xFFFFu) | ((uint32_t)dy << 16);
  h = h * 0x9E3779B1u ^ cnt * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

typedef struct {
  int32_t dx, dy;
  uint32_t valid;          /* == std::optional has_value of kpm::match (src/kpm.hpp:395-415) */
  uint32_t tie_sensitive;  /* declared result could depend on the undefined tie order        */
  uint32_t active;
  int32_t top_dx[2], top_dy[2];
  uint32_t top_score[2];
  uint32_t ntop;
} ro_match_result;

/* ---- a1: luminance-ordered colour LUTs (src/cpl.hpp:77-92,116-120,163-217) ----------------- */
static const uint32_t ro_palette[16] = { /* src/cpl.hpp:77-92, 0x00RRGGBB */
    0x00000000, 0x00FFFFFF, 0x0068372B, 0x0070A4B2, 0x006F3D86, 0x00588D43, 0x00352879, 0x00B8C76F,
    0x006F4F25, 0x00433900, 0x009A6759, 0x00444444, 0x006C6C6C, 0x009AD284, 0x006C5EB5, 0x00959595};

void ro_luts(uint8_t native_to_ordered[16], uint8_t ordered_to_native[16]) {
  /* intensity = (0.3 R + 0.59 G + 0.11 B) / 255 in float (src/cpl.hpp:116-120); blend_to_pack
   * (src/cpl.hpp:99-103) takes byte0 as "red", byte1 "green", byte2 "blue" of the 0x00RRGGBB word,
   * i.e. the weights are applied to (B, G, R) of the nominal colour -- reproduced literally. */
  float inten[16];
  int idx[16];
  for (int i = 0; i < 16; ++i) {
    uint32_t c = ro_palette[i];
    float red = (float)(c & 0xff), green = (float)((c >> 8) & 0xff), blue = (float)((c >> 16) & 0xff);
    inten[i] = (0.3f * red + 0.59f * green + 0.11f * blue) / 255.0f;
    idx[i] = i;
  }
  for (int i = 1; i < 16; ++i) { /* insertion sort, ascending intensity; no ties exist */
    int k = idx[i], j = i - 1;
    while (j >= 0 && inten[idx[j]] > inten[k]) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = k;
  }
  for (int o = 0; o < 16; ++o) {
    ordered_to_native[o] = (uint8_t)idx[o];
    native_to_ordered[idx[o]] = (uint8_t)o;
  }
}

/* ---- a2/A.4: region sections (src/kpe.hpp:84-90,157-192,235-277; src/kpr.hpp:71-74) -------- */
/* colsect[x] / rowsect[y]: bit s set <=> column/row belongs to grid section s; 0 outside the
 * keypoint domain x in [2, W-3], y in [2, H-5]. */
void ro_sections(const ro_config* cfg, uint32_t* colsect, uint32_t* rowsect) {
  const uint32_t W = cfg->width, H = cfg->height, gw = cfg->grid_w, gh = cfg->grid_h, O = cfg->overlap;
  const uint32_t rw = W / gw - O / 2, rh = H / gh - O / 2; /* src/kpe.hpp:86-87 */
  memset(colsect, 0, W * sizeof(uint32_t));
  memset(rowsect, 0, H * sizeof(uint32_t));
  uint32_t c = 2; /* col_out_gen<0> starts at column kernel_half (src/kpe.hpp:190) */
  for (uint32_t s = 0; s + 1 < gw; ++s) {
    for (uint32_t x = c; x < c + rw && x < W; ++x) colsect[x] |= 1u << s;                      /* low_t */
    for (uint32_t x = c + rw; x < c + rw + O && x < W; ++x) colsect[x] |= (1u << s) | (1u << (s + 1)); /* mid_t */
    c += rw + O;
  }
  for (uint32_t x = c; x + 2 < W; ++x) colsect[x] |= 1u << (gw - 1); /* last = W - kernel_half (src/kpe.hpp:183-184) */
  /* rows: y == 2 is handled by col_in itself with Inner = {0} (src/kpe.hpp:223-230) */
  if (H > 2) rowsect[2] |= 1u;
  uint32_t r = 3; /* col_in_gen<0> returns col + kernel_size, centre y = 3 (src/kpe.hpp:275,299) */
  for (uint32_t s = 0; s + 1 < gh; ++s) {
    for (uint32_t y = r; y < r + rh && y < H; ++y) rowsect[y] |= 1u << s;
    for (uint32_t y = r + rh; y < r + rh + O && y < H; ++y) rowsect[y] |= (1u << s) | (1u << (s + 1));
    r += rh + O;
  }
  for (uint32_t y = r; y + 4 < H; ++y) rowsect[y] |= 1u << (gh - 1); /* last = col + H - 2, centre = first - 2 (src/kpe.hpp:268,299) */
  /* rows/cols beyond the domain keep 0; clip sections that ran past the domain */
  for (uint32_t x = 0; x < W; ++x) if (x < 2 || x + 2 >= W) colsect[x] = 0;
  for (uint32_t y = 0; y < H; ++y) if (y < 2 || y + 4 >= H) rowsect[y] = 0;
}

static uint32_t ro_region_mask(const ro_config* cfg, uint32_t cs, uint32_t rs) {
  uint32_t m = 0; /* idx = grid_h * colsect + rowsect (src/kpr.hpp:71-74, src/kpe.hpp:79) */
  for (uint32_t a = 0; a < cfg->grid_w; ++a)
    if (cs & (1u << a))
      for (uint32_t b = 0; b < cfg->grid_h; ++b)
        if (rs & (1u << b)) m |= 1u << (cfg->grid_h * a + b);
  return m;
}

/* ---- a7: rank filter (src/kpe.hpp:308-340) -------------------------------------------------- */
static uint8_t ro_rank(const uint8_t hist[16], uint32_t half) {
  uint32_t total = 0; /* median_pixel: scan 15 -> 0, first i with cumulative >= half */
  for (int i = 15; i >= 0; --i) {
    total += hist[i];
    if (total >= half) return (uint8_t)i;
  }
  return 0;
}

/* ---- a3..a9: kpe::extractor::extract (src/kpe.hpp:92-108 and everything it calls) ---------- */
/* frame: W*H bytes 0..15.  median: W*H bytes, fully written (0 outside the domain, as the
 * reference's zero-initialised matrix keeps it).  kps: up to cap keypoints in the reference's
 * insertion order (column-major: x outer, y inner).  Returns the total number of keypoints (may
 * exceed cap; only cap are stored). */
size_t ro_extract(const ro_config* cfg, const uint8_t* frame, uint8_t* median, ro_keypoint* kps, size_t cap) {
  const uint32_t W = cfg->width, H = cfg->height;
  uint8_t n2o[16], o2n[16];
  ro_luts(n2o, o2n);
  uint32_t* colsect = (uint32_t*)malloc(W * sizeof(uint32_t));
  uint32_t* rowsect = (uint32_t*)malloc(H * sizeof(uint32_t));
  ro_sections(cfg, colsect, rowsect);
  if (median) memset(median, 0, (size_t)W * H);
  size_t n = 0;
  if (W >= 5 && H >= 7) {
    for (uint32_t x = 2; x + 2 < W; ++x) {
      for (uint32_t y = 2; y + 4 < H; ++y) { /* rows end at H-5 (src/kpe.hpp:268) */
        uint8_t h3[16] = {0}, h5[16] = {0};
        for (int j = -2; j <= 2; ++j)
          for (int i = -2; i <= 2; ++i) {
            uint8_t o = n2o[frame[(size_t)(y + j) * W + (x + i)] & 15];
            ++h5[o];
            if (i >= -1 && i <= 1 && j >= -1 && j <= 1) ++h3[o];
          }
        uint8_t p1 = n2o[frame[(size_t)y * W + x] & 15];
        uint8_t p3 = ro_rank(h3, 4);  /* src/kpe.hpp:313 */
        if (median) median[(size_t)y * W + x] = o2n[p3]; /* src/kpe.hpp:314 */
        if (p1 == p3) continue;
        uint8_t p5 = ro_rank(h5, 12); /* src/kpe.hpp:317 */
        if (p3 == p5) continue;
        uint8_t weight = (p1 != p5) ? 2 : 1; /* src/kpe.hpp:319 */
        if (n < cap && kps) {
          ro_keypoint* k = &kps[n];
          uint8_t v[5][5];
          for (int r = 0; r < 5; ++r)
            for (int c = 0; c < 5; ++c) v[r][c] = frame[(size_t)(y - 2 + r) * W + (x - 2 + c)] & 15;
          /* src/kpe.hpp:342-379: rows 0,2,4 "even" (2.5 bytes starting on a byte), rows 1,3 "odd" */
          k->code[0] = (uint8_t)(v[0][0] | (v[0][1] << 4));
          k->code[1] = (uint8_t)(v[0][2] | (v[0][3] << 4));
          k->code[2] = (uint8_t)(v[1][0] | (v[0][4] << 4));
          k->code[3] = (uint8_t)(v[1][1] | (v[1][2] << 4));
          k->code[4] = (uint8_t)(v[1][3] | (v[1][4] << 4));
          k->code[5] = (uint8_t)(v[2][0] | (v[2][1] << 4));
          k->code[6] = (uint8_t)(v[2][2] | (v[2][3] << 4));
          k->code[7] = (uint8_t)(v[3][0] | (v[2][4] << 4));
          k->code[8] = (uint8_t)(v[3][1] | (v[3][2] << 4));
          k->code[9] = (uint8_t)(v[3][3] | (v[3][4] << 4));
          k->code[10] = (uint8_t)(v[4][0] | (v[4][1] << 4));
          k->code[11] = (uint8_t)(v[4][2] | (v[4][3] << 4));
          k->code[12] = (uint8_t)(weight | (v[4][4] << 4));
          k->weight = weight;
          k->x = (uint16_t)x;
          k->y = (uint16_t)y;
          k->region_mask = ro_region_mask(cfg, colsect[x], rowsect[y]);
        }
        ++n;
      }
    }
  }
  free(colsect);
  free(rowsect);
  return n;
}

/* ---- a10..a13: kpm::match, grid variant (src/kpm.hpp:395-415) ------------------------------ */
typedef struct { uint8_t code[RO_CODE_LEN]; uint16_t x, y; } ro_entry;

static int ro_entry_cmp(const void* a, const void* b) {
  return memcmp(((const ro_entry*)a)->code, ((const ro_entry*)b)->code, RO_CODE_LEN);
}
static int ro_u64_cmp(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}
static int ro_bin_before(const ro_bin* a, const ro_bin* b) { /* defined order */
  if (a->cnt != b->cnt) return a->cnt > b->cnt;
  if (a->dx != b->dx) return a->dx < b->dx;
  return a->dy < b->dy;
}

/* One region: histogram of prev - curr over all pairs with equal code (src/kpm.hpp:91-125), then
 * the ticket (src/kpm.hpp:127-159).  bins_out (may be NULL) receives the histogram sorted by
 * (dx, dy); *nbins_out its size (bins_cap entries stored at most). */
static void ro_region(const ro_config* cfg, const ro_keypoint* prev, size_t np, const ro_keypoint* curr, size_t nc,
                      uint32_t region, ro_region_vote* out, ro_bin* bins_out, size_t bins_cap) {
  memset(out, 0, sizeof(*out));
  const uint32_t bit = 1u << region;
  ro_entry* pe = (ro_entry*)malloc((np + 1) * sizeof(ro_entry));
  size_t npe = 0;
  for (size_t i = 0; i < np; ++i)
    if (prev[i].region_mask & bit) {
      ++out->n_prev;
      if (prev[i].weight == 2) ++out->w2_prev;
    }
  for (size_t i = 0; i < nc; ++i)
    if (curr[i].region_mask & bit) {
      ++out->n_curr;
      if (curr[i].weight == 2) ++out->w2_curr;
    }
  /* src/kpm.hpp:219-220: note '<' on previous, '<=' on current */
  out->use_all = (out->w2_prev < cfg->weight_switch) || (out->w2_curr <= cfg->weight_switch);
  for (size_t i = 0; i < np; ++i)
    if (prev[i].region_mask & bit) {
      memcpy(pe[npe].code, prev[i].code, RO_CODE_LEN);
      pe[npe].x = prev[i].x;
      pe[npe].y = prev[i].y;
      ++npe;
    }
  qsort(pe, npe, sizeof(ro_entry), ro_entry_cmp);

  size_t vcap = 1024, nv = 0;
  uint64_t* votes = (uint64_t*)malloc(vcap * sizeof(uint64_t));
  for (size_t i = 0; i < nc; ++i) {
    if (!(curr[i].region_mask & bit)) continue;
    if (!out->use_all && curr[i].weight != 2) continue; /* src/kpm.hpp:113-117 */
    /* equal range of curr[i].code in pe */
    size_t lo = 0, hi = npe;
    while (lo < hi) {
      size_t mid = (lo + hi) / 2;
      if (memcmp(pe[mid].code, curr[i].code, RO_CODE_LEN) < 0) lo = mid + 1; else hi = mid;
    }
    for (size_t j = lo; j < npe && memcmp(pe[j].code, curr[i].code, RO_CODE_LEN) == 0; ++j) {
      int32_t dx = (int32_t)pe[j].x - (int32_t)curr[i].x; /* prev - curr (src/kpm.hpp:96-98) */
      int32_t dy = (int32_t)pe[j].y - (int32_t)curr[i].y;
      if (nv == vcap) { vcap *= 2; votes = (uint64_t*)realloc(votes, vcap * sizeof(uint64_t)); }
      votes[nv++] = ((uint64_t)(uint32_t)(dx + 0x40000000) << 32) | (uint32_t)(dy + 0x40000000);
    }
  }
  qsort(votes, nv, sizeof(uint64_t), ro_u64_cmp);
  /* run-length -> bins sorted by (dx, dy); ticket = top region_votes in the defined order */
  uint32_t rv = cfg->region_votes > 3 ? 3 : cfg->region_votes;
  size_t nb = 0;
  ro_bin* all = (ro_bin*)malloc((nv + 1) * sizeof(ro_bin));
  for (size_t i = 0; i < nv;) {
    size_t j = i;
    while (j < nv && votes[j] == votes[i]) ++j;
    all[nb].dx = (int32_t)(uint32_t)(votes[i] >> 32) - 0x40000000;
    all[nb].dy = (int32_t)(uint32_t)(votes[i] & 0xffffffffu) - 0x40000000;
    all[nb].cnt = (uint32_t)(j - i);
    ++nb;
    i = j;
  }
  out->nbins = (uint32_t)nb;
  for (size_t i = 0; i < nb; ++i) out->hist_hash += ro_bin_hash(all[i].dx, all[i].dy, all[i].cnt);
  if (bins_out)
    for (size_t i = 0; i < nb && i < bins_cap; ++i) bins_out[i] = all[i];
  out->nticket = nb < rv ? (uint32_t)nb : rv;
  uint32_t filled = 0;
  for (size_t i = 0; i < nb; ++i) { /* insertion into the top list, defined order */
    uint32_t pos = filled;
    while (pos > 0 && ro_bin_before(&all[i], &out->ticket[pos - 1])) --pos;
    if (pos >= rv) continue;
    uint32_t last = filled < rv ? filled : rv - 1;
    for (uint32_t k = last; k > pos; --k) out->ticket[k] = out->ticket[k - 1];
    out->ticket[pos] = all[i];
    if (filled < rv) ++filled;
  }
  for (uint32_t k = 0; k < out->nticket; ++k) {
    uint32_t c = out->ticket[k].cnt;
    for (size_t i = 0; i < nb; ++i) {
      if (all[i].cnt > c) ++out->ngt[k];
      if (all[i].cnt >= c) ++out->nge[k];
    }
  }
  free(all);
  free(votes);
  free(pe);
}

/* Borda count + declare (src/kpm.hpp:172-184,199-211) with the tie-sensitivity analysis described
 * at the top of this file.  votes: one ro_region_vote per region. */
void ro_declare(const ro_config* cfg, const ro_region_vote* votes, ro_match_result* res) {
  const uint32_t nreg = cfg->grid_w * cfg->grid_h;
  const uint32_t rv = cfg->region_votes > 3 ? 3 : cfg->region_votes;
  memset(res, 0, sizeof(*res));
  uint32_t active = 0;
  for (uint32_t r = 0; r < nreg; ++r)
    if (votes[r].n_curr > 0) ++active; /* get_active looks at the CURRENT grid only (src/kpm.hpp:400) */
  res->active = active;
  if (active < nreg / 4) return; /* src/kpm.hpp:401 */

  ro_bin cand[RO_MAX_REGIONS * 3];
  uint32_t ncand = 0;
  for (uint32_t r = 0; r < nreg; ++r)
    for (uint32_t k = 0; k < votes[r].nticket; ++k) { /* total[off] += rank-- (src/kpm.hpp:176-181) */
      uint32_t pts = rv - k, j;
      for (j = 0; j < ncand; ++j)
        if (cand[j].dx == votes[r].ticket[k].dx && cand[j].dy == votes[r].ticket[k].dy) break;
      if (j == ncand) { cand[ncand].dx = votes[r].ticket[k].dx; cand[ncand].dy = votes[r].ticket[k].dy; cand[ncand].cnt = 0; ++ncand; }
      cand[j].cnt += pts;
    }
  if (ncand == 0) return; /* top.empty() (src/kpm.hpp:202-204) */
  /* top two in the defined order */
  int b0 = -1, b1 = -1;
  for (uint32_t j = 0; j < ncand; ++j) {
    if (b0 < 0 || ro_bin_before(&cand[j], &cand[b0])) { b1 = b0; b0 = (int)j; }
    else if (b1 < 0 || ro_bin_before(&cand[j], &cand[b1])) { b1 = (int)j; }
  }
  res->ntop = b1 >= 0 ? 2 : 1;
  res->top_dx[0] = cand[b0].dx; res->top_dy[0] = cand[b0].dy; res->top_score[0] = cand[b0].cnt;
  if (b1 >= 0) { res->top_dx[1] = cand[b1].dx; res->top_dy[1] = cand[b1].dy; res->top_score[1] = cand[b1].cnt; }
  const uint32_t half = active / 2; /* src/kpm.hpp:206 */
  const uint32_t S0 = cand[b0].cnt, S1 = b1 >= 0 ? cand[b1].cnt : 0;
  if (b1 >= 0 && S0 < S1 + half) res->valid = 0;
  else { res->valid = 1; res->dx = cand[b0].dx; res->dy = cand[b0].dy; }

  /* Tie sensitivity.  Under ANY order of count-tied bins, an offset o scores between
   *   lo(o) = sum_r minpts_r(o)   and   hi(o) = sum_r maxpts_r(o)
   * where, in region r: o on the ticket at k -> best position ngt[k], worst position nge[k]-1;
   * o off the ticket -> it can only enter if the ticket is full and its last entry's count is shared
   * with bins outside the ticket (nge[rv-1] > rv), and then at best at position ngt[rv-1].
   * An offset on no ticket at all scores at most hi_out.  Declared w is tie-proof iff
   * lo(w) >= max_{o != w} hi(o) + max(half, 1); "none" is tie-proof iff the two largest lo() keep two
   * candidates alive (second >= 1) and max hi() < second lo + half. */
  {
    uint32_t lo[RO_MAX_REGIONS * 3], hi[RO_MAX_REGIONS * 3], hi_out = 0;
    int ambiguous = 0;
    for (uint32_t j = 0; j < ncand; ++j) lo[j] = hi[j] = 0;
    for (uint32_t r = 0; r < nreg; ++r) {
      const ro_region_vote* v = &votes[r];
      uint32_t og = 0;
      if (v->nticket == rv && v->nge[rv - 1] > rv) og = v->ngt[rv - 1] < rv ? rv - v->ngt[rv - 1] : 0;
      hi_out += og;
      for (uint32_t j = 0; j < ncand; ++j) {
        uint32_t k;
        for (k = 0; k < v->nticket; ++k)
          if (v->ticket[k].dx == cand[j].dx && v->ticket[k].dy == cand[j].dy) break;
        if (k < v->nticket) {
          uint32_t worst = v->nge[k] - 1;
          hi[j] += v->ngt[k] < rv ? rv - v->ngt[k] : 0;
          lo[j] += worst < rv ? rv - worst : 0;
          if (v->nge[k] != v->ngt[k] + 1) ambiguous = 1; /* shares its count with another bin */
        } else {
          hi[j] += og;
        }
      }
    }
    if (!ambiguous) res->tie_sensitive = 0;
    else if (res->valid) {
      uint32_t H = hi_out;
      for (uint32_t j = 0; j < ncand; ++j) if ((int)j != b0 && hi[j] > H) H = hi[j];
      res->tie_sensitive = !(lo[b0] >= H + (half > 1 ? half : 1));
    } else {
      uint32_t H = hi_out, l1 = 0, l2 = 0;
      for (uint32_t j = 0; j < ncand; ++j) {
        if (hi[j] > H) H = hi[j];
        if (lo[j] > l1) { l2 = l1; l1 = lo[j]; } else if (lo[j] > l2) l2 = lo[j];
      }
      res->tie_sensitive = !(l2 >= 1 && H < l2 + half);
    }
  }
}

/* kpm::match for one consecutive pair.  region_votes_out: nreg entries (may be NULL). */
void ro_match(const ro_config* cfg, const ro_keypoint* prev, size_t np, const ro_keypoint* curr, size_t nc,
              ro_match_result* res, ro_region_vote* region_votes_out) {
  const uint32_t nreg = cfg->grid_w * cfg->grid_h;
  ro_region_vote votes[RO_MAX_REGIONS];
  for (uint32_t r = 0; r < nreg; ++r) ro_region(cfg, prev, np, curr, nc, r, &votes[r], NULL, 0);
  ro_declare(cfg, votes, res);
  if (region_votes_out) memcpy(region_votes_out, votes, nreg * sizeof(ro_region_vote));
}

/* Parity tap: the full offset histogram of one region, sorted by (dx, dy). */
size_t ro_region_bins(const ro_config* cfg, const ro_keypoint* prev, size_t np, const ro_keypoint* curr, size_t nc,
                      uint32_t region, ro_bin* bins, size_t cap) {
  ro_region_vote v;
  ro_region(cfg, prev, np, curr, nc, region, &v, bins, cap);
  return v.nbins;
}

/* ---- a14: the frc::collector loop (src/frc.hpp:55-68,83-127) ------------------------------- */
/* frames: n*H*W.  results: n-1 entries (pair i-1 -> i at index i-1).  positions: n*(fragment,x,y)
 * int32 -- position_ += off, or a new fragment at (0,0) when kpm::match returned nothing
 * (src/frc.hpp:109-115,124-127).  medians: n*H*W or NULL.  Returns total keypoints. */
size_t ro_register(const ro_config* cfg, const uint8_t* frames, size_t n, ro_match_result* results,
                   int32_t* positions, uint8_t* medians, uint32_t* kp_counts) {
  const size_t px = (size_t)cfg->width * cfg->height;
  size_t cap = px, total = 0;
  ro_keypoint* a = (ro_keypoint*)malloc(cap * sizeof(ro_keypoint));
  ro_keypoint* b = (ro_keypoint*)malloc(cap * sizeof(ro_keypoint));
  size_t na = 0, nb = 0;
  int32_t frag = 0, x = 0, y = 0;
  for (size_t i = 0; i < n; ++i) {
    nb = ro_extract(cfg, frames + i * px, medians ? medians + i * px : NULL, b, cap);
    total += nb;
    if (kp_counts) kp_counts[i] = (uint32_t)nb;
    if (i > 0) {
      ro_match_result res;
      ro_match(cfg, a, na, b, nb, &res, NULL);
      if (results) results[i - 1] = res;
      if (res.valid) { x += res.dx; y += res.dy; }
      else { ++frag; x = 0; y = 0; }
    }
    if (positions) { positions[3 * i] = frag; positions[3 * i + 1] = x; positions[3 * i + 2] = y; }
    ro_keypoint* t = a; a = b; b = t;
    na = nb;
  }
  free(a);
  free(b);
  return total;
}

/* ---- a15: fde::details::generate_mask (src/fde.hpp:19-55; idx from src/fde.hpp:87) --------- */
void ro_foreground_mask(const uint8_t* bg, uint32_t bgW, uint32_t bgH, int32_t px, int32_t py,
                        const uint8_t* frame, uint32_t W, uint32_t H, uint8_t* mask) {
  (void)bgH;
  const ptrdiff_t idx = (ptrdiff_t)bgW * py + px; /* cdt::to_index (src/cdt.hpp:173-177) */
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t x = 0; x < W; ++x)
      mask[(size_t)y * W + x] = bg[idx + (ptrdiff_t)y * bgW + x] == frame[(size_t)y * W + x] ? 0xFF : 0x00;
}

/* ---- f2: pass-2 foreground extraction: fde::extractor::extract + fde::mask (src/fde.hpp:83-146) over
 * cte::extractor (src/cte.hpp:60-166) and ctr::contour (src/ctr.hpp:121-235) -------------------------
 *
 * What the reference does per frame (src/fdf.hpp:58-66):
 *   1. generate_mask: eq[p] = background window == frame (src/fde.hpp:87,106-114);
 *   2. cte::extractor::extract(median, pred = eq[p] == 0): scan the INTERIOR pixels (rows 1..H-2, columns
 *      1..W-2, src/cte.hpp:66-72,92) in row-major order; a pixel that has no contour id yet and satisfies
 *      the predicate starts a breadth-first fill over 4-connected pixels of EQUAL MEDIAN colour
 *      (push_pixel, src/cte.hpp:143-157).  Row 0, columns 0 and W-1 and the LAST TWO rows (H-2 and H-1) are
 *      pre-marked with the horizon id (clear_outline, src/cte.hpp:159-177: the side-column loop stops at row
 *      H-3 and the closing loop marks everything from row H-2 on) and are never entered or seeded;
 *   3. contours whose area (pixel count, ctr::contour::add_point src/ctr.hpp:139-149) exceeds
 *      frame area / 5 are dropped (src/fde.hpp:94-100);
 *   4. fde::mask paints every kept contour's pixels (contour::recover fills each horizontal run between
 *      its left-edge and right-edge pixel, src/ctr.hpp:151-170) and then every kept contour's enclosure
 *      with EXCLUSIVE right and bottom bounds (src/fde.hpp:133-143; enclosure = min/max column of the
 *      edge pixels, first/last edge row, src/ctr.hpp:96-107).
 * The fill order inside a contour does not influence any of this, so the restatement uses a stack.
 * Contours come out in the reference's order (first seed in row-major order).  The reference's contour
 * ids are uint16 (src/cte.hpp:20,96-98) and wrap after 65,534 contours in one frame; that case is outside
 * the contract (returns (size_t)-1). */
typedef struct { uint32_t area, left, top, right, bottom, colour; } ro_contour;

static int ro_u32_cmp(const void* a, const void* b) {
  const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return x < y ? -1 : x > y;
}

size_t ro_foreground(const uint8_t* bg, uint32_t bgW, uint32_t bgH, int32_t px, int32_t py, const uint8_t* frame,
                     const uint8_t* median, uint32_t W, uint32_t H, uint8_t* out_mask, ro_contour* contours,
                     size_t cap) {
  (void)bgH;
  const size_t npx = (size_t)W * H;
  const ptrdiff_t idx = (ptrdiff_t)bgW * py + px;
  uint32_t* id = (uint32_t*)calloc(npx, sizeof(uint32_t));
  uint32_t* stack = (uint32_t*)malloc(npx * sizeof(uint32_t));
  uint32_t* members = (uint32_t*)malloc(npx * sizeof(uint32_t)); /* pixels of the contour being filled */
  const uint32_t area_limit = (uint32_t)(npx / 5);                /* src/fde.hpp:94 */
  size_t ncont = 0, nkept = 0;
  memset(out_mask, 0, npx);
  if (W < 3 || H < 4) { free(id); free(stack); free(members); return 0; }
  /* kept contours are remembered so that the enclosure pass runs after ALL pixel passes (src/fde.hpp:129-143) */
  ro_contour* kept = (ro_contour*)malloc((npx + 1) * sizeof(ro_contour));
  size_t kept_cap = npx + 1;
  for (uint32_t y = 1; y + 2 < H; ++y) { /* row H-2 is scanned by the reference but is all horizon */
    for (uint32_t x = 1; x + 1 < W; ++x) {
      const size_t p = (size_t)y * W + x;
      if (id[p] != 0) continue;
      if (bg[idx + (ptrdiff_t)y * bgW + x] == frame[p]) continue; /* pred: mask value == 0 <=> differs */
      if (++ncont >= 65535) { free(id); free(stack); free(members); free(kept); return (size_t)-1; }
      const uint8_t colour = median[p];
      ro_contour c = {0, 0, 0, 0, 0, colour};
      size_t sp = 0, nm = 0;
      stack[sp++] = (uint32_t)p;
      id[p] = (uint32_t)ncont;
      while (sp) {
        const uint32_t q = stack[--sp];
        members[nm++] = q;
        ++c.area;
        const int32_t nb[4] = {-1, 1, -(int32_t)W, (int32_t)W};
        for (int k = 0; k < 4; ++k) {
          const uint32_t r = (uint32_t)((int32_t)q + nb[k]);
          const uint32_t rx = r % W, ry = r / W;
          if (rx == 0 || ry == 0 || rx == W - 1 || ry >= H - 2) continue; /* horizon */
          if (median[r] != colour || id[r] != 0) continue;
          id[r] = (uint32_t)ncont;
          stack[sp++] = r;
        }
      }
      /* enclosure (src/ctr.hpp:96-107, cdt::limits::update src/cdt.hpp:183-190): over the pixels that carry
       * a left or right edge (a horizontal neighbour of another colour, or the horizon), in ascending
       * position order; note the `else if` -- a value that raises the upper bound never lowers the lower
       * one, and the lower bound starts at SIZE_MAX (here: UINT32_MAX, which also makes the fill empty). */
      qsort(members, nm, sizeof(uint32_t), ro_u32_cmp);
      {
        uint32_t lower = UINT32_MAX, upper = 0, first = UINT32_MAX, last = 0;
        for (size_t k = 0; k < nm; ++k) {
          const uint32_t q = members[k], qx = q % W;
          const int le = qx == 1 || median[q - 1] != colour, re = qx == W - 2 || median[q + 1] != colour;
          if (!le && !re) continue;
          if (first == UINT32_MAX) first = q;
          last = q;
          if (qx > upper) upper = qx; else if (qx < lower) lower = qx;
        }
        c.left = lower; c.right = upper; c.top = first / W; c.bottom = last / W;
      }
      if (c.area > area_limit) continue;
      for (size_t k = 0; k < nm; ++k) out_mask[members[k]] = 1;
      if (nkept < kept_cap) kept[nkept] = c;
      if (contours && nkept < cap) contours[nkept] = c;
      ++nkept;
    }
  }
  for (size_t k = 0; k < nkept && k < kept_cap; ++k)
    for (uint32_t y = kept[k].top; y < kept[k].bottom; ++y)
      for (uint32_t x = kept[k].left; x < kept[k].right; ++x) out_mask[(size_t)y * W + x] = 1;
  free(id); free(stack); free(members); free(kept);
  return nkept;
}
size_t ro_sizeof_contour(void) { return sizeof(ro_contour); }

/* ---- f3: the cellular kpm::match used by fragment splicing (src/kpm.hpp:225-393, called from
 * fgs::details::match_partial, src/fgs.hpp:119-134, with cell_size {15, 15}) --------------------------
 *
 *   count_offsets (:249-262) / get_offsets (:231-247): for every code present in both regions, ALL pairs
 *     (prev point, curr point) vote for offset prev - curr; the vote is filed under the cell
 *     (min(px, cx) / cw, min(py, cy) / ch) (to_cell, :225-229);
 *   find_best (:280-299): per offset matched_cells = distinct cells, matched_keypoints = votes; the offset with
 *     the most votes wins -- std::max_element over an unordered_map, i.e. the FIRST maximum in an
 *     implementation-defined order.  Here ties go to the smallest (dy, dx) and `ties` reports how many offsets
 *     share the maximum (> 1: the reference's choice is not defined by the language);
 *   count_active_cells (:349-369) -> filter_keypoints (:321-347): curr points inside clim (get_limits, :301-315,
 *     in size_t arithmetic) whose position + offset hits a set pixel of prev's mask; distinct cells relative
 *     to clim's corner;
 *   match (:371-393): no common code -> nullopt; matched_cells < active * 0.66f -> nullopt; else the vote. */
typedef struct {
  uint32_t valid;
  int32_t dx, dy;
  uint32_t matched_keypoints, matched_cells, active_cells, offsets, ties;
  uint64_t pairs;
} ro_cell_match_result;

typedef struct { int32_t oy, ox; uint32_t cy, cx; } ro_cell_vote;

static int ro_cell_vote_cmp(const void* a, const void* b) {
  const ro_cell_vote* x = (const ro_cell_vote*)a;
  const ro_cell_vote* y = (const ro_cell_vote*)b;
  if (x->oy != y->oy) return x->oy < y->oy ? -1 : 1;
  if (x->ox != y->ox) return x->ox < y->ox ? -1 : 1;
  if (x->cy != y->cy) return x->cy < y->cy ? -1 : 1;
  if (x->cx != y->cx) return x->cx < y->cx ? -1 : 1;
  return 0;
}
static int ro_kp_code_cmp(const void* a, const void* b) {
  return memcmp(((const ro_keypoint*)a)->code, ((const ro_keypoint*)b)->code, RO_CODE_LEN);
}
static void ro_cell_limits(int32_t delta, uint64_t previous, uint64_t current, uint64_t* lo, uint64_t* hi) {
  if (delta < 0) { /* src/kpm.hpp:304-309: the `second` (curr) span */
    const uint64_t d = (uint64_t)(-(int64_t)delta);
    *lo = d;
    *hi = current < previous + d ? current : previous + d;
  } else {         /* src/kpm.hpp:312-313 */
    *lo = 0;
    *hi = current < previous - (uint64_t)delta ? current : previous - (uint64_t)delta;
  }
}

void ro_cell_match(const ro_keypoint* prev, size_t np, const uint8_t* pmask, uint32_t pW, uint32_t pH,
                   const ro_keypoint* curr, size_t nc, uint32_t cW, uint32_t cH, uint32_t cell_w, uint32_t cell_h,
                   ro_cell_match_result* res) {
  memset(res, 0, sizeof(*res));
  ro_keypoint* ps = (ro_keypoint*)malloc((np + 1) * sizeof(ro_keypoint));
  memcpy(ps, prev, np * sizeof(ro_keypoint));
  qsort(ps, np, sizeof(ro_keypoint), ro_kp_code_cmp);
  size_t cap = 1024, nv = 0;
  ro_cell_vote* votes = (ro_cell_vote*)malloc(cap * sizeof(ro_cell_vote));
  for (size_t j = 0; j < nc; ++j) {
    size_t lo = 0, hi = np; /* first prev entry with code >= curr's */
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (memcmp(ps[mid].code, curr[j].code, RO_CODE_LEN) < 0) lo = mid + 1; else hi = mid;
    }
    for (size_t i = lo; i < np && memcmp(ps[i].code, curr[j].code, RO_CODE_LEN) == 0; ++i) {
      if (nv == cap) { cap *= 2; votes = (ro_cell_vote*)realloc(votes, cap * sizeof(ro_cell_vote)); }
      const int32_t px = ps[i].x, py = ps[i].y, cx = curr[j].x, cy = curr[j].y;
      votes[nv].ox = px - cx; votes[nv].oy = py - cy;
      votes[nv].cx = (uint32_t)((px < cx ? px : cx) / (int32_t)cell_w);
      votes[nv].cy = (uint32_t)((py < cy ? py : cy) / (int32_t)cell_h);
      ++nv;
    }
  }
  res->pairs = nv;
  if (nv == 0) { free(ps); free(votes); return; }
  qsort(votes, nv, sizeof(ro_cell_vote), ro_cell_vote_cmp);
  uint32_t best_kp = 0, best_cells = 0;
  for (size_t a = 0; a < nv;) {
    size_t b = a;
    uint32_t cells = 0;
    while (b < nv && votes[b].oy == votes[a].oy && votes[b].ox == votes[a].ox) {
      if (b == a || votes[b].cy != votes[b - 1].cy || votes[b].cx != votes[b - 1].cx) ++cells;
      ++b;
    }
    const uint32_t kp = (uint32_t)(b - a);
    ++res->offsets;
    if (kp > best_kp) { best_kp = kp; best_cells = cells; res->dx = votes[a].ox; res->dy = votes[a].oy; res->ties = 1; }
    else if (kp == best_kp) ++res->ties; /* sorted by (oy, ox): the first one met stays */
    a = b;
  }
  res->matched_keypoints = best_kp;
  res->matched_cells = best_cells;
  uint64_t l, r, t, bt;
  ro_cell_limits(res->dx, pW, cW, &l, &r);
  ro_cell_limits(res->dy, pH, cH, &t, &bt);
  const uint32_t AW = cW / cell_w + 1, AH = cH / cell_h + 1;
  uint8_t* act = (uint8_t*)calloc((size_t)AW * AH, 1);
  for (size_t j = 0; j < nc; ++j) {
    const uint64_t x = curr[j].x, y = curr[j].y;
    if (!(x >= l && x < r && y >= t && y < bt)) continue;
    const int64_t idx = (int64_t)pW * ((int64_t)y + res->dy) + ((int64_t)x + res->dx);
    if (idx < 0 || idx >= (int64_t)pW * pH || pmask[idx] == 0) continue;
    act[((y - t) / cell_h) * AW + (x - l) / cell_w] = 1;
  }
  for (size_t k = 0; k < (size_t)AW * AH; ++k) res->active_cells += act[k];
  res->valid = !((float)res->matched_cells < (float)res->active_cells * 0.66f);
  free(act); free(ps); free(votes);
}
size_t ro_sizeof_cell_match(void) { return sizeof(ro_cell_match_result); }

size_t ro_sizeof_keypoint(void) { return sizeof(ro_keypoint); }
size_t ro_sizeof_region_vote(void) { return sizeof(ro_region_vote); }
size_t ro_sizeof_match_result(void) { return sizeof(ro_match_result); }
