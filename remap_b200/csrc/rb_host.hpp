// rb_host.hpp -- host-side geometry shared by the C-ABI (rb_api.cu) and the test harness.
#pragma once

#include <string.h>

#include "rb_common.cuh"

static inline uint32_t rb_align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// Region sections exactly as the reference lays them out (src/kpe.hpp:84-90,157-192,235-277;
// SURVEY.md Appendix A.4).  Returns 0 on success, <0 if the frame is too small for the grid.
static inline int rb_make_geom(uint32_t W, uint32_t H, uint32_t grid_w, uint32_t grid_h, uint32_t overlap,
                               uint32_t weight_switch, uint32_t region_votes, RbGeom* g) {
  memset(g, 0, sizeof(*g));
  if (grid_w == 0 || grid_h == 0 || grid_w > 8 || grid_h > 8 || grid_w * grid_h > RB_MAX_REGIONS) return -1;
  if (region_votes < 1 || region_votes > 3) return -1;
  if (W / grid_w <= overlap / 2 || H / grid_h <= overlap / 2) return -1;
  if (W < 8 || H < 8 || W > 65535 || H > 65535) return -1;
  const uint32_t rw = W / grid_w - overlap / 2, rh = H / grid_h - overlap / 2;  // src/kpe.hpp:86-87
  g->W = W;
  g->H = H;
  g->pitch = rb_align_up(W, 16);
  g->frame_stride = (uint64_t)g->pitch * H;
  g->NS = (W - 4 + RB_STRIP_OUT - 1) / RB_STRIP_OUT;
  g->mpitch = rb_align_up(RB_STRIP_OUT * g->NS + 4, 16);
  g->median_stride = (uint64_t)g->mpitch * H;
  g->grid_w = grid_w;
  g->grid_h = grid_h;
  g->nreg = grid_w * grid_h;
  g->weight_switch = weight_switch;
  g->region_votes = region_votes;
  // columns: section s owns [c, c+rw) alone and shares [c+rw, c+rw+O) with s+1; the last runs to W-2
  uint32_t c = 2;
  for (uint32_t s = 0; s < grid_w; ++s) {
    g->col0[s] = s == 0 ? 2 : c - overlap;                 // start of the shared band with s-1
    g->col1[s] = s + 1 < grid_w ? c + rw + overlap : W - 2;
    c += rw + overlap;
  }
  // rows: y == 2 belongs to section 0; then from y == 3 the same pattern; the last runs to H-4
  uint32_t r = 3;
  for (uint32_t s = 0; s < grid_h; ++s) {
    g->row0[s] = s == 0 ? 2 : r - overlap;
    g->row1[s] = s + 1 < grid_h ? r + rh + overlap : H - 4;
    r += rh + overlap;
  }
  // Small frames: a shared band may run past the keypoint domain (the reference then walks columns
  // that hold no data); clip to the domain like oracle/remap_oracle.c does.
  for (uint32_t s = 0; s < grid_w; ++s) {
    if (g->col1[s] > W - 2) g->col1[s] = W - 2;
    if (g->col0[s] >= g->col1[s]) return -1;
  }
  for (uint32_t s = 0; s < grid_h; ++s) {
    if (g->row1[s] > H - 4) g->row1[s] = H - 4;
    if (g->row0[s] >= g->row1[s]) return -1;
  }
  return 0;
}
