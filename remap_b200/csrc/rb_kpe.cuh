// rb_kpe.cuh -- K1: keypoint extraction (replaces kpe::extractor::extract, src/kpe.hpp:92-108).
//
// The reference slides 16-bin byte histograms in AVX2 registers and scans them per pixel
// (src/kpe.hpp:111-147,207-340).  Here the two rank filters are computed BIT-SLICED: one 32-bit
// word holds one bit of 32 horizontally adjacent pixels, so a LOP3 works on 32 pixels at once.
//
//   p3 = 4th largest of the 3x3 window, p5 = 12th largest of the 5x5 window, in luminance-ordered
//   space (src/kpe.hpp:313,317,326-340).  For a threshold t in 1..15 let T_t = [ordered >= t] (a
//   bit plane).  The number of window pixels >= t is a box sum of T_t, and
//       p3 = #{t : boxsum3x3(T_t) >= 4},   p5 = #{t : boxsum5x5(T_t) >= 12}
//   because the box sums are monotone in t.  Box sums of bit planes are carry-save adder trees:
//   vertical 3/5-row sums first (shared between the two filters), then the horizontal sums via
//   in-register shifts, reduced directly to the ">= k" predicate.
//
// Work decomposition: one thread owns one STRIP = a 32-pixel-wide window [28j, 28j+32) of one frame
// that yields the 28 output columns [28j+2, 28j+30) (2-pixel halo on both sides, so horizontal
// neighbours are plain shifts and threads never talk to each other), and marches down a segment of
// rows keeping the threshold planes of the last four rows in registers.  Grid = frames x segments
// x strips; nothing is shared, so the grid is sized freely in multiples of the SM count.
//
// Outputs per frame (fixed size, no atomics, deterministic):
//   kpbits[y][j]  bit i set <=> pixel (28j + i, y) is a keypoint        (src/kpe.hpp:316-320)
//   w2bits[y][j]  bit i set <=> that keypoint has weight 2 (p1 != p5)   (src/kpe.hpp:319)
//   median[y][x]  = ordered_to_native(p3), stored at byte x + 2 of the row (src/kpe.hpp:314)
// The 13-byte code (src/kpe.hpp:342-379) is a pure function of the 5x5 patch, so it is never
// materialised in HBM: K2 compares patches directly, and the rb_keypoints tap formats codes on
// demand.
#pragma once

#include "rb_common.cuh"

struct RbKpeParams {
  RbGeom g;
  const uint8_t* frames;
  uint8_t* median;    // nullptr: skip
  uint32_t* kpbits;   // [nframes][H][NS]
  uint32_t* w2bits;   // [nframes][H][NS]
  uint32_t nframes;
  uint32_t nseg;      // row segments per strip
  uint32_t seg_rows;  // output rows per segment
};

namespace rbk {

constexpr int N2O[16] = {RB_N2O_LIST};
constexpr int O2N[16] = {RB_O2N_LIST};

// 8 words of 4 byte-pixels (values 0..15; high nibbles ignored) -> 4 bit planes, bit i of plane k =
// bit k of pixel i.
RB_HD void bytes_to_planes(const uint32_t w[8], uint32_t pl[4]) {
  uint32_t c[4];
#pragma unroll
  for (int m = 0; m < 4; ++m)  // byte b of c[m] = pixel(8m+b) | pixel(8m+4+b) << 4
    c[m] = (w[2 * m] & 0x0F0F0F0Fu) | ((w[2 * m + 1] << 4) & 0xF0F0F0F0u);
  // 4x4 byte transpose: d[b] byte m = c[m] byte b, i.e. nibble i of d[b] = pixel 4i + b
  const uint32_t t0 = rb_prmt(c[0], c[1], 0x5140), t1 = rb_prmt(c[0], c[1], 0x7362);
  const uint32_t t2 = rb_prmt(c[2], c[3], 0x5140), t3 = rb_prmt(c[2], c[3], 0x7362);
  uint32_t d0 = rb_prmt(t0, t2, 0x5410), d1 = rb_prmt(t0, t2, 0x7632);
  uint32_t d2 = rb_prmt(t1, t3, 0x5410), d3 = rb_prmt(t1, t3, 0x7632);
  // bit address (word b = p1 p0 | nibble p4 p3 p2 | bit k1 k0) -> (word k1 k0 | p4..p0): two rounds of
  // 2x2 block swaps across words
  uint32_t t;
  t = ((d0 >> 1) ^ d1) & 0x55555555u; d1 ^= t; d0 ^= t << 1;
  t = ((d2 >> 1) ^ d3) & 0x55555555u; d3 ^= t; d2 ^= t << 1;
  t = ((d0 >> 2) ^ d2) & 0x33333333u; d2 ^= t; d0 ^= t << 2;
  t = ((d1 >> 2) ^ d3) & 0x33333333u; d3 ^= t; d1 ^= t << 2;
  pl[0] = d0; pl[1] = d1; pl[2] = d2; pl[3] = d3;
}

// inverse of bytes_to_planes
RB_HD void planes_to_bytes(const uint32_t pl[4], uint32_t w[8]) {
  uint32_t d0 = pl[0], d1 = pl[1], d2 = pl[2], d3 = pl[3], t;
  t = ((d0 >> 2) ^ d2) & 0x33333333u; d2 ^= t; d0 ^= t << 2;
  t = ((d1 >> 2) ^ d3) & 0x33333333u; d3 ^= t; d1 ^= t << 2;
  t = ((d0 >> 1) ^ d1) & 0x55555555u; d1 ^= t; d0 ^= t << 1;
  t = ((d2 >> 1) ^ d3) & 0x55555555u; d3 ^= t; d2 ^= t << 1;
  const uint32_t t0 = rb_prmt(d0, d1, 0x5140), t1 = rb_prmt(d0, d1, 0x7362);
  const uint32_t t2 = rb_prmt(d2, d3, 0x5140), t3 = rb_prmt(d2, d3, 0x7362);
  uint32_t c[4];
  c[0] = rb_prmt(t0, t2, 0x5410); c[1] = rb_prmt(t0, t2, 0x7632);
  c[2] = rb_prmt(t1, t3, 0x5410); c[3] = rb_prmt(t1, t3, 0x7632);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    w[2 * m] = c[m] & 0x0F0F0F0Fu;
    w[2 * m + 1] = (c[m] >> 4) & 0x0F0F0F0Fu;
  }
}

template <int BIT>
RB_HD uint32_t n2o_plane(const uint32_t n[4]) { return rb_bool4<rb_lut_tt(N2O, BIT)>(n[0], n[1], n[2], n[3]); }
template <int BIT>
RB_HD uint32_t o2n_plane(const uint32_t o[4]) { return rb_bool4<rb_lut_tt(O2N, BIT)>(o[0], o[1], o[2], o[3]); }

// T[t-1] = [ordered >= t], t = 1..15, from the four ordered bit planes.
RB_HD void thresholds(const uint32_t o[4], uint32_t T[15]) {
  const uint32_t o0 = o[0], o1 = o[1], o2 = o[2], o3 = o[3];
  T[7] = o3;                    // >= 8
  T[11] = o3 & o2;              // >= 12
  T[3] = o3 | o2;               // >= 4
  T[13] = o3 & o2 & o1;         // >= 14
  T[9] = o3 & (o2 | o1);        // >= 10
  T[5] = o3 | (o2 & o1);        // >= 6
  T[1] = o3 | o2 | o1;          // >= 2
  T[14] = T[13] & o0;           // >= 15
  T[12] = T[13] | (T[11] & o0); // >= 13
  T[10] = T[11] | (T[9] & o0);  // >= 11
  T[8] = T[9] | (T[7] & o0);    // >= 9
  T[6] = T[7] | (T[5] & o0);    // >= 7
  T[4] = T[5] | (T[3] & o0);    // >= 5
  T[2] = T[3] | (T[1] & o0);    // >= 3
  T[0] = T[1] | o0;             // >= 1
}

// One threshold plane, five consecutive rows t4 (y-2) .. t0 (y+2) -> q3 = [3x3 count >= 4],
// q5 = [5x5 count >= 12], both SHIFTED LEFT BY TWO: bit i of q3 / q5 belongs to pixel i - 2 of the
// word (exact for bits 4..31).  All horizontal neighbours are taken with LEFT shifts only: ptxas
// turns those into IMAD.SHL on the FMA pipe, which idles here, while right shifts (SHF.R) would
// compete with the LOP3s for the ALU pipe that bounds this kernel (profiles/README.md, r1b).
RB_HD void rank_planes(uint32_t t4, uint32_t t3, uint32_t t2, uint32_t t1, uint32_t t0, uint32_t& q3, uint32_t& q5) {
  // vertical sums: V3 = t3+t2+t1 = a1 a0 ; V5 = V3 + t4 + t0 = b2 b1 b0
  const uint32_t a0 = rb_xor3(t3, t2, t1), a1 = rb_maj(t3, t2, t1);
  const uint32_t b0 = rb_xor3(a0, t4, t0), c = rb_maj(a0, t4, t0);
  const uint32_t b1 = a1 ^ c, b2 = a1 & c;
  // 3x3: count = u + 2 v, u = a0(x-1)+a0(x)+a0(x+1), v likewise on a1;  count >= 4 <=> v>=2 | (v>=1 & u>=2).
  // (a, a << 1, a << 2) centres the window on bit i - 1; one more shift brings it to i - 2.
  {
    const uint32_t m0 = a0 << 1, n0 = a0 << 2, m1 = a1 << 1, n1 = a1 << 2;
    q3 = (rb_maj(a1, m1, n1) | ((a1 | m1 | n1) & rb_maj(a0, m0, n0))) << 1;
  }
  // 5x5: S = sum over columns i-4 .. i of the column count b2 b1 b0.  Only S >= 12 is wanted, i.e.
  // (S >> 2) >= 3, so each bit position is reduced with full adders just far enough to hand its
  // carries up; sum bits that cannot reach bit 2 are never formed.
  {
    // weight 1: five b0 bits -> two carries of weight 2
    uint32_t e1 = b0 << 1, e2 = b0 << 2, e3 = b0 << 3, e4 = b0 << 4;
    const uint32_t s = rb_xor3(b0, e1, e2), k = rb_maj(b0, e1, e2), k2 = rb_maj(s, e3, e4);
    // weight 2: five b1 bits + k + k2 -> three carries of weight 4
    e1 = b1 << 1; e2 = b1 << 2; e3 = b1 << 3; e4 = b1 << 4;
    const uint32_t t1 = rb_xor3(b1, e1, e2), c1 = rb_maj(b1, e1, e2);
    const uint32_t t2 = rb_xor3(t1, e3, e4), c2 = rb_maj(t1, e3, e4);
    const uint32_t c3 = rb_maj(t2, k, k2);
    // weight 4: five b2 bits + c1 + c2 + c3 = N ones; S >= 12 <=> N >= 3
    e1 = b2 << 1; e2 = b2 << 2; e3 = b2 << 3; e4 = b2 << 4;
    const uint32_t g1 = rb_xor3(b2, e1, e2), d1 = rb_maj(b2, e1, e2);
    const uint32_t g2 = rb_xor3(e3, e4, c1), d2 = rb_maj(e3, e4, c1);
    const uint32_t g3 = rb_xor3(c2, c3, g1), d3 = rb_maj(c2, c3, g1);
    // N = g2 + g3 + 2 (d1 + d2 + d3) >= 3  <=>  two of the d's, or one d and one g
    q5 = rb_maj(d1, d2, d3) | ((d1 | d2 | d3) & (g2 | g3));
  }
}

// thermometer q[t-1] (t = 1..15, monotone) -> 4 binary planes of p = #{t : q_t}
RB_HD void thermo_to_binary(const uint32_t q[15], uint32_t p[4]) {
  p[3] = q[7];
  p[2] = q[11] | (q[3] & ~q[7]);
  p[1] = q[13] | (q[9] & ~q[11]) | (q[5] & ~q[7]) | (q[1] & ~q[3]);
  p[0] = q[14] | (q[12] & ~q[13]) | (q[10] & ~q[11]) | (q[8] & ~q[9]) | (q[6] & ~q[7]) | (q[4] & ~q[5]) |
         (q[2] & ~q[3]) | (q[0] & ~q[1]);
}

RB_HD uint32_t ld32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

struct StripState {
  uint32_t T[4][15];  // threshold planes of rows y-2 .. y+1 (slot 0 = oldest)
  uint32_t O[4][4];   // ordered planes of the same rows
};

RB_HD void load_row(const uint8_t* rowptr, uint32_t w[8]) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(rowptr);
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = ld32(p + k);
}

RB_HD void row_to_planes(const uint32_t w[8], uint32_t O[4], uint32_t T[15]) {
  uint32_t n[4];
  bytes_to_planes(w, n);
  O[0] = n2o_plane<0>(n); O[1] = n2o_plane<1>(n); O[2] = n2o_plane<2>(n); O[3] = n2o_plane<3>(n);
  thresholds(O, T);
}

// One output row y: st holds rows y-2 .. y+1, wnew is row y+2.  Afterwards st holds y-1 .. y+2.
// The body is deliberately NOT unrolled over rows: one step is ~1000 SASS instructions (16 KB) and
// must stay resident in the instruction cache (an earlier 4x-unrolled version was instruction-fetch
// bound, profiles/README.md); the register rotation at the end is plain MOVs.
RB_HD void strip_step(StripState& st, const uint32_t wnew[8], uint32_t vmask, uint32_t* kp_out, uint32_t* w2_out,
                      uint8_t* med_out) {
  uint32_t On[4], Tn[15];
  row_to_planes(wnew, On, Tn);
  uint32_t q3[15], q5[15];
#pragma unroll
  for (int t = 0; t < 15; ++t) rank_planes(st.T[0][t], st.T[1][t], st.T[2][t], st.T[3][t], Tn[t], q3[t], q5[t]);
  // q3 / q5 and everything derived from them live two bits to the left of the input planes
  uint32_t p3[4], p5[4];
  thermo_to_binary(q3, p3);
  thermo_to_binary(q5, p5);
  const uint32_t p1[4] = {st.O[2][0] << 2, st.O[2][1] << 2, st.O[2][2] << 2, st.O[2][3] << 2};  // centre row
  const uint32_t ne13 = (p1[0] ^ p3[0]) | (p1[1] ^ p3[1]) | (p1[2] ^ p3[2]) | (p1[3] ^ p3[3]);
  const uint32_t ne35 = (p3[0] ^ p5[0]) | (p3[1] ^ p5[1]) | (p3[2] ^ p5[2]) | (p3[3] ^ p5[3]);
  const uint32_t ne15 = (p1[0] ^ p5[0]) | (p1[1] ^ p5[1]) | (p1[2] ^ p5[2]) | (p1[3] ^ p5[3]);
  const uint32_t kp = (ne13 & ne35) >> 2 & vmask;  // src/kpe.hpp:316-318
  *kp_out = kp;
  *w2_out = kp & (ne15 >> 2);               // src/kpe.hpp:319
  if (med_out) {
    uint32_t m[4], w[8];
    m[0] = (o2n_plane<0>(p3) >> 2) & vmask; m[1] = (o2n_plane<1>(p3) >> 2) & vmask;
    m[2] = (o2n_plane<2>(p3) >> 2) & vmask; m[3] = (o2n_plane<3>(p3) >> 2) & vmask;
    planes_to_bytes(m, w);
    uint32_t* dst = reinterpret_cast<uint32_t*>(med_out);  // pixel 28j+2 -> byte 28j+4 of the row
#pragma unroll
    for (int k = 0; k < 7; ++k) dst[k] = rb_prmt(w[k], w[k + 1], 0x5432);
  }
#pragma unroll
  for (int t = 0; t < 15; ++t) { st.T[0][t] = st.T[1][t]; st.T[1][t] = st.T[2][t]; st.T[2][t] = st.T[3][t]; st.T[3][t] = Tn[t]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) { st.O[0][k] = st.O[1][k]; st.O[1][k] = st.O[2][k]; st.O[2][k] = st.O[3][k]; st.O[3][k] = On[k]; }
}

// The whole segment of one strip.  f = frame, s = segment, j = strip.
RB_HD void kpe_strip(const RbKpeParams& p, uint32_t f, uint32_t s, uint32_t j) {
  const RbGeom& g = p.g;
  const uint32_t x0 = RB_STRIP_OUT * j;
  uint32_t vmask = 0x3FFFFFFCu;  // bits 2..29 = output columns of this strip
  {
    const int maxi = (int)g.W - 3 - (int)x0;  // x <= W-3 (src/kpe.hpp:183-184)
    if (maxi < 2) return;
    if (maxi < 29) vmask &= (2u << maxi) - 1u;
  }
  const uint32_t ya = 2 + s * p.seg_rows;                                   // first output row
  const uint32_t yend = g.H - 4;                                            // rows end at H-5 (src/kpe.hpp:268)
  const uint32_t yb = ya + p.seg_rows < yend ? ya + p.seg_rows : yend;      // one past last output row
  if (ya >= yb) return;
  const uint8_t* fbase = p.frames + (uint64_t)f * g.frame_stride + x0;
  uint32_t* kprow = p.kpbits + ((uint64_t)f * g.H + ya) * g.NS + j;
  uint32_t* w2row = p.w2bits + ((uint64_t)f * g.H + ya) * g.NS + j;
  uint8_t* medrow = p.median ? p.median + (uint64_t)f * g.median_stride + (uint64_t)ya * g.mpitch + x0 + 4 : nullptr;

  StripState st;
  uint32_t w[8];
  // warm-up: rows ya-2 .. ya+1
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    load_row(fbase + (uint64_t)(ya - 2 + k) * g.pitch, w);
    uint32_t On[4], Tn[15];
    row_to_planes(w, On, Tn);
#pragma unroll
    for (int t = 0; t < 15; ++t) { st.T[0][t] = st.T[1][t]; st.T[1][t] = st.T[2][t]; st.T[2][t] = st.T[3][t]; st.T[3][t] = Tn[t]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { st.O[0][q] = st.O[1][q]; st.O[1][q] = st.O[2][q]; st.O[2][q] = st.O[3][q]; st.O[3][q] = On[q]; }
  }
  const uint32_t rlast = yb + 1;  // last input row
  load_row(fbase + (uint64_t)(ya + 2) * g.pitch, w);
#pragma unroll 1
  for (uint32_t r = ya + 2; r <= rlast; ++r) {  // input row r enters, output row y = r - 2
    uint32_t wn[8];
    const uint32_t rn = r + 1 <= rlast ? r + 1 : rlast;  // prefetch the next row (clamped)
    load_row(fbase + (uint64_t)rn * g.pitch, wn);
    strip_step(st, w, vmask, kprow, w2row, medrow);
    kprow += g.NS; w2row += g.NS;
    if (medrow) medrow += g.mpitch;
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = wn[k];
  }
}

}  // namespace rbk

#if defined(__CUDACC__)
__global__ void __launch_bounds__(128) rb_kpe_kernel(const RbKpeParams p) {
  const uint32_t items = p.nframes * p.nseg * p.g.NS;
  for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const uint32_t j = it % p.g.NS;
    const uint32_t fs = it / p.g.NS;
    rbk::kpe_strip(p, fs / p.nseg, fs % p.nseg, j);
  }
}
#endif
