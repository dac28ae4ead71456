// oracle/ref_harness.cpp -- harness around the REAL reference (kataklinger/remap) headers.
//
// TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py).  This file is ours; it #includes the
// reference's own headers (patched for GCC in a temp dir at build time) and calls the
// reference's own functions:
//   kpe::extractor<frc::grid_type, frc::grid_overlap>::extract   (src/kpe.hpp:92-108)
//   kpm::match(cfg, prev_grid, curr_grid)                        (src/kpm.hpp:395-415)
//   kpm::details::count_offsets / top_offsets                    (src/kpm.hpp:105-159)
//   frc::collector::collect                                      (src/frc.hpp:55-68)
//   fde::details::generate_mask                                  (src/fde.hpp:19-55)
//   nic::compress                                                (src/nic.hpp:8-105)
//
// Modes
//   dump  <frames.bin> W H N <out.bin>      canonical dump of every intermediate (see below)
//   bench <frames.bin> W H N <reg|frc> T R  time the reference on T host threads, R repeats
//   mask  <bg.bin> bgW bgH px py <frame.bin> W H <out.bin>
//   filter <frames.bin> W H N <out.bin> R   frc::collector::collect + complete, then fdf::filter
//                                           (src/fdf.hpp:40-91) over the collected fragments: dumps every
//                                           frame's foreground contours and fde::mask, the backgrounds and
//                                           the filtered fragments' dots; R > 1 repeats fdf::filter for timing
//   heat  <frames.bin> W H N <out.bin>      aws::details::compare (src/aws.hpp:37-60) over every consecutive pair,
//                                           from aws::scan's initial heat map of ones; dumps the map after each pair
//   digest <frames.bin> W H N <out.bin> T [C]  FULL-SEQUENCE parity record, T host threads on contiguous shards (one-frame
//                                           overlap).  Per frame: order-independent 64-bit digests of the median image
//                                           and of every (region, point, code) insertion of the grid; per pair: the
//                                           reference's kpm::match result and, per region, the weight switch, the
//                                           number of bins, a digest of the whole offset histogram and the ticket in
//                                           the REFERENCE'S OWN order (top_offsets over libstdc++'s unordered_map).
//                                           C = 1 adds (fragment, x, y) of every frame from the unmodified
//                                           frc::collector run on the same shards (stitched at the shard borders).
//   pairs <frames.bin> W H N <pairs.bin> <out.bin>  kpm::match of the listed consecutive pairs only (pairs.bin: uint32 indices
//                                           i = pair (i, i + 1)); out: (valid, dx, dy) int32 per listed pair.  What
//                                           tools/resolve_flagged.py replays the reference's own tie order with.
//   splice <frames.bin> W H N <out.bin>     frc::collector::collect + complete, then (a) every fragment as a
//                                           fgs snippet (blend, kpe with a 1x1 grid, src/fgs.hpp:80-89) and the
//                                           cellular kpm::match (src/kpm.hpp:371-393) of every snippet pair with
//                                           its intermediates, (b) fgs::splice (src/fgs.hpp:187-213) itself
//
// frames.bin = N*H*W bytes, row-major, values 0..15 (== nil::read_raw format, src/nil.hpp:24).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <list>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "aws.hpp"
#include "fde.hpp"
#include "fdf.hpp"
#include "fgs.hpp"
#include "frc.hpp"
#include "nic.hpp"

namespace {

using pixel_alloc_t = frc::allocator_t<cpl::nat_cc>;
using extractor_t = kpe::extractor<frc::grid_type, frc::grid_overlap>;

// Same constants as the private frc::collector::match_config (src/frc.hpp:30-44).
struct match_config {
  using allocator_type = frc::allocator_t<char>;
  static constexpr std::size_t weight_switch{10};
  static constexpr std::size_t region_votes{3};
  explicit match_config(allocator_type const& a) noexcept : allocator_{a} {}
  [[nodiscard]] allocator_type get_allocator() const noexcept { return allocator_; }
  allocator_type allocator_;
};

std::vector<std::uint8_t> read_file(char const* path, std::size_t expect) {
  std::vector<std::uint8_t> buf(expect);
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
  std::size_t got = std::fread(buf.data(), 1, expect, f);
  std::fclose(f);
  if (got != expect) { std::fprintf(stderr, "%s: short read %zu < %zu\n", path, got, expect); std::exit(2); }
  return buf;
}

// In-memory feed satisfying ifd::feeder (src/ifd.hpp:20-28).
class memory_feed {
public:
  memory_feed(std::uint8_t const* data, std::size_t w, std::size_t h, std::size_t first, std::size_t last)
      : data_{data}, dim_{w, h}, next_{first}, last_{last} {}
  [[nodiscard]] bool has_more() const noexcept { return next_ < last_; }
  template<typename Alloc>
  [[nodiscard]] auto produce(Alloc const& alloc) {
    using image_type = sid::nat::aimg_t<Alloc>;
    image_type img{dim_, alloc};
    std::memcpy(img.data(), data_ + next_ * dim_.area(), dim_.area());
    return ifd::frame<image_type>{next_++, std::move(img)};
  }
private:
  std::uint8_t const* data_;
  mrl::dimensions_t dim_;
  std::size_t next_, last_;
};

struct null_compression {
  template<typename Alloc>
  [[nodiscard]] icd::compressed_t operator()(sid::nat::aimg_t<Alloc> const&) const { return {}; }
};
struct native_compression {  // what main.cpp uses (src/main.cpp:112-125)
  template<typename Alloc>
  [[nodiscard]] icd::compressed_t operator()(sid::nat::aimg_t<Alloc> const& image) const {
    return nic::compress(image);
  }
};

void put32(FILE* f, std::uint32_t v) { std::fwrite(&v, 4, 1, f); }
void puti32(FILE* f, std::int32_t v) { std::fwrite(&v, 4, 1, f); }

struct kp_rec { std::uint16_t x, y; std::uint8_t code[13]; };

void dump_grid(FILE* out, frc::grid_type const& grid) {
  for (auto& region : grid.regions()) {
    std::vector<kp_rec> recs;
    for (auto& [code, pts] : region.points()) {
      for (auto& p : pts) {
        kp_rec r;
        r.x = static_cast<std::uint16_t>(p.x_);
        r.y = static_cast<std::uint16_t>(p.y_);
        std::memcpy(r.code, code.data(), 13);
        recs.push_back(r);
      }
    }
    std::sort(recs.begin(), recs.end(), [](kp_rec const& a, kp_rec const& b) {
      return a.x != b.x ? a.x < b.x : a.y < b.y;
    });
    put32(out, static_cast<std::uint32_t>(recs.size()));
    put32(out, static_cast<std::uint32_t>(region.counts()[1]));
    put32(out, static_cast<std::uint32_t>(region.counts()[2]));
    for (auto& r : recs) {
      std::fwrite(&r.x, 2, 1, out);
      std::fwrite(&r.y, 2, 1, out);
      std::fwrite(r.code, 1, 13, out);
    }
  }
}

template<typename Alloc>
void dump_pair(FILE* out, frc::grid_type const& prev, frc::grid_type const& curr, Alloc const& alloc) {
  match_config cfg{alloc};
  auto off = kpm::match(cfg, prev, curr);  // the reference's own declaration
  put32(out, off ? 1u : 0u);
  puti32(out, off ? off->x_ : 0);
  puti32(out, off ? off->y_ : 0);
  put32(out, static_cast<std::uint32_t>(kpm::details::get_active(curr)));

  auto pregs{prev.regions()}, cregs{curr.regions()};
  for (std::size_t i = 0; i < frc::grid_type::region_count; ++i) {
    // same switch as kpm::details::cast_vote (src/kpm.hpp:217-222)
    bool use_all = pregs[i].counts()[2] < match_config::weight_switch ||
                   cregs[i].counts()[2] <= match_config::weight_switch;
    auto total = use_all ? kpm::details::count_offsets<true>(cfg, pregs[i], cregs[i])
                         : kpm::details::count_offsets<false>(cfg, pregs[i], cregs[i]);
    auto ticket = kpm::details::top_offsets(cfg, total, match_config::region_votes);

    std::map<std::pair<std::int32_t, std::int32_t>, std::uint32_t> sorted;
    for (auto& [o, c] : total) sorted[{o.x_, o.y_}] = static_cast<std::uint32_t>(c);
    put32(out, use_all ? 1u : 0u);
    put32(out, static_cast<std::uint32_t>(sorted.size()));
    for (auto& [o, c] : sorted) { puti32(out, o.first); puti32(out, o.second); put32(out, c); }
    put32(out, static_cast<std::uint32_t>(ticket.size()));
    for (auto& v : ticket) {
      puti32(out, v.offset_.x_); puti32(out, v.offset_.y_);
      put32(out, static_cast<std::uint32_t>(v.count_));
    }
  }
}

int run_dump(int argc, char** argv) {
  if (argc < 7) return 1;
  std::size_t W = std::atoll(argv[3]), H = std::atoll(argv[4]), N = std::atoll(argv[5]);
  auto frames = read_file(argv[2], N * W * H);
  FILE* out = std::fopen(argv[6], "wb");
  if (!out) return 2;
  std::fwrite("RMDP", 1, 4, out);
  put32(out, W); put32(out, H); put32(out, N);

  // section 1: per frame median + grid; per pair votes and the declared offset
  {
    extractor_t extractor{mrl::dimensions_t{W, H}};
    all::memory_stack<cpl::nat_cc> memory{};
    memory_feed feed{frames.data(), W, H, 0, N};

    auto first_alloc{memory.previous()};
    auto frame{feed.produce(first_alloc)};
    frc::image_type median{frame.image_.dimensions(), first_alloc};
    auto pkeys{extractor.extract(frame.image_, median, first_alloc)};
    std::fwrite(median.data(), 1, W * H, out);
    dump_grid(out, pkeys);

    while (feed.has_more()) {
      all::memory_swing swing{memory};
      pixel_alloc_t alloc{swing};
      auto fr{feed.produce(alloc)};
      frc::image_type med{fr.image_.dimensions(), alloc};
      auto keys{extractor.extract(fr.image_, med, alloc)};
      std::fwrite(med.data(), 1, W * H, out);
      dump_grid(out, keys);
      dump_pair(out, pkeys, keys, alloc);
      pkeys = std::move(keys);
    }
  }

  // section 2: the unmodified frc::collector loop -> (fragment index, position) per frame
  {
    frc::collector collector{mrl::dimensions_t{W, H}};
    memory_feed feed{frames.data(), W, H, 0, N};
    std::vector<std::int32_t> rec(3 * N, 0);  // fragment, x, y per frame (frame 0 is (0,0,0))
    collector.collect(feed, null_compression{},
                      [&](fgm::fragment const& frag, frc::frame_type const& fr,
                          frc::image_type const&, frc::grid_type const&) {
                        auto& pos = frag.frames().back().position_;  // raw, before normalize()
                        rec[3 * fr.number_ + 1] = pos.x_;
                        rec[3 * fr.number_ + 2] = pos.y_;
                      });
    auto frags = collector.complete();
    std::int32_t fi = 0;
    for (auto& f : frags) {
      for (auto& fr : f.frames()) rec[3 * fr.number_] = fi;
      ++fi;
    }
    put32(out, static_cast<std::uint32_t>(rec.size() / 3));
    std::fwrite(rec.data(), 4, rec.size(), out);
  }
  std::fclose(out);
  return 0;
}

// ---- bench ------------------------------------------------------------------------------
struct shard_result { double seconds; std::size_t frames; std::size_t keypoints; std::int64_t checksum; };

shard_result bench_reg(std::uint8_t const* frames, std::size_t W, std::size_t H, std::size_t first, std::size_t last) {
  // kpe + kpm + position accumulation only (the registration path this repo replaces)
  auto t0 = std::chrono::steady_clock::now();
  extractor_t extractor{mrl::dimensions_t{W, H}};
  all::memory_stack<cpl::nat_cc> memory{};
  memory_feed feed{frames, W, H, first, last};
  auto first_alloc{memory.previous()};
  auto frame{feed.produce(first_alloc)};
  frc::image_type median{frame.image_.dimensions(), first_alloc};
  auto pkeys{extractor.extract(frame.image_, median, first_alloc)};
  std::int64_t px = 0, py = 0, checksum = 0;
  std::size_t kps = 0;
  while (feed.has_more()) {
    all::memory_swing swing{memory};
    pixel_alloc_t alloc{swing};
    auto fr{feed.produce(alloc)};
    frc::image_type med{fr.image_.dimensions(), alloc};
    auto keys{extractor.extract(fr.image_, med, alloc)};
    if (auto off{kpm::match(match_config{alloc}, pkeys, keys)}; off) { px += off->x_; py += off->y_; }
    else { px = py = 0; checksum += 1000003; }
    checksum += px * 31 + py;
    for (auto& r : keys.regions()) kps += r.total_count();
    pkeys = std::move(keys);
  }
  auto t1 = std::chrono::steady_clock::now();
  return {std::chrono::duration<double>(t1 - t0).count(), last - first, kps, checksum};
}

shard_result bench_frc(std::uint8_t const* frames, std::size_t W, std::size_t H, std::size_t first, std::size_t last) {
  // the reference's whole per-frame loop: kpe + kpm + nic compression x2 + fragment blit
  auto t0 = std::chrono::steady_clock::now();
  frc::collector collector{mrl::dimensions_t{W, H}};
  memory_feed feed{frames, W, H, first, last};
  std::size_t kps = 0;
  collector.collect(feed, native_compression{},
                    [&](fgm::fragment const&, frc::frame_type const&, frc::image_type const&,
                        frc::grid_type const& g) { for (auto& r : g.regions()) kps += r.total_count(); });
  auto frags = collector.complete();
  auto t1 = std::chrono::steady_clock::now();
  std::int64_t checksum = static_cast<std::int64_t>(frags.size());
  return {std::chrono::duration<double>(t1 - t0).count(), last - first, kps, checksum};
}

int run_bench(int argc, char** argv) {
  if (argc < 9) return 1;
  std::size_t W = std::atoll(argv[3]), H = std::atoll(argv[4]), N = std::atoll(argv[5]);
  std::string mode = argv[6];
  std::size_t T = std::atoll(argv[7]), R = std::atoll(argv[8]);
  auto frames = read_file(argv[2], N * W * H);
  if (T < 1) T = 1;
  if (T > N / 2) T = std::max<std::size_t>(1, N / 2);
  double best = 1e30, total = 0;
  std::size_t kps = 0, pairs = 0;
  std::int64_t checksum = 0;
  for (std::size_t rep = 0; rep < R; ++rep) {
    std::vector<shard_result> res(T);
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (std::size_t t = 0; t < T; ++t) {
      // contiguous shards with a one-frame overlap: every consecutive pair belongs to one shard
      std::size_t lo = t * N / T, hi = (t + 1) * N / T;
      std::size_t first = lo == 0 ? 0 : lo - 1;
      th.emplace_back([&, t, first, hi] {
        res[t] = mode == "frc" ? bench_frc(frames.data(), W, H, first, hi)
                               : bench_reg(frames.data(), W, H, first, hi);
      });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();
    best = std::min(best, wall);
    total += wall;
    kps = 0; checksum = 0;
    for (auto& r : res) { kps += r.keypoints; checksum += r.checksum; }
    pairs = N - 1;
  }
  std::printf("{\"mode\": \"%s\", \"threads\": %zu, \"frames\": %zu, \"pairs\": %zu, \"reps\": %zu, "
              "\"best_s\": %.6f, \"mean_s\": %.6f, \"fps_best\": %.3f, \"fps_mean\": %.3f, "
              "\"keypoint_insertions_per_frame\": %.1f, \"checksum\": %lld}\n",
              mode.c_str(), T, N, pairs, R, best, total / R, N / best, N / (total / R),
              double(kps) / double(N > 1 ? N - 1 : 1), static_cast<long long>(checksum));
  return 0;
}


// ---- digest -----------------------------------------------------------------------------
// Record layouts (mirrored by tests/digest_check.py and, for the digests themselves, by rb_digest_kernel):
//   splitmix(z): z += 0x9E3779B97F4A7C15; z = (z ^ z >> 30) * 0xBF58476D1CE4E5B9; z = (z ^ z >> 27) * 0x94D049BB133111EB; z ^ z >> 31
//   median_hash = sum over pixels with value v != 0 at index i = y * W + x of splitmix(i << 8 | v)
//   kp_hash     = sum over regions r and insertions (code, (x, y)) of
//                 splitmix((x | y << 16 | r << 32) ^ splitmix(code[0..7] ^ splitmix(code[8..12])))   (little-endian words)
//   hist_hash   = 32-bit sum over the bins of a region's totalizator_t of the bin digest of include/remap_b200.h
static inline std::uint64_t splitmix(std::uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline std::uint32_t bin_hash(std::int32_t dx, std::int32_t dy, std::uint32_t cnt) {
  std::uint32_t h = (static_cast<std::uint32_t>(dx) & 0xFFFFu) | (static_cast<std::uint32_t>(dy) << 16);
  h = h * 0x9E3779B1u ^ cnt * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

#pragma pack(push, 1)
struct frame_digest { std::uint64_t median_hash, kp_hash; std::uint32_t insertions; std::uint32_t n[8], w2[8]; };
struct region_digest { std::uint32_t use_all, nbins, hist_hash, nticket; std::int32_t tdx[3], tdy[3]; std::uint32_t tcnt[3]; };
struct pair_digest { std::uint32_t valid; std::int32_t dx, dy; std::uint32_t active; region_digest r[8]; };
#pragma pack(pop)

template<typename Image>
frame_digest digest_frame(Image const& median, frc::grid_type const& grid, std::size_t W, std::size_t H) {
  frame_digest d{};
  auto const* m = reinterpret_cast<std::uint8_t const*>(median.data());
  for (std::size_t i = 0; i < W * H; ++i)
    if (m[i]) d.median_hash += splitmix((static_cast<std::uint64_t>(i) << 8) | m[i]);
  std::size_t r = 0;
  for (auto& region : grid.regions()) {
    for (auto& [code, pts] : region.points()) {
      std::uint64_t lo = 0, hi = 0;
      std::memcpy(&lo, code.data(), 8);
      std::memcpy(&hi, reinterpret_cast<std::uint8_t const*>(code.data()) + 8, 5);
      std::uint64_t const ch = splitmix(lo ^ splitmix(hi));
      for (auto& p : pts) {
        std::uint64_t const a = static_cast<std::uint64_t>(p.x_) | (static_cast<std::uint64_t>(p.y_) << 16) |
                                (static_cast<std::uint64_t>(r) << 32);
        d.kp_hash += splitmix(a ^ ch);
        ++d.insertions;
      }
    }
    if (r < 8) {
      d.n[r] = static_cast<std::uint32_t>(region.counts()[1] + region.counts()[2]);
      d.w2[r] = static_cast<std::uint32_t>(region.counts()[2]);
    }
    ++r;
  }
  return d;
}

template<typename Alloc>
pair_digest digest_pair(frc::grid_type const& prev, frc::grid_type const& curr, Alloc const& alloc) {
  pair_digest d{};
  match_config cfg{alloc};
  auto off = kpm::match(cfg, prev, curr);  // the reference's own declaration
  d.valid = off ? 1u : 0u;
  d.dx = off ? off->x_ : 0;
  d.dy = off ? off->y_ : 0;
  d.active = static_cast<std::uint32_t>(kpm::details::get_active(curr));
  auto pregs{prev.regions()}, cregs{curr.regions()};
  for (std::size_t i = 0; i < frc::grid_type::region_count && i < 8; ++i) {
    bool use_all = pregs[i].counts()[2] < match_config::weight_switch || cregs[i].counts()[2] <= match_config::weight_switch;
    auto total = use_all ? kpm::details::count_offsets<true>(cfg, pregs[i], cregs[i])
                         : kpm::details::count_offsets<false>(cfg, pregs[i], cregs[i]);
    auto& rd = d.r[i];
    rd.use_all = use_all ? 1u : 0u;
    rd.nbins = static_cast<std::uint32_t>(total.size());
    for (auto& [o, c] : total) rd.hist_hash += bin_hash(o.x_, o.y_, static_cast<std::uint32_t>(c));
    auto ticket = kpm::details::top_offsets(cfg, total, match_config::region_votes);  // consumes `total`
    rd.nticket = static_cast<std::uint32_t>(ticket.size());
    std::size_t k = 0;
    for (auto& v : ticket) {
      if (k < 3) { rd.tdx[k] = v.offset_.x_; rd.tdy[k] = v.offset_.y_; rd.tcnt[k] = static_cast<std::uint32_t>(v.count_); }
      ++k;
    }
  }
  return d;
}

// frames [first, last) (first = own range minus the overlap frame): frame digests for all of them, pair digests for
// pairs (i - 1, i), i in (first, last)
void digest_shard(std::uint8_t const* frames, std::size_t W, std::size_t H, std::size_t first, std::size_t last,
                  frame_digest* fd, pair_digest* pd) {
  extractor_t extractor{mrl::dimensions_t{W, H}};
  all::memory_stack<cpl::nat_cc> memory{};
  memory_feed feed{frames, W, H, first, last};
  auto first_alloc{memory.previous()};
  auto frame{feed.produce(first_alloc)};
  frc::image_type median{frame.image_.dimensions(), first_alloc};
  auto pkeys{extractor.extract(frame.image_, median, first_alloc)};
  fd[first] = digest_frame(median, pkeys, W, H);
  std::size_t i = first + 1;
  while (feed.has_more()) {
    all::memory_swing swing{memory};
    pixel_alloc_t alloc{swing};
    auto fr{feed.produce(alloc)};
    frc::image_type med{fr.image_.dimensions(), alloc};
    auto keys{extractor.extract(fr.image_, med, alloc)};
    fd[i] = digest_frame(med, keys, W, H);
    pd[i - 1] = digest_pair(pkeys, keys, alloc);
    pkeys = std::move(keys);
    ++i;
  }
}

// the unmodified collector on frames [first, last): (fragment, x, y) per frame relative to the shard's first frame
void collect_shard(std::uint8_t const* frames, std::size_t W, std::size_t H, std::size_t first, std::size_t last,
                   std::size_t own, std::int32_t* rec /* 3 per frame, indexed by global frame number; only frames >= own */) {
  frc::collector collector{mrl::dimensions_t{W, H}};
  memory_feed feed{frames, W, H, first, last};
  collector.collect(feed, null_compression{},
                    [&](fgm::fragment const& frag, frc::frame_type const& fr, frc::image_type const&, frc::grid_type const&) {
                      auto& pos = frag.frames().back().position_;  // raw, before normalize()
                      rec[3 * fr.number_ + 1] = pos.x_;
                      rec[3 * fr.number_ + 2] = pos.y_;
                    });
  auto frags = collector.complete();
  std::int32_t fi = 0;
  for (auto& f : frags) {
    for (auto& fr : f.frames())
      if (fr.number_ >= own) rec[3 * fr.number_] = fi;  // the overlap frame belongs to the previous shard
    ++fi;
  }
}

int run_digest(int argc, char** argv) {
  if (argc < 8) return 1;
  std::size_t W = std::atoll(argv[3]), H = std::atoll(argv[4]), N = std::atoll(argv[5]);
  std::size_t T = std::atoll(argv[7]);
  bool const with_collector = argc > 8 && std::atoi(argv[8]) != 0;
  auto frames = read_file(argv[2], N * W * H);
  if (T < 1) T = 1;
  if (T > N / 2) T = std::max<std::size_t>(1, N / 2);
  std::vector<frame_digest> fd(N);
  std::vector<pair_digest> pd(N > 1 ? N - 1 : 0);
  std::vector<std::int32_t> rec(with_collector ? 3 * N : 0, 0), loc(with_collector ? 3 * N : 0, 0);
  std::vector<std::size_t> starts;
  auto t0 = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> th;
    for (std::size_t t = 0; t < T; ++t) {
      std::size_t lo = t * N / T, hi = (t + 1) * N / T;
      std::size_t first = lo == 0 ? 0 : lo - 1;
      starts.push_back(first);
      th.emplace_back([&, first, lo, hi] {
        digest_shard(frames.data(), W, H, first, hi, fd.data(), pd.data());
        if (with_collector) collect_shard(frames.data(), W, H, first, hi, lo, loc.data());
      });
    }
    for (auto& x : th) x.join();
  }
  if (with_collector) {
    // stitch: a shard's first frame is the previous shard's last (already global); frames of the shard's fragment 0
    // continue that frame's fragment at its position, later fragments of the shard are new ones
    for (std::size_t t = 0; t < T; ++t) {
      std::size_t lo = t * N / T, hi = (t + 1) * N / T;
      std::size_t first = starts[t];
      std::int32_t bf = 0, bx = 0, by = 0;
      if (t > 0) { bf = rec[3 * first]; bx = rec[3 * first + 1]; by = rec[3 * first + 2]; }
      for (std::size_t i = (t == 0 ? 0 : lo); i < hi; ++i) {
        std::int32_t f = loc[3 * i], x = loc[3 * i + 1], y = loc[3 * i + 2];
        if (f == 0) { rec[3 * i] = bf; rec[3 * i + 1] = bx + x; rec[3 * i + 2] = by + y; }
        else { rec[3 * i] = bf + f; rec[3 * i + 1] = x; rec[3 * i + 2] = y; }
      }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  FILE* out = std::fopen(argv[6], "wb");
  if (!out) return 2;
  std::fwrite("RMDG", 1, 4, out);
  put32(out, W); put32(out, H); put32(out, N); put32(out, with_collector ? 1u : 0u);
  std::fwrite(fd.data(), sizeof(frame_digest), fd.size(), out);
  std::fwrite(pd.data(), sizeof(pair_digest), pd.size(), out);
  if (with_collector) std::fwrite(rec.data(), 4, rec.size(), out);
  std::fclose(out);
  std::printf("{\"mode\": \"digest\", \"threads\": %zu, \"frames\": %zu, \"seconds\": %.3f}\n", T, N,
              std::chrono::duration<double>(t1 - t0).count());
  return 0;
}

int run_pairs(int argc, char** argv) {
  if (argc < 8) return 1;
  std::size_t W = std::atoll(argv[3]), H = std::atoll(argv[4]), N = std::atoll(argv[5]);
  auto frames = read_file(argv[2], N * W * H);
  FILE* pf = std::fopen(argv[6], "rb");
  if (!pf) return 2;
  std::vector<std::uint32_t> pairs;
  for (std::uint32_t v; std::fread(&v, 4, 1, pf) == 1;) pairs.push_back(v);
  std::fclose(pf);
  FILE* out = std::fopen(argv[7], "wb");
  if (!out) return 2;
  extractor_t extractor{mrl::dimensions_t{W, H}};
  for (auto p : pairs) {
    if (p + 1 >= N) { puti32(out, 0); puti32(out, 0); puti32(out, 0); continue; }
    all::memory_stack<cpl::nat_cc> memory{};
    memory_feed feed{frames.data(), W, H, p, p + 2};
    auto a0{memory.previous()};
    auto f0{feed.produce(a0)};
    frc::image_type m0{f0.image_.dimensions(), a0};
    auto k0{extractor.extract(f0.image_, m0, a0)};
    all::memory_swing swing{memory};
    pixel_alloc_t a1{swing};
    auto f1{feed.produce(a1)};
    frc::image_type m1{f1.image_.dimensions(), a1};
    auto k1{extractor.extract(f1.image_, m1, a1)};
    auto off = kpm::match(match_config{a1}, k0, k1);
    puti32(out, off ? 1 : 0); puti32(out, off ? off->x_ : 0); puti32(out, off ? off->y_ : 0);
  }
  std::fclose(out);
  return 0;
}

int run_mask(int argc, char** argv) {
  if (argc < 11) return 1;
  std::size_t bw = std::atoll(argv[3]), bh = std::atoll(argv[4]);
  std::int32_t px = std::atoi(argv[5]), py = std::atoi(argv[6]);
  std::size_t W = std::atoll(argv[8]), H = std::atoll(argv[9]);
  auto bg = read_file(argv[2], bw * bh);
  auto fr = read_file(argv[7], W * H);
  sid::nat::dimg_t background{mrl::dimensions_t{bw, bh}};
  std::memcpy(background.data(), bg.data(), bg.size());
  sid::nat::dimg_t frame{mrl::dimensions_t{W, H}};
  std::memcpy(frame.data(), fr.data(), fr.size());
  sid::mon::dimg_t mask{mrl::dimensions_t{W, H}};
  auto idx = cdt::to_index(fgm::point_t{px, py}, background.dimensions());  // src/fde.hpp:87
  if (idx % fde::details::mm_size == 0)                                     // src/fde.hpp:107-114
    fde::details::generate_mask(background, frame, mask, idx, std::true_type{});
  else
    fde::details::generate_mask(background, frame, mask, idx, std::false_type{});
  FILE* out = std::fopen(argv[10], "wb");
  if (!out) return 2;
  std::fwrite(mask.data(), 1, W * H, out);
  std::fclose(out);
  return 0;
}

// ---- pass 2: fdf::filter ------------------------------------------------------------------
struct native_codec {  // both faces of main.cpp's native_compression (src/main.cpp:112-125)
  template<typename Alloc>
  [[nodiscard]] icd::compressed_t operator()(sid::nat::aimg_t<Alloc> const& image) const {
    return nic::compress(image);
  }
  [[nodiscard]] sid::nat::dimg_t operator()(icd::compressed_t const& c, mrl::dimensions_t const& dim) const {
    return nic::decompress(c, dim);
  }
};

int run_filter(int argc, char** argv) {
  if (argc < 7) return 1;
  std::size_t W = std::atoll(argv[3]), H = std::atoll(argv[4]), N = std::atoll(argv[5]);
  std::size_t R = argc > 7 ? std::atoll(argv[7]) : 1;
  auto frames = read_file(argv[2], N * W * H);
  mrl::dimensions_t const dim{W, H};

  std::vector<fgm::fragment> fragments;
  {
    frc::collector collector{dim};
    memory_feed feed{frames.data(), W, H, 0, N};
    collector.collect(feed, native_codec{},
                      [](fgm::fragment const&, frc::frame_type const&, frc::image_type const&,
                         frc::grid_type const&) {});
    auto list = collector.complete();
    for (auto& f : list) fragments.push_back(std::move(f));
  }

  FILE* out = std::fopen(argv[6], "wb");
  if (!out) return 2;
  std::fwrite("RMF2", 1, 4, out);
  put32(out, W); put32(out, H); put32(out, N);
  put32(out, static_cast<std::uint32_t>(fragments.size()));

  // backgrounds exactly as the 4-argument fdf::filter makes them (src/fdf.hpp:21-34,79-89)
  auto backgrounds = fdf::details::get_background(fragments);
  for (auto& b : backgrounds) {
    put32(out, static_cast<std::uint32_t>(b.image_.width()));
    put32(out, static_cast<std::uint32_t>(b.image_.height()));
    puti32(out, b.zero_.x_); puti32(out, b.zero_.y_);
    std::fwrite(b.image_.data(), 1, b.image_.size(), out);
  }

  double best = 1e30;
  std::vector<fgm::fragment> filtered;
  for (std::size_t rep = 0; rep < R; ++rep) {
    bool const dump = rep == 0;
    auto t0 = std::chrono::steady_clock::now();
    filtered = fdf::filter(
        fragments, backgrounds, dim, native_codec{},
        [&](fgm::fragment const& result, std::size_t frag, sid::nat::dimg_t const& image, std::size_t no,
            sid::nat::dimg_t const& median, fgm::point_t const& pos, fdf::contours_t const& foreground,
            auto const& mask) {
          if (!dump) return;
          put32(out, static_cast<std::uint32_t>(frag));
          put32(out, static_cast<std::uint32_t>(no));
          puti32(out, pos.x_ - result.zero().x_);
          puti32(out, pos.y_ - result.zero().y_);
          put32(out, static_cast<std::uint32_t>(foreground.size()));
          for (auto& c : foreground) {
            auto& e = c.enclosure();
            put32(out, c.area());
            put32(out, static_cast<std::uint32_t>(e.left_)); put32(out, static_cast<std::uint32_t>(e.top_));
            put32(out, static_cast<std::uint32_t>(e.right_)); put32(out, static_cast<std::uint32_t>(e.bottom_));
            put32(out, static_cast<std::uint32_t>(value(c.color())));
          }
          std::fwrite(mask.data(), 1, W * H, out);
        });
    auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  for (auto& f : filtered) {
    put32(out, static_cast<std::uint32_t>(f.dots().width()));
    put32(out, static_cast<std::uint32_t>(f.dots().height()));
    puti32(out, f.zero().x_); puti32(out, f.zero().y_);
    std::fwrite(f.dots().data(), sizeof(fgm::dot_type), f.dots().size(), out);
  }
  std::fclose(out);
  std::printf("{\"mode\": \"filter\", \"frames\": %zu, \"fragments\": %zu, \"reps\": %zu, \"best_s\": %.6f, "
              "\"fps_best\": %.3f}\n", N, fragments.size(), R, best, N / best);
  return 0;
}

// ---- fragment splicing: fgs::splice ---------------------------------------------------------
void dump_fragment(FILE* out, fgm::fragment const& f) {
  put32(out, static_cast<std::uint32_t>(f.dots().width()));
  put32(out, static_cast<std::uint32_t>(f.dots().height()));
  puti32(out, f.zero().x_); puti32(out, f.zero().y_);
  put32(out, static_cast<std::uint32_t>(f.frames().size()));
  for (auto& fr : f.frames()) {
    put32(out, static_cast<std::uint32_t>(fr.number_));
    puti32(out, fr.position_.x_); puti32(out, fr.position_.y_);
  }
  std::fwrite(f.dots().data(), sizeof(fgm::dot_type), f.dots().size(), out);
}

int run_splice(int argc, char** argv) {
  if (argc < 7) return 1;
  std::size_t W = std::atoll(argv[3]), H = std::atoll(argv[4]), N = std::atoll(argv[5]);
  auto frames = read_file(argv[2], N * W * H);
  mrl::dimensions_t const dim{W, H};
  std::vector<fgm::fragment> fragments;
  {
    frc::collector collector{dim};
    memory_feed feed{frames.data(), W, H, 0, N};
    collector.collect(feed, native_codec{},
                      [](fgm::fragment const&, frc::frame_type const&, frc::image_type const&, frc::grid_type const&) {});
    auto list = collector.complete();
    for (auto& f : list) fragments.push_back(std::move(f));
  }
  FILE* out = std::fopen(argv[6], "wb");
  if (!out) return 2;
  std::fwrite("RMSP", 1, 4, out);
  put32(out, W); put32(out, H); put32(out, N);
  put32(out, static_cast<std::uint32_t>(fragments.size()));
  for (auto& f : fragments) dump_fragment(out, f);

  // (a) the snippets and every pairwise cellular match, in the order fgs::details::match_all makes them
  std::vector<fgs::details::snippet> snips;
  for (auto& f : fragments) snips.push_back(fgs::details::extract_single(fgm::fragment{f}));
  for (auto& sn : snips) {
    auto& reg = sn.grid_[0];
    std::vector<kp_rec> recs;
    for (auto& [code, pts] : reg.points())
      for (auto& p : pts) {
        kp_rec r;
        r.x = static_cast<std::uint16_t>(p.x_); r.y = static_cast<std::uint16_t>(p.y_);
        std::memcpy(r.code, code.data(), 13);
        recs.push_back(r);
      }
    std::sort(recs.begin(), recs.end(), [](kp_rec const& a, kp_rec const& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });
    put32(out, static_cast<std::uint32_t>(sn.mask_.width()));
    put32(out, static_cast<std::uint32_t>(sn.mask_.height()));
    put32(out, static_cast<std::uint32_t>(recs.size()));
    for (auto& r : recs) { std::fwrite(&r.x, 2, 1, out); std::fwrite(&r.y, 2, 1, out); std::fwrite(r.code, 1, 13, out); }
    std::fwrite(sn.mask_.data(), 1, sn.mask_.size(), out);
  }
  constexpr kpm::cell_size_t cell{15, 15};  // src/fgs.hpp:121
  for (std::size_t i = 0; i < snips.size(); ++i)
    for (std::size_t j = i + 1; j < snips.size(); ++j) {
      auto& a = snips[i];
      auto& b = snips[j];
      auto totals = kpm::details::count_offsets(a.grid_[0], b.grid_[0], cell);
      put32(out, static_cast<std::uint32_t>(i)); put32(out, static_cast<std::uint32_t>(j));
      put32(out, static_cast<std::uint32_t>(totals.size()));
      std::uint64_t pairs = 0;
      std::size_t best_kp = 0, ties = 0;
      for (auto& [off, cells] : totals) {
        std::size_t kp = 0;
        for (auto& [c, n] : cells) kp += n;
        pairs += kp;
        if (kp > best_kp) { best_kp = kp; ties = 1; }
        else if (kp == best_kp) ++ties;
      }
      std::fwrite(&pairs, 8, 1, out);
      put32(out, static_cast<std::uint32_t>(ties));  // offsets sharing the largest matched_keypoints
      if (!totals.empty()) {
        auto best = kpm::details::find_best(totals);
        auto active = kpm::details::count_active_cells(a.grid_[0], a.mask_, b.grid_[0], b.mask_, best.offset_, cell);
        puti32(out, best.offset_.x_); puti32(out, best.offset_.y_);
        put32(out, static_cast<std::uint32_t>(best.matched_keypoints_));
        put32(out, static_cast<std::uint32_t>(best.matched_cells_));
        put32(out, static_cast<std::uint32_t>(active));
      }
      auto vote = kpm::match(a.grid_[0], a.mask_, b.grid_[0], b.mask_, cell);  // the reference's own verdict
      put32(out, vote ? 1u : 0u);
      puti32(out, vote ? vote->offset_.x_ : 0); puti32(out, vote ? vote->offset_.y_ : 0);
      put32(out, vote ? static_cast<std::uint32_t>(vote->count_) : 0u);
    }

  // (b) fgs::splice on copies of the fragments
  auto t0 = std::chrono::steady_clock::now();
  std::vector<fgm::fragment> copy{fragments};
  auto spliced = fgs::splice(copy.begin(), copy.end());
  auto t1 = std::chrono::steady_clock::now();
  put32(out, static_cast<std::uint32_t>(spliced.size()));
  for (auto& f : spliced) dump_fragment(out, f);
  std::fclose(out);
  std::printf("{\"mode\": \"splice\", \"frames\": %zu, \"fragments\": %zu, \"spliced\": %zu, \"splice_s\": %.6f}\n", N,
              fragments.size(), spliced.size(), std::chrono::duration<double>(t1 - t0).count());
  return 0;
}

int run_heat(int argc, char** argv) {
  if (argc < 7) return 1;
  std::size_t W = std::atoll(argv[3]), H = std::atoll(argv[4]), N = std::atoll(argv[5]);
  auto frames = read_file(argv[2], N * W * H);
  FILE* out = std::fopen(argv[6], "wb");
  if (!out) return 2;
  mrl::dimensions_t const dim{W, H};
  aws::heatmap_type heatmap{dim, {1}};  // src/aws.hpp:113
  // padded copies: compare's vector loop runs to the 32-byte boundary past the image (src/aws.hpp:45-52)
  sid::nat::dimg_t prev{dim}, cur{dim};
  std::memcpy(prev.data(), frames.data(), W * H);
  for (std::size_t i = 1; i < N; ++i) {
    std::memcpy(cur.data(), frames.data() + i * W * H, W * H);
    aws::details::compare(prev, cur, heatmap);
    std::fwrite(heatmap.data(), 1, W * H, out);
    std::swap(prev, cur);
  }
  std::fclose(out);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: ref_harness dump|bench|mask ...\n"); return 1; }
  std::string mode = argv[1];
  if (mode == "dump") return run_dump(argc, argv);
  if (mode == "bench") return run_bench(argc, argv);
  if (mode == "digest") return run_digest(argc, argv);
  if (mode == "pairs") return run_pairs(argc, argv);
  if (mode == "mask") return run_mask(argc, argv);
  if (mode == "filter") return run_filter(argc, argv);
  if (mode == "splice") return run_splice(argc, argv);
  if (mode == "heat") return run_heat(argc, argv);
  return 1;
}
