// oracle/shim_harness.cpp -- drop-in check of include/frc_b200.hpp against the REAL frc::collector.
//
// TEST INFRASTRUCTURE ONLY (built by oracle/build_ref.py into oracle/_ref/shim_harness, which links
// remap_b200/libremap_b200.so).  This file is ours.  It includes the reference's own headers
// (patched for GCC in a temp dir at build time), runs the same frame sequence through
//   frc::collector::collect        (src/frc.hpp:55-68)   -- the reference, CPU
//   frc_b200::collector::collect   (include/frc_b200.hpp) -- the C-ABI / B200 path
// with the reference's feeder concept (src/ifd.hpp:20-28), the reference's nic::compress
// (src/nic.hpp:8-105) as the compressor and a recording callback, and compares everything the rest
// of the pipeline (mpb/fgs/fdf) can observe of the result:
//   fragment count; per fragment: zero, dots dimensions, every dot histogram, and per frame its
//   number, position and the compressed image + compressed MEDIAN bytes; the callback sequence;
//   and (fill_keys = 1) the kpr::grid contents handed to the callback.
//
// With filter = 1 the reference's fragments then go through pass 2 twice,
//   fdf::filter        (src/fdf.hpp:77-89)       -- the reference, CPU
//   fdf_b200::filter   (include/fdf_b200.hpp)    -- rb_filter_fragment on the GPU
// and the results are compared: per fragment zero, dimensions, every dot histogram, the frame list; per
// callback the fragment index, frame number, position and the fde::mask image.
//
// With splice = 1 the reference's fragments are also spliced twice,
//   fgs::splice        (src/fgs.hpp:187-213)     -- the reference, CPU
//   fgs_b200::splice   (include/fgs_b200.hpp)    -- rb_snippet_create / rb_snippet_match on the GPU
// and the resulting fragments compared (count, order, zero, dimensions, dots, frame lists).
//
// gpu_blit = 2: as 1, and the collector keeps no compressed copies and fetches no medians (options::keep_packed,
// fetch_medians = false): the compressed-image comparisons are skipped, the callback medians too; with filter = 1
// only the resident pass 2 can run on such fragments.
//
// usage: shim_harness <frames.bin> W H N <batch> <fill_keys 0|1> [gpu_blit 0|1|2] [filter 0|1] [splice 0|1] [devices d0,d1,...]
//        exit 0 = identical.  devices with more than one entry: frc_b200::options::devices (rb_group, one context per device)

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
#include <memory>
#include <string>
#include <vector>

#include "frc.hpp"
#include "nic.hpp"

#include "fdf.hpp"
#include "fgs.hpp"

#include "fdf_b200.hpp"
#include "fgs_b200.hpp"
#include "frc_b200.hpp"

namespace {

std::vector<std::uint8_t> read_file(char const* path, std::size_t expect) {
  std::vector<std::uint8_t> buf(expect);
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
  std::size_t got = std::fread(buf.data(), 1, expect, f);
  std::fclose(f);
  if (got != expect) { std::fprintf(stderr, "%s: short read\n", path); std::exit(2); }
  return buf;
}

class memory_feed {  // models ifd::feeder
public:
  memory_feed(std::uint8_t const* data, std::size_t w, std::size_t h, std::size_t n)
      : data_{data}, dim_{w, h}, last_{n} {}
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
  std::size_t next_{0}, last_;
};

struct native_compression {  // what main.cpp plugs in (src/main.cpp:112-125)
  template<typename Alloc>
  [[nodiscard]] icd::compressed_t operator()(sid::nat::aimg_t<Alloc> const& image) const {
    return nic::compress(image);
  }
  [[nodiscard]] sid::nat::dimg_t operator()(icd::compressed_t const& c, mrl::dimensions_t const& dim) const {
    return nic::decompress(c, dim);
  }
};

struct filter_rec {
  std::size_t fragment, frame_no;
  fgm::point_t pos;
  std::vector<std::uint8_t> mask;
};
struct filter_recorder {
  std::vector<filter_rec>* calls;
  template<typename Contours, typename Mask>
  void operator()(fgm::fragment const&, std::size_t frag, sid::nat::dimg_t const&, std::size_t no, sid::nat::dimg_t const&,
                  fgm::point_t const& pos, Contours const&, Mask const& mask) {
    filter_rec r{frag, no, pos, {}};
    r.mask.assign(reinterpret_cast<std::uint8_t const*>(mask.data()), reinterpret_cast<std::uint8_t const*>(mask.end()));
    calls->push_back(std::move(r));
  }
};

struct key_rec {
  std::uint32_t region; std::uint16_t x, y; std::uint8_t code[13];
  bool operator<(key_rec const& o) const { return std::memcmp(this, &o, sizeof(*this)) < 0; }
  bool operator==(key_rec const& o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; }
};

struct call_rec {
  std::size_t frame_no;
  std::size_t fragment_frames;  // frames blitted into the current fragment when called
  std::vector<std::uint8_t> median;
  std::vector<key_rec> keys;
  std::vector<std::size_t> weights;
};

struct recorder {
  std::vector<call_rec>* calls;
  bool keep_keys;
  template<typename Frame, typename Image, typename Grid>
  void operator()(fgm::fragment const& frag, Frame const& frame, Image const& median, Grid const& keys) {
    call_rec r;
    r.frame_no = frame.number_;
    r.fragment_frames = frag.frames().size();
    r.median.assign(reinterpret_cast<std::uint8_t const*>(median.data()),
                    reinterpret_cast<std::uint8_t const*>(median.end()));
    if (keep_keys) {
      std::uint32_t ri = 0;
      for (auto& region : keys.regions()) {
        for (auto& [code, pts] : region.points())
          for (auto& p : pts) {
            key_rec k;
            std::memset(&k, 0, sizeof(k));
            k.region = ri; k.x = static_cast<std::uint16_t>(p.x_); k.y = static_cast<std::uint16_t>(p.y_);
            std::memcpy(k.code, code.data(), 13);
            r.keys.push_back(k);
          }
        r.weights.push_back(region.counts()[1]);
        r.weights.push_back(region.counts()[2]);
        ++ri;
      }
      std::sort(r.keys.begin(), r.keys.end());
    }
    calls->push_back(std::move(r));
  }
};

int fail(char const* what, std::size_t a = 0, std::size_t b = 0) {
  std::printf("MISMATCH: %s (%zu, %zu)\n", what, a, b);
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: shim_harness frames.bin W H N batch fill_keys\n"); return 2; }
  std::size_t const w = std::strtoul(argv[2], nullptr, 10), h = std::strtoul(argv[3], nullptr, 10);
  std::size_t const n = std::strtoul(argv[4], nullptr, 10), batch = std::strtoul(argv[5], nullptr, 10);
  bool const fill = std::atoi(argv[6]) != 0;
  bool const gpu_blit = argc > 7 && std::atoi(argv[7]) != 0;
  bool const lean = argc > 7 && std::atoi(argv[7]) == 2;
  bool const filter = argc > 8 && std::atoi(argv[8]) != 0;
  bool const splice = argc > 9 && std::atoi(argv[9]) != 0;
  std::vector<int> devices;
  if (argc > 10)
    for (char const* q = argv[10]; *q;) { devices.push_back(std::atoi(q)); while (*q && *q != ',') ++q; if (*q) ++q; }
  auto data = read_file(argv[1], w * h * n);

  std::vector<call_rec> ref_calls, gpu_calls;
  std::list<fgm::fragment> ref_frags, gpu_frags;

  auto t0 = std::chrono::steady_clock::now();
  {
    frc::collector collector{mrl::dimensions_t{w, h}};
    memory_feed feed{data.data(), w, h, n};
    collector.collect(feed, native_compression{}, recorder{&ref_calls, fill});
    ref_frags = collector.complete();
  }
  auto t1 = std::chrono::steady_clock::now();
  std::unique_ptr<frc_b200::collector> gcol;  // stays alive: filter = 2 runs pass 2 on its resident frames
  {
    frc_b200::options opt;
    opt.batch = batch;
    opt.fill_keys = fill;
    opt.gpu_blit = gpu_blit;
    opt.max_frames = n;
    opt.keep_packed = !lean;
    opt.fetch_medians = !lean;
    opt.devices = devices;
    gcol = std::make_unique<frc_b200::collector>(mrl::dimensions_t{w, h}, opt);
    memory_feed feed{data.data(), w, h, n};
    gcol->collect(feed, native_compression{}, recorder{&gpu_calls, fill});
    gpu_frags = gcol->complete();
  }
  auto t2 = std::chrono::steady_clock::now();

  if (ref_frags.size() != gpu_frags.size()) return fail("fragment count", ref_frags.size(), gpu_frags.size());
  auto gi = gpu_frags.begin();
  std::size_t fi = 0, nframes = 0;
  for (auto& rf : ref_frags) {
    auto& gf = *gi++;
    if (!(rf.zero() == gf.zero())) return fail("fragment zero", fi);
    if (rf.dots().width() != gf.dots().width() || rf.dots().height() != gf.dots().height())
      return fail("fragment dimensions", fi);
    if (std::memcmp(rf.dots().data(), gf.dots().data(), rf.dots().size() * sizeof(fgm::dot_type)) != 0)
      return fail("fragment dots", fi);
    if (rf.frames().size() != gf.frames().size()) return fail("fragment frame count", fi);
    for (std::size_t k = 0; k < rf.frames().size(); ++k) {
      auto& a = rf.frames()[k];
      auto& b = gf.frames()[k];
      if (a.number_ != b.number_) return fail("frame number", fi, k);
      if (!(a.position_ == b.position_)) return fail("frame position", fi, k);
      if (!lean && a.data_.image_ != b.data_.image_) return fail("compressed image", fi, k);
      if (!lean && a.data_.median_ != b.data_.median_) return fail("compressed median", fi, k);
      ++nframes;
    }
    ++fi;
  }
  if (ref_calls.size() != gpu_calls.size()) return fail("callback count", ref_calls.size(), gpu_calls.size());
  for (std::size_t k = 0; k < ref_calls.size(); ++k) {
    auto& a = ref_calls[k];
    auto& b = gpu_calls[k];
    if (a.frame_no != b.frame_no) return fail("callback frame", k);
    if (!gpu_blit && a.fragment_frames != b.fragment_frames) return fail("callback fragment state", k);
    if (!lean && a.median != b.median) return fail("callback median", k);
    if (fill && !(a.keys == b.keys)) return fail("callback keys", k, a.keys.size());
    if (fill && a.weights != b.weights) return fail("callback weight counts", k);
  }
  if (splice) {
    std::vector<fgm::fragment> ra, ga;
    for (auto& f : ref_frags) { ra.push_back(f); ga.push_back(f); }
    auto s0 = std::chrono::steady_clock::now();
    auto rout = fgs::splice(ra.begin(), ra.end());
    auto s1 = std::chrono::steady_clock::now();
    std::size_t tied = 0;
    fgs_b200::options sopt;
    sopt.tied_matches = &tied;
    auto gout = fgs_b200::splice(ga.begin(), ga.end(), sopt);
    auto s2 = std::chrono::steady_clock::now();
    if (rout.size() != gout.size()) return fail("splice: fragment count", rout.size(), gout.size());
    for (std::size_t k = 0; k < rout.size(); ++k) {
      auto& a = rout[k];
      auto& b = gout[k];
      if (!(a.zero() == b.zero())) return fail("splice: zero", k);
      if (a.dots().width() != b.dots().width() || a.dots().height() != b.dots().height()) return fail("splice: dimensions", k);
      if (std::memcmp(a.dots().data(), b.dots().data(), a.dots().size() * sizeof(fgm::dot_type)) != 0) return fail("splice: dots", k);
      if (a.frames().size() != b.frames().size()) return fail("splice: frame count", k);
      for (std::size_t j = 0; j < a.frames().size(); ++j)
        if (a.frames()[j].number_ != b.frames()[j].number_ || !(a.frames()[j].position_ == b.frames()[j].position_))
          return fail("splice: frame record", k, j);
    }
    std::printf("SPLICE IDENTICAL: %zu fragments -> %zu, %zu tied matches; fgs::splice %.1f ms, fgs_b200::splice %.1f ms\n",
                ref_frags.size(), rout.size(), tied, std::chrono::duration<double, std::milli>(s1 - s0).count(),
                std::chrono::duration<double, std::milli>(s2 - s1).count());
  }
  if (filter) {
    std::vector<fgm::fragment> frags;
    for (auto& f : ref_frags) frags.push_back(std::move(f));
    std::vector<filter_rec> rcalls, gcalls;
    auto f0 = std::chrono::steady_clock::now();
    auto rout = fdf::filter(frags, mrl::dimensions_t{w, h}, native_compression{}, filter_recorder{&rcalls});
    auto f1 = std::chrono::steady_clock::now();
    auto gout = fdf_b200::filter(frags, mrl::dimensions_t{w, h}, native_compression{}, filter_recorder{&gcalls});
    auto f2 = std::chrono::steady_clock::now();
    if (rout.size() != gout.size()) return fail("filter: fragment count", rout.size(), gout.size());
    for (std::size_t k = 0; k < rout.size(); ++k) {
      auto& a = rout[k];
      auto& b = gout[k];
      if (!(a.zero() == b.zero())) return fail("filter: zero", k);
      if (a.dots().width() != b.dots().width() || a.dots().height() != b.dots().height()) return fail("filter: dimensions", k);
      if (std::memcmp(a.dots().data(), b.dots().data(), a.dots().size() * sizeof(fgm::dot_type)) != 0) return fail("filter: dots", k);
      if (a.frames().size() != b.frames().size()) return fail("filter: frame count", k);
      for (std::size_t j = 0; j < a.frames().size(); ++j)
        if (a.frames()[j].number_ != b.frames()[j].number_ || !(a.frames()[j].position_ == b.frames()[j].position_))
          return fail("filter: frame record", k, j);
    }
    if (rcalls.size() != gcalls.size()) return fail("filter: callback count", rcalls.size(), gcalls.size());
    for (std::size_t k = 0; k < rcalls.size(); ++k) {
      auto& a = rcalls[k];
      auto& b = gcalls[k];
      if (a.fragment != b.fragment || a.frame_no != b.frame_no || !(a.pos == b.pos)) return fail("filter: callback order", k);
      if (a.mask != b.mask) return fail("filter: fde::mask", k, a.frame_no);
    }
    std::printf("FILTER IDENTICAL: %zu fragments, %zu masks; fdf::filter %.1f ms, fdf_b200::filter %.1f ms\n", rout.size(),
                rcalls.size(), std::chrono::duration<double, std::milli>(f1 - f0).count(),
                std::chrono::duration<double, std::milli>(f2 - f1).count());
    if (gpu_blit) {  // pass 2 in place on the frames the collector left on the device: no decompression, no upload
      fdf_b200::options ropt;
      ropt.callback = false;
      ropt.resident_ctx = gcol->context();
      ropt.resident_numbers = &gcol->resident_numbers();
      std::vector<filter_rec> none;
      auto f3 = std::chrono::steady_clock::now();
      // lean mode: the fragments of the B200 collector itself (no compressed copies in their frame records)
      std::vector<fgm::fragment> own;
      if (lean) for (auto& f : gpu_frags) own.push_back(f);
      auto res = fdf_b200::filter(lean ? own : frags, mrl::dimensions_t{w, h}, native_compression{}, filter_recorder{&none}, ropt);
      auto f4 = std::chrono::steady_clock::now();
      if (res.size() != rout.size()) return fail("resident filter: fragment count", res.size(), rout.size());
      for (std::size_t k = 0; k < rout.size(); ++k)
        if (std::memcmp(rout[k].dots().data(), res[k].dots().data(), rout[k].dots().size() * sizeof(fgm::dot_type)) != 0)
          return fail("resident filter: dots", k);
      std::printf("RESIDENT FILTER IDENTICAL: %zu fragments; fdf_b200::filter on the collector's resident frames %.1f ms\n",
                  res.size(), std::chrono::duration<double, std::milli>(f4 - f3).count());
    }
    for (auto& f : frags) ref_frags.push_back(std::move(f));
  }
  std::printf("IDENTICAL: %zu fragments, %zu frames, %zu callbacks%s%s; reference %.1f ms, frc_b200 %.1f ms\n",
              ref_frags.size(), nframes, ref_calls.size(), fill ? " (keys compared)" : "", gpu_blit ? " (dots from rb_blit_blend)" : "",
              std::chrono::duration<double, std::milli>(t1 - t0).count(),
              std::chrono::duration<double, std::milli>(t2 - t1).count());
  return 0;
}
