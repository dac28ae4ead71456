// oracle/pipeline_harness.cpp -- the reference's WHOLE pipeline, once as it is and once with the three B200 shims.
//
// TEST INFRASTRUCTURE ONLY (built by oracle/build_ref.py into oracle/_ref/pipeline_harness, linked against
// remap_b200/libremap_b200.so).  This file is ours.  It runs
//   mpb::builder<adapter>::build()        (src/mpb.hpp:28-41: aws::scan -> frc -> fgs -> fdf -> arf)  -- the reference
//   mpb_subst::builder<adapter>::build()  -- the SAME header after the three substitutions of INTEGRATION.md
//                                            (frc::collector -> frc_b200::collector, fgs::splice -> fgs_b200::splice,
//                                            fdf::filter -> fdf_b200::filter), made mechanically at build time
//                                            in a temp dir by oracle/build_ref.py; nothing else differs
// on the same screen-sized frames through an in-memory adapter shaped like build_adapter (src/main.cpp:194-244),
// and compares the final maps byte for byte, plus the fragments reported to the "frc" / "spl" / "fdf" callbacks.
//
// mode fast additionally runs mpb_b200::fast_builder (include/mpb_b200.hpp: the same stages wired so that the
// frames stay in HBM between them) and compares its stage hand-overs and final maps with the reference's too.
//
// usage: pipeline_harness <frames.bin> screenW screenH N [mode: both|ref|b200|fast|fast-time]       exit 0 = identical
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
#include <optional>
#include <string>
#include <vector>

#include "mpb.hpp"
#include "nic.hpp"

#include "mpb_subst.hpp"  // generated at build time: mpb.hpp with the three names substituted (namespace mpb_subst)
#include "mpb_b200.hpp"   // include/: mpb_b200::fast_builder (frames resident between the stages)

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

class memory_feed {  // models file_feed (src/main.cpp:16-52): whole screens, optionally cropped to the action window
public:
  memory_feed(std::uint8_t const* data, mrl::dimensions_t dim, std::size_t n, std::optional<mrl::region_t> crop = {})
      : data_{data}, dim_{dim}, last_{n}, crop_{crop} {}
  [[nodiscard]] bool has_more() const noexcept { return next_ < last_; }
  template<typename Alloc>
  [[nodiscard]] auto produce(Alloc const& alloc) {
    using image_type = sid::nat::aimg_t<Alloc>;
    image_type img{dim_, alloc};
    std::memcpy(img.data(), data_ + next_ * dim_.area(), dim_.area());
    auto no{next_++};
    return ifd::frame<image_type>{no, crop_ ? img.crop(*crop_) : img};
  }
private:
  std::uint8_t const* data_;
  mrl::dimensions_t dim_;
  std::size_t next_{0}, last_;
  std::optional<mrl::region_t> crop_;
};

struct native_compression {  // src/main.cpp:112-125
  template<typename Alloc>
  [[nodiscard]] icd::compressed_t operator()(sid::nat::aimg_t<Alloc> const& image) const { return nic::compress(image); }
  [[nodiscard]] sid::nat::dimg_t operator()(icd::compressed_t const& c, mrl::dimensions_t const& dim) const {
    return nic::decompress(c, dim);
  }
};

struct stage_record {
  std::string tag;
  double at_ms{0};  // when the stage handed its fragments over, since the first callback of the run
  std::vector<std::vector<std::uint16_t>> dots;
  std::vector<std::array<std::int64_t, 4>> geom;  // width, height, zero x, zero y
};

struct callbacks {  // every signature the stages call (src/main.cpp:127-192); records the fragment hand-overs
  std::vector<stage_record>* log{nullptr};
  std::optional<mrl::region_t> window;
  std::chrono::steady_clock::time_point t0{std::chrono::steady_clock::now()};
  double window_ms{0};
  void operator()(aws::frame_type const&, aws::heatmap_type const&, aws::contour_type const&, std::size_t) noexcept {}
  template<typename Frame, typename Image, typename Grid>
  void operator()(fgm::fragment const&, Frame const&, Image const&, Grid const&) noexcept {}
  template<typename Contours, typename Mask>
  void operator()(fgm::fragment const&, std::size_t, sid::nat::dimg_t const&, std::size_t, sid::nat::dimg_t const&,
                  fgm::point_t const&, Contours const&, Mask const&) noexcept {}
  void operator()(sid::nat::dimg_t const&, mrl::matrix<float> const&) const noexcept {}
  void operator()(std::optional<aws::window_info> const& w) noexcept {
    if (w) window = w->bounds();
    window_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  template<typename Container>
  void operator()(std::string const& tag, Container const& fragments) {
    if (!log) return;
    stage_record r;
    r.tag = tag;
    r.at_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (auto& f : fragments) {
      auto p{reinterpret_cast<std::uint16_t const*>(f.dots().data())};
      r.dots.emplace_back(p, p + f.dots().size() * fgm::depth);
      r.geom.push_back({static_cast<std::int64_t>(f.dots().width()), static_cast<std::int64_t>(f.dots().height()),
                        f.zero().x_, f.zero().y_});
    }
    log->push_back(std::move(r));
  }
};

class adapter {
public:
  using callbacks_type = callbacks;
  using feed_type = memory_feed;
  using artifact_filter_size = arf::filter_size<15>;  // src/main.cpp:201
  adapter(std::uint8_t const* data, mrl::dimensions_t screen, std::size_t n, std::vector<stage_record>* log)
      : data_{data}, screen_{screen}, n_{n} { callbacks_.log = log; }
  [[nodiscard]] feed_type get_feed() const { return {data_, screen_, n_}; }
  [[nodiscard]] feed_type get_feed(mrl::region_t crop) const { return {data_, screen_, n_, crop}; }
  [[nodiscard]] native_compression get_compression() const { return {}; }
  [[nodiscard]] mrl::dimensions_t get_screen_dimensions() const noexcept { return screen_; }
  [[nodiscard]] float get_artifact_filter_dev() const noexcept { return 2.0f; }  // src/main.cpp:200
  [[nodiscard]] callbacks_type& get_callbacks() noexcept { return callbacks_; }
private:
  std::uint8_t const* data_;
  mrl::dimensions_t screen_;
  std::size_t n_;
  callbacks_type callbacks_{};
};

int fail(char const* what, std::size_t a = 0, std::size_t b = 0) {
  std::printf("MISMATCH: %s (%zu, %zu)\n", what, a, b);
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: pipeline_harness frames.bin screenW screenH N [both|ref|b200]\n"); return 2; }
  std::size_t const w = std::strtoul(argv[2], nullptr, 10), h = std::strtoul(argv[3], nullptr, 10);
  std::size_t const n = std::strtoul(argv[4], nullptr, 10);
  std::string const mode = argc > 5 ? argv[5] : "both";
  auto data = read_file(argv[1], w * h * n);
  mrl::dimensions_t const screen{w, h};

  std::vector<stage_record> rlog, glog;
  std::vector<sid::nat::dimg_t> rmaps, gmaps;
  double rms = 0, gms = 0;
  std::optional<mrl::region_t> rwin, gwin;
  bool const fast_only = mode == "fast-time";  // timing run: nothing else in the process, no comparison
  if (mode != "b200" && !fast_only) {
    auto t0 = std::chrono::steady_clock::now();
    mpb::builder builder{adapter{data.data(), screen, n, &rlog}};
    rmaps = builder.build();
    rms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  if (mode != "ref" && mode != "fast") {
    auto t0 = std::chrono::steady_clock::now();
    mpb_subst::builder builder{adapter{data.data(), screen, n, &glog}};
    gmaps = builder.build();
    gms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  auto describe = [](std::vector<stage_record> const& log) {
    for (auto& r : log) std::printf("  %s: %zu fragment(s) @ %.0f ms", r.tag.c_str(), r.dots.size(), r.at_ms);
    std::printf("\n");
  };
  if (mode == "both") {
    if (rlog.size() != glog.size()) return fail("stage count", rlog.size(), glog.size());
    for (std::size_t s = 0; s < rlog.size(); ++s) {
      if (rlog[s].tag != glog[s].tag) return fail("stage order", s);
      if (rlog[s].dots.size() != glog[s].dots.size()) return fail(rlog[s].tag.c_str(), rlog[s].dots.size(), glog[s].dots.size());
      for (std::size_t k = 0; k < rlog[s].dots.size(); ++k) {
        if (rlog[s].geom[k] != glog[s].geom[k]) return fail((rlog[s].tag + " geometry").c_str(), k);
        if (rlog[s].dots[k] != glog[s].dots[k]) return fail((rlog[s].tag + " dots").c_str(), k);
      }
    }
    if (rmaps.size() != gmaps.size()) return fail("map count", rmaps.size(), gmaps.size());
    for (std::size_t k = 0; k < rmaps.size(); ++k) {
      if (rmaps[k].width() != gmaps[k].width() || rmaps[k].height() != gmaps[k].height()) return fail("map size", k);
      if (std::memcmp(rmaps[k].data(), gmaps[k].data(), rmaps[k].size()) != 0) return fail("map pixels", k);
    }
  }
  if (mode == "fast" || fast_only) {
    std::vector<stage_record> flog;
    auto t0 = std::chrono::steady_clock::now();
    mpb_b200::options fo;
    fo.max_frames = n;
    mpb_b200::fast_builder builder{adapter{data.data(), screen, n, &flog}, fo};
    auto fmaps = builder.build();
    double fms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (fast_only) {
      std::printf("FAST PIPELINE (timing only): mpb_b200::fast_builder %.1f ms, %zu final map(s)\n", fms, fmaps.size());
      describe(flog);
      return 0;
    }
    if (rlog.size() != flog.size()) return fail("fast: stage count", rlog.size(), flog.size());
    for (std::size_t s = 0; s < rlog.size(); ++s) {
      if (rlog[s].dots.size() != flog[s].dots.size()) return fail(("fast " + rlog[s].tag).c_str(), rlog[s].dots.size(), flog[s].dots.size());
      for (std::size_t k = 0; k < rlog[s].dots.size(); ++k)
        if (rlog[s].geom[k] != flog[s].geom[k] || rlog[s].dots[k] != flog[s].dots[k]) return fail(("fast " + rlog[s].tag + " dots").c_str(), k);
    }
    if (rmaps.size() != fmaps.size()) return fail("fast: map count", rmaps.size(), fmaps.size());
    for (std::size_t k = 0; k < rmaps.size(); ++k)
      if (rmaps[k].width() != fmaps[k].width() || rmaps[k].height() != fmaps[k].height() ||
          std::memcmp(rmaps[k].data(), fmaps[k].data(), rmaps[k].size()) != 0)
        return fail("fast: map pixels", k);
    std::printf("FAST PIPELINE IDENTICAL: mpb_b200::fast_builder %.1f ms\n", fms);
    describe(flog);
  }
  auto& maps = mode == "b200" ? gmaps : rmaps;
  std::printf("%s: %zu frames of %zux%zu, %zu final map(s)", mode == "both" ? "PIPELINE IDENTICAL" : "PIPELINE", n, w, h, maps.size());
  for (auto& m : maps) std::printf(" %zux%zu", static_cast<std::size_t>(m.width()), static_cast<std::size_t>(m.height()));
  std::printf("; mpb::builder %.1f ms, with the B200 shims %.1f ms\n", rms, gms);
  describe(mode == "b200" ? glog : rlog);
  return 0;
}
