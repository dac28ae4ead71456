// mpb_b200.hpp -- the reference's map builder (src/mpb.hpp:17-102) with its three per-frame stages on the B200
// and the frames RESIDENT on the device between them.
//
// mpb::builder::build (src/mpb.hpp:28-41) is five calls: get_window (aws::scan), collect (frc::collector), splice
// (fgs::splice), filter (fdf::filter), clean (arf::filter).  Substituting frc_b200::collector, fgs_b200::splice and
// fdf_b200::filter into that header one name at a time gives identical maps (checked
// by oracle/_ref/pipeline_harness), but every stage then still compresses, decompresses and re-uploads frames
// the way the reference's stages hand them to each other.  This builder keeps the same Adapter contract
// (src/main.cpp:194-244: get_feed, get_feed(crop), get_compression, get_screen_dimensions,
// get_artifact_filter_dev, artifact_filter_size, get_callbacks) and the same stage order, and wires the shims the
// fast way: one collector that outlives collect() (gpu_blit, no compressed copies), splice on its fragments,
// pass 2 in place on the frames it left in HBM.  aws::scan and arf::filter are the reference's own code.
//
// Callbacks: as mpb::builder, except that fdf's per-frame callback is not invoked (options::filter_callback) and
// the frc callback receives a zeroed median unless options::fetch_medians.
#pragma once

#include "fdf_b200.hpp"
#include "fgs_b200.hpp"
#include "frc_b200.hpp"

#include "arf.hpp"
#include "aws.hpp"

#include <execution>
#include <memory>

namespace mpb_b200 {

struct options {
  int device{0};
  std::size_t max_frames{1u << 16};  // capacity of the resident frame store
  std::size_t batch{512};            // frames per registration call (pinned staging is sized by it)
  bool fetch_medians{false};
  bool filter_callback{false};
};

template<typename Adapter>
class fast_builder {
public:
  using adapter_type = Adapter;

  explicit fast_builder(adapter_type const& adapter, options opt = {})
      : adapter_{adapter}
      , opt_{opt} {
  }

  [[nodiscard]] std::vector<sid::nat::dimg_t> build() {
    auto window{aws::scan(adapter_.get_feed(), adapter_.get_screen_dimensions(), cb())};  // src/mpb.hpp:45-50
    cb()(window);
    if (!window) {
      return {};
    }
    auto dimensions{window->bounds().dimensions()};
    auto feed{adapter_.get_feed(window->margins())};

    frc_b200::options copt;
    copt.device = opt_.device;
    copt.batch = std::min(opt_.batch, std::max<std::size_t>(opt_.max_frames, 2));
    copt.gpu_blit = true;
    copt.max_frames = opt_.max_frames;
    copt.keep_packed = false;
    copt.fetch_medians = opt_.fetch_medians;
    frc_b200::collector collector{dimensions, copt};  // lives until the maps are done: its HBM store is pass 2's input
    collector.collect(feed, adapter_.get_compression(), cb());
    auto fragments{collector.complete()};
    cb()("frc", fragments);

    fgs_b200::options sopt;
    sopt.device = opt_.device;
    auto spliced{fgs_b200::splice(fragments.begin(), fragments.end(), sopt)};
    cb()("spl", spliced);

    fdf_b200::options fopt;
    fopt.device = opt_.device;
    fopt.callback = opt_.filter_callback;
    fopt.resident_ctx = collector.context();
    fopt.resident_numbers = &collector.resident_numbers();
    auto filtered{fdf_b200::filter(spliced, dimensions, adapter_.get_compression(), cb(), fopt)};
    cb()("fdf", filtered);

    std::vector<sid::nat::dimg_t> result{filtered.size()};  // clean, src/mpb.hpp:79-94
    std::transform(std::execution::par, filtered.begin(), filtered.end(), result.begin(),
                   [this, dev = adapter_.get_artifact_filter_dev()](auto& fragment) {
                     return arf::filter(fragment, cb(), dev, typename adapter_type::artifact_filter_size{});
                   });
    return result;
  }

private:
  [[nodiscard]] inline auto& cb() noexcept {
    return adapter_.get_callbacks();
  }

  adapter_type adapter_;
  options opt_;
};

}  // namespace mpb_b200
