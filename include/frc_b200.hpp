// frc_b200.hpp -- a frc::collector-shaped front of the B200 registration path.
//
// Drop-in for the reference's frame collector (src/frc.hpp:27-143): same constructor, same
// collect(feed, comp, cb) / current() / complete() surface, same callback contract
// (cb(fragment, frame, median, keys), src/frc.hpp:119), same result (the fragment list with every
// frame blitted at its accumulated position).  mpb::builder::collect (src/mpb.hpp:52-61) switches by
// naming frc_b200::collector instead of frc::collector; see INTEGRATION.md.
//
// What moves to the GPU: kpe::extractor::extract (src/frc.hpp:90,105) and kpm::match (src/frc.hpp:107)
// for a whole BATCH of frames per call, through the C ABI of remap_b200.h.  What stays here, in the
// reference's own types and order: feed.produce, position_ += off / add_fragment
// (src/frc.hpp:108-115,124-127), fragment::blit with the compressed image + median (:129-135), the
// callback.  This header is compiled in the REFERENCE's translation unit, so it includes the
// reference headers by the names the reference uses; it contains no CUDA.
//
// Differences a caller can observe (both documented in INTEGRATION.md):
//  * frames are pulled from the feeder `batch` at a time before the first callback of the batch;
//  * `keys` handed to the callback is populated only when `fill_keys` is set (the reference's only
//    callback ignores it, src/main.cpp:139-149); populating costs one device round trip per frame.
#pragma once

#include "remap_b200.h"

#include "fgm.hpp"
#include "ifd.hpp"
#include "kpr.hpp"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

namespace frc_b200 {

template<typename Ty>
using allocator_t = all::frame_allocator<Ty>;

using image_type = sid::nat::aimg_t<allocator_t<cpl::nat_cc>>;
using frame_type = ifd::frame<image_type>;

// src/frc.hpp:22-24
static inline constexpr std::size_t grid_horizontal{4};
static inline constexpr std::size_t grid_vertical{2};
static inline constexpr std::size_t grid_overlap{16};

using grid_type = kpr::grid<grid_horizontal, grid_vertical, allocator_t<char>>;

struct options {
  std::size_t batch{2048};  // frames registered per rb_register call
  int device{0};
  // More than one entry: every batch is cut into contiguous frame ranges, one per device, and registered on all of
  // them at once (rb_group, one context per device in this process; pair results gathered on devices[0]).  Choose
  // `batch` as a multiple of what one GPU should see per call (e.g. 2048 x devices).  Not combined with gpu_blit.
  std::vector<int> devices{};
  bool fill_keys{false};    // rebuild the kpr::grid for the callback from rb_keypoints
  // Map assembly on the GPU (rb_blit_blend) instead of fgm::fragment::blit per frame on the host.  All
  // frames then stay resident in the device frame store, so the sequence must fit `max_frames`; a
  // fragment's dots are filled when the fragment ends (next fragment opens, or complete()), i.e. the
  // fragment handed to the callback has its frames but not yet its dots.
  bool gpu_blit{false};
  std::size_t max_frames{0};  // capacity of the frame store with gpu_blit
  // gpu_blit only.  false: the frame records carry no compressed image / median (fgm::packed_data stays empty):
  // the frames and medians live on in the device store, where fdf_b200::filter's resident mode finds them.  The
  // two comp() calls per frame (src/frc.hpp:134) are what the host spends most of its time on otherwise.
  bool keep_packed{true};
  // false: medians are not copied back, the callback receives a zeroed median image (the reference's callback
  // only writes it to a PNG when asked to, src/main.cpp:139-149).
  bool fetch_medians{true};
};

class collector {
  using pixel_alloc_t = allocator_t<cpl::nat_cc>;

public:
  explicit collector(mrl::dimensions_t dimensions, options opt = {})
      : dimensions_{dimensions}
      , opt_{opt} {
    if (opt_.batch < 2) {
      opt_.batch = 2;
    }

    rb_config cfg;
    rb_default_config(&cfg,
                      static_cast<std::uint32_t>(dimensions.width_),
                      static_cast<std::uint32_t>(dimensions.height_),
                      static_cast<std::uint32_t>(opt_.gpu_blit ? std::max(opt_.max_frames, opt_.batch + 1)
                                                               : opt_.batch + 1));
    cfg.grid_w = grid_horizontal;
    cfg.grid_h = grid_vertical;
    cfg.overlap = grid_overlap;
    cfg.weight_switch = 10; // frc::collector::match_config, src/frc.hpp:32
    cfg.region_votes = 3;   // src/frc.hpp:33
    cfg.device = opt_.devices.size() == 1 ? opt_.devices[0] : opt_.device;

    if (opt_.devices.size() > 1) {
      if (opt_.gpu_blit) {
        throw std::runtime_error("frc_b200::collector: options::devices and options::gpu_blit cannot be combined");
      }
      std::vector<std::int32_t> devs(opt_.devices.begin(), opt_.devices.end());
      if (auto rc{rb_group_create(&cfg, devs.data(), devs.size(), &group_)}; rc != RB_OK) {
        std::string msg{group_ != nullptr ? rb_group_last_error(group_) : "no CUDA device"};
        rb_group_destroy(group_);
        group_ = nullptr;
        throw std::runtime_error("frc_b200::collector: " + msg);
      }
    }
    else if (auto rc{rb_create(&cfg, &ctx_)}; rc != RB_OK) {
      std::string msg{ctx_ != nullptr ? rb_last_error(ctx_) : "no CUDA device"};
      rb_destroy(ctx_);
      ctx_ = nullptr;
      throw std::runtime_error("frc_b200::collector: " + msg);
    }

    auto pixels{dimensions.area()};
    stage_ = static_cast<std::uint8_t*>(rb_alloc_host((opt_.batch + 1) * pixels));
    medians_ = static_cast<std::uint8_t*>(rb_alloc_host((opt_.batch + 1) * pixels));
    offsets_ = static_cast<rb_offset*>(rb_alloc_host((opt_.batch + 1) * sizeof(rb_offset)));
    if (stage_ == nullptr || medians_ == nullptr || offsets_ == nullptr) {
      release();
      throw std::runtime_error("frc_b200::collector: pinned host allocation failed");
    }
  }

  collector(collector const&) = delete;
  collector& operator=(collector const&) = delete;

  ~collector() {
    release();
  }

  template<typename Feeder, typename Comp, typename Callback>
  void collect(Feeder&& feed, Comp&& comp, Callback&& cb) requires(
      ifd::feeder<std::decay_t<Feeder>, pixel_alloc_t>&&
          icd::compressor<std::decay_t<Comp>, pixel_alloc_t>) {
    if (!feed.has_more()) {
      return;
    }

    auto pixels{dimensions_.area()};
    all::memory_stack<cpl::nat_cc> memory{};
    bool first_batch{true};

    while (feed.has_more()) {
      // one pool per batch; the previous batch's pool stays alive (two-pool swing, src/all.hpp:151-205)
      all::memory_swing swing{memory};
      pixel_alloc_t alloc{swing};

      std::vector<frame_type> frames;
      frames.reserve(opt_.batch);
      while (frames.size() < opt_.batch && feed.has_more()) {
        frames.push_back(feed.produce(alloc));
      }

      // slot 0 of the device frame store = last frame of the previous batch (the `previous` grid
      // of src/frc.hpp:64-65); slots 1.. = this batch.  With gpu_blit every frame keeps its own slot
      // (its index in the sequence) and the previous frame is simply still there.
      auto base{first_batch ? std::size_t{0} : std::size_t{1}};
      if (!first_batch && !opt_.gpu_blit) {
        std::memcpy(stage_, carry_.data(), pixels);
      }
      auto skip{opt_.gpu_blit ? base : std::size_t{0}};  // staged frames that need no upload
      for (std::size_t i{0}; i < frames.size(); ++i) {
        std::memcpy(stage_ + (base + i) * pixels, frames[i].image_.data(), pixels);
      }

      auto total{base + frames.size()};
      auto slot0{opt_.gpu_blit ? count_ - base : std::size_t{0}};  // store slot of staged frame 0
      if (opt_.gpu_blit && count_ + frames.size() > std::max(opt_.max_frames, opt_.batch + 1)) {
        throw std::runtime_error("frc_b200::collector: sequence longer than options::max_frames (gpu_blit)");
      }
      if (group_ != nullptr) {
        check_group(rb_group_register_host(group_, stage_, total, offsets_));
        if (opt_.fetch_medians) {
          check_group(rb_group_fetch_medians(group_, 0, total, medians_));
        }
      }
      else {
        check(rb_upload(ctx_, stage_ + skip * pixels, slot0 + skip, total - skip));
        check(rb_register(ctx_, slot0, total, offsets_, opt_.fetch_medians ? medians_ : nullptr));
      }

      for (std::size_t i{0}; i < frames.size(); ++i) {
        auto& frame{frames[i]};
        auto& dim{frame.image_.dimensions()};

        image_type median{dim, alloc};
        if (opt_.fetch_medians) {
          std::memcpy(median.data(), medians_ + (base + i) * pixels, pixels);
        }

        bool init{first_batch && i == 0};
        if (init) {
          add_fragment(dim); // process_init, src/frc.hpp:83-95
        }
        else {
          auto const& off{offsets_[base + i - 1]};
          if ((off.flags & RB_OFFSET_VALID) != 0) {
            position_.x_ += off.dx; // src/frc.hpp:108-111
            position_.y_ += off.dy;
          }
          else {
            add_fragment(dim); // src/frc.hpp:113-115
          }
        }

        auto& [no, image]{frame};
        if (opt_.gpu_blit) { // same record as fragment::blit leaves (src/fgm.hpp:96), the dots come later
          pending_.emplace_back(no, position_,
                                opt_.keep_packed ? fgm::packed_data{comp(image), comp(median)} : fgm::packed_data{});
          slots_.push_back(static_cast<std::uint32_t>(count_));
          numbers_.push_back(no);
        }
        else {
          current_->blit(position_, image, {comp(image), comp(median)}, no); // src/frc.hpp:129-135
        }
        ++count_;

        if (!init) { // the reference does not call back for the first frame (src/frc.hpp:83-95)
          grid_type keys{allocator_t<char>{alloc}};
          if (opt_.fill_keys) {
            fill(keys, slot0 + base + i);  // the frame's slot in the device store (== base + i unless gpu_blit)
          }
          cb(*current_, frame, median, keys);
        }
      }

      auto last{frames.back().image_.data()};
      carry_.assign(reinterpret_cast<std::uint8_t const*>(last),
                    reinterpret_cast<std::uint8_t const*>(last) + pixels);
      first_batch = false;
    }
  }

  [[nodiscard]] inline fgm::fragment const& current() const noexcept {
    return *current_;
  }

  // gpu_blit only: the device context that still holds every collected frame and its median, and the frame
  // number stored in each slot.  fdf_b200::filter can run pass 2 on them in place (fdf_b200::options::resident_*),
  // without decompressing and uploading anything.  Valid while this collector lives.
  [[nodiscard]] inline rb_ctx* context() const noexcept {
    return ctx_;
  }
  [[nodiscard]] inline std::vector<std::size_t> const& resident_numbers() const noexcept {
    return numbers_;
  }

  [[nodiscard]] inline std::list<fgm::fragment> complete() {
    finish_fragment();
    for (auto& fragment : fragments_) {
      fragment.normalize();
    }

    return std::move(fragments_);
  }

private:
  inline void add_fragment(mrl::dimensions_t dimension) {
    finish_fragment();
    current_ = &fragments_.emplace_back(dimension);
    position_.x_ = position_.y_ = 0;
  }

  // gpu_blit: the fragment that just ended gets its dots from rb_blit_blend.  The map geometry is what
  // fragment::ensure / extend would have produced frame by frame (src/fgm.hpp:190-233): the map grows in
  // whole frame-sized steps and zero_ moves with it.
  void finish_fragment() {
    if (!opt_.gpu_blit || current_ == nullptr || pending_.empty()) {
      return;
    }

    std::int64_t zero[2]{0, 0};
    std::int64_t dim[2]{static_cast<std::int64_t>(dimensions_.width_), static_cast<std::int64_t>(dimensions_.height_)};
    std::int64_t const step[2]{dim[0], dim[1]};
    auto round{[](std::int64_t change, std::int64_t st) { // fragment::get_step, src/fgm.hpp:228-233
      auto rest{change % st};
      return (change - rest) + (rest != 0 ? st : 0);
    }};
    for (auto const& f : pending_) {
      std::int64_t const p[2]{f.position_.x_, f.position_.y_};
      for (int k{0}; k < 2; ++k) {
        std::int64_t lo{0}, hi{0};
        if (p[k] < zero[k]) {
          lo = round(zero[k] - p[k], step[k]);
        }
        if (auto required{p[k] + step[k]}; required > 0 && required > zero[k] + dim[k]) {
          hi = round(required - (zero[k] + dim[k]), step[k]);
        }
        zero[k] -= lo;
        dim[k] += lo + hi;
      }
    }

    std::vector<rb_placement> places(pending_.size());
    for (std::size_t i{0}; i < pending_.size(); ++i) {
      places[i].frame = slots_[i];
      places[i].x = static_cast<std::int32_t>(pending_[i].position_.x_ - zero[0]);
      places[i].y = static_cast<std::int32_t>(pending_[i].position_.y_ - zero[1]);
    }

    fgm::fragment::matrix_type dots{
        mrl::dimensions_t{static_cast<std::size_t>(dim[0]), static_cast<std::size_t>(dim[1])}};
    check(rb_blit_blend(ctx_,
                        places.data(),
                        places.size(),
                        static_cast<std::uint32_t>(dim[0]),
                        static_cast<std::uint32_t>(dim[1]),
                        reinterpret_cast<std::uint16_t*>(dots.data()),
                        nullptr,
                        nullptr));

    *current_ = fgm::fragment{std::move(dots),
                              dimensions_,
                              fgm::point_t{static_cast<std::int32_t>(zero[0]), static_cast<std::int32_t>(zero[1])},
                              std::move(pending_)};
    pending_.clear();
    slots_.clear();
  }

  // kpr::grid as kpe::extractor would have filled it (src/kpe.hpp:225-229,301-303)
  void fill(grid_type& keys, std::size_t slot) {
    kps_.resize(dimensions_.area());
    std::size_t count{0};
    if (group_ != nullptr) { // the member that owns the frame, and the frame's slot there
      std::size_t member{0}, local{0};
      check_group(rb_group_locate(group_, slot, &member, &local));
      auto ctx{rb_group_context(group_, member)};
      if (rb_keypoints(ctx, local, kps_.data(), kps_.size(), &count) != RB_OK) {
        throw std::runtime_error(std::string{"frc_b200::collector: "} + rb_last_error(ctx));
      }
    }
    else {
      check(rb_keypoints(ctx_, slot, kps_.data(), kps_.size(), &count));
    }
    for (std::size_t k{0}; k < count; ++k) {
      auto const& kp{kps_[k]};
      kpr::code code;
      std::memcpy(code.data(), kp.code, sizeof(kp.code));
      for (std::size_t r{0}; r < grid_type::region_count; ++r) {
        if ((kp.region_mask >> r) & 1u) {
          keys[r].add(code, mrl::point_t{kp.x, kp.y});
        }
      }
    }
  }

  void check(int rc) {
    if (rc != RB_OK) {
      throw std::runtime_error(std::string{"frc_b200::collector: "} + rb_last_error(ctx_));
    }
  }

  void check_group(int rc) {
    if (rc != RB_OK) {
      throw std::runtime_error(std::string{"frc_b200::collector: "} + rb_group_last_error(group_));
    }
  }

  void release() noexcept {
    rb_free_host(stage_);
    rb_free_host(medians_);
    rb_free_host(offsets_);
    stage_ = medians_ = nullptr;
    offsets_ = nullptr;
    rb_destroy(ctx_);
    ctx_ = nullptr;
    rb_group_destroy(group_);
    group_ = nullptr;
  }

private:
  mrl::dimensions_t dimensions_;
  options opt_;

  rb_ctx* ctx_{nullptr};
  rb_group* group_{nullptr};  // options::devices with more than one entry
  std::uint8_t* stage_{nullptr};
  std::uint8_t* medians_{nullptr};
  rb_offset* offsets_{nullptr};
  std::vector<std::uint8_t> carry_;
  std::vector<rb_keypoint> kps_;
  std::vector<std::size_t> numbers_;  // gpu_blit: frame number held by each slot of the device store

  std::size_t count_{0};               // frames collected so far (== store slot with gpu_blit)
  std::vector<fgm::frame> pending_;    // gpu_blit: frames of the current fragment, dots still to come
  std::vector<std::uint32_t> slots_;

  fgm::point_t position_{};

  std::list<fgm::fragment> fragments_;
  fgm::fragment* current_{nullptr};
};

} // namespace frc_b200
