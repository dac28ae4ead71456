// fdf_b200.hpp -- a fdf::filter-shaped front of the B200 pass-2 foreground filter.
//
// Drop-in for the reference's fdf::filter(fragments, frame_dim, comp, cb) (src/fdf.hpp:77-89), which
// mpb::builder::filter calls (src/mpb.hpp:71-77): same arguments, same result -- one fragment per input
// fragment, of the background's dimensions and zero, holding every frame blitted where its foreground mask
// is 0 (src/fdf.hpp:51-72) -- and the callback is invoked once per frame with the same arguments.
//
// What moves to the GPU, per fragment, through rb_filter_fragment (remap_b200.h): the background
// (fragment::blend, src/fdf.hpp:21-34), fde::extractor::extract and fde::mask for every frame
// (src/fdf.hpp:62-63), the masked fragment::blit (src/fdf.hpp:64).  What stays here in the reference's own
// types: decompression of the stored frames and medians with the caller's decompressor (src/fdf.hpp:59-60).
// This header is compiled in the REFERENCE's translation unit and contains no CUDA.
//
// Differences a caller can observe:
//  * the callback runs after its fragment is complete, so `result` is the finished fragment for every
//    frame of it (the reference passes the fragment as blitted so far);
//  * `foreground` (the contour list) is passed EMPTY: the device produces the mask and the number of kept
//    contours, not the contours' edge lists.  The reference's callback only draws the mask
//    (src/main.cpp:151-170).  options::callback_masks = false skips the per-frame mask download as well.
#pragma once

#include "remap_b200.h"

#include "fde.hpp"
#include "fdf.hpp"
#include "fgm.hpp"

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace fdf_b200 {

struct options {
  int device{0};
  bool callback{true};        // false: no per-frame callback at all (then nothing is decompressed either, see below)
  bool callback_masks{true};  // download every frame's fde::mask for the callback
  // Resident mode: a context that still holds the frames AND their medians (frc_b200::collector with
  // options::gpu_blit: collector.context(), collector.resident_numbers()).  Pass 2 then runs on the device
  // store in place: no decompression, no upload; with callback = false the host only receives the dots.
  rb_ctx* resident_ctx{nullptr};
  std::vector<std::size_t> const* resident_numbers{nullptr};  // (*resident_numbers)[slot] = frame number
};

namespace details {
  struct context {  // one rb_ctx per (frame size, capacity); rebuilt when a larger fragment arrives
    rb_ctx* ctx{nullptr};
    std::size_t capacity{0};
    ~context() { rb_destroy(ctx); }
    void ensure(mrl::dimensions_t const& dim, std::size_t frames, int device) {
      if (ctx != nullptr && frames <= capacity) return;
      rb_destroy(ctx);
      ctx = nullptr;
      rb_config cfg;
      rb_default_config(&cfg, static_cast<std::uint32_t>(dim.width_), static_cast<std::uint32_t>(dim.height_),
                        static_cast<std::uint32_t>(frames < 2 ? 2 : frames));
      cfg.device = device;
      if (auto rc{rb_create(&cfg, &ctx)}; rc != RB_OK) {
        std::string msg{ctx != nullptr ? rb_last_error(ctx) : "no CUDA device"};
        rb_destroy(ctx);
        ctx = nullptr;
        throw std::runtime_error("fdf_b200::filter: " + msg);
      }
      capacity = frames < 2 ? 2 : frames;
    }
    void check(int rc) const {
      if (rc != RB_OK) throw std::runtime_error(std::string{"fdf_b200::filter: "} + rb_last_error(ctx));
    }
  };
}  // namespace details

template<typename Comp, typename Callback>
[[nodiscard]] std::vector<fgm::fragment> filter(std::vector<fgm::fragment> const& fragments,
                                                mrl::dimensions_t const& frame_dim,
                                                Comp&& comp,
                                                Callback&& cb,
                                                options const& opt = options{})
    requires(icd::decompressor<std::decay_t<Comp>, std::allocator<cpl::nat_cc>>) {
  std::vector<fgm::fragment> results{};
  details::context dev;
  auto const pixels{frame_dim.area()};
  bool const resident{opt.resident_ctx != nullptr && opt.resident_numbers != nullptr};
  std::unordered_map<std::size_t, std::uint32_t> slot_of;
  if (resident) {
    for (std::size_t s{0}; s < opt.resident_numbers->size(); ++s) slot_of[(*opt.resident_numbers)[s]] = static_cast<std::uint32_t>(s);
  }

  std::size_t i{0};
  for (auto& fragment : fragments) {
    auto const& frames{fragment.frames()};
    auto const n{frames.size()};
    auto const map_dim{fragment.dots().dimensions()};
    auto const zero{fragment.zero()};

    rb_ctx* ctx{opt.resident_ctx};
    if (!resident) {
      dev.ensure(frame_dim, n, opt.device);
      ctx = dev.ctx;
    }
    auto check{[&](int rc) {
      if (rc != RB_OK) throw std::runtime_error(std::string{"fdf_b200::filter: "} + rb_last_error(ctx));
    }};
    std::vector<sid::nat::dimg_t> images, medians;  // kept for the callback, as the reference hands them over
    std::vector<rb_placement> places(n);
    std::uint8_t* stage{nullptr};
    bool const unpack{!resident || opt.callback};  // somebody needs the frames on the host
    if (unpack) {
      images.reserve(n);
      medians.reserve(n);
    }
    if (!resident) {
      stage = static_cast<std::uint8_t*>(rb_alloc_host(2 * n * pixels + 16));
      if (stage == nullptr) throw std::runtime_error("fdf_b200::filter: rb_alloc_host failed");
    }
    auto mstage{stage + n * pixels};
    std::size_t k{0};
    for (auto& [no, pos, data] : frames) {  // src/fdf.hpp:58-60
      std::uint32_t slot{static_cast<std::uint32_t>(k)};
      if (unpack) {
        images.push_back(comp(data.image_, frame_dim));
        medians.push_back(comp(data.median_, frame_dim));
      }
      if (resident) {
        auto it{slot_of.find(no)};
        if (it == slot_of.end()) throw std::runtime_error("fdf_b200::filter: frame not in the resident store");
        slot = it->second;
      }
      else {
        std::memcpy(stage + k * pixels, images.back().data(), pixels);
        std::memcpy(mstage + k * pixels, medians.back().data(), pixels);
      }
      places[k] = {slot, pos.x_ - zero.x_, pos.y_ - zero.y_};  // src/fgm.hpp:177
      ++k;
    }
    fgm::fragment::matrix_type dots{map_dim};
    bool const want_masks{opt.callback && opt.callback_masks};
    std::vector<std::uint8_t> masks(want_masks ? n * pixels : 0);
    if (!resident && n != 0) {
      check(rb_upload(ctx, stage, 0, n));
      check(rb_upload_medians(ctx, mstage, 0, n));
    }
    check(rb_filter_fragment(ctx, places.data(), n, static_cast<std::uint32_t>(map_dim.width_),
                             static_cast<std::uint32_t>(map_dim.height_), nullptr,
                             reinterpret_cast<std::uint16_t*>(dots.data()), nullptr, nullptr,
                             want_masks ? masks.data() : nullptr, nullptr));
    rb_free_host(stage);

    std::vector<fgm::frame> placed;  // what fragment::blit(pos, image, mask, no) records (src/fgm.hpp:84)
    placed.reserve(n);
    for (auto& f : frames) placed.emplace_back(f.number_, f.position_);
    auto& result{results.emplace_back(std::move(dots), mrl::dimensions_t{1, 1}, zero, std::move(placed))};

    k = 0;
    for (auto& [no, pos, data] : frames) {  // src/fdf.hpp:66
      if (!opt.callback) break;
      sid::mon::dimg_t mask{frame_dim};
      if (want_masks) std::memcpy(mask.data(), masks.data() + k * pixels, pixels);
      fdf::contours_t foreground{};
      cb(result, i, images[k], no, medians[k], pos, foreground, mask);
      ++k;
    }
    ++i;
  }
  return results;
}

}  // namespace fdf_b200
