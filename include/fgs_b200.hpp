// fgs_b200.hpp -- a fgs::splice-shaped front of the B200 fragment splicer.
//
// Drop-in for the reference's fgs::splice(first, last) (src/fgs.hpp:187-213), which mpb::builder::splice calls
// (src/mpb.hpp:63-69): same arguments, same result -- the fragments merged greedily, best vote first, until no
// two of them match.
//
// What moves to the GPU (remap_b200.h): fgs::details::extract_single for every fragment and every merged
// fragment (rb_snippet_create: fragment.blend() + kpe with a 1 x 1 grid over the whole map, src/fgs.hpp:80-89),
// every cellular kpm::match (rb_snippet_match, src/kpm.hpp:371-393; src/fgs.hpp:119-134) AND the merge itself:
// fgm::fragment::blit(pos, fragment&&) (src/fgm.hpp:99-113) adds one resident dot map into another on the device
// (rb_snippet_merge) and re-extracts there -- no 32 byte / pixel map crosses PCIe between the merges; a merged
// fragment's dots are read back once, at the end.  What stays here: the bookkeeping of who matched whom, the map
// geometry (fragment::ensure / extend, src/fgm.hpp:190-233, replayed on the four integers it needs) and the frame
// records (src/fgm.hpp:108-112, normalize :137-143).  The selection rule is the
// reference's (src/fgs.hpp:142-168): among the edges recorded from the earlier snippet of each matching pair,
// the first one with the largest vote count, walking the snippets in list order and each snippet's edges in
// the order they were found; the merged fragment goes to the FRONT of the list and is matched against all
// the others (src/fgs.hpp:170-183).  This header is compiled in the REFERENCE's translation unit and
// contains no CUDA.
#pragma once

#include "remap_b200.h"

#include "fgm.hpp"
#include "kpm.hpp"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iterator>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

namespace fgs_b200 {

struct options {
  int device{0};
  std::uint8_t cell_width{15}, cell_height{15};  // src/fgs.hpp:121
  std::size_t* tied_matches{nullptr};            // out: matches whose best offset shared its vote count (see remap_b200.h)
};

namespace details {

  struct edge {
    std::size_t other;  // id of the later snippet of the pair
    kpm::vote vote;
  };

  struct node {
    std::size_t id;
    fgm::fragment fragment;  // the caller's fragment, untouched, until the node takes part in a merge
    bool merged{false};      // true: `fragment` is empty; dots live in the snippet, the rest below
    mrl::dimensions_t step{}, dim{};
    fgm::point_t zero{};
    std::vector<fgm::frame> frames;
    rb_snippet* snippet{nullptr};
    std::vector<edge> edges;  // "primary" edges only: this node was the head of the match (src/fgs.hpp:68-71)
  };

  inline void fail(rb_snippet* s, char const* what) {
    throw std::runtime_error(std::string{"fgs_b200::splice: "} + what + ": " + (s != nullptr ? rb_snippet_last_error(s) : "no snippet"));
  }

  inline void extract(node& n, options const& opt) {  // fgs::details::extract_single
    auto const& dots{n.fragment.dots()};
    if (auto rc{rb_snippet_create(opt.device,
                                  reinterpret_cast<std::uint16_t const*>(dots.data()),
                                  static_cast<std::uint32_t>(dots.width()),
                                  static_cast<std::uint32_t>(dots.height()),
                                  &n.snippet)};
        rc != RB_OK) {
      std::string msg{n.snippet != nullptr ? rb_snippet_last_error(n.snippet) : "no CUDA device"};
      rb_snippet_destroy(n.snippet);
      n.snippet = nullptr;
      throw std::runtime_error("fgs_b200::splice: " + msg);
    }
  }

  // a node's fragment as plain records (the dots stay where they are: in the snippet)
  inline void open(node& n) {
    if (n.merged) return;
    n.step = n.fragment.step();
    n.dim = n.fragment.dimensions();
    n.zero = n.fragment.zero();
    n.frames = n.fragment.frames();  // the one copy of the frame records (fgm::fragment only hands out a const view)
    n.fragment = fgm::fragment{};
    n.merged = true;
  }

  // fragment::extend<Idx> (src/fgm.hpp:199-226) on one axis: how far the map grows below (lo) and above (hi) so that
  // [pos, pos + size) fits, in whole steps (get_step, :228-233); zero moves down by lo
  inline void grow(std::int32_t pos, std::size_t size, std::int32_t& zero, std::size_t dim, std::size_t step, std::size_t& lo,
                   std::size_t& hi) {
    auto round{[step](std::size_t change) {
      auto rest{change % step};
      return (change - rest) + (rest != 0 ? step : 0);
    }};
    lo = hi = 0;
    if (pos < zero) {
      lo = round(static_cast<std::size_t>(zero - pos));
    }
    if (auto required{pos + static_cast<std::int32_t>(size)}; required > 0) {
      if (auto limit{zero + dim}; static_cast<std::size_t>(required) > limit) {  // (the reference's mixed-sign sum, :215)
        hi = round(static_cast<std::size_t>(required) - limit);
      }
    }
    zero -= static_cast<std::int32_t>(lo);
  }

  // kpm::match(head, first) and head->bind on success (src/fgs.hpp:123-131)
  inline void match(node& head, node& other, options const& opt) {
    rb_cell_match m;
    if (rb_snippet_match(head.snippet, other.snippet, opt.cell_width, opt.cell_height, &m) != RB_OK) {
      throw std::runtime_error(std::string{"fgs_b200::splice: "} + rb_snippet_last_error(head.snippet));
    }
    if (m.valid != 0) {
      head.edges.push_back({other.id, kpm::vote{cdt::offset_t{m.dx, m.dy}, m.matched_keypoints}});
      if (m.ties > 1 && opt.tied_matches != nullptr) ++*opt.tied_matches;
    }
  }

}  // namespace details

template<typename Iter>
[[nodiscard]] std::vector<fgm::fragment> splice(Iter first, Iter last, options const& opt = options{}) {
  using namespace details;
  // RB_SPLICE_TRACE=1: wall time per phase on stderr (where a splice spends its time; measurements only)
  bool const trace{std::getenv("RB_SPLICE_TRACE") != nullptr};
  double t_extract{0}, t_match{0}, t_merge{0}, t_fetch{0};
  auto now{[] { return std::chrono::steady_clock::now(); }};
  auto since{[](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }};
  std::list<node> nodes;
  std::size_t next_id{0};
  struct cleanup {
    std::list<node>& nodes;
    ~cleanup() {
      for (auto& n : nodes) rb_snippet_destroy(n.snippet);
    }
  } guard{nodes};

  auto t0{now()};
  for (; first != last; ++first) {  // extract_all
    nodes.push_back(node{next_id++, std::move(*first)});
    extract(nodes.back(), opt);
  }
  t_extract += since(t0);
  t0 = now();
  for (auto head{nodes.begin()}; head != nodes.end(); ++head) {  // match_all
    for (auto other{std::next(head)}; other != nodes.end(); ++other) match(*head, *other, opt);
  }
  t_match += since(t0);

  while (true) {
    // select_match: first edge with the largest count, snippets in list order, edges in creation order
    auto left{nodes.end()};
    edge const* pick{nullptr};
    for (auto it{nodes.begin()}; it != nodes.end(); ++it) {
      for (auto const& e : it->edges) {
        if (pick == nullptr || pick->vote.count_ < e.vote.count_) {
          pick = &e;
          left = it;
        }
      }
    }
    if (pick == nullptr) break;

    // splice_single: dst.blit(dst.zero() + offset, std::move(right.fragment)); dst.normalize()  (src/fgs.hpp:146-150)
    t0 = now();
    auto right{nodes.begin()};
    while (right->id != pick->other) ++right;
    auto offset{pick->vote.offset_};
    open(*left);
    open(*right);
    fgm::point_t const pos{left->zero.x_ + offset.x_, left->zero.y_ + offset.y_};
    node merged{next_id++, fgm::fragment{}};
    merged.merged = true;
    merged.step = left->step;
    merged.zero = left->zero;
    std::size_t lo_x, hi_x, lo_y, hi_y;  // fragment::ensure (src/fgm.hpp:190-197)
    grow(pos.x_, right->dim.width_, merged.zero.x_, left->dim.width_, left->step.width_, lo_x, hi_x);
    grow(pos.y_, right->dim.height_, merged.zero.y_, left->dim.height_, left->step.height_, lo_y, hi_y);
    merged.dim = mrl::dimensions_t{left->dim.width_ + lo_x + hi_x, left->dim.height_ + lo_y + hi_y};
    if (rb_snippet_merge(left->snippet, static_cast<std::uint32_t>(lo_x), static_cast<std::uint32_t>(lo_y), right->snippet,
                         static_cast<std::uint32_t>(pos.x_ - merged.zero.x_), static_cast<std::uint32_t>(pos.y_ - merged.zero.y_),
                         static_cast<std::uint32_t>(merged.dim.width_), static_cast<std::uint32_t>(merged.dim.height_),
                         &merged.snippet) != RB_OK) {
      std::string msg{merged.snippet != nullptr ? rb_snippet_last_error(merged.snippet) : rb_snippet_last_error(left->snippet)};
      rb_snippet_destroy(merged.snippet);
      throw std::runtime_error("fgs_b200::splice: " + msg);
    }
    merged.frames = std::move(left->frames);
    merged.frames.reserve(merged.frames.size() + right->frames.size());
    for (auto& f : right->frames) {  // src/fgm.hpp:108-112
      merged.frames.emplace_back(f.number_, f.position_ - right->zero + pos, std::move(f.data_));
    }
    for (auto& f : merged.frames) {  // normalize, src/fgm.hpp:137-143
      f.position_ -= merged.zero;
    }
    merged.zero = {0, 0};

    auto const gone_a{left->id}, gone_b{right->id};
    rb_snippet_destroy(left->snippet);
    rb_snippet_destroy(right->snippet);
    nodes.erase(right);
    nodes.erase(left);
    for (auto& n : nodes) {  // unbind: nobody keeps an edge to a snippet that is gone
      std::erase_if(n.edges, [&](edge const& e) { return e.other == gone_a || e.other == gone_b; });
    }
    nodes.push_front(std::move(merged));
    t_merge += since(t0);
    t0 = now();
    for (auto other{std::next(nodes.begin())}; other != nodes.end(); ++other) match(nodes.front(), *other, opt);
    t_match += since(t0);
  }
  t0 = now();

  std::vector<fgm::fragment> result{};
  result.reserve(nodes.size());
  for (auto& n : nodes) {
    if (!n.merged) {
      result.push_back(std::move(n.fragment));
      continue;
    }
    fgm::fragment::matrix_type dots{n.dim};  // a merged fragment's dots come back once, here
    if (rb_snippet_fetch_dots(n.snippet, reinterpret_cast<std::uint16_t*>(dots.data())) != RB_OK) {
      fail(n.snippet, "rb_snippet_fetch_dots");
    }
    result.emplace_back(std::move(dots), n.step, n.zero, std::move(n.frames));
  }
  t_fetch += since(t0);
  if (trace) {
    std::fprintf(stderr, "fgs_b200::splice: extract (upload + blend + kpe) %.1f ms, matches %.1f ms, merges %.1f ms, read-back %.1f ms\n",
                 t_extract, t_match, t_merge, t_fetch);
  }
  return result;
}

}  // namespace fgs_b200
