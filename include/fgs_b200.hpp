// fgs_b200.hpp -- a fgs::splice-shaped front of the B200 fragment splicer.
//
// Drop-in for the reference's fgs::splice(first, last) (src/fgs.hpp:187-213), which mpb::builder::splice calls
// (src/mpb.hpp:63-69): same arguments, same result -- the fragments merged greedily, best vote first, until no
// two of them match.
//
// What moves to the GPU (remap_b200.h): fgs::details::extract_single for every fragment and every merged
// fragment (rb_snippet_create: fragment.blend() + kpe with a 1 x 1 grid over the whole map, src/fgs.hpp:80-89)
// and every cellular kpm::match (rb_snippet_match, src/kpm.hpp:371-393; src/fgs.hpp:119-134).  What stays
// here: the bookkeeping of who matched whom and the merge itself, fgm::fragment::blit(pos, fragment&&) +
// normalize (src/fgs.hpp:146-150), in the reference's own fragment type.  The selection rule is the
// reference's (src/fgs.hpp:142-168): among the edges recorded from the earlier snippet of each matching pair,
// the first one with the largest vote count, walking the snippets in list order and each snippet's edges in
// the order they were found; the merged fragment goes to the FRONT of the list and is matched against all
// the others (src/fgs.hpp:170-183).  This header is compiled in the REFERENCE's translation unit and
// contains no CUDA.
#pragma once

#include "remap_b200.h"

#include "fgm.hpp"
#include "kpm.hpp"

#include <cstdint>
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
    fgm::fragment fragment;
    rb_snippet* snippet{nullptr};
    std::vector<edge> edges;  // "primary" edges only: this node was the head of the match (src/fgs.hpp:68-71)
  };

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
  std::list<node> nodes;
  std::size_t next_id{0};
  struct cleanup {
    std::list<node>& nodes;
    ~cleanup() {
      for (auto& n : nodes) rb_snippet_destroy(n.snippet);
    }
  } guard{nodes};

  for (; first != last; ++first) {  // extract_all
    nodes.push_back(node{next_id++, std::move(*first)});
    extract(nodes.back(), opt);
  }
  for (auto head{nodes.begin()}; head != nodes.end(); ++head) {  // match_all
    for (auto other{std::next(head)}; other != nodes.end(); ++other) match(*head, *other, opt);
  }

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

    // splice_single
    auto right{nodes.begin()};
    while (right->id != pick->other) ++right;
    auto offset{pick->vote.offset_};
    auto& dst{left->fragment};
    dst.blit(dst.zero() + offset, std::move(right->fragment));
    dst.normalize();

    node merged{next_id++, std::move(dst)};
    auto const gone_a{left->id}, gone_b{right->id};
    rb_snippet_destroy(left->snippet);
    rb_snippet_destroy(right->snippet);
    nodes.erase(right);
    nodes.erase(left);
    for (auto& n : nodes) {  // unbind: nobody keeps an edge to a snippet that is gone
      std::erase_if(n.edges, [&](edge const& e) { return e.other == gone_a || e.other == gone_b; });
    }
    nodes.push_front(std::move(merged));
    extract(nodes.front(), opt);
    for (auto other{std::next(nodes.begin())}; other != nodes.end(); ++other) match(nodes.front(), *other, opt);
  }

  std::vector<fgm::fragment> result{};
  result.reserve(nodes.size());
  for (auto& n : nodes) result.push_back(std::move(n.fragment));
  return result;
}

}  // namespace fgs_b200
