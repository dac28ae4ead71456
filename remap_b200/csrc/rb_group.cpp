// rb_group.cpp -- several GPUs of ONE process behind the C ABI (SURVEY.md 8(b) "one ctx per device in one process",
// 8(e)): the consumer is mpb::builder::collect (src/mpb.hpp:52-61), a single C++ caller that owns the whole feed.
//
// A group owns one rb_ctx per device.  rb_group_register_host cuts the caller's frame sequence into contiguous
// ranges, one per device, with a one-frame overlap (member i also takes the last frame of member i - 1, so every
// consecutive pair belongs to exactly one member -- kpe is independent per frame and kpm per pair, src/frc.hpp:105-107),
// runs rb_register_host_async on every member from its own host thread (each with its share of the packer threads
// and its own PCIe link), then gathers the 12-byte pair results to the lead device over NVLink
// (cudaMemcpyPeerAsync, device to device) and hands them to the caller with ONE device-to-host copy.
// No data-path collective: the members never talk while registering.
//
// Host code only (no kernels): everything device-side goes through the rb_* entry points of rb_api.cu.
#include <cuda_runtime.h>
#include <string.h>

#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/remap_b200.h"

struct rb_group {
  std::vector<rb_ctx*> ctx;
  std::vector<int> device;
  std::vector<size_t> first, end;  // member i registered frames [first[i], end[i]) of the last call as its slots 0..
  std::vector<size_t> own;         // ... and owns frames [own[i], end[i]) (the overlap frame is its predecessor's)
  rb_config cfg;
  size_t n_last;
  rb_offset* d_gather;  // on the lead device: the whole sequence's pair results
  size_t gather_cap;
  std::string err;
};

extern "C" {

void rb_group_destroy(rb_group* g) {
  if (!g) return;
  for (rb_ctx* c : g->ctx) rb_destroy(c);
  if (g->d_gather && !g->device.empty()) { cudaSetDevice(g->device[0]); cudaFree(g->d_gather); }
  delete g;
}

const char* rb_group_last_error(rb_group* g) { return g ? g->err.c_str() : "null group"; }
size_t rb_group_size(rb_group* g) { return g ? g->ctx.size() : 0; }
rb_ctx* rb_group_context(rb_group* g, size_t i) { return g && i < g->ctx.size() ? g->ctx[i] : nullptr; }

int rb_group_create(const rb_config* cfg, const int32_t* devices, size_t ndev, rb_group** out) {
  if (!cfg || !devices || !out || ndev == 0 || ndev > 64) return RB_ERR_INVALID;
  *out = nullptr;
  rb_group* g = new (std::nothrow) rb_group();
  if (!g) return RB_ERR_INVALID;
  *out = g;  // returned even on failure so that rb_group_last_error can be read; caller rb_group_destroy()s it
  g->cfg = *cfg;
  g->n_last = 0;
  g->d_gather = nullptr;
  g->gather_cap = 0;
  if (cfg->stream) { g->err = "rb_group_create: a group creates its own streams (cfg.stream must be NULL)"; return RB_ERR_INVALID; }
  // every member holds its share of the sequence plus the overlap frame
  const size_t per = (cfg->max_frames + ndev - 1) / ndev + 1;
  for (size_t i = 0; i < ndev; ++i) {
    rb_config c = *cfg;
    c.device = devices[i];
    c.max_frames = (uint32_t)(per < 2 ? 2 : per);
    rb_ctx* ctx = nullptr;
    const int rc = rb_create(&c, &ctx);
    if (rc != RB_OK) {
      g->err = std::string("rb_group_create: device ") + std::to_string(devices[i]) + ": " + (ctx ? rb_last_error(ctx) : "no usable CUDA device");
      rb_destroy(ctx);
      return rc;
    }
    g->ctx.push_back(ctx);
    g->device.push_back(devices[i]);
  }
  g->first.assign(ndev, 0);
  g->end.assign(ndev, 0);
  g->own.assign(ndev, 0);
  return RB_OK;
}

// The frame range member i holds after the last rb_group_register_host: frame f of the sequence, first <= f < end,
// lives in slot f - first of rb_group_context(g, i) (frames, medians, keypoint taps, rb_blit_blend placements).
// `own` = the first frame the member owns (first + 1 for every member but the one that starts the sequence: the
// overlap frame belongs to its predecessor).
int rb_group_range(rb_group* g, size_t i, size_t* first, size_t* end, size_t* own) {
  if (!g || i >= g->ctx.size()) return RB_ERR_INVALID;
  if (first) *first = g->first[i];
  if (end) *end = g->end[i];
  if (own) *own = g->own[i];
  return RB_OK;
}

int rb_group_register_host(rb_group* g, const uint8_t* frames, size_t n, rb_offset* out) {
  if (!g || !frames || (n > 1 && !out)) return RB_ERR_INVALID;
  const size_t nd = g->ctx.size();
  if (n < 1 || n > g->cfg.max_frames) { g->err = "rb_group_register_host: more frames than the group was created for"; return RB_ERR_CAPACITY; }
  const size_t px = (size_t)g->cfg.width * g->cfg.height;
  // contiguous ranges; a member whose predecessors are all empty starts the sequence (no overlap frame)
  for (size_t i = 0; i < nd; ++i) {
    const size_t lo = i * n / nd, hi = (i + 1) * n / nd;
    g->first[i] = hi > lo && lo > 0 ? lo - 1 : lo;
    g->end[i] = hi > lo ? hi : lo;
    g->own[i] = lo;
  }
  g->n_last = n;
  std::vector<int> rc(nd, RB_OK);
  std::vector<std::thread> th;
  for (size_t i = 0; i < nd; ++i) {
    if (g->end[i] <= g->first[i]) continue;
    th.emplace_back([g, i, frames, px, &rc] {
      rc[i] = rb_register_host_async(g->ctx[i], frames + g->first[i] * px, 0, g->end[i] - g->first[i]);
    });
  }
  for (auto& t : th) t.join();
  for (size_t i = 0; i < nd; ++i)
    if (rc[i] != RB_OK) { g->err = std::string("member ") + std::to_string(i) + ": " + rb_last_error(g->ctx[i]); return rc[i]; }
  if (n < 2) return RB_OK;
  // gather: member i's pairs (first[i] + k, first[i] + k + 1) -> slot first[i] + k of the lead device's array, device to
  // device (NVLink where the devices are peers), each copy on its member's stream, i.e. after that member's kernels
  const int lead = g->device[0];
  cudaError_t e = cudaSetDevice(lead);
  if (e == cudaSuccess && g->gather_cap < n - 1) {
    if (g->d_gather) cudaFree(g->d_gather);
    g->d_gather = nullptr;
    g->gather_cap = 0;
    e = cudaMalloc(reinterpret_cast<void**>(&g->d_gather), (size_t)g->cfg.max_frames * sizeof(rb_offset));
    if (e == cudaSuccess) g->gather_cap = g->cfg.max_frames;
  }
  for (size_t i = 0; i < nd && e == cudaSuccess; ++i) {
    const size_t np = g->end[i] > g->first[i] ? g->end[i] - g->first[i] - 1 : 0;
    if (np == 0) continue;
    e = cudaSetDevice(g->device[i]);
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(g->d_gather + g->first[i], lead, rb_offsets_device(g->ctx[i]), g->device[i], np * sizeof(rb_offset),
                              static_cast<cudaStream_t>(rb_stream(g->ctx[i])));
  }
  for (size_t i = 0; i < nd && e == cudaSuccess; ++i) {
    e = cudaSetDevice(g->device[i]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(static_cast<cudaStream_t>(rb_stream(g->ctx[i])));
  }
  if (e == cudaSuccess) e = cudaSetDevice(lead);
  if (e == cudaSuccess) e = cudaMemcpy(out, g->d_gather, (n - 1) * sizeof(rb_offset), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { g->err = std::string("rb_group_register_host: ") + cudaGetErrorString(e); cudaGetLastError(); return RB_ERR_CUDA; }
  // a member's matcher error word travels with rb_fetch_offsets; ask every member (0 pairs: the check only)
  for (size_t i = 0; i < nd; ++i) {
    if (g->end[i] <= g->first[i]) continue;
    const int r = rb_fetch_offsets(g->ctx[i], nullptr, 0);
    if (r != RB_OK) { g->err = std::string("member ") + std::to_string(i) + ": " + rb_last_error(g->ctx[i]); return r; }
  }
  return RB_OK;
}

// The gathered pair results of the last rb_group_register_host on the lead device (n - 1 records), for consumers that
// keep working on the GPU (map assembly).
const rb_offset* rb_group_offsets_device(rb_group* g) { return g ? g->d_gather : nullptr; }

// kpe's median images of frames [first, first + n) of the last rb_group_register_host, each from the member that owns
// the frame: n * H * W bytes.
int rb_group_fetch_medians(rb_group* g, size_t first, size_t n, uint8_t* out) {
  if (!g || !out) return RB_ERR_INVALID;
  if (first + n > g->n_last) { g->err = "rb_group_fetch_medians: frames not registered"; return RB_ERR_STATE; }
  const size_t px = (size_t)g->cfg.width * g->cfg.height;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    if (g->end[i] <= g->first[i]) continue;
    const size_t own = g->own[i];
    const size_t a = first > own ? first : own, b = first + n < g->end[i] ? first + n : g->end[i];
    if (a >= b) continue;
    const int rc = rb_fetch_medians(g->ctx[i], a - g->first[i], b - a, out + (a - first) * px);
    if (rc != RB_OK) { g->err = std::string("member ") + std::to_string(i) + ": " + rb_last_error(g->ctx[i]); return rc; }
  }
  return RB_OK;
}

// The member and slot that hold frame `frame` of the last rb_group_register_host (the member that OWNS it).
int rb_group_locate(rb_group* g, size_t frame, size_t* member, size_t* slot) {
  if (!g || frame >= g->n_last) return RB_ERR_INVALID;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    if (g->end[i] <= g->first[i]) continue;
    const size_t own = g->own[i];
    if (frame >= own && frame < g->end[i]) {
      if (member) *member = i;
      if (slot) *slot = frame - g->first[i];
      return RB_OK;
    }
  }
  return RB_ERR_INVALID;
}

}  // extern "C"
