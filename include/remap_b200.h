/* remap_b200.h -- C ABI of the B200-native registration path for kataklinger/remap.
 *
 * This is the drop-in boundary.  The reference (header-only C++20, /root/reference/src) has no
 * plugin or FFI layer; the seam for its per-frame registration hot path is the public surface of
 * frc::collector (src/frc.hpp:51-80), which mpb::builder::collect names concretely
 * (src/mpb.hpp:52-61).  include/frc_b200.hpp is a collector-shaped C++ shim over this ABI; the
 * entry points below are what that shim (or a ctypes / cgo / JNI stub) binds.  INTEGRATION.md shows
 * the one-line change in mpb.hpp.
 *
 * All functions return 0 on success and a negative rb_status on failure; none throws.  The library
 * has NO host compute path: every result is produced by the sm_100a kernels, and rb_create fails
 * with RB_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Conventions: caller allocates every buffer; the library never frees caller memory; host pointers
 * may be pageable or pinned (pinned makes the copies asynchronous); one CUDA stream per context; a
 * context is not thread-safe; one context per GPU.
 */
#ifndef REMAP_B200_H
#define REMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_ABI_VERSION 4

typedef enum rb_status {
  RB_OK = 0,
  RB_ERR_INVALID = -1,   /* bad argument / configuration                    */
  RB_ERR_NO_DEVICE = -2, /* no usable CUDA device (there is no CPU fallback) */
  RB_ERR_CUDA = -3,      /* a CUDA call failed; see rb_last_error            */
  RB_ERR_CAPACITY = -4,  /* more frames than the context was created for     */
  RB_ERR_STATE = -5      /* call sequence error (e.g. register before upload) */
} rb_status;

typedef struct rb_ctx rb_ctx;

/* Replaces the compile-time constants of frc (src/frc.hpp:22-24: grid 4x2, overlap 16;
 * src/frc.hpp:32-33: weight_switch 10, region_votes 3) and the collector's window dimensions
 * (src/frc.hpp:51-53).  rb_default_config fills in the reference's values. */
typedef struct rb_config {
  uint32_t width, height;   /* frame (action window) size in pixels                       */
  uint32_t grid_w, grid_h;  /* kpr::grid<4, 2>                                             */
  uint32_t overlap;         /* frc::grid_overlap                                           */
  uint32_t weight_switch;   /* match_config::weight_switch                                 */
  uint32_t region_votes;    /* match_config::region_votes (1..3)                           */
  int32_t device;           /* CUDA device ordinal                                         */
  uint32_t max_frames;      /* capacity of the HBM-resident frame store of this context    */
  uint32_t compute_median;  /* 1: K1 also writes the median image (kpe's second output)    */
  uint32_t code_slots;      /* 0 = auto; shared-memory code table slots (power of two)     */
  uint32_t offset_slots;    /* 0 = auto; shared-memory offset table slots (power of two)   */
  uint32_t profile;         /* 1: record CUDA events around every kernel (rb_kernel_times) */
  void* stream;             /* cudaStream_t to use, or NULL to create one                  */
  uint32_t kpm_mode;        /* 0 = pipelined matchers (which of the two takes the regions first follows the
                               region size), general kernel for what they defer; 1 = general kernel only (one
                               CTA per pair and region); 2 / 3 = large-region / fast pipelined matcher first */
  uint32_t list_cap;        /* 0 = auto; per-region keypoint list capacity of the first-pass pipelined
                               matcher (<= 2047); longer lists go to the large-region matcher (second pass,
                               up to 2 x list_cap), then to the general kernel                              */
  uint32_t run_pairs;       /* 0 = auto (8..40, per launch); consecutive pairs per work item of the matcher */
  uint32_t upload_chunk;    /* 0 = auto; frames per host->device chunk of rb_register_host_async */
  uint32_t overlap_batches; /* 0 = auto (1 = K1 then matcher, one after the other); > 1: batches per call, K1 of
                               batch b + 1 on a second stream concurrently with the matcher of batch b (slower
                               on B200 as measured; kept for experiments)                                     */
  uint32_t host_threads;    /* 0 = auto; host threads of rb_register_host_async's packer (auto: processors /
                               max(ranks on this node, contexts of this process))                            */
} rb_config;

/* == std::optional<cdt::offset_t> returned by kpm::match (src/kpm.hpp:395-415), plus flags. */
typedef struct rb_offset {
  int32_t dx, dy;  /* prev - curr keypoint offset (src/kpm.hpp:96-98)                         */
  uint32_t flags;  /* RB_OFFSET_VALID: has_value(); RB_OFFSET_TIE_SENSITIVE: see DESIGN.md    */
} rb_offset;
#define RB_OFFSET_VALID 1u
#define RB_OFFSET_TIE_SENSITIVE 2u

/* One keypoint as kpe::extractor hands it to kpr::grid::add (src/kpe.hpp:225-229,301-303):
 * the 13-byte kpr::code (src/kpr.hpp:20-23, layout src/kpe.hpp:342-379), the point, and the set of
 * grid regions it is inserted into (bit i = region i, src/kpr.hpp:71-74). */
typedef struct rb_keypoint {
  uint8_t code[13];
  uint8_t weight;
  uint16_t x, y;
  uint32_t region_mask;
} rb_keypoint;

/* One offset-histogram bin of one region: an entry of kpm's totalizator_t (src/kpm.hpp:70-76). */
typedef struct rb_bin {
  int32_t dx, dy;
  uint32_t count;
} rb_bin;

/* One region's ballot: kpm::details::cast_vote's ticket (src/kpm.hpp:213-223) + tie statistics. */
typedef struct rb_region_vote {
  uint32_t use_all;          /* weight switch outcome (src/kpm.hpp:219-220)        */
  uint32_t n_prev, n_curr;   /* insertions into the region (prev / curr frame)     */
  uint32_t w2_prev, w2_curr; /* weight-2 insertions                                */
  uint32_t nbins;            /* distinct offsets                                   */
  uint32_t nticket;          /* min(region_votes, nbins)                           */
  rb_bin ticket[4];          /* count desc, dx asc, dy asc                         */
  uint32_t ngt[4], nge[4];   /* #bins with count > / >= ticket[k].count            */
  uint32_t hist_hash;        /* digest of the whole histogram (totalizator_t, src/kpm.hpp:70-76): wrapping sum over
                                its bins of the bin digest below                    */
} rb_region_vote;
/* bin digest: h = (dx & 0xFFFF) | dy << 16; h = h * 0x9E3779B1 ^ count * 0x85EBCA77; h ^= h >> 15; h *= 0x2C1B3C6D;
 * h ^= h >> 12; h *= 0x297A2D39; h ^= h >> 15   (all uint32, wrapping) */

void rb_default_config(rb_config* cfg, uint32_t width, uint32_t height, uint32_t max_frames);

/* frc::collector::collector(dimensions) (src/frc.hpp:51-53): allocates the HBM frame store, the
 * per-frame keypoint bit maps, the median store and the per-pair ballots. */
int rb_create(const rb_config* cfg, rb_ctx** out);
void rb_destroy(rb_ctx* ctx);

/* feed.produce() for n frames at once (src/frc.hpp:86,102; src/ifd.hpp:20-28): frames = n*H*W
 * bytes, row-major, one C64 colour 0..15 per byte (src/nil.hpp:14-31).  Copies them into the
 * context's frame store starting at slot `first` (asynchronously on the context's stream). */
int rb_upload(rb_ctx* ctx, const uint8_t* frames, size_t first, size_t n);

/* The registration of frames [first, first+n): kpe::extractor::extract on every frame
 * (src/frc.hpp:90,105) and kpm::match on every consecutive pair (src/frc.hpp:107).  Enqueues the
 * kernels on the context's stream and returns; results stay in HBM until fetched. */
int rb_register_async(rb_ctx* ctx, size_t first, size_t n);

/* rb_upload + rb_register_async for frames in HOST memory, pipelined: the frames go to the device in
 * chunks on a second stream and each chunk is registered as soon as it has landed, so the copies of
 * later frames run under the kernels of earlier ones.  Same results, same fetch calls afterwards.
 * `frames` must stay valid until the next synchronising call (rb_fetch_*, rb_synchronize); pinned
 * memory (rb_alloc_host) is what makes the copies overlap. */
int rb_register_host_async(rb_ctx* ctx, const uint8_t* frames, size_t first, size_t n);
/* How the chunks travel is decided per chunk from measured rates: host threads pack a chunk
 * to 4 bit/pixel (half the bytes cross PCIe), or -- page-locked `frames` only -- the chunk is copied as it is and
 * packed on the device; both lanes work at the same time.  rb_host_lane_stats reports the last call's split. */
int rb_host_lane_stats(rb_ctx* ctx, uint64_t* raw_chunks, uint64_t* packed_chunks, double* link_GBps, double* pack_fps,
                       int* threads, uint64_t* h2d_bytes);
/* The same for a caller that already holds 4 bit/pixel frames (the layout a packed capture would have: pixel x in
 * nibble x & 1 of byte x >> 1; rows of row_bytes >= ceil(W / 2) bytes, frames back to back): one copy per chunk,
 * no host work. */
int rb_register_host_packed4(rb_ctx* ctx, const uint8_t* packed, size_t row_bytes, size_t first, size_t n);

/* Copies the n-1 pair results of the last rb_register_async to the host and waits for them.
 * out[i] belongs to frames (first+i, first+i+1). */
int rb_fetch_offsets(rb_ctx* ctx, rb_offset* out, size_t n_pairs);

/* kpe's median image (src/kpe.hpp:314) of frames [first, first+n): n*H*W bytes. */
int rb_fetch_medians(rb_ctx* ctx, size_t first, size_t n, uint8_t* out);

/* rb_register_async + rb_fetch_offsets (+ rb_fetch_medians when out_median != NULL). */
int rb_register(rb_ctx* ctx, size_t first, size_t n, rb_offset* out, uint8_t* out_median);

/* Parity taps (not on the hot path). */
int rb_keypoints(rb_ctx* ctx, size_t frame, rb_keypoint* out, size_t cap, size_t* count);
int rb_region_ballots(rb_ctx* ctx, size_t pair, rb_region_vote* out /* grid_w*grid_h entries */);
int rb_region_votes(rb_ctx* ctx, size_t pair, uint32_t region, rb_bin* out, size_t cap, size_t* count);
/* Bulk parity taps for full-sequence comparisons: the ballots of pairs [pair, pair + n_pairs) of the last
 * registration (n_pairs * grid_w * grid_h records), and per-frame digests of kpe's outputs:
 *   splitmix(z): z += 0x9E3779B97F4A7C15; z = (z ^ z >> 30) * 0xBF58476D1CE4E5B9; z = (z ^ z >> 27) * 0x94D049BB133111EB; z ^ z >> 31
 *   median_hash = sum over pixels with value v != 0 at index i = y * W + x of splitmix(i << 8 | v)
 *   kp_hash     = sum over every insertion (region r, point (x, y), 13-byte code) into the grid (src/kpe.hpp:225-229,301-303)
 *                 of splitmix((x | y << 16 | r << 32) ^ splitmix(code[0..7] ^ splitmix(code[8..12])))    (little-endian)
 * Together with rb_region_vote.hist_hash they let a test compare EVERY frame and pair of a long sequence with the
 * reference (oracle/_ref/ref_harness digest) without moving images or keypoint lists. */
typedef struct rb_frame_digest {
  uint64_t median_hash, kp_hash;
  uint32_t keypoints;   /* unique pixels                        */
  uint32_t insertions;  /* sum over regions (overlaps count twice or four times) */
} rb_frame_digest;
int rb_frame_digests(rb_ctx* ctx, size_t first, size_t n, rb_frame_digest* out);
int rb_fetch_ballots(rb_ctx* ctx, size_t pair, size_t n_pairs, rb_region_vote* out);

/* fde::details::generate_mask (src/fde.hpp:19-55): mask[y][x] = 0xFF where the background map
 * equals the frame placed at (px, py), else 0.  bg = bgH*bgW bytes, frame/out_mask = H*W bytes. */
int rb_foreground_mask(rb_ctx* ctx, const uint8_t* bg, uint32_t bgW, uint32_t bgH, int32_t px, int32_t py,
                       const uint8_t* frame, uint8_t* out_mask);
/* Same, frame taken from the resident frame store; mask left in HBM unless out_mask != NULL. */
int rb_foreground_mask_resident(rb_ctx* ctx, const uint8_t* bg, uint32_t bgW, uint32_t bgH, int32_t px,
                                int32_t py, size_t frame, uint8_t* out_mask);

/* Map assembly.  fgm::fragment::blit of n frames of the resident store into one
 * fragment's dot map (src/fgm.hpp:87-97,176-188: ++dots[pos + xy][colour], uint16 counters that wrap) and
 * fgm::fragment::blend of the result (src/fgm.hpp:115-135: first colour with the largest count, mask =
 * pixel ever covered).  placements[i] = (frame slot, x, y) with (x, y) the frame's position inside the
 * map, i.e. fgm::frame::position_ minus the fragment's zero (the caller replays fragment::ensure,
 * src/fgm.hpp:190-233, to get zero and the map size; include/frc_b200.hpp does).  Any output may be NULL.
 * out_dots: mapH*mapW*16 uint16, out_image / out_mask: mapH*mapW bytes. */
typedef struct rb_placement {
  uint32_t frame;
  int32_t x, y;
} rb_placement;
int rb_blit_blend(rb_ctx* ctx, const rb_placement* placements, size_t n, uint32_t mapW, uint32_t mapH, uint16_t* out_dots,
                  uint8_t* out_image, uint8_t* out_mask);

/* Multi-GPU map assembly (per-rank partial dot maps, one reduction to rank 0, blend there).
 * Every rank calls rb_blit_blend (or rb_filter_fragment) with ITS frames and the fragment's full map geometry;
 * rb_map_device hands out the device addresses of the result so that the partial dot maps can be summed across
 * ranks in place (NCCL reduce; uint16 counters wrap, so sum in a wider type and keep the low 16 bits --
 * remap_b200.shard.reduce_fragment_map does); rb_blend_map then runs fgm::fragment::blend (src/fgm.hpp:115-135)
 * over the reduced dots on the destination rank.  Pointers stay valid until the context's next map call. */
int rb_map_device(rb_ctx* ctx, uint16_t** dots, uint8_t** image, uint8_t** mask, uint32_t* mapW, uint32_t* mapH);
int rb_blend_map(rb_ctx* ctx, uint16_t* out_dots, uint8_t* out_image, uint8_t* out_mask);
/* The same as ONE kernel over peer memory: every rank exports its partial map (rb_map_export: a CUDA-IPC handle of
 * the map scratch; ship the 80 bytes to the destination rank any way you like), the destination rank calls
 * rb_blend_map_peers with the other ranks' handles: its kernel reads the peers' dot maps in place over NVLink,
 * adds them to its own (uint16 lanes, wrapping) and blends in the same pass -- no staging buffers, no second pass.
 * The peers must not touch their map scratch until the call has returned (a barrier on the caller's side). */
typedef struct rb_map_handle {
  uint8_t opaque[64];  /* cudaIpcMemHandle_t */
  uint32_t map_w, map_h, device, reserved;
} rb_map_handle;
int rb_map_export(rb_ctx* ctx, rb_map_handle* out);
int rb_blend_map_peers(rb_ctx* ctx, const rb_map_handle* peers, size_t npeers, uint16_t* out_dots, uint8_t* out_image,
                       uint8_t* out_mask);
/* The all-links form for many GPUs (reduce-scatter, then gather).  ranks[r] = rank r's handle (ranks[self] is
 * ignored), world <= 15.  Step 1, on EVERY rank: rb_sum_map_slice sums slice `self` of the map (pixels
 * [self * ceil(px / world), ...)) over all other ranks into this rank's scratch -- every NVLink carries 1/world of a
 * map per peer at the same time.  Barrier.  Step 2, on the destination rank: rb_blend_map_slices pulls each slice from
 * the rank that reduced it and blends.  Barrier before anyone touches its scratch again. */
int rb_sum_map_slice(rb_ctx* ctx, const rb_map_handle* ranks, size_t world, size_t self);
int rb_blend_map_slices(rb_ctx* ctx, const rb_map_handle* ranks, size_t world, size_t self, uint16_t* out_dots,
                        uint8_t* out_image, uint8_t* out_mask);

/* Pass-2 foreground filtering of one fragment: fdf::filter (src/fdf.hpp:40-75).  For every
 * placed frame: fde::extractor::extract (src/fde.hpp:83-103: generate_mask against the background window, the
 * contours of the frame's MEDIAN image that hold a differing pixel -- cte::extractor, src/cte.hpp:60-166 --
 * minus those larger than a fifth of the frame) and fde::mask (src/fde.hpp:122-146: the kept contours' pixels
 * and enclosures); then fgm::fragment::blit with that mask (src/fgm.hpp:71-85: a pixel counts where the mask
 * is 0) into a fresh dot map of the background's size, and blend.
 *   placements   as for rb_blit_blend; the frames must have been registered (their median images are read
 *                from the context's median store, as fdf decompresses frc's stored medians, src/fdf.hpp:60)
 *   background   mapH*mapW bytes = fdf::background::image_, or NULL for what the 4-argument fdf::filter
 *                computes itself: blend() of the plain blit of these placements (src/fdf.hpp:21-34,79-89)
 *   out_fgmasks  n*H*W bytes, 1 = foreground (== fde::mask), or NULL     (parity tap)
 *   out_ncontours n counts of kept contours per frame, or NULL          (parity tap)
 * The reference's contour ids are uint16 and wrap after 65,534 contours in one frame (src/cte.hpp:20,96-98);
 * frames beyond that are outside the contract. */
int rb_filter_fragment(rb_ctx* ctx, const rb_placement* placements, size_t n, uint32_t mapW, uint32_t mapH,
                       const uint8_t* background, uint16_t* out_dots, uint8_t* out_image, uint8_t* out_mask,
                       uint8_t* out_fgmasks, uint32_t* out_ncontours);
/* Median images from the caller instead of from K1: n*H*W bytes into slots [first, first+n) of the median
 * store.  fdf::filter reads the medians frc stored with every frame (src/fdf.hpp:60, src/frc.hpp:134); a
 * caller that runs pass 2 in a context other than the one that registered the frames hands them over here. */
int rb_upload_medians(rb_ctx* ctx, const uint8_t* medians, size_t first, size_t n);
/* Last rb_filter_fragment: ms[0] background blit + blend, ms[1] foreground bit maps, ms[2] masked blit + blend;
 * frames_deferred = frames whose runs or contours did not fit the shared-memory tables and took the
 * global-memory variant of the same kernel. */
int rb_filter_times(rb_ctx* ctx, float* ms, size_t n, uint32_t* frames_deferred);

/* Fragment splicing: the device side of fgs::splice (src/fgs.hpp:187-213).  A snippet is what
 * fgs::details::extract_single (src/fgs.hpp:80-89) makes of a fragment: blend() image + mask and the keypoints of
 * kpe::extractor<kpr::grid<1, 1>, 0> over the whole map image; it lives on the device.  dots = H*W*16 uint16
 * (fgm::fragment::dots()).  Snippets have no rb_ctx; each owns a stream on `device`. */
typedef struct rb_snippet rb_snippet;
int rb_snippet_create(int device, const uint16_t* dots, uint32_t W, uint32_t H, rb_snippet** out);
/* fgm::fragment::blit(pos, fragment&&) (src/fgm.hpp:99-113) + extract_single of the result (src/fgs.hpp:146-152) without
 * leaving the device: the new snippet's W x H dot map is a's map at (ax, ay) plus b's map at (bx, by) (uint16 counters,
 * wrapping), everything else zero.  The caller replays fragment::ensure (src/fgm.hpp:190-233) for the geometry;
 * include/fgs_b200.hpp does.  a and b are left untouched.  rb_snippet_fetch_dots reads a snippet's map back
 * (H * W * 16 uint16 = fgm::fragment::dots()). */
int rb_snippet_merge(rb_snippet* a, uint32_t ax, uint32_t ay, rb_snippet* b, uint32_t bx, uint32_t by, uint32_t W, uint32_t H,
                     rb_snippet** out);
int rb_snippet_fetch_dots(rb_snippet* s, uint16_t* out);
void rb_snippet_destroy(rb_snippet* s);
const char* rb_snippet_last_error(rb_snippet* s);
/* Parity tap: keypoint count, blend image / mask (H*W bytes each), keypoint records (unordered). Any may be NULL. */
int rb_snippet_fetch(rb_snippet* s, uint32_t* nkeypoints, uint8_t* out_image, uint8_t* out_mask, rb_keypoint* out_kps, size_t cap);

/* kpm::match, cellular variant (src/kpm.hpp:371-393), of two snippets: every pair of equal codes votes
 * prev - curr (count_offsets, :231-262), the offset with the most votes wins (find_best, :280-299), and it is
 * accepted iff its votes fall into at least 0.66 x the active cells (count_active_cells, :349-369, :387-389).
 * find_best takes the FIRST maximum in std::unordered_map iteration order, which is implementation-defined;
 * here ties go to the smallest (dy, dx) and `ties` > 1 reports that the reference's choice is not defined. */
typedef struct rb_cell_match {
  uint32_t valid;              /* the std::optional<kpm::vote> has a value                    */
  int32_t dx, dy;              /* vote.offset_ (prev - curr)                                   */
  uint32_t matched_keypoints;  /* vote.count_                                                  */
  uint32_t matched_cells, active_cells;
  uint32_t offsets;            /* distinct offsets that received a vote                        */
  uint32_t ties;               /* offsets sharing the largest vote count                       */
  uint64_t pairs;              /* all votes                                                    */
} rb_cell_match;
int rb_snippet_match(rb_snippet* prev, rb_snippet* curr, uint32_t cell_w, uint32_t cell_h, rb_cell_match* out);

/* aws::details::compare (src/aws.hpp:37-60; called once per frame by aws::scan, :121) for every consecutive pair
 * of the resident frames [first, first + n), in one streaming pass: heat (H*W bytes, in/out, may be NULL) is
 * cleared wherever a pair differs, exactly as n - 1 compare calls leave it; first_change (H*W uint32, may be
 * NULL) receives the index i of the first pair (first + i, first + i + 1) that differs at that pixel, 0xFFFFFFFF
 * if none -- the heat map after k pairs is heat0 & (first_change >= k), so aws::scan's per-frame states can be
 * replayed from one call.  (The context's width/height are the SCREEN dimensions here, src/aws.hpp:98-101.) */
int rb_aws_compare(rb_ctx* ctx, size_t first, size_t n, uint8_t* heat, uint32_t* first_change);

/* Several GPUs of ONE process: the reference's consumer, mpb::builder::collect
 * (src/mpb.hpp:52-61), is a single C++ caller that owns the whole feed.  A group owns one rb_ctx per device;
 * rb_group_register_host cuts the caller's n frames into contiguous ranges, one per device, with a one-frame
 * overlap (every consecutive pair belongs to exactly one member), registers them concurrently (one host thread,
 * one share of the packer threads and one PCIe link per member; no data-path collective), gathers the 12-byte pair
 * results on the lead device (devices[0]) with device-to-device copies over NVLink and returns all n - 1 of them
 * in `out` with one device-to-host copy.  cfg->device is ignored, cfg->max_frames is the capacity of the WHOLE
 * sequence, cfg->stream must be NULL.  The frames stay resident: rb_group_range / rb_group_locate say where, and
 * rb_group_context hands out the member for per-member calls (rb_keypoints, rb_blit_blend, rb_filter_fragment...). */
typedef struct rb_group rb_group;
int rb_group_create(const rb_config* cfg, const int32_t* devices, size_t ndev, rb_group** out);
void rb_group_destroy(rb_group* g);
const char* rb_group_last_error(rb_group* g);
size_t rb_group_size(rb_group* g);
rb_ctx* rb_group_context(rb_group* g, size_t member);
int rb_group_register_host(rb_group* g, const uint8_t* frames, size_t n, rb_offset* out);
/* member i holds frames [first, end) of the last call in its slots 0.. and OWNS frames [own, end) */
int rb_group_range(rb_group* g, size_t member, size_t* first, size_t* end, size_t* own);
int rb_group_locate(rb_group* g, size_t frame, size_t* member, size_t* slot);
int rb_group_fetch_medians(rb_group* g, size_t first, size_t n, uint8_t* out);
const rb_offset* rb_group_offsets_device(rb_group* g);  /* the gathered results on the lead device */

/* Device-side access for callers that keep working on the GPU (multi-GPU gather with NCCL, map
 * assembly): the n-1 rb_offset records of the last rb_register_async, in HBM. */
const rb_offset* rb_offsets_device(rb_ctx* ctx);
/* Number of keypoints (unique pixels, not region insertions) K1 found in frames [first, first+n). */
int rb_count_keypoints(rb_ctx* ctx, size_t first, size_t n, uint64_t* total);

/* Page-locked host memory for the caller's staging buffers (frames in, offsets / medians out): with
 * pinned buffers rb_upload and the fetches are true asynchronous DMA.  Plain malloc'ed memory works
 * too, only slower.  (The reference allocates frames from all::memory_pool, src/all.hpp:14-106.) */
void* rb_alloc_host(size_t bytes);
void rb_free_host(void* p);

/* Introspection for benchmarks and tests. */
int rb_synchronize(rb_ctx* ctx);
void* rb_stream(rb_ctx* ctx);                          /* the cudaStream_t the kernels run on       */
int rb_kernel_times(rb_ctx* ctx, float* ms, size_t n); /* last rb_register_async: kpe, matcher (all), declare[, lists, match, deferred] */
uint64_t rb_kernel_launches(rb_ctx* ctx);              /* kernels launched by this context so far   */
/* (pair, region) ballots of the last rb_register_async that the pipelined matcher deferred to the
 * general kernel (waits for the stream). */
int rb_deferred_count(rb_ctx* ctx, uint32_t* count);
size_t rb_device_bytes(rb_ctx* ctx);                   /* HBM held by this context                  */
const char* rb_last_error(rb_ctx* ctx);
const char* rb_matcher_kernel(rb_ctx* ctx);           /* the kernel that takes the regions first (introspection) */
uint32_t rb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* REMAP_B200_H */
