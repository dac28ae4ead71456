"""Thin Python face of the C ABI (include/remap_b200.h): one Registrar per GPU.

All computation happens in libremap_b200.so's sm_100a kernels; this file only marshals numpy
buffers.  Record layouts (numpy dtypes) mirror the C structs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

OFFSET_DTYPE = np.dtype([("dx", "<i4"), ("dy", "<i4"), ("flags", "<u4")])
KEYPOINT_DTYPE = np.dtype([("code", "u1", (13,)), ("weight", "u1"), ("x", "<u2"), ("y", "<u2"),
                           ("region_mask", "<u4")], align=True)
PLACEMENT_DTYPE = np.dtype([("frame", "<u4"), ("x", "<i4"), ("y", "<i4")])
BIN_DTYPE = np.dtype([("dx", "<i4"), ("dy", "<i4"), ("cnt", "<u4")])
VOTE_DTYPE = np.dtype([("use_all", "<u4"), ("n_prev", "<u4"), ("n_curr", "<u4"), ("w2_prev", "<u4"),
                       ("w2_curr", "<u4"), ("nbins", "<u4"), ("nticket", "<u4"),
                       ("ticket", BIN_DTYPE, (4,)), ("ngt", "<u4", (4,)), ("nge", "<u4", (4,)), ("hist_hash", "<u4")])
FRAME_DIGEST_DTYPE = np.dtype([("median_hash", "<u8"), ("kp_hash", "<u8"), ("keypoints", "<u4"), ("insertions", "<u4")])
assert KEYPOINT_DTYPE.itemsize == 24 and OFFSET_DTYPE.itemsize == 12 and VOTE_DTYPE.itemsize == 112


class RemapError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"remap_b200 error {code}: {msg}")
        self.code = code


class Registrar:
    """Owns one rb_ctx: an HBM-resident frame store plus the registration kernels.

    Mirrors the reference constants by default: grid 4x2, overlap 16 (src/frc.hpp:22-24),
    weight_switch 10, region_votes 3 (src/frc.hpp:32-33).
    """

    def __init__(self, width, height, max_frames, device=0, compute_median=True, profile=False, stream=None,
                 code_slots=0, offset_slots=0, grid=(4, 2), overlap=16, weight_switch=10, region_votes=3,
                 kpm_mode=0, list_cap=0, run_pairs=0, upload_chunk=0, overlap_batches=0, host_threads=0):
        self._lib = _lib.load()
        cfg = _lib.RbConfig()
        self._lib.rb_default_config(C.byref(cfg), width, height, max_frames)
        cfg.grid_w, cfg.grid_h = grid
        cfg.overlap, cfg.weight_switch, cfg.region_votes = overlap, weight_switch, region_votes
        cfg.device = device
        cfg.compute_median = int(bool(compute_median))
        cfg.profile = int(bool(profile))
        cfg.code_slots, cfg.offset_slots = code_slots, offset_slots
        cfg.stream = stream
        cfg.kpm_mode, cfg.list_cap, cfg.run_pairs, cfg.upload_chunk = kpm_mode, list_cap, run_pairs, upload_chunk
        cfg.overlap_batches = overlap_batches
        cfg.host_threads = host_threads
        self.width, self.height, self.max_frames = width, height, max_frames
        self.nreg = grid[0] * grid[1]
        self._ctx = C.c_void_p()
        rc = self._lib.rb_create(C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.rb_last_error(self._ctx).decode() if self._ctx else "no CUDA device (no CPU fallback exists)"
            if self._ctx:
                self._lib.rb_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise RemapError(rc, msg)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise RemapError(rc, self._lib.rb_last_error(self._ctx).decode())

    # -- the path ---------------------------------------------------------------------------
    def upload(self, frames, first=0):
        """frames: (n, H, W) uint8, values 0..15 (nil::read_raw format)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 3 and frames.shape[1:] == (self.height, self.width), frames.shape
        self._check(self._lib.rb_upload(self._ctx, frames.ctypes.data_as(C.c_void_p), first, frames.shape[0]))
        self._keep = frames  # the copy is asynchronous for pinned memory

    def upload_ptr(self, ptr, n, first=0):
        self._check(self._lib.rb_upload(self._ctx, C.c_void_p(ptr), first, n))

    def register_async(self, n, first=0):
        self._check(self._lib.rb_register_async(self._ctx, first, n))

    def register_host_async(self, frames, first=0):
        """upload + register_async for host frames, copies overlapped with the kernels (rb_register_host_async)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 3 and frames.shape[1:] == (self.height, self.width), frames.shape
        self._check(self._lib.rb_register_host_async(self._ctx, frames.ctypes.data_as(C.c_void_p), first, frames.shape[0]))
        self._keep = frames

    def register_host_packed4(self, packed, first=0):
        """register frames the caller holds as 4 bit/pixel: packed (n, H, row_bytes) uint8, pixel x in nibble x & 1
        of byte x >> 1 (rb_register_host_packed4)."""
        packed = np.ascontiguousarray(packed, np.uint8)
        assert packed.ndim == 3 and packed.shape[1] == self.height and packed.shape[2] >= (self.width + 1) // 2, packed.shape
        self._check(self._lib.rb_register_host_packed4(self._ctx, packed.ctypes.data_as(C.c_void_p), packed.shape[2], first,
                                                       packed.shape[0]))
        self._keep = packed

    @property
    def matcher_kernel(self) -> str:
        return self._lib.rb_matcher_kernel(self._ctx).decode()

    @property
    def host_lane_stats(self):
        """Last register_host_async: chunks sent raw / packed, and the measured rates behind the choice."""
        raw, pk, link, fps, th, nb = C.c_uint64(), C.c_uint64(), C.c_double(), C.c_double(), C.c_int(), C.c_uint64()
        self._check(self._lib.rb_host_lane_stats(self._ctx, C.byref(raw), C.byref(pk), C.byref(link), C.byref(fps), C.byref(th),
                                                 C.byref(nb)))
        return dict(raw_chunks=raw.value, packed_chunks=pk.value, link_GBps=link.value, pack_fps=fps.value, threads=th.value,
                    h2d_bytes=nb.value)

    def fetch_offsets(self, n_pairs, out=None):
        if out is None:
            out = np.zeros(n_pairs, OFFSET_DTYPE)
        self._check(self._lib.rb_fetch_offsets(self._ctx, out.ctypes.data_as(C.c_void_p), n_pairs))
        return out

    def fetch_medians(self, n, first=0):
        out = np.zeros((n, self.height, self.width), np.uint8)
        self._check(self._lib.rb_fetch_medians(self._ctx, first, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def register(self, n, first=0, want_medians=False):
        """-> (offsets (n-1,) OFFSET_DTYPE, medians (n, H, W) or None)"""
        self.register_async(n, first)
        off = self.fetch_offsets(n - 1)
        return off, (self.fetch_medians(n, first) if want_medians else None)

    def synchronize(self):
        self._check(self._lib.rb_synchronize(self._ctx))

    # -- taps ---------------------------------------------------------------------------------
    def keypoints(self, frame):
        cap = self.width * self.height
        out = np.zeros(cap, KEYPOINT_DTYPE)
        n = C.c_size_t()
        self._check(self._lib.rb_keypoints(self._ctx, frame, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return out[:n.value].copy()

    def region_ballots(self, pair):
        out = np.zeros(self.nreg, VOTE_DTYPE)
        self._check(self._lib.rb_region_ballots(self._ctx, pair, out.ctypes.data_as(C.c_void_p)))
        return out

    def fetch_ballots(self, n_pairs, pair=0):
        """(n_pairs, nreg) VOTE_DTYPE: every region ballot of pairs [pair, pair + n_pairs) in one copy."""
        out = np.zeros((n_pairs, self.nreg), VOTE_DTYPE)
        self._check(self._lib.rb_fetch_ballots(self._ctx, pair, n_pairs, out.ctypes.data_as(C.c_void_p)))
        return out

    def frame_digests(self, n, first=0):
        """(n,) FRAME_DIGEST_DTYPE: digests of kpe's outputs per frame (rb_frame_digests)."""
        out = np.zeros(n, FRAME_DIGEST_DTYPE)
        self._check(self._lib.rb_frame_digests(self._ctx, first, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def region_votes(self, pair, region):
        cap = 1 << 20
        out = np.zeros(cap, BIN_DTYPE)
        n = C.c_size_t()
        self._check(self._lib.rb_region_votes(self._ctx, pair, region, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        b = out[:n.value]
        return b[np.lexsort((b["dy"], b["dx"]))].copy()

    def foreground_mask(self, bg, px, py, frame):
        bg = np.ascontiguousarray(bg, np.uint8)
        frame = np.ascontiguousarray(frame, np.uint8)
        assert frame.shape == (self.height, self.width)
        out = np.zeros_like(frame)
        self._check(self._lib.rb_foreground_mask(self._ctx, bg.ctypes.data_as(C.c_void_p), bg.shape[1], bg.shape[0],
                                                 px, py, frame.ctypes.data_as(C.c_void_p),
                                                 out.ctypes.data_as(C.c_void_p)))
        return out

    def foreground_mask_resident(self, bg, px, py, frame_index):
        bg = np.ascontiguousarray(bg, np.uint8)
        out = np.zeros((self.height, self.width), np.uint8)
        self._check(self._lib.rb_foreground_mask_resident(self._ctx, bg.ctypes.data_as(C.c_void_p), bg.shape[1],
                                                          bg.shape[0], px, py, frame_index,
                                                          out.ctypes.data_as(C.c_void_p)))
        return out

    def blit_blend(self, placements, map_w, map_h, want_dots=True):
        """fgm::fragment::blit of the placed resident frames + blend.  placements: PLACEMENT_DTYPE array
        (frame slot, x, y inside the map).  -> (dots (map_h, map_w, 16) uint16 | None, image, mask)"""
        pl = np.ascontiguousarray(placements, PLACEMENT_DTYPE)
        dots = np.zeros((map_h, map_w, 16), np.uint16) if want_dots else None
        image = np.zeros((map_h, map_w), np.uint8)
        mask = np.zeros((map_h, map_w), np.uint8)
        self._check(self._lib.rb_blit_blend(self._ctx, pl.ctypes.data_as(C.c_void_p), len(pl), map_w, map_h,
                                            dots.ctypes.data_as(C.c_void_p) if want_dots else None,
                                            image.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p)))
        return dots, image, mask

    def filter_fragment(self, placements, map_w, map_h, background=None, want_dots=True, want_fgmasks=False):
        """fdf::filter for one fragment (src/fdf.hpp:40-75): per placed frame the pass-2 foreground mask
        (fde::extractor::extract + fde::mask on the frame's median image), then the masked blit and blend.
        background: (map_h, map_w) uint8 or None (= blend of the plain blit of the placements).
        -> dict(dots|None, image, mask, fgmasks (n, H, W)|None, ncontours (n,), times_ms, frames_deferred)"""
        pl = np.ascontiguousarray(placements, PLACEMENT_DTYPE)
        n = len(pl)
        dots = np.zeros((map_h, map_w, 16), np.uint16) if want_dots else None
        image = np.zeros((map_h, map_w), np.uint8)
        mask = np.zeros((map_h, map_w), np.uint8)
        fg = np.zeros((n, self.height, self.width), np.uint8) if want_fgmasks else None
        nc = np.zeros(n, np.uint32)
        bg = None
        if background is not None:
            bg = np.ascontiguousarray(background, np.uint8)
            assert bg.shape == (map_h, map_w)
        vp = C.c_void_p
        self._check(self._lib.rb_filter_fragment(
            self._ctx, pl.ctypes.data_as(vp), n, map_w, map_h, bg.ctypes.data_as(vp) if bg is not None else None,
            dots.ctypes.data_as(vp) if want_dots else None, image.ctypes.data_as(vp), mask.ctypes.data_as(vp),
            fg.ctypes.data_as(vp) if want_fgmasks else None, nc.ctypes.data_as(vp)))
        ms = (C.c_float * 3)()
        deferred = C.c_uint32()
        self._check(self._lib.rb_filter_times(self._ctx, ms, 3, C.byref(deferred)))
        return dict(dots=dots, image=image, mask=mask, fgmasks=fg, ncontours=nc,
                    times_ms=dict(background=ms[0], foreground=ms[1], masked_blit=ms[2]),
                    frames_deferred=int(deferred.value))

    def map_device(self):
        """Device addresses of the map the last blit_blend / filter_fragment left in the context's scratch.
        -> dict(dots, image, mask (ints), width, height)"""
        d, i, m = C.c_void_p(), C.c_void_p(), C.c_void_p()
        w, h = C.c_uint32(), C.c_uint32()
        self._check(self._lib.rb_map_device(self._ctx, C.byref(d), C.byref(i), C.byref(m), C.byref(w), C.byref(h)))
        return dict(dots=d.value, image=i.value, mask=m.value, width=w.value, height=h.value)

    def blend_map(self, want_dots=True):
        """fgm::fragment::blend over the dots currently in the map scratch (after a cross-rank reduction).
        -> (dots | None, image, mask)"""
        md = self.map_device()
        h, w = md["height"], md["width"]
        dots = np.zeros((h, w, 16), np.uint16) if want_dots else None
        image = np.zeros((h, w), np.uint8)
        mask = np.zeros((h, w), np.uint8)
        self._check(self._lib.rb_blend_map(self._ctx, dots.ctypes.data_as(C.c_void_p) if want_dots else None,
                                           image.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p)))
        return dots, image, mask

    def map_export(self) -> bytes:
        """CUDA-IPC handle (rb_map_handle, 80 bytes) of this context's map scratch, for rb_blend_map_peers on
        another rank."""
        buf = (C.c_uint8 * 80)()
        self._check(self._lib.rb_map_export(self._ctx, buf))
        return bytes(buf)

    def blend_map_peers(self, handles, want_dots=True):
        """Sum this rank's partial dot map with the peers' (read in place over NVLink) and blend, in one kernel.
        handles: list of bytes from the other ranks' map_export().  -> (dots | None, image, mask)"""
        md = self.map_device()
        h, w = md["height"], md["width"]
        raw = b"".join(handles)
        arr = (C.c_uint8 * max(len(raw), 1)).from_buffer_copy(raw or b"\0")
        dots = np.zeros((h, w, 16), np.uint16) if want_dots else None
        image = np.zeros((h, w), np.uint8)
        mask = np.zeros((h, w), np.uint8)
        self._check(self._lib.rb_blend_map_peers(self._ctx, arr, len(handles), dots.ctypes.data_as(C.c_void_p) if want_dots else None,
                                                 image.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p)))
        return dots, image, mask

    def sum_map_slice(self, handles, self_index):
        """Reduce-scatter step of the all-links map reduction: sum slice `self_index` over all other ranks' maps."""
        raw = b"".join(handles)
        arr = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        self._check(self._lib.rb_sum_map_slice(self._ctx, arr, len(handles), self_index))

    def blend_map_slices(self, handles, self_index, want_dots=True):
        """Gather step on the destination rank: pull every reduced slice from its rank and blend."""
        md = self.map_device()
        h, w = md["height"], md["width"]
        raw = b"".join(handles)
        arr = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        dots = np.zeros((h, w, 16), np.uint16) if want_dots else None
        image = np.zeros((h, w), np.uint8)
        mask = np.zeros((h, w), np.uint8)
        self._check(self._lib.rb_blend_map_slices(self._ctx, arr, len(handles), self_index,
                                                  dots.ctypes.data_as(C.c_void_p) if want_dots else None,
                                                  image.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p)))
        return dots, image, mask

    def aws_compare(self, n, first=0, heat=None):
        """aws::details::compare (src/aws.hpp:37-60) over every consecutive pair of resident frames
        [first, first + n).  heat: (H, W) uint8 to continue from, or None for aws::scan's initial map of ones
        (src/aws.hpp:113).  -> (heat after all pairs, first_change (H, W) uint32)"""
        h = np.ones((self.height, self.width), np.uint8) if heat is None else np.ascontiguousarray(heat, np.uint8).copy()
        fc = np.zeros((self.height, self.width), np.uint32)
        self._check(self._lib.rb_aws_compare(self._ctx, first, n, h.ctypes.data_as(C.c_void_p), fc.ctypes.data_as(C.c_void_p)))
        return h, fc

    # -- introspection ------------------------------------------------------------------------
    @property
    def stream(self):
        return self._lib.rb_stream(self._ctx)

    def kernel_times(self):
        ms = (C.c_float * 6)()
        self._check(self._lib.rb_kernel_times(self._ctx, ms, 6))
        return dict(kpe_ms=ms[0], kpm_ms=ms[1], declare_ms=ms[2], list_ms=ms[3], match_ms=ms[4], deferred_ms=ms[5])

    @property
    def offsets_device_ptr(self):
        return self._lib.rb_offsets_device(self._ctx)

    def count_keypoints(self, n, first=0):
        t = C.c_uint64()
        self._check(self._lib.rb_count_keypoints(self._ctx, first, n, C.byref(t)))
        return int(t.value)

    @property
    def deferred_count(self):
        """(pair, region) ballots the pipelined matcher handed to the general kernel in the last register."""
        n = C.c_uint32()
        self._check(self._lib.rb_deferred_count(self._ctx, C.byref(n)))
        return int(n.value)

    @property
    def kernel_launches(self):
        return int(self._lib.rb_kernel_launches(self._ctx))

    @property
    def device_bytes(self):
        return int(self._lib.rb_device_bytes(self._ctx))


CELL_MATCH_DTYPE = np.dtype([("valid", "<u4"), ("dx", "<i4"), ("dy", "<i4"), ("matched_keypoints", "<u4"),
                             ("matched_cells", "<u4"), ("active_cells", "<u4"), ("offsets", "<u4"), ("ties", "<u4"),
                             ("pairs", "<u8")])


class Group:
    """rb_group: one process, several GPUs (one rb_ctx per device), contiguous frame ranges with a one-frame
    overlap, pair results gathered on the lead device -- the C-ABI counterpart of shard.py's torchrun path."""

    def __init__(self, width, height, max_frames, devices, **kw):
        self._lib = _lib.load()
        cfg = _lib.RbConfig()
        self._lib.rb_default_config(C.byref(cfg), width, height, max_frames)
        for k, v in kw.items():
            setattr(cfg, k, v)
        self.width, self.height, self.max_frames = width, height, max_frames
        devs = (C.c_int32 * len(devices))(*devices)
        self._g = C.c_void_p()
        rc = self._lib.rb_group_create(C.byref(cfg), devs, len(devices), C.byref(self._g))
        if rc != 0:
            msg = self._lib.rb_group_last_error(self._g).decode() if self._g else "allocation failed"
            if self._g:
                self._lib.rb_group_destroy(self._g)
                self._g = C.c_void_p()
            raise RemapError(rc, msg)

    def close(self):
        if getattr(self, "_g", None):
            self._lib.rb_group_destroy(self._g)
            self._g = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RemapError(rc, self._lib.rb_group_last_error(self._g).decode())

    def __len__(self):
        return int(self._lib.rb_group_size(self._g))

    def register_host(self, frames, out=None):
        """(n, H, W) uint8 host frames -> (n - 1,) OFFSET_DTYPE of the whole sequence."""
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 3 and frames.shape[1:] == (self.height, self.width), frames.shape
        n = frames.shape[0]
        if out is None:
            out = np.zeros(max(n - 1, 0), OFFSET_DTYPE)
        self._check(self._lib.rb_group_register_host(self._g, frames.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)))
        self._keep = frames
        return out

    def range(self, member):
        a, b, o = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._check(self._lib.rb_group_range(self._g, member, C.byref(a), C.byref(b), C.byref(o)))
        return a.value, b.value, o.value

    def fetch_medians(self, n, first=0):
        out = np.zeros((n, self.height, self.width), np.uint8)
        self._check(self._lib.rb_group_fetch_medians(self._g, first, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def member(self, i) -> "Registrar":
        """A borrowed Registrar view of member i's context (do not close it)."""
        r = Registrar.__new__(Registrar)
        r._lib = self._lib
        r._ctx = C.c_void_p(self._lib.rb_group_context(self._g, i))
        r.width, r.height, r.max_frames, r.nreg = self.width, self.height, self.max_frames, 8
        r.close = lambda: None
        return r


class Snippet:
    """fgs::details::extract_single (src/fgs.hpp:80-89) on the device: the blend of a fragment's dot map and the
    keypoints of kpe with a 1 x 1 grid over the whole map image.  dots: (H, W, 16) uint16."""

    def __init__(self, dots, device=0):
        self._lib = _lib.load()
        dots = np.ascontiguousarray(dots, np.uint16)
        assert dots.ndim == 3 and dots.shape[2] == 16
        self.height, self.width = dots.shape[:2]
        self._s = C.c_void_p()
        rc = self._lib.rb_snippet_create(device, dots.ctypes.data_as(C.c_void_p), self.width, self.height, C.byref(self._s))
        if rc != 0:
            msg = self._lib.rb_snippet_last_error(self._s).decode() if self._s else "no CUDA device (no CPU fallback exists)"
            self.close()
            raise RemapError(rc, msg)

    def close(self):
        if getattr(self, "_s", None):
            self._lib.rb_snippet_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise RemapError(rc, self._lib.rb_snippet_last_error(self._s).decode())

    def fetch(self):
        """-> dict(image, mask, kps (KEYPOINT_DTYPE, unordered))"""
        n = C.c_uint32()
        self._check(self._lib.rb_snippet_fetch(self._s, C.byref(n), None, None, None, 0))
        image = np.zeros((self.height, self.width), np.uint8)
        mask = np.zeros((self.height, self.width), np.uint8)
        kps = np.zeros(n.value, KEYPOINT_DTYPE)
        self._check(self._lib.rb_snippet_fetch(self._s, C.byref(n), image.ctypes.data_as(C.c_void_p),
                                               mask.ctypes.data_as(C.c_void_p), kps.ctypes.data_as(C.c_void_p), len(kps)))
        return dict(image=image, mask=mask, kps=kps)

    def match(self, curr, cell=(15, 15)):
        """The cellular kpm::match (src/kpm.hpp:371-393) with self as `previous`.  -> CELL_MATCH_DTYPE record"""
        res = np.zeros(1, CELL_MATCH_DTYPE)
        self._check(self._lib.rb_snippet_match(self._s, curr._s, cell[0], cell[1], res.ctypes.data_as(C.c_void_p)))
        return res[0]
