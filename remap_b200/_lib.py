"""Loads libremap_b200.so (the C ABI of include/remap_b200.h) with ctypes.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded this module raises,
and rb_create fails with RB_ERR_NO_DEVICE on a machine without a usable GPU.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libremap_b200.so")

RB_OK = 0
RB_ERR_INVALID, RB_ERR_NO_DEVICE, RB_ERR_CUDA, RB_ERR_CAPACITY, RB_ERR_STATE = -1, -2, -3, -4, -5
RB_OFFSET_VALID = 1
RB_OFFSET_TIE_SENSITIVE = 2

# every symbol include/remap_b200.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = [
    "rb_default_config", "rb_create", "rb_destroy", "rb_upload", "rb_register_async", "rb_fetch_offsets",
    "rb_fetch_medians", "rb_register", "rb_keypoints", "rb_region_ballots", "rb_region_votes",
    "rb_foreground_mask", "rb_foreground_mask_resident", "rb_synchronize", "rb_stream", "rb_kernel_times",
    "rb_kernel_launches", "rb_device_bytes", "rb_last_error", "rb_abi_version", "rb_offsets_device",
    "rb_count_keypoints", "rb_alloc_host", "rb_free_host", "rb_deferred_count", "rb_register_host_async", "rb_blit_blend",
    "rb_filter_fragment", "rb_filter_times", "rb_upload_medians",
    "rb_aws_compare", "rb_map_device", "rb_blend_map", "rb_map_export", "rb_blend_map_peers", "rb_sum_map_slice", "rb_blend_map_slices", "rb_snippet_create", "rb_snippet_destroy", "rb_snippet_last_error", "rb_snippet_fetch", "rb_snippet_match",
    "rb_host_lane_stats", "rb_register_host_packed4", "rb_frame_digests", "rb_fetch_ballots", "rb_matcher_kernel",
    "rb_snippet_merge", "rb_snippet_fetch_dots", "rb_group_create", "rb_group_destroy", "rb_group_last_error", "rb_group_size", "rb_group_context", "rb_group_register_host",
    "rb_group_range", "rb_group_locate", "rb_group_fetch_medians", "rb_group_offsets_device",
]


class RbConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("grid_w", C.c_uint32), ("grid_h", C.c_uint32),
                ("overlap", C.c_uint32), ("weight_switch", C.c_uint32), ("region_votes", C.c_uint32),
                ("device", C.c_int32), ("max_frames", C.c_uint32), ("compute_median", C.c_uint32),
                ("code_slots", C.c_uint32), ("offset_slots", C.c_uint32), ("profile", C.c_uint32),
                ("stream", C.c_void_p), ("kpm_mode", C.c_uint32), ("list_cap", C.c_uint32),
                ("run_pairs", C.c_uint32), ("upload_chunk", C.c_uint32), ("overlap_batches", C.c_uint32),
                ("host_threads", C.c_uint32)]


class RemapLibraryMissing(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False):
    """Returns the ctypes handle of libremap_b200.so; raises RemapLibraryMissing if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build()
        else:
            raise RemapLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m remap_b200.build` (nvcc, sm_100a). "
                "remap_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32
    lib.rb_default_config.restype = None
    lib.rb_default_config.argtypes = [C.POINTER(RbConfig), u32, u32, u32]
    lib.rb_create.restype = C.c_int
    lib.rb_create.argtypes = [C.POINTER(RbConfig), C.POINTER(vp)]
    lib.rb_destroy.restype = None
    lib.rb_destroy.argtypes = [vp]
    lib.rb_upload.restype = C.c_int
    lib.rb_upload.argtypes = [vp, vp, sz, sz]
    lib.rb_register_async.restype = C.c_int
    lib.rb_register_async.argtypes = [vp, sz, sz]
    lib.rb_register_host_async.restype = C.c_int
    lib.rb_register_host_async.argtypes = [vp, vp, sz, sz]
    lib.rb_register_host_packed4.restype = C.c_int
    lib.rb_register_host_packed4.argtypes = [vp, vp, sz, sz, sz]
    lib.rb_host_lane_stats.restype = C.c_int
    lib.rb_host_lane_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    lib.rb_frame_digests.restype = C.c_int
    lib.rb_frame_digests.argtypes = [vp, sz, sz, vp]
    lib.rb_fetch_ballots.restype = C.c_int
    lib.rb_fetch_ballots.argtypes = [vp, sz, sz, vp]
    lib.rb_blit_blend.restype = C.c_int
    lib.rb_blit_blend.argtypes = [vp, vp, sz, u32, u32, vp, vp, vp]
    lib.rb_filter_fragment.restype = C.c_int
    lib.rb_filter_fragment.argtypes = [vp, vp, sz, u32, u32, vp, vp, vp, vp, vp, vp]
    lib.rb_snippet_create.restype = C.c_int
    lib.rb_snippet_create.argtypes = [i32, vp, u32, u32, C.POINTER(vp)]
    lib.rb_snippet_merge.restype = C.c_int
    lib.rb_snippet_merge.argtypes = [vp, u32, u32, vp, u32, u32, u32, u32, C.POINTER(vp)]
    lib.rb_snippet_fetch_dots.restype = C.c_int
    lib.rb_snippet_fetch_dots.argtypes = [vp, vp]
    lib.rb_snippet_destroy.restype = None
    lib.rb_snippet_destroy.argtypes = [vp]
    lib.rb_snippet_last_error.restype = C.c_char_p
    lib.rb_snippet_last_error.argtypes = [vp]
    lib.rb_snippet_fetch.restype = C.c_int
    lib.rb_snippet_fetch.argtypes = [vp, C.POINTER(u32), vp, vp, vp, sz]
    lib.rb_snippet_match.restype = C.c_int
    lib.rb_snippet_match.argtypes = [vp, vp, u32, u32, vp]
    lib.rb_map_device.restype = C.c_int
    lib.rb_map_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u32), C.POINTER(u32)]
    lib.rb_map_export.restype = C.c_int
    lib.rb_map_export.argtypes = [vp, vp]
    lib.rb_blend_map_peers.restype = C.c_int
    lib.rb_blend_map_peers.argtypes = [vp, vp, sz, vp, vp, vp]
    lib.rb_sum_map_slice.restype = C.c_int
    lib.rb_sum_map_slice.argtypes = [vp, vp, sz, sz]
    lib.rb_blend_map_slices.restype = C.c_int
    lib.rb_blend_map_slices.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    lib.rb_blend_map.restype = C.c_int
    lib.rb_blend_map.argtypes = [vp, vp, vp, vp]
    lib.rb_aws_compare.restype = C.c_int
    lib.rb_aws_compare.argtypes = [vp, sz, sz, vp, vp]
    lib.rb_upload_medians.restype = C.c_int
    lib.rb_upload_medians.argtypes = [vp, vp, sz, sz]
    lib.rb_filter_times.restype = C.c_int
    lib.rb_filter_times.argtypes = [vp, C.POINTER(C.c_float), sz, C.POINTER(u32)]
    lib.rb_fetch_offsets.restype = C.c_int
    lib.rb_fetch_offsets.argtypes = [vp, vp, sz]
    lib.rb_fetch_medians.restype = C.c_int
    lib.rb_fetch_medians.argtypes = [vp, sz, sz, vp]
    lib.rb_register.restype = C.c_int
    lib.rb_register.argtypes = [vp, sz, sz, vp, vp]
    lib.rb_keypoints.restype = C.c_int
    lib.rb_keypoints.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
    lib.rb_region_ballots.restype = C.c_int
    lib.rb_region_ballots.argtypes = [vp, sz, vp]
    lib.rb_region_votes.restype = C.c_int
    lib.rb_region_votes.argtypes = [vp, sz, u32, vp, sz, C.POINTER(sz)]
    lib.rb_foreground_mask.restype = C.c_int
    lib.rb_foreground_mask.argtypes = [vp, vp, u32, u32, i32, i32, vp, vp]
    lib.rb_foreground_mask_resident.restype = C.c_int
    lib.rb_foreground_mask_resident.argtypes = [vp, vp, u32, u32, i32, i32, sz, vp]
    lib.rb_synchronize.restype = C.c_int
    lib.rb_synchronize.argtypes = [vp]
    lib.rb_stream.restype = vp
    lib.rb_stream.argtypes = [vp]
    lib.rb_kernel_times.restype = C.c_int
    lib.rb_kernel_times.argtypes = [vp, C.POINTER(C.c_float), sz]
    lib.rb_kernel_launches.restype = C.c_uint64
    lib.rb_kernel_launches.argtypes = [vp]
    lib.rb_device_bytes.restype = sz
    lib.rb_device_bytes.argtypes = [vp]
    lib.rb_last_error.restype = C.c_char_p
    lib.rb_last_error.argtypes = [vp]
    lib.rb_group_create.restype = C.c_int
    lib.rb_group_create.argtypes = [C.POINTER(RbConfig), C.POINTER(i32), sz, C.POINTER(vp)]
    lib.rb_group_destroy.restype = None
    lib.rb_group_destroy.argtypes = [vp]
    lib.rb_group_last_error.restype = C.c_char_p
    lib.rb_group_last_error.argtypes = [vp]
    lib.rb_group_size.restype = sz
    lib.rb_group_size.argtypes = [vp]
    lib.rb_group_context.restype = vp
    lib.rb_group_context.argtypes = [vp, sz]
    lib.rb_group_register_host.restype = C.c_int
    lib.rb_group_register_host.argtypes = [vp, vp, sz, vp]
    lib.rb_group_range.restype = C.c_int
    lib.rb_group_range.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
    lib.rb_group_locate.restype = C.c_int
    lib.rb_group_locate.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz)]
    lib.rb_group_fetch_medians.restype = C.c_int
    lib.rb_group_fetch_medians.argtypes = [vp, sz, sz, vp]
    lib.rb_group_offsets_device.restype = vp
    lib.rb_group_offsets_device.argtypes = [vp]
    lib.rb_matcher_kernel.restype = C.c_char_p
    lib.rb_matcher_kernel.argtypes = [vp]
    lib.rb_offsets_device.restype = vp
    lib.rb_offsets_device.argtypes = [vp]
    lib.rb_count_keypoints.restype = C.c_int
    lib.rb_count_keypoints.argtypes = [vp, sz, sz, C.POINTER(C.c_uint64)]
    lib.rb_deferred_count.restype = C.c_int
    lib.rb_deferred_count.argtypes = [vp, C.POINTER(u32)]
    lib.rb_alloc_host.restype = vp
    lib.rb_alloc_host.argtypes = [sz]
    lib.rb_free_host.restype = None
    lib.rb_free_host.argtypes = [vp]
    lib.rb_abi_version.restype = u32
    lib.rb_abi_version.argtypes = []
    _lib = lib
    return lib
