// rb_hostpack.hpp -- see rb_hostpack.cpp
#pragma once
#include <stddef.h>
#include <stdint.h>

// frames: n * H rows of W bytes (dense) -> dst: n * H rows of pitch4 bytes, 4 bit/pixel, on `threads` host threads
// (<= 0: all processors).  The context passes its share of the host (rb_api.cu, packer_threads).
// C linkage only so that the CPU test-suite can reach it through ctypes; not part of the public ABI.
extern "C" void rb_hostpack_frames(const uint8_t* frames, uint32_t W, uint32_t H, size_t n, uint8_t* dst, uint32_t pitch4,
                                   int threads);

// The same; `pre(arg)` (may be null) runs on the calling thread at the start, while the other threads already pack.
extern "C" void rb_hostpack_frames_cb(const uint8_t* frames, uint32_t W, uint32_t H, size_t n, uint8_t* dst, uint32_t pitch4,
                                      int threads, void (*pre)(void*), void* arg);
