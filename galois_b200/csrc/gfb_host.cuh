// gfb_host.cuh — chunked, multi-stream host-buffer pipeline used by every `*_host` entry point.
//
// The reference's FieldArray is a host np.ndarray, so the literal drop-in
// moves every operand over PCIe.  The pipeline splits the work into chunks and rotates them over
// kSlots streams, each with its own device staging buffers, so the H2D copy of chunk c+1, the
// kernel of chunk c and the D2H copy of chunk c-1 overlap (copy engines are full duplex).
#pragma once
#include <cstddef>
#include <cstdint>

#include "gfb_common.cuh"

namespace gfb {

constexpr int kSlots = 3;
constexpr int kMaxBufs = 6;
constexpr size_t kChunkBytes = (size_t)32 << 20;  // per staging buffer

struct HostPipeline {
    int device = -1;
    cudaStream_t streams[kSlots] = {nullptr, nullptr, nullptr};
    void* bufs[kSlots][kMaxBufs] = {};
    size_t caps[kSlots][kMaxBufs] = {};

    // the pipeline of the calling thread for the current device (created on first use)
    static HostPipeline* get();
    // device staging buffer `idx` of `slot`, grown to at least `bytes`
    void* buffer(int slot, int idx, size_t bytes);
    // wait for everything queued on all slots
    cudaError_t drain();
};

// Any early return out of a pipeline loop (a failed copy or launch) must not leave earlier chunks' asynchronous copies
// running against the caller's buffers: every `*_host` function holds one of these for its whole body.
struct DrainOnExit {
    HostPipeline* hp;
    explicit DrainOnExit(HostPipeline* p) : hp(p) {}
    ~DrainOnExit() {
        if (hp) hp->drain();
    }
    DrainOnExit(const DrainOnExit&) = delete;
    DrainOnExit& operator=(const DrainOnExit&) = delete;
};

}  // namespace gfb
