// gfb_host.cuh — chunked, multi-stream host-buffer pipeline used by every `*_host` entry point.
//
// The reference's FieldArray is a host np.ndarray, so the literal drop-in
// moves every operand over PCIe.  The pipeline splits the work into chunks and rotates them over
// kSlots streams, each with its own device staging buffers, so the H2D copy of chunk c+1, the
// kernel of chunk c and the D2H copy of chunk c-1 overlap (copy engines are full duplex).
//
// Host buffers come in two kinds:
//   * page-locked (cudaHostAlloc / gfb_host_alloc): DMA'd directly;
//   * ordinary pageable memory (what a galois FieldArray lives in): the driver would bounce every
//     cudaMemcpyAsync through its own small pinned buffer on the calling thread (~15 GB/s, no
//     overlap).  Instead the chunk is copied by a pool of host threads into a pinned staging
//     buffer of the slot (memory-bandwidth bound, several times PCIe) and DMA'd from there while
//     the threads already fill the next slot; results take the reverse route.
#pragma once
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <vector>

#include "gfb_common.cuh"

namespace gfb {

constexpr int kSlots = 4;
constexpr int kMaxBufs = 6;
constexpr size_t kChunkBytes = (size_t)32 << 20;  // per staging buffer

// dst[0..bytes) = src[0..bytes) on the host-thread pool (GFB_COPY_THREADS, default min(16, cores/4))
void parallel_memcpy(void* dst, const void* src, size_t bytes);
// true if `p` is page-locked memory known to the CUDA driver (cudaHostAlloc / cudaHostRegister)
bool is_pinned_host(const void* p);

struct HostPipeline {
    int device = -1;
    std::mutex mu;  // one `*_host` call at a time per device (the pipeline is shared by all threads)
    cudaStream_t streams[kSlots] = {};
    void* bufs[kSlots][kMaxBufs] = {};     // device staging
    size_t caps[kSlots][kMaxBufs] = {};
    void* hbufs[kSlots][kMaxBufs] = {};    // pinned host staging (pageable user buffers only)
    size_t hcaps[kSlots][kMaxBufs] = {};
    struct PendingOut {
        void* dst;
        const void* staging;
        size_t bytes;
    };
    std::vector<PendingOut> pending[kSlots];  // D2H results parked in host staging, to be copied out
    bool host_staged[kSlots] = {};            // the slot's host staging buffers are in flight

    ~HostPipeline();
    // the pipeline of the current device (created on first use; shared, guarded by `mu`)
    static HostPipeline* get();
    // device staging buffer `idx` of `slot`, grown to at least `bytes`
    void* buffer(int slot, int idx, size_t bytes);
    // call before queueing a new chunk on `slot`: waits for the slot's previous chunk when it used host staging
    // (so the staging buffers can be refilled) and copies its parked results out
    cudaError_t begin_chunk(int slot);
    // host -> device staging buffer (through pinned host staging when `src_pinned` is false)
    cudaError_t copy_in(int slot, int idx, void* dev_dst, const void* host_src, size_t bytes, bool src_pinned);
    // device -> host (through pinned host staging when `dst_pinned` is false; finished by begin_chunk / drain)
    cudaError_t copy_out(int slot, int idx, void* host_dst, const void* dev_src, size_t bytes, bool dst_pinned);
    // wait for everything queued on all slots and flush parked results
    cudaError_t drain();

  private:
    void* host_buffer(int slot, int idx, size_t bytes);
    void flush(int slot);
};

// Holds the pipeline for the whole body of a `*_host` function: serialises callers and guarantees that any early
// return (a failed copy or launch) does not leave earlier chunks' asynchronous copies running against the caller's
// buffers.
struct PipelineScope {
    HostPipeline* hp;
    std::unique_lock<std::mutex> lock;
    explicit PipelineScope(HostPipeline* p) : hp(p) {
        if (hp) lock = std::unique_lock<std::mutex>(hp->mu);
    }
    ~PipelineScope() {
        if (hp) hp->drain();
    }
    PipelineScope(const PipelineScope&) = delete;
    PipelineScope& operator=(const PipelineScope&) = delete;
};

}  // namespace gfb
