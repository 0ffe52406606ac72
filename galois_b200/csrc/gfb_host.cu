// gfb_host.cu — host-buffer (`*_host`) variants of the element-wise ufuncs.
//
// These are what the NumPy-facing FieldArray mirror calls: the operands are host ndarrays
// exactly as in the reference's `UFunc.__call__` (/root/reference/src/galois/_domains/_ufunc.py:339-441),
// and the result is written to a host buffer.  See gfb_host.cuh for the pipeline.
#include <map>
#include <memory>
#include <mutex>

#include "gfb_host.cuh"

namespace gfb {

HostPipeline* HostPipeline::get() {
    static thread_local std::map<int, std::unique_ptr<HostPipeline>> pipes;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    auto it = pipes.find(dev);
    if (it != pipes.end()) return it->second.get();
    auto p = std::make_unique<HostPipeline>();
    p->device = dev;
    for (int s = 0; s < kSlots; ++s) {
        if (cudaStreamCreateWithFlags(&p->streams[s], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    }
    HostPipeline* raw = p.get();
    pipes[dev] = std::move(p);
    return raw;
}

void* HostPipeline::buffer(int slot, int idx, size_t bytes) {
    if (caps[slot][idx] >= bytes && bufs[slot][idx]) return bufs[slot][idx];
    if (bufs[slot][idx]) {
        cudaStreamSynchronize(streams[slot]);
        cudaFree(bufs[slot][idx]);
        bufs[slot][idx] = nullptr;
        caps[slot][idx] = 0;
    }
    size_t want = bytes < kChunkBytes ? kChunkBytes : bytes;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        set_last_error("cudaMalloc(staging)", e);
        return nullptr;
    }
    bufs[slot][idx] = p;
    caps[slot][idx] = want;
    return p;
}

cudaError_t HostPipeline::drain() {
    cudaError_t first = cudaSuccess;
    for (int s = 0; s < kSlots; ++s) {
        cudaError_t e = cudaStreamSynchronize(streams[s]);
        if (e != cudaSuccess && first == cudaSuccess) first = e;
    }
    return first;
}

// Generic element-wise host driver.
//   n_in   : number of input arrays (1 or 2), in_bytes[i] = bytes per element of input i
//   launch : issues the device kernel for `count` elements on `stream` with device pointers
template <typename Launch>
static gfb_status ewise_host(gfb_field* f, int n_in, const void* const* in, const size_t* in_bytes, void* out,
                             size_t out_bytes, size_t n, Launch launch) {
    if (!f) return GFB_ERR_BAD_ARG;
    if (n == 0) return GFB_OK;
    for (int i = 0; i < n_in; ++i)
        if (!in[i]) return GFB_ERR_BAD_ARG;
    if (!out) return GFB_ERR_BAD_ARG;
    HostPipeline* hp = HostPipeline::get();
    if (!hp) return GFB_ERR_CUDA;
    DrainOnExit drain_guard(hp);
    size_t widest = out_bytes;
    for (int i = 0; i < n_in; ++i) widest = in_bytes[i] > widest ? in_bytes[i] : widest;
    // chunk length: a multiple of 16 elements keeps every chunk start 16-byte aligned
    size_t chunk = (kChunkBytes / widest) & ~(size_t)15;
    size_t n_chunks = (n + chunk - 1) / chunk;
    gfb_status status = GFB_OK;
    for (size_t c = 0; c < n_chunks && status == GFB_OK; ++c) {
        const int slot = (int)(c % kSlots);
        cudaStream_t st = hp->streams[slot];
        const size_t off = c * chunk;
        const size_t cnt = (n - off) < chunk ? (n - off) : chunk;
        void* din[2] = {nullptr, nullptr};
        for (int i = 0; i < n_in; ++i) {
            din[i] = hp->buffer(slot, i, chunk * in_bytes[i]);
            if (!din[i]) return GFB_ERR_CUDA;
            GFB_CUDA_TRY(cudaMemcpyAsync(din[i], (const char*)in[i] + off * in_bytes[i], cnt * in_bytes[i],
                                         cudaMemcpyHostToDevice, st));
        }
        void* dout = hp->buffer(slot, 2, chunk * out_bytes);
        if (!dout) return GFB_ERR_CUDA;
        status = launch(din[0], din[1], dout, cnt, (void*)st);
        if (status != GFB_OK) break;
        GFB_CUDA_TRY(cudaMemcpyAsync((char*)out + off * out_bytes, dout, cnt * out_bytes, cudaMemcpyDeviceToHost, st));
    }
    cudaError_t e = hp->drain();
    if (e != cudaSuccess) {
        set_last_error("host pipeline drain", e);
        return GFB_ERR_CUDA;
    }
    if (status != GFB_OK) return status;
    int flag = *(volatile int*)f->h_err;
    *(volatile int*)f->h_err = 0;
    return flag ? GFB_ERR_ZERO_DIVISION : GFB_OK;
}

}  // namespace gfb

extern "C" {

gfb_status gfb_binary_host(gfb_field* f, int op, int eb, const void* a, const void* b, void* out, size_t n) {
    const void* in[2] = {a, b};
    size_t ib[2] = {(size_t)eb, (size_t)eb};
    return gfb::ewise_host(f, 2, in, ib, out, (size_t)eb, n,
                           [&](void* da, void* db, void* dout, size_t cnt, void* st) {
                               return gfb_binary(f, op, eb, da, db, dout, cnt, st);
                           });
}

gfb_status gfb_binary_scalar_host(gfb_field* f, int op, int eb, const void* a, uint64_t s_lo, uint64_t s_hi,
                                  int scalar_on_left, void* out, size_t n) {
    const void* in[1] = {a};
    size_t ib[1] = {(size_t)eb};
    return gfb::ewise_host(f, 1, in, ib, out, (size_t)eb, n,
                           [&](void* da, void*, void* dout, size_t cnt, void* st) {
                               return gfb_binary_scalar(f, op, eb, da, s_lo, s_hi, scalar_on_left, dout, cnt, st);
                           });
}

gfb_status gfb_unary_host(gfb_field* f, int op, int eb, const void* a, void* out, size_t n) {
    const void* in[1] = {a};
    size_t ib[1] = {(size_t)eb};
    return gfb::ewise_host(f, 1, in, ib, out, (size_t)eb, n,
                           [&](void* da, void*, void* dout, size_t cnt, void* st) {
                               return gfb_unary(f, op, eb, da, dout, cnt, st);
                           });
}

gfb_status gfb_power_host(gfb_field* f, int eb, const void* a, const int64_t* e, void* out, size_t n) {
    const void* in[2] = {a, e};
    size_t ib[2] = {(size_t)eb, sizeof(int64_t)};
    return gfb::ewise_host(f, 2, in, ib, out, (size_t)eb, n,
                           [&](void* da, void* de, void* dout, size_t cnt, void* st) {
                               return gfb_power(f, eb, da, (const int64_t*)de, dout, cnt, st);
                           });
}

gfb_status gfb_power_scalar_host(gfb_field* f, int eb, const void* a, int64_t e, void* out, size_t n) {
    const void* in[1] = {a};
    size_t ib[1] = {(size_t)eb};
    return gfb::ewise_host(f, 1, in, ib, out, (size_t)eb, n,
                           [&](void* da, void*, void* dout, size_t cnt, void* st) {
                               return gfb_power_scalar(f, eb, da, e, dout, cnt, st);
                           });
}

}  // extern "C"
