// gfb_host.cu — host-buffer (`*_host`) variants of the element-wise ufuncs.
//
// These are what the NumPy-facing FieldArray mirror calls: the operands are host ndarrays
// exactly as in the reference's `UFunc.__call__` (/root/reference/src/galois/_domains/_ufunc.py:339-441),
// and the result is written to a host buffer.  See gfb_host.cuh for the pipeline.
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "gfb_host.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// host-thread pool for pageable <-> pinned staging copies
// ------------------------------------------------------------------------------------------------
namespace {

class CopyPool {
  public:
    static CopyPool& instance() {
        static CopyPool pool;
        return pool;
    }
    void run(char* dst, const char* src, size_t bytes) {
        const size_t kMinSlice = (size_t)1 << 20;
        if (workers_.empty() || bytes < 2 * kMinSlice) {
            memcpy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mu_);
        const size_t parts_max = workers_.size() + 1;
        size_t parts = bytes / kMinSlice;
        if (parts > parts_max) parts = parts_max;
        size_t slice = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst;
            src_ = src;
            bytes_ = bytes;
            slice_ = slice;
            next_ = 1;  // the caller copies slice 0
            parts_ = (bytes + slice - 1) / slice;
            remaining_ = parts_ - 1;
            ++generation_;
        }
        cv_work_.notify_all();
        memcpy(dst, src, slice < bytes ? slice : bytes);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return remaining_ == 0; });
    }

  private:
    CopyPool() {
        int n = 0;
        if (const char* e = getenv("GFB_COPY_THREADS")) n = atoi(e);
        if (n <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            n = (int)(hw / 4);
            if (n > 16) n = 16;
        }
        if (n < 1) n = 1;
        for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void loop() {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_work_.wait(lk, [&] { return stop_ || (generation_ != seen && next_ < parts_); });
            if (stop_) return;
            while (next_ < parts_) {
                const size_t part = next_++;
                const size_t off = part * slice_;
                const size_t len = bytes_ - off < slice_ ? bytes_ - off : slice_;
                char* d = dst_ + off;
                const char* s = src_ + off;
                lk.unlock();
                memcpy(d, s, len);
                lk.lock();
                if (--remaining_ == 0) cv_done_.notify_all();
            }
            seen = generation_;
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_work_, cv_done_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0, next_ = 0, parts_ = 0, remaining_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

}  // namespace

void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    if (bytes) CopyPool::instance().run((char*)dst, (const char*)src, bytes);
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();  // unregistered memory reports an error on older drivers: clear it
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// ------------------------------------------------------------------------------------------------
// the pipeline
// ------------------------------------------------------------------------------------------------
HostPipeline* HostPipeline::get() {
    static std::mutex registry_mu;
    static std::map<int, std::unique_ptr<HostPipeline>> pipes;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(registry_mu);
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

HostPipeline::~HostPipeline() {
    // process teardown: the CUDA context may already be gone, so failures are ignored
    for (int s = 0; s < kSlots; ++s) {
        for (int i = 0; i < kMaxBufs; ++i) {
            if (bufs[s][i]) cudaFree(bufs[s][i]);
            if (hbufs[s][i]) cudaFreeHost(hbufs[s][i]);
        }
        if (streams[s]) cudaStreamDestroy(streams[s]);
    }
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

void* HostPipeline::host_buffer(int slot, int idx, size_t bytes) {
    if (hcaps[slot][idx] >= bytes && hbufs[slot][idx]) return hbufs[slot][idx];
    if (hbufs[slot][idx]) {
        cudaStreamSynchronize(streams[slot]);
        flush(slot);
        cudaFreeHost(hbufs[slot][idx]);
        hbufs[slot][idx] = nullptr;
        hcaps[slot][idx] = 0;
    }
    size_t want = bytes < kChunkBytes ? kChunkBytes : bytes;
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_last_error("cudaHostAlloc(staging)", e);
        return nullptr;
    }
    hbufs[slot][idx] = p;
    hcaps[slot][idx] = want;
    return p;
}

void HostPipeline::flush(int slot) {
    for (const PendingOut& po : pending[slot]) parallel_memcpy(po.dst, po.staging, po.bytes);
    pending[slot].clear();
    host_staged[slot] = false;
}

cudaError_t HostPipeline::begin_chunk(int slot) {
    if (!host_staged[slot]) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(streams[slot]);
    flush(slot);
    return e;
}

cudaError_t HostPipeline::copy_in(int slot, int idx, void* dev_dst, const void* host_src, size_t bytes,
                                  bool src_pinned) {
    if (bytes == 0) return cudaSuccess;
    if (!src_pinned) {
        void* h = host_buffer(slot, idx, bytes);
        if (!h) return cudaErrorMemoryAllocation;
        parallel_memcpy(h, host_src, bytes);
        host_staged[slot] = true;
        host_src = h;
    }
    return cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, streams[slot]);
}

cudaError_t HostPipeline::copy_out(int slot, int idx, void* host_dst, const void* dev_src, size_t bytes,
                                   bool dst_pinned) {
    if (bytes == 0) return cudaSuccess;
    if (dst_pinned) return cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, streams[slot]);
    void* h = host_buffer(slot, idx, bytes);
    if (!h) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaMemcpyAsync(h, dev_src, bytes, cudaMemcpyDeviceToHost, streams[slot]);
    if (e == cudaSuccess) {
        pending[slot].push_back({host_dst, h, bytes});
        host_staged[slot] = true;
    }
    return e;
}

cudaError_t HostPipeline::drain() {
    cudaError_t first = cudaSuccess;
    for (int s = 0; s < kSlots; ++s) {
        cudaError_t e = cudaStreamSynchronize(streams[s]);
        if (e != cudaSuccess && first == cudaSuccess) first = e;
        if (e == cudaSuccess)
            flush(s);
        else {
            pending[s].clear();
            host_staged[s] = false;
        }
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
    PipelineScope scope(hp);
    size_t widest = out_bytes;
    for (int i = 0; i < n_in; ++i) widest = in_bytes[i] > widest ? in_bytes[i] : widest;
    bool in_pinned[2] = {true, true};
    for (int i = 0; i < n_in; ++i) in_pinned[i] = is_pinned_host(in[i]);
    const bool out_pinned = is_pinned_host(out);
    // chunk length: a multiple of 16 elements keeps every chunk start 16-byte aligned
    size_t chunk = (kChunkBytes / widest) & ~(size_t)15;
    size_t n_chunks = (n + chunk - 1) / chunk;
    gfb_status status = GFB_OK;
    for (size_t c = 0; c < n_chunks && status == GFB_OK; ++c) {
        const int slot = (int)(c % kSlots);
        cudaStream_t st = hp->streams[slot];
        const size_t off = c * chunk;
        const size_t cnt = (n - off) < chunk ? (n - off) : chunk;
        GFB_CUDA_TRY(hp->begin_chunk(slot));
        void* din[2] = {nullptr, nullptr};
        for (int i = 0; i < n_in; ++i) {
            din[i] = hp->buffer(slot, i, chunk * in_bytes[i]);
            if (!din[i]) return GFB_ERR_CUDA;
            GFB_CUDA_TRY(hp->copy_in(slot, i, din[i], (const char*)in[i] + off * in_bytes[i], cnt * in_bytes[i],
                                     in_pinned[i]));
        }
        void* dout = hp->buffer(slot, 2, chunk * out_bytes);
        if (!dout) return GFB_ERR_CUDA;
        status = launch(din[0], din[1], dout, cnt, (void*)st);
        if (status != GFB_OK) break;
        GFB_CUDA_TRY(hp->copy_out(slot, 2, (char*)out + off * out_bytes, dout, cnt * out_bytes, out_pinned));
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
