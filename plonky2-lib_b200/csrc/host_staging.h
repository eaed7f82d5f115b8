// Page-able host memory <-> HBM at PCIe speed.
//
// The reference's buffers are ordinary Rust `Vec`s: one per polynomial (`Vec<PolynomialValues<F>>`,
// fri/oracle.rs from_values), page-able.  A cudaMemcpyAsync from such memory is staged by the driver on the
// calling thread at 6-10 GB/s and serialises with everything else that thread would enqueue.  Here the
// staging is explicit: a small pool of helper threads copies between the caller's arrays and a ring of
// page-locked slots, and the DMA engine moves the slots, so both PCIe directions run at link speed while
// the transforms run (SURVEY.md 8b "Host buffers may be pageable; library pins/stages internally").
//
//   H2D: the calling thread packs the next slot (parallel memcpy), enqueues its DMA and goes on.
//   D2H: a per-context worker thread waits for "data ready" events, DMAs into slots and unpacks them
//        into the caller's arrays; the calling thread only waits for it at the end of the entry point.
//
// Page-locked caller memory (gl_host_alloc, cudaHostRegister) bypasses all of this.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace staging {

struct CopyTask {
    void* dst;
    const void* src;
    size_t bytes;
};

// Process-wide helper threads (GL_B200_HOST_THREADS, default min(8, cores / 2)); the caller takes part.
void parallel_copy(const CopyTask* tasks, size_t count);

// true when `p` is page-locked host memory the DMA engine can use directly
bool is_pinned(const void* p);

struct HostSeg {  // one caller array; consecutive segments map to consecutive device bytes
    void* ptr;
    size_t bytes;
};

constexpr size_t kSlotBytes = (size_t)4 << 20;
constexpr int kSlots = 4;

struct Ring {
    char* slot[kSlots] = {};
    cudaEvent_t ev[kSlots] = {};
    bool used[kSlots] = {};
    int next = 0;
    cudaError_t init();
    void destroy();
};

// Copies the segments to `dev_dst` (contiguously) through `ring` on `stream`.  Returns once every byte has
// left the caller's arrays (the last DMAs may still be in flight: they read the ring, not the caller).
cudaError_t h2d_gather(Ring& ring, void* dev_dst, const HostSeg* segs, size_t count, cudaStream_t stream);

// D2H worker of one context.
class Downloader {
public:
    Downloader(int device, cudaStream_t stream) : device_(device), stream_(stream) {}
    ~Downloader();
    // `ready` must already be recorded; the segments receive dev_src[0 .. sum(bytes)) in order.
    void submit(const void* dev_src, std::vector<HostSeg> segs, cudaEvent_t ready);
    // blocks until every submitted download has landed in the caller's arrays
    cudaError_t wait();

private:
    struct Request {
        const char* src;
        std::vector<HostSeg> segs;
        cudaEvent_t ready;
    };
    void loop();
    cudaError_t serve(Request& r);
    int device_;
    cudaStream_t stream_;
    Ring ring_;
    bool ring_ok_ = false;
    std::thread thread_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Request> queue_;
    size_t pending_ = 0;
    bool stop_ = false;
    cudaError_t error_ = cudaSuccess;
};

}  // namespace staging
