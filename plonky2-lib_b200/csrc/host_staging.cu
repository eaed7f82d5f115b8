// Page-able host memory <-> HBM through page-locked slot rings (see host_staging.h).  Host code only.
#include "host_staging.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>

namespace staging {

// ------------------------------------------------------------------------------------------------
// helper threads
// ------------------------------------------------------------------------------------------------
namespace {

constexpr size_t kPiece = (size_t)256 << 10;   // one unit of work for a helper thread

struct Job {
    std::vector<CopyTask> pieces;
    std::atomic<size_t> next{0};
    std::atomic<size_t> done{0};
    std::mutex mu;
    std::condition_variable cv;
};

// takes pieces until none are left; returns how many this thread copied
size_t work_on(Job& job) {
    size_t mine = 0;
    const size_t total = job.pieces.size();
    for (;;) {
        size_t i = job.next.fetch_add(1, std::memory_order_relaxed);
        if (i >= total) break;
        const CopyTask& t = job.pieces[i];
        std::memcpy(t.dst, t.src, t.bytes);
        mine++;
    }
    if (mine && job.done.fetch_add(mine, std::memory_order_acq_rel) + mine == total) {
        std::lock_guard<std::mutex> lk(job.mu);
        job.cv.notify_all();
    }
    return mine;
}

class Pool {
public:
    static Pool& get() {
        static Pool* p = new Pool();   // never destroyed: helper threads may outlive static destructors
        return *p;
    }
    void run(const std::shared_ptr<Job>& job) {
        if (!threads_.empty() && job->pieces.size() > 1) {
            {
                std::lock_guard<std::mutex> lk(mu_);
                jobs_.push_back(job);
            }
            cv_.notify_all();
        }
        work_on(*job);
        std::unique_lock<std::mutex> lk(job->mu);
        job->cv.wait(lk, [&] { return job->done.load(std::memory_order_acquire) == job->pieces.size(); });
    }

private:
    Pool() {
        unsigned hw = std::thread::hardware_concurrency();
        unsigned want = hw / 2 < 8 ? hw / 2 : 8;
        if (const char* s = std::getenv("GL_B200_HOST_THREADS")) {
            int v = std::atoi(s);
            if (v >= 1 && v <= 64) want = (unsigned)v;
        }
        if (want < 1) want = 1;
        for (unsigned i = 1; i < want; i++) {   // the submitting thread is worker 0
            threads_.emplace_back([this] { loop(); });
            threads_.back().detach();
        }
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] {
                    while (!jobs_.empty() && jobs_.front()->next.load(std::memory_order_relaxed) >= jobs_.front()->pieces.size())
                        jobs_.pop_front();
                    return !jobs_.empty();
                });
                job = jobs_.front();
            }
            work_on(*job);
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> jobs_;
    std::vector<std::thread> threads_;
};

}  // namespace

void parallel_copy(const CopyTask* tasks, size_t count) {
    size_t total = 0;
    for (size_t i = 0; i < count; i++) total += tasks[i].bytes;
    if (total == 0) return;
    if (total <= kPiece) {
        for (size_t i = 0; i < count; i++) std::memcpy(tasks[i].dst, tasks[i].src, tasks[i].bytes);
        return;
    }
    auto job = std::make_shared<Job>();
    for (size_t i = 0; i < count; i++) {
        size_t off = 0;
        while (off < tasks[i].bytes) {
            size_t len = tasks[i].bytes - off < kPiece ? tasks[i].bytes - off : kPiece;
            job->pieces.push_back({(char*)tasks[i].dst + off, (const char*)tasks[i].src + off, len});
            off += len;
        }
    }
    Pool::get().run(job);
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// ------------------------------------------------------------------------------------------------
// slot ring
// ------------------------------------------------------------------------------------------------
cudaError_t Ring::init() {
    if (slot[0]) return cudaSuccess;
    char* base = nullptr;
    cudaError_t e = cudaHostAlloc((void**)&base, kSlotBytes * kSlots, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < kSlots; i++) {
        slot[i] = base + (size_t)i * kSlotBytes;
        e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        used[i] = false;
    }
    next = 0;
    return cudaSuccess;
}

void Ring::destroy() {
    if (!slot[0]) return;
    for (int i = 0; i < kSlots; i++)
        if (ev[i]) cudaEventDestroy(ev[i]);
    cudaFreeHost(slot[0]);
    for (int i = 0; i < kSlots; i++) {
        slot[i] = nullptr;
        ev[i] = nullptr;
    }
}

cudaError_t h2d_gather(Ring& ring, void* dev_dst, const HostSeg* segs, size_t count, cudaStream_t stream) {
    cudaError_t e = ring.init();
    if (e != cudaSuccess) return e;
    char* dst = (char*)dev_dst;
    std::vector<CopyTask> tasks;
    size_t fill = 0;          // bytes packed into the current slot
    auto flush = [&]() -> cudaError_t {
        if (!fill) return cudaSuccess;
        const int s = ring.next;
        parallel_copy(tasks.data(), tasks.size());
        cudaError_t err = cudaMemcpyAsync(dst, ring.slot[s], fill, cudaMemcpyHostToDevice, stream);
        if (err == cudaSuccess) err = cudaEventRecord(ring.ev[s], stream);
        ring.used[s] = true;
        ring.next = (s + 1) % kSlots;
        dst += fill;
        fill = 0;
        tasks.clear();
        return err;
    };
    auto open_slot = [&]() -> cudaError_t {   // the slot about to be packed must have left the host
        const int s = ring.next;
        return ring.used[s] ? cudaEventSynchronize(ring.ev[s]) : cudaSuccess;
    };
    for (size_t i = 0; i < count; i++) {
        const char* src = (const char*)segs[i].ptr;
        size_t left = segs[i].bytes;
        if (left >= kSlotBytes && is_pinned(src)) {   // page-locked caller memory: no staging
            if ((e = flush()) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(dst, src, left, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
            dst += left;
            continue;
        }
        while (left) {
            if (fill == 0 && (e = open_slot()) != cudaSuccess) return e;
            size_t len = kSlotBytes - fill < left ? kSlotBytes - fill : left;
            tasks.push_back({ring.slot[ring.next] + fill, src, len});
            fill += len;
            src += len;
            left -= len;
            if (fill == kSlotBytes && (e = flush()) != cudaSuccess) return e;
        }
    }
    return flush();
}

// ------------------------------------------------------------------------------------------------
// D2H worker
// ------------------------------------------------------------------------------------------------
Downloader::~Downloader() {
    if (thread_.joinable()) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        thread_.join();
    }
    ring_.destroy();
}

void Downloader::submit(const void* dev_src, std::vector<HostSeg> segs, cudaEvent_t ready) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (!thread_.joinable()) thread_ = std::thread([this] { loop(); });
        queue_.push_back(Request{(const char*)dev_src, std::move(segs), ready});
        pending_++;
    }
    cv_.notify_all();
}

cudaError_t Downloader::wait() {
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    cudaError_t e = error_;
    error_ = cudaSuccess;
    return e;
}

void Downloader::loop() {
    cudaSetDevice(device_);
    for (;;) {
        Request r;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || !queue_.empty(); });
            if (queue_.empty()) return;   // stop_ and drained
            r = std::move(queue_.front());
            queue_.pop_front();
        }
        cudaError_t e = serve(r);
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (e != cudaSuccess && error_ == cudaSuccess) error_ = e;
            pending_--;
        }
        done_cv_.notify_all();
    }
}

// DMA chunk k+1 .. k+kSlots-1 are in flight while chunk k is unpacked into the caller's arrays.
cudaError_t Downloader::serve(Request& r) {
    cudaError_t e = ring_.init();
    if (e != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(stream_, r.ready, 0)) != cudaSuccess) return e;
    size_t total = 0;
    for (auto& s : r.segs) total += s.bytes;
    const size_t chunks = (total + kSlotBytes - 1) / kSlotBytes;
    size_t issued = 0;
    size_t seg = 0, seg_off = 0;   // unpack cursor
    std::vector<CopyTask> tasks;
    for (size_t k = 0; k < chunks; k++) {
        while (issued < chunks && issued < k + kSlots) {
            const size_t off = issued * kSlotBytes;
            const size_t len = total - off < kSlotBytes ? total - off : kSlotBytes;
            const int s = (int)(issued % kSlots);
            if ((e = cudaMemcpyAsync(ring_.slot[s], r.src + off, len, cudaMemcpyDeviceToHost, stream_)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(ring_.ev[s], stream_)) != cudaSuccess) return e;
            issued++;
        }
        const int s = (int)(k % kSlots);
        if ((e = cudaEventSynchronize(ring_.ev[s])) != cudaSuccess) return e;
        size_t left = total - k * kSlotBytes < kSlotBytes ? total - k * kSlotBytes : kSlotBytes;
        const char* from = ring_.slot[s];
        tasks.clear();
        while (left) {
            HostSeg& hs = r.segs[seg];
            size_t len = hs.bytes - seg_off < left ? hs.bytes - seg_off : left;
            if (len) tasks.push_back({(char*)hs.ptr + seg_off, from, len});
            from += len;
            seg_off += len;
            left -= len;
            if (seg_off == hs.bytes) {
                seg++;
                seg_off = 0;
            }
        }
        parallel_copy(tasks.data(), tasks.size());
    }
    return cudaSuccess;
}

}  // namespace staging
