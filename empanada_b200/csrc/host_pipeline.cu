// host_pipeline.cu — emp_panoptic_batched_host: the fused get_panoptic_segmentation pipeline
// (postprocess.py:298-356) for callers whose tensors live in HOST memory.  Tiles are streamed
// host -> device -> host through kSlots device slots on independent streams, so the H2D copy of
// tile b+1, the kernels of tile b and the D2H copy of tile b-1 overlap (PCIe is full duplex).
// This is the path the benchmark's end-to-end figure times.
//
// The link is the bound (20 B/px in, 8 B/px out, full duplex), so the int64 class map is narrowed on the
// host before it crosses: worker threads pack each tile's sem to one byte per pixel into a pinned staging
// buffer while the previous tile's copies are in flight (13 B/px in instead of 20).  The kernels take uint8
// sem natively.  A tile holding a class id outside [0, 255] is sent as int64 instead.
#include <stdio.h>
#include <stdlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <sched.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"

extern "C" int emp_panoptic_batched(int, const void*, int, const float*, const float*, int, int, const int64_t*, int,
                                    int64_t, int64_t, int64_t, float, int, int64_t*, int64_t*, int, int, void*,
                                    size_t, void*);

namespace emp {

constexpr int kSlots = 3;                       // 4 and 6 slots measure the same (the link is the bound)

// Worker threads that live as long as the pipe (a thread per tile and chunk cost ~50 us each to spawn and join):
// run(n, f) executes f(0) .. f(n-1), f(0) on the caller, and returns when all are done.
class WorkerPool {
public:
    explicit WorkerPool(int workers)
    {
        for (int t = 0; t < workers; ++t) threads_.emplace_back([this, t] { loop(t + 1); });
    }
    ~WorkerPool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++generation_;
        }
        cv_.notify_all();
        for (auto& th : threads_) th.join();
    }
    int size() const { return (int)threads_.size() + 1; }
    void run(int n, const std::function<void(int)>& f)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &f; n_ = n; pending_ = std::min(n, size()) - 1;
            ++generation_;
        }
        cv_.notify_all();
        for (int i = 0; i < n; i += size()) f(i);               // the caller is worker 0
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ <= 0; });
        job_ = nullptr;
    }

private:
    void loop(int id)
    {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            int n;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                job = job_; n = n_;
                if (id >= n) continue;                          // fewer items than workers: not counted in pending_
            }
            for (int i = id; i < n; i += size()) (*job)(i);
            {
                std::lock_guard<std::mutex> g(m_);
                --pending_;
            }
            done_.notify_one();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    int n_ = 0, pending_ = 0;
    unsigned long generation_ = 0;
    bool stop_ = false;
};

// One pipe per device: its streams, events and pinned staging belong to that device's context.  A pipe serves one
// caller at a time (mutex held for the whole call).
struct HostPipe {
    std::mutex busy;
    cudaStream_t streams[kSlots] = {};
    int32_t* status_pinned = nullptr;   // B * EMP_ST_WORDS
    int status_cap = 0;
    uint8_t* sem8[kSlots] = {};     // pinned staging: one narrowed tile each
    cudaEvent_t sem8_free[kSlots] = {};
    size_t sem8_cap = 0;
    int threads = 0;                // 0: not decided yet, < 0: narrowing off
    std::unique_ptr<WorkerPool> pool;
};

static std::mutex g_pipes_mutex;
static std::map<int, std::unique_ptr<HostPipe>> g_pipes;
static std::atomic<double> g_sem_bytes_per_px{8.0};     // of the last call (any device)

// int64 -> uint8 over [i0, i1); returns the OR of everything seen (any bit above 0xff: not narrowable)
static uint64_t narrow_range(const int64_t* __restrict__ src, uint8_t* __restrict__ dst, size_t i0, size_t i1)
{
    uint64_t seen = 0;
    size_t i = i0;
    for (; i < i1 && (i & 7); ++i) { const uint64_t v = (uint64_t)src[i]; seen |= v; dst[i] = (uint8_t)v; }
#if defined(__SSE2__)
    // 16 ids per step: OR everything for the range check, take the low dwords (shuffle_ps), then the saturating
    // packs 32 -> 16 -> 8 (saturation only touches values the range check rejects anyway)
    for (; i < i1 && (i & 15); ++i) { const uint64_t v = (uint64_t)src[i]; seen |= v; dst[i] = (uint8_t)v; }
    __m128i acc = _mm_setzero_si128();
    for (; i + 16 <= i1; i += 16) {
        const __m128i* p = reinterpret_cast<const __m128i*>(src + i);
        const __m128i v0 = _mm_loadu_si128(p), v1 = _mm_loadu_si128(p + 1), v2 = _mm_loadu_si128(p + 2), v3 = _mm_loadu_si128(p + 3);
        const __m128i v4 = _mm_loadu_si128(p + 4), v5 = _mm_loadu_si128(p + 5), v6 = _mm_loadu_si128(p + 6), v7 = _mm_loadu_si128(p + 7);
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(_mm_or_si128(v0, v1), _mm_or_si128(v2, v3)),
                                             _mm_or_si128(_mm_or_si128(v4, v5), _mm_or_si128(v6, v7))));
        auto low = [](const __m128i a, const __m128i b) {
            return _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(a), _mm_castsi128_ps(b), _MM_SHUFFLE(2, 0, 2, 0)));
        };
        const __m128i w0 = _mm_packs_epi32(low(v0, v1), low(v2, v3)), w1 = _mm_packs_epi32(low(v4, v5), low(v6, v7));
        _mm_store_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packus_epi16(w0, w1));      // dst + i is 16-byte aligned
    }
    seen |= (uint64_t)_mm_cvtsi128_si64(acc) | (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(acc, acc));
#else
    for (; i + 8 <= i1; i += 8) {
        uint64_t packed = 0, any = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint64_t v = (uint64_t)src[i + k];
            any |= v;
            packed |= (v & 0xffull) << (8 * k);
        }
        seen |= any;
        *reinterpret_cast<uint64_t*>(dst + i) = packed;          // dst + i is 8-byte aligned (i % 8 == 0, base pinned)
    }
#endif
    for (; i < i1; ++i) { const uint64_t v = (uint64_t)src[i]; seen |= v; dst[i] = (uint8_t)v; }
    return seen;
}

// true if the whole tile fitted into bytes
static bool narrow_tile(const int64_t* src, uint8_t* dst, size_t n, WorkerPool* pool)
{
    const int threads = pool ? pool->size() : 1;
    if (threads <= 1 || n < (1u << 16)) return (narrow_range(src, dst, 0, n) >> 8) == 0;
    std::atomic<uint64_t> seen{0};
    const size_t chunk = ((n + threads - 1) / threads + 15) & ~(size_t)15;
    pool->run(threads, [&](int t) {
        const size_t a = std::min(n, chunk * t), b = std::min(n, chunk * (t + 1));
        if (a < b) seen.fetch_or(narrow_range(src, dst, a, b), std::memory_order_relaxed);
    });
    return (seen.load() >> 8) == 0;
}

// CPUs this process may run on (its affinity mask): a launcher that gives every rank its own slice of the host
// thereby sizes every rank's worker pool
static int usable_cpus()
{
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    return std::max(1, (int)std::thread::hardware_concurrency());
}

static int get_pipe(int B, HostPipe** out)
{
    int dev = 0;
    EMP_CUDA_CHECK(cudaGetDevice(&dev));
    HostPipe* p;
    {
        std::lock_guard<std::mutex> g(g_pipes_mutex);
        auto& slot = g_pipes[dev];
        if (!slot) slot.reset(new HostPipe());
        p = slot.get();
    }
    *out = p;
    return EMP_OK;
}

// called with p->busy held and p's device current
static int prepare_pipe(HostPipe* p, int B)
{
    for (int i = 0; i < kSlots; ++i) {
        if (!p->streams[i]) EMP_CUDA_CHECK(cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking));
        if (!p->sem8_free[i]) EMP_CUDA_CHECK(cudaEventCreateWithFlags(&p->sem8_free[i], cudaEventDisableTiming));
    }
    if (p->threads == 0) {
        // EMP_HOST_THREADS: workers narrowing sem (0 disables narrowing); default: the CPUs of this process's affinity
        // mask — shared by the visible GPUs when the mask is the whole machine (one process per GPU) — at most 8
        const char* e = getenv("EMP_HOST_THREADS");
        int ndev = 1;
        cudaGetDeviceCount(&ndev);
        const int cpus = usable_cpus();
        const bool whole_machine = cpus >= (int)std::thread::hardware_concurrency();
        p->threads = e ? atoi(e) : std::max(1, std::min(8, whole_machine ? cpus / std::max(ndev, 1) : cpus));
        if (p->threads <= 0) p->threads = -1;                   // "decided: off"
        if (p->threads > 1) p->pool.reset(new WorkerPool(p->threads - 1));
    }
    if (p->status_cap < B) {
        if (p->status_pinned) cudaFreeHost(p->status_pinned);
        p->status_pinned = nullptr;
        p->status_cap = 0;
        EMP_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&p->status_pinned),
                                     sizeof(int32_t) * EMP_ST_WORDS * (size_t)B, cudaHostAllocDefault));
        p->status_cap = B;
    }
    return EMP_OK;
}

// Whatever way the tile loop ends, no copy may still be in flight into the caller's buffers when we return.
struct DrainStreams {
    HostPipe* p;
    ~DrainStreams()
    {
        for (int i = 0; i < kSlots; ++i)
            if (p->streams[i]) cudaStreamSynchronize(p->streams[i]);
    }
};

struct SlotLayout {
    size_t sem, hm, off, pan, ws, total;
};

static SlotLayout slot_layout(int H, int W, int k_cap, int n_things)
{
    const size_t n = (size_t)H * W;
    SlotLayout s;
    size_t o = 0;
    s.sem = o; o = align_up(o + 8 * n, 256);
    s.hm = o;  o = align_up(o + 4 * n, 256);
    s.off = o; o = align_up(o + 8 * n, 256);
    s.pan = o; o = align_up(o + 8 * n, 256);
    s.ws = o;  o = align_up(o + ws_layout(H, W, k_cap, n_things).total, 256);
    s.total = o;
    return s;
}

}  // namespace emp

using namespace emp;

EMP_API double emp_host_sem_bytes_per_px(void) { return g_sem_bytes_per_px.load(); }

EMP_API size_t emp_host_scratch_bytes(int H, int W, int k_cap, int n_things)
{
    if (H <= 0 || W <= 0 || k_cap < 1) return 0;
    return slot_layout(H, W, k_cap, n_things).total * kSlots;
}

// sem_elt: 8 (int64 class maps, narrowed on the host when workers are on) or 1 (the caller already holds bytes)
static int batched_host(int B, const void* sem_h, int sem_elt, const float* hm_h, const float* off_h, int H, int W,
                        const int64_t* thing_list, int n_things, int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                        float threshold, int nms_kernel, int64_t* pan_out_h, int32_t* k_out, int32_t* flags_out, int k_cap,
                        void* dev_scratch, size_t dev_scratch_bytes)
{
    EMP_REQUIRE(B >= 1, EMP_ERR_INVALID, "bad batch %d", B);
    EMP_REQUIRE(sem_h && hm_h && off_h && pan_out_h, EMP_ERR_INVALID, "null host pointer");
    EMP_REQUIRE(H > 0 && W > 0 && k_cap >= 1, EMP_ERR_INVALID, "bad shape / k_cap");
    const SlotLayout S = slot_layout(H, W, k_cap, n_things);
    EMP_REQUIRE(dev_scratch && (reinterpret_cast<uintptr_t>(dev_scratch) & 255u) == 0, EMP_ERR_WORKSPACE,
                "device scratch must be 256-byte aligned");
    EMP_REQUIRE(dev_scratch_bytes >= S.total * kSlots, EMP_ERR_WORKSPACE, "device scratch too small: %zu < %zu",
                dev_scratch_bytes, S.total * kSlots);
    HostPipe* p = nullptr;
    int rc = get_pipe(B, &p);
    if (rc) return rc;
    std::lock_guard<std::mutex> busy(p->busy);                  // one caller per device at a time
    if ((rc = prepare_pipe(p, B))) return rc;
    DrainStreams drain{p};                                      // every return below waits for the copies already enqueued
    const size_t n = (size_t)H * W;
    const size_t ws_bytes = ws_layout(H, W, k_cap, n_things).total;
    char* base = static_cast<char*>(dev_scratch);
    const bool narrowing = sem_elt == 8 && p->threads > 0;
    if (narrowing && p->sem8_cap < n) {
        for (int i = 0; i < kSlots; ++i) {
            if (p->sem8[i]) cudaFreeHost(p->sem8[i]);
            p->sem8[i] = nullptr;
        }
        p->sem8_cap = 0;
        for (int i = 0; i < kSlots; ++i)
            EMP_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&p->sem8[i]), n, cudaHostAllocDefault));
        p->sem8_cap = n;
    }

    int narrowed = 0;
    for (int b = 0; b < B; ++b) {
        const int s = b % kSlots;
        cudaStream_t st = p->streams[s];
        char* slot = base + (size_t)s * S.total;
        int sem_u8 = sem_elt == 1 ? 1 : 0;
        const char* sem_b = static_cast<const char*>(sem_h) + (size_t)b * n * sem_elt;
        if (narrowing) {
            if (b >= kSlots) EMP_CUDA_CHECK(cudaEventSynchronize(p->sem8_free[s]));   // its last copy has left the buffer
            sem_u8 = narrow_tile(reinterpret_cast<const int64_t*>(sem_b), p->sem8[s], n, p->pool.get()) ? 1 : 0;
        }
        narrowed += sem_u8;
        if (narrowing && sem_u8) {
            EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.sem, p->sem8[s], n, cudaMemcpyHostToDevice, st));
            EMP_CUDA_CHECK(cudaEventRecord(p->sem8_free[s], st));
        } else {
            if (narrowing) EMP_CUDA_CHECK(cudaEventRecord(p->sem8_free[s], st));
            EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.sem, sem_b, (size_t)sem_elt * n, cudaMemcpyHostToDevice, st));
        }
        EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.hm, hm_h + (size_t)b * n, 4 * n, cudaMemcpyHostToDevice, st));
        EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.off, off_h + (size_t)b * 2 * n, 8 * n, cudaMemcpyHostToDevice, st));
        rc = emp_panoptic_batched(1, slot + S.sem, sem_u8, reinterpret_cast<const float*>(slot + S.hm),
                                  reinterpret_cast<const float*>(slot + S.off), H, W, thing_list, n_things,
                                  label_divisor, stuff_area, void_label, threshold, nms_kernel,
                                  reinterpret_cast<int64_t*>(slot + S.pan), nullptr, 0, k_cap, slot + S.ws,
                                  align_up(ws_bytes, 256), st);
        if (rc) return rc;
        EMP_CUDA_CHECK(cudaMemcpyAsync(pan_out_h + (size_t)b * n, slot + S.pan, 8 * n, cudaMemcpyDeviceToHost, st));
        EMP_CUDA_CHECK(cudaMemcpyAsync(p->status_pinned + (size_t)b * EMP_ST_WORDS, slot + S.ws,
                                       sizeof(int32_t) * EMP_ST_WORDS, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kSlots; ++i) EMP_CUDA_CHECK(cudaStreamSynchronize(p->streams[i]));
    g_sem_bytes_per_px.store((1.0 * narrowed + 8.0 * (B - narrowed)) / B);
    for (int b = 0; b < B; ++b) {
        if (k_out) k_out[b] = p->status_pinned[(size_t)b * EMP_ST_WORDS + EMP_ST_K];
        if (flags_out) flags_out[b] = p->status_pinned[(size_t)b * EMP_ST_WORDS + EMP_ST_FLAGS];
    }
    return EMP_OK;
}

EMP_API int emp_panoptic_batched_host(int B, const int64_t* sem_h, const float* hm_h, const float* off_h,
                                      int H, int W, const int64_t* thing_list, int n_things,
                                      int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                                      float threshold, int nms_kernel, int64_t* pan_out_h, int32_t* k_out,
                                      int32_t* flags_out, int k_cap, void* dev_scratch, size_t dev_scratch_bytes)
{
    return batched_host(B, sem_h, 8, hm_h, off_h, H, W, thing_list, n_things, label_divisor, stuff_area, void_label, threshold,
                        nms_kernel, pan_out_h, k_out, flags_out, k_cap, dev_scratch, dev_scratch_bytes);
}

EMP_API int emp_panoptic_batched_host_u8(int B, const uint8_t* sem8_h, const float* hm_h, const float* off_h,
                                         int H, int W, const int64_t* thing_list, int n_things,
                                         int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                                         float threshold, int nms_kernel, int64_t* pan_out_h, int32_t* k_out,
                                         int32_t* flags_out, int k_cap, void* dev_scratch, size_t dev_scratch_bytes)
{
    return batched_host(B, sem8_h, 1, hm_h, off_h, H, W, thing_list, n_things, label_divisor, stuff_area, void_label, threshold,
                        nms_kernel, pan_out_h, k_out, flags_out, k_cap, dev_scratch, dev_scratch_bytes);
}
