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
#include <atomic>
#include <thread>
#include <vector>
#include "common.cuh"

extern "C" int emp_panoptic_batched(int, const void*, int, const float*, const float*, int, int, const int64_t*, int,
                                    int64_t, int64_t, int64_t, float, int, int64_t*, int64_t*, int, int, void*,
                                    size_t, void*);

namespace emp {

constexpr int kSlots = 3;                       // 4 and 6 slots measure the same (the link is the bound)

struct HostPipe {
    int device = -1;
    cudaStream_t streams[kSlots] = {};
    int32_t* status_pinned = nullptr;   // B * EMP_ST_WORDS
    int status_cap = 0;
    uint8_t* sem8[kSlots] = {};     // pinned staging: one narrowed tile each
    cudaEvent_t sem8_free[kSlots] = {};
    size_t sem8_cap = 0;
    int threads = 0;
    double sem_bytes_per_px = 8.0;  // of the last call
};

static HostPipe g_pipe;

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
static bool narrow_tile(const int64_t* src, uint8_t* dst, size_t n, int threads)
{
    if (threads <= 1 || n < (1u << 16)) return (narrow_range(src, dst, 0, n) >> 8) == 0;
    std::atomic<uint64_t> seen{0};
    std::vector<std::thread> pool;
    const size_t chunk = ((n + threads - 1) / threads + 15) & ~(size_t)15;
    auto work = [&](int t) {
        const size_t a = std::min(n, chunk * t), b = std::min(n, chunk * (t + 1));
        if (a < b) seen.fetch_or(narrow_range(src, dst, a, b), std::memory_order_relaxed);
    };
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    return (seen.load() >> 8) == 0;
}

static int ensure_pipe(int B)
{
    int dev = 0;
    EMP_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_pipe.device != dev) {
        for (int i = 0; i < kSlots; ++i) {
            if (g_pipe.streams[i]) cudaStreamDestroy(g_pipe.streams[i]);
            EMP_CUDA_CHECK(cudaStreamCreateWithFlags(&g_pipe.streams[i], cudaStreamNonBlocking));
        }
        g_pipe.device = dev;
    }
    if (g_pipe.threads == 0) {
        // EMP_HOST_THREADS: workers narrowing sem (0 disables narrowing); default: the host's threads shared by
        // the visible GPUs (one process per GPU), at most 8
        const char* e = getenv("EMP_HOST_THREADS");
        int ndev = 1;
        cudaGetDeviceCount(&ndev);
        const int hw = (int)std::thread::hardware_concurrency();
        g_pipe.threads = e ? atoi(e) : std::max(1, std::min(8, hw / std::max(ndev, 1)));
        if (g_pipe.threads < 0) g_pipe.threads = 0;
        if (g_pipe.threads == 0) g_pipe.threads = -1;           // "decided: off"
    }
    if (g_pipe.status_cap < B) {
        if (g_pipe.status_pinned) cudaFreeHost(g_pipe.status_pinned);
        g_pipe.status_pinned = nullptr;
        EMP_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&g_pipe.status_pinned),
                                     sizeof(int32_t) * EMP_ST_WORDS * (size_t)B, cudaHostAllocDefault));
        g_pipe.status_cap = B;
    }
    return EMP_OK;
}

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

EMP_API double emp_host_sem_bytes_per_px(void) { return g_pipe.sem_bytes_per_px; }

EMP_API size_t emp_host_scratch_bytes(int H, int W, int k_cap, int n_things)
{
    if (H <= 0 || W <= 0 || k_cap < 1) return 0;
    return slot_layout(H, W, k_cap, n_things).total * kSlots;
}

EMP_API int emp_panoptic_batched_host(int B, const int64_t* sem_h, const float* hm_h, const float* off_h,
                                      int H, int W, const int64_t* thing_list, int n_things,
                                      int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                                      float threshold, int nms_kernel, int64_t* pan_out_h, int32_t* k_out,
                                      int32_t* flags_out, int k_cap, void* dev_scratch, size_t dev_scratch_bytes)
{
    EMP_REQUIRE(B >= 1, EMP_ERR_INVALID, "bad batch %d", B);
    EMP_REQUIRE(sem_h && hm_h && off_h && pan_out_h, EMP_ERR_INVALID, "null host pointer");
    EMP_REQUIRE(H > 0 && W > 0 && k_cap >= 1, EMP_ERR_INVALID, "bad shape / k_cap");
    const SlotLayout S = slot_layout(H, W, k_cap, n_things);
    EMP_REQUIRE(dev_scratch && (reinterpret_cast<uintptr_t>(dev_scratch) & 255u) == 0, EMP_ERR_WORKSPACE,
                "device scratch must be 256-byte aligned");
    EMP_REQUIRE(dev_scratch_bytes >= S.total * kSlots, EMP_ERR_WORKSPACE, "device scratch too small: %zu < %zu",
                dev_scratch_bytes, S.total * kSlots);
    int rc = ensure_pipe(B);
    if (rc) return rc;
    const size_t n = (size_t)H * W;
    const size_t ws_bytes = ws_layout(H, W, k_cap, n_things).total;
    char* base = static_cast<char*>(dev_scratch);
    const bool narrowing = g_pipe.threads > 0;
    if (narrowing && g_pipe.sem8_cap < n) {
        for (int i = 0; i < kSlots; ++i) {
            if (g_pipe.sem8[i]) cudaFreeHost(g_pipe.sem8[i]);
            g_pipe.sem8[i] = nullptr;
            EMP_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&g_pipe.sem8[i]), n, cudaHostAllocDefault));
            if (!g_pipe.sem8_free[i]) EMP_CUDA_CHECK(cudaEventCreateWithFlags(&g_pipe.sem8_free[i], cudaEventDisableTiming));
        }
        g_pipe.sem8_cap = n;
    }

    int narrowed = 0;
    for (int b = 0; b < B; ++b) {
        const int s = b % kSlots;
        cudaStream_t st = g_pipe.streams[s];
        char* slot = base + (size_t)s * S.total;
        int sem_u8 = 0;
        if (narrowing) {
            if (b >= kSlots) EMP_CUDA_CHECK(cudaEventSynchronize(g_pipe.sem8_free[s]));   // its last copy has left the buffer
            sem_u8 = narrow_tile(sem_h + (size_t)b * n, g_pipe.sem8[s], n, g_pipe.threads) ? 1 : 0;
        }
        narrowed += sem_u8;
        if (sem_u8) {
            EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.sem, g_pipe.sem8[s], n, cudaMemcpyHostToDevice, st));
            EMP_CUDA_CHECK(cudaEventRecord(g_pipe.sem8_free[s], st));
        } else {
            if (narrowing) EMP_CUDA_CHECK(cudaEventRecord(g_pipe.sem8_free[s], st));
            EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.sem, sem_h + (size_t)b * n, 8 * n, cudaMemcpyHostToDevice, st));
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
        EMP_CUDA_CHECK(cudaMemcpyAsync(g_pipe.status_pinned + (size_t)b * EMP_ST_WORDS, slot + S.ws,
                                       sizeof(int32_t) * EMP_ST_WORDS, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kSlots; ++i) EMP_CUDA_CHECK(cudaStreamSynchronize(g_pipe.streams[i]));
    g_pipe.sem_bytes_per_px = (1.0 * narrowed + 8.0 * (B - narrowed)) / B;
    for (int b = 0; b < B; ++b) {
        if (k_out) k_out[b] = g_pipe.status_pinned[(size_t)b * EMP_ST_WORDS + EMP_ST_K];
        if (flags_out) flags_out[b] = g_pipe.status_pinned[(size_t)b * EMP_ST_WORDS + EMP_ST_FLAGS];
    }
    return EMP_OK;
}
