// host_pipeline.cu — emp_panoptic_batched_host: the fused get_panoptic_segmentation pipeline
// (postprocess.py:298-356) for callers whose tensors live in HOST memory.  Tiles are streamed
// host -> device -> host through kSlots device slots on independent streams, so the H2D copy of
// tile b+1, the kernels of tile b and the D2H copy of tile b-1 overlap (PCIe is full duplex).
// This is the path the benchmark's end-to-end figure times.
#include <stdio.h>
#include "common.cuh"

extern "C" int emp_panoptic_batched(int, const void*, int, const float*, const float*, int, int, const int64_t*, int,
                                    int64_t, int64_t, int64_t, float, int, int64_t*, int64_t*, int, int, void*,
                                    size_t, void*);

namespace emp {

constexpr int kSlots = 3;

struct HostPipe {
    int device = -1;
    cudaStream_t streams[kSlots] = {nullptr, nullptr, nullptr};
    int32_t* status_pinned = nullptr;   // B * EMP_ST_WORDS
    int status_cap = 0;
};

static HostPipe g_pipe;

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

    for (int b = 0; b < B; ++b) {
        const int s = b % kSlots;
        cudaStream_t st = g_pipe.streams[s];
        char* slot = base + (size_t)s * S.total;
        EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.sem, sem_h + (size_t)b * n, 8 * n, cudaMemcpyHostToDevice, st));
        EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.hm, hm_h + (size_t)b * n, 4 * n, cudaMemcpyHostToDevice, st));
        EMP_CUDA_CHECK(cudaMemcpyAsync(slot + S.off, off_h + (size_t)b * 2 * n, 8 * n, cudaMemcpyHostToDevice, st));
        rc = emp_panoptic_batched(1, slot + S.sem, 0, reinterpret_cast<const float*>(slot + S.hm),
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
    for (int b = 0; b < B; ++b) {
        if (k_out) k_out[b] = g_pipe.status_pinned[(size_t)b * EMP_ST_WORDS + EMP_ST_K];
        if (flags_out) flags_out[b] = g_pipe.status_pinned[(size_t)b * EMP_ST_WORDS + EMP_ST_FLAGS];
    }
    return EMP_OK;
}
