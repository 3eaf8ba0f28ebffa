// stack.cu — emp_stack_slice: one z-slice of the stack path behind the median queue, in ONE host call.
//   harden (engines.py:114-121) -> coarse centers + nearest-center ids (:257-272) -> upsample-fused merge (:274-292)
//   -> crop to the unpadded size -> pan_seg -> RLE tables (rle.py:26-86)
// It only chains the library's own entry points on the caller's stream and gathers their three status blocks into
// the caller's device array, so a z-block can be enqueued without a single host synchronisation and with one
// foreign-function call per slice instead of eight plus the allocator traffic around them.
#include <stdio.h>
#include "common.cuh"

extern "C" {
int emp_median_harden(const float* const*, int, int, int, int, float, float*, void*, int, void*);
int emp_coarse_ids(const float*, const float*, int, int, float, int, float, int32_t*, int, void*, size_t, void*);
int emp_merge_coarse(const void*, int, const int32_t*, int, int, int, int, int, int64_t, const int64_t*, int, int64_t, int64_t,
                     int64_t, const int32_t*, int64_t*, void*, size_t, void*);
int emp_rle(const int64_t*, int, int, const int64_t*, int, int64_t, const int64_t*, int, int, int64_t*, int, int64_t*, int,
            void*, size_t, void*);
size_t emp_workspace_bytes(int, int, int, int);
size_t emp_rle_workspace_bytes(int, int, int, int, int64_t);
}

namespace emp {

struct SliceScratch {
    size_t sem8, ids, pan, crop, ws_coarse, ws_merge, ws_rle, total;
    size_t n_coarse, n_merge, n_rle;
};

static SliceScratch slice_scratch(int H, int W, int h, int w, int k_cap, int n_things, int run_cap, int n_labels,
                                  int64_t label_divisor)
{
    SliceScratch s;
    size_t o = 0;
    s.n_coarse = emp_workspace_bytes(h, w, k_cap, 1);
    s.n_merge = emp_workspace_bytes(H, W, k_cap, n_things > 0 ? n_things : 1);
    s.n_rle = emp_rle_workspace_bytes(H, W, run_cap, n_labels, label_divisor);       // the crop is never larger
    s.sem8 = o;      o = align_up(o + (size_t)H * W, 256);
    s.ids = o;       o = align_up(o + sizeof(int32_t) * (size_t)h * w, 256);
    s.pan = o;       o = align_up(o + sizeof(int64_t) * (size_t)H * W, 256);
    s.crop = o;      o = align_up(o + sizeof(int64_t) * (size_t)H * W, 256);
    s.ws_coarse = o; o = align_up(o + s.n_coarse, 256);
    s.ws_merge = o;  o = align_up(o + s.n_merge, 256);
    s.ws_rle = o;    o = align_up(o + s.n_rle, 256);
    s.total = o;
    return s;
}

}  // namespace emp

using namespace emp;

EMP_API size_t emp_stack_slice_scratch_bytes(int H, int W, int h, int w, int k_cap, int n_things, int run_cap,
                                             int n_labels, int64_t label_divisor)
{
    if (H <= 0 || W <= 0 || h <= 0 || w <= 0 || k_cap < 1 || run_cap < 1 || n_labels < 1) return 0;
    const SliceScratch s = slice_scratch(H, W, h, w, k_cap, n_things, run_cap, n_labels, label_divisor);
    if (!s.n_coarse || !s.n_merge || !s.n_rle) return 0;
    return s.total;
}

EMP_API int emp_stack_slice(const float* sem_prob, int C, int H, int W, float confidence_thr, const float* hm,
                            const float* off, int h, int w, float nms_threshold, int nms_kernel, float step, int shift,
                            const int64_t* thing_list, int n_things, int64_t label_divisor, int64_t stuff_area,
                            int64_t void_label, int k_cap, int crop_h, int crop_w, const int64_t* labels, int n_labels,
                            int force_connected, void* scratch, size_t scratch_bytes, int64_t* pan_out,
                            int64_t* runs_out, int run_cap, int64_t* inst_out, int inst_cap, int32_t* status_out,
                            void* stream)
{
    EMP_REQUIRE(sem_prob && hm && off && scratch && runs_out && inst_out && status_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(C >= 1 && H > 0 && W > 0 && h > 0 && w > 0, EMP_ERR_INVALID, "bad shape");
    EMP_REQUIRE(shift >= 0 && ((long long)h << shift) >= H && ((long long)w << shift) >= W, EMP_ERR_INVALID,
                "coarse map %d x %d << %d does not cover %d x %d", h, w, shift, H, W);
    EMP_REQUIRE(crop_h >= 1 && crop_h <= H && crop_w >= 1 && crop_w <= W, EMP_ERR_INVALID, "bad crop %d x %d", crop_h, crop_w);
    EMP_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255u) == 0, EMP_ERR_WORKSPACE, "scratch must be 256-byte aligned");
    const SliceScratch S = slice_scratch(H, W, h, w, k_cap, n_things, run_cap, n_labels, label_divisor);
    EMP_REQUIRE(S.n_coarse && S.n_merge && S.n_rle, EMP_ERR_INVALID, "bad k_cap / run_cap / labels");
    EMP_REQUIRE(scratch_bytes >= S.total, EMP_ERR_WORKSPACE, "scratch too small: %zu < %zu", scratch_bytes, S.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(scratch);
    int rc;
    const float* planes[1] = {sem_prob};
    if ((rc = emp_median_harden(planes, 1, C, H, W, confidence_thr, nullptr, base + S.sem8, 1, stream))) return rc;
    int32_t* ids = reinterpret_cast<int32_t*>(base + S.ids);
    if ((rc = emp_coarse_ids(hm, off, h, w, nms_threshold, nms_kernel, step, ids, k_cap, base + S.ws_coarse, S.n_coarse, stream)))
        return rc;
    int64_t* pan = pan_out ? pan_out : reinterpret_cast<int64_t*>(base + S.pan);
    if ((rc = emp_merge_coarse(base + S.sem8, 1, ids, h, w, shift, H, W, label_divisor, thing_list, n_things, stuff_area,
                               void_label, k_cap, reinterpret_cast<const int32_t*>(base + S.ws_coarse), pan, base + S.ws_merge,
                               S.n_merge, stream)))
        return rc;
    const int64_t* rle_in = pan;
    if (crop_h != H || crop_w != W) {                   // the engine crops the padded map before it is encoded
        int64_t* crop = reinterpret_cast<int64_t*>(base + S.crop);
        EMP_CUDA_CHECK(cudaMemcpy2DAsync(crop, sizeof(int64_t) * (size_t)crop_w, pan, sizeof(int64_t) * (size_t)W,
                                         sizeof(int64_t) * (size_t)crop_w, (size_t)crop_h, cudaMemcpyDeviceToDevice, st));
        rle_in = crop;
    }
    if ((rc = emp_rle(rle_in, crop_h, crop_w, labels, n_labels, label_divisor, thing_list, n_things, force_connected, runs_out,
                      run_cap, inst_out, inst_cap, base + S.ws_rle, S.n_rle, stream)))
        return rc;
    const size_t nb = sizeof(int32_t) * EMP_ST_WORDS;
    EMP_CUDA_CHECK(cudaMemcpyAsync(status_out, base + S.ws_coarse, nb, cudaMemcpyDeviceToDevice, st));
    EMP_CUDA_CHECK(cudaMemcpyAsync(status_out + EMP_ST_WORDS, base + S.ws_merge, nb, cudaMemcpyDeviceToDevice, st));
    EMP_CUDA_CHECK(cudaMemcpyAsync(status_out + 2 * EMP_ST_WORDS, base + S.ws_rle, nb, cudaMemcpyDeviceToDevice, st));
    return EMP_OK;
}
