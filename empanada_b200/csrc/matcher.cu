// matcher.cu — the arithmetic behind the cross-slice matcher (reference empanada/inference/matcher.py
// :136-232 rle_matcher; array_utils.py rle_intersection :371-403, rle_iou :405-429, rle_ioa :431-449):
// pixel overlaps between the instances of consecutive slices, straight from the run tables emp_rle
// leaves in HBM.  The reference sorts and sweeps the two run lists of every box-overlapping instance
// pair on the host (the documented multi-GPU bottleneck, docs/plugin/best-practice.rst:30-34); here one
// launch covers every pair of consecutive slices of a z-block: a thread takes one run of slice p+1,
// binary-searches the (start-ordered, disjoint) runs of slice p for the first one that ends after its
// start and walks while they begin before its end, emitting (pair, slot_a, slot_b, overlap) rows.
// Hungarian assignment and the label bookkeeping stay on the host (they work on n x m matrices).
#include "common.cuh"

namespace emp {

__global__ void __launch_bounds__(256)
rle_pair_overlaps_kernel(const long long* __restrict__ runs, size_t run_stride, const int32_t* __restrict__ n_runs,
                         int32_t* __restrict__ out, int cap, int32_t* __restrict__ count)
{
    const int p = blockIdx.y;                                   // pair: A = slice p, B = slice p + 1
    const int nA = __ldg(n_runs + p), nB = __ldg(n_runs + p + 1);
    const long long* A = runs + (size_t)p * run_stride * 3;
    const long long* B = runs + (size_t)(p + 1) * run_stride * 3;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nB; j += gridDim.x * blockDim.x) {
        const long long bs = B[3 * (size_t)j], be = bs + B[3 * (size_t)j + 1];
        const int slot_b = (int)B[3 * (size_t)j + 2];
        if (slot_b < 0) continue;                               // a run of a label that is not selected
        int lo = 0, hi = nA;                                    // first A run with end > bs (ends ascend too)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (A[3 * (size_t)mid] + A[3 * (size_t)mid + 1] > bs) hi = mid; else lo = mid + 1;
        }
        for (int i = lo; i < nA; ++i) {
            const long long as = A[3 * (size_t)i];
            if (as >= be) break;
            const long long ae = as + A[3 * (size_t)i + 1];
            const long long ov = min(ae, be) - max(as, bs);
            if (ov > 0 && A[3 * (size_t)i + 2] >= 0) {
                const int pos = atomicAdd(count, 1);
                if (pos < cap) {
                    int4 row = make_int4(p, (int)A[3 * (size_t)i + 2], slot_b, (int)ov);
                    reinterpret_cast<int4*>(out)[pos] = row;
                }
            }
        }
    }
}

}  // namespace emp

using namespace emp;

EMP_API int emp_rle_pair_overlaps(const int64_t* runs, size_t run_stride, const int32_t* n_runs, int n_slices,
                                  int max_runs, int32_t* out, int cap, int32_t* count, void* stream)
{
    EMP_REQUIRE(runs && n_runs && out && count, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(n_slices >= 1 && max_runs >= 0 && cap >= 0, EMP_ERR_INVALID, "bad sizes");
    EMP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, EMP_ERR_INVALID, "out must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EMP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    if (n_slices < 2 || max_runs == 0) return EMP_OK;
    int bx = (max_runs + 255) / 256;
    if (bx > 1024) bx = 1024;
    dim3 grid(bx, n_slices - 1);
    rle_pair_overlaps_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(runs), run_stride, n_runs, out, cap, count);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

// ---------------------------------------------------------------------------------------------
// Dense fill of a z-block — reference array_utils.numpy_fill_instances (array_utils.py:725-736) applied
// to what the tracker holds after matching: every run (start, length, slot) of slice s paints
// labels[s][slot] over [start, start + length) of plane s.  Voxels no run covers keep their value (the
// reference leaves them untouched too); a label < 0 skips the run.  One warp per run, 16-byte stores in
// the aligned middle of long runs.
// ---------------------------------------------------------------------------------------------
namespace emp {

template <typename T>
__global__ void __launch_bounds__(256)
fill_runs_kernel(const long long* __restrict__ runs, size_t run_stride, const int32_t* __restrict__ n_runs,
                 const long long* __restrict__ labels, size_t label_stride, T* __restrict__ out, size_t plane)
{
    const int s = blockIdx.y;
    const int n = __ldg(n_runs + s);
    const long long* R = runs + (size_t)s * run_stride * 3;
    const long long* lab = labels + (size_t)s * label_stride;
    T* dst = out + (size_t)s * plane;
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp_global; j < n; j += n_warps) {
        long long start = __ldg(R + 3 * (size_t)j), end = start + __ldg(R + 3 * (size_t)j + 1);
        const long long slot = __ldg(R + 3 * (size_t)j + 2);
        if (slot < 0 || (size_t)slot >= label_stride) continue;            // no such instance: nothing to paint
        const long long v = __ldg(lab + slot);
        if (v < 0) continue;
        // numpy's volume[s:e] = id clips silently: so does this (runs of a tracker loaded for another shape ...)
        start = max(start, 0ll);
        end = min(end, (long long)plane);
        const T tv = (T)v;
        for (long long i = start + lane; i < end; i += 32) dst[i] = tv;     // consecutive lanes, consecutive voxels
    }
}

}  // namespace emp

EMP_API int emp_fill_runs(const int64_t* runs, size_t run_stride, const int32_t* n_runs, int n_slices, int max_runs,
                          const int64_t* labels, size_t label_stride, void* out, int elem_bytes, size_t plane, void* stream)
{
    EMP_REQUIRE(runs && n_runs && labels && out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(n_slices >= 1 && max_runs >= 0, EMP_ERR_INVALID, "bad sizes");
    EMP_REQUIRE(elem_bytes == 4 || elem_bytes == 8, EMP_ERR_INVALID, "out must hold 4- or 8-byte integers");
    if (max_runs == 0) return EMP_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int bx = (max_runs + 7) / 8;                        // one warp per run
    if (bx > 2048) bx = 2048;
    dim3 grid(bx, n_slices);
    if (elem_bytes == 8)
        fill_runs_kernel<long long><<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(runs), run_stride, n_runs,
                                                          reinterpret_cast<const long long*>(labels), label_stride,
                                                          static_cast<long long*>(out), plane);
    else
        fill_runs_kernel<unsigned><<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(runs), run_stride, n_runs,
                                                         reinterpret_cast<const long long*>(labels), label_stride,
                                                         static_cast<unsigned*>(out), plane);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

// ---------------------------------------------------------------------------------------------
// Overlaps between two arbitrary run lists — the intersections behind consensus.object_iou_graph
// (reference empanada/consensus.py:233-287: rle_iou of every box-overlapping pair of objects from
// different trackers).  A = all runs of one tracker's objects, B = another's, each as (start, length,
// slot) rows in ascending start order.  Unlike the slices of a stack, objects of one tracker may
// overlap each other, so ends are not monotone: a thread takes one run of B, finds the first run of A
// that starts after (its start - lmax_a) — no earlier run can reach it, lmax_a being A's longest run —
// and walks while runs start before its end.
// ---------------------------------------------------------------------------------------------
namespace emp {

__global__ void __launch_bounds__(256)
rle_list_overlaps_kernel(const long long* __restrict__ A, int nA, long long lmax_a, const long long* __restrict__ B, int nB,
                         int32_t* __restrict__ out, int cap, int32_t* __restrict__ count)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nB; j += gridDim.x * blockDim.x) {
        const long long bs = B[3 * (size_t)j], be = bs + B[3 * (size_t)j + 1];
        const int slot_b = (int)B[3 * (size_t)j + 2];
        if (slot_b < 0) continue;                               // a run of a label that is not selected
        int lo = 0, hi = nA;                                    // first A run with start > bs - lmax_a
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (A[3 * (size_t)mid] > bs - lmax_a) hi = mid; else lo = mid + 1;
        }
        for (int i = lo; i < nA; ++i) {
            const long long as = A[3 * (size_t)i];
            if (as >= be) break;
            const long long ov = min(as + A[3 * (size_t)i + 1], be) - max(as, bs);
            if (ov > 0) {
                const int pos = atomicAdd(count, 1);
                if (pos < cap) reinterpret_cast<int4*>(out)[pos] = make_int4(0, (int)A[3 * (size_t)i + 2], slot_b, (int)ov);
            }
        }
    }
}

}  // namespace emp

EMP_API int emp_rle_list_overlaps(const int64_t* runs_a, int n_a, int64_t lmax_a, const int64_t* runs_b, int n_b,
                                  int32_t* out, int cap, int32_t* count, void* stream)
{
    EMP_REQUIRE(runs_a && runs_b && out && count, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(n_a >= 0 && n_b >= 0 && cap >= 0 && lmax_a >= 0, EMP_ERR_INVALID, "bad sizes");
    EMP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, EMP_ERR_INVALID, "out must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EMP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    if (n_a == 0 || n_b == 0) return EMP_OK;
    int bx = (n_b + 255) / 256;
    if (bx > device_sm_count() * 8) bx = device_sm_count() * 8;
    rle_list_overlaps_kernel<<<bx, 256, 0, st>>>(reinterpret_cast<const long long*>(runs_a), n_a, (long long)lmax_a,
                                                 reinterpret_cast<const long long*>(runs_b), n_b, out, cap, count);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}
