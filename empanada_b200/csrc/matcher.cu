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
            if (ov > 0) {
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
