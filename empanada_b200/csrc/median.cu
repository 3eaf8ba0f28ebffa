// median.cu — _MedianQueue.get_median + _harden_seg (reference empanada/inference/engines.py:59-66,
// :114-121) as one streaming kernel: per element the middle order statistic of ks (odd) planes,
// optionally written back (the reference stores the filtered plane into the queued entry, which
// is what makes its median recursive), then hardened into a class map (C == 1: p >= thr,
// C > 1: first arg-max over channels).  HBM-bound: ks*4*C B/px in, 4*C (+8 or +1) B/px out.
#include <math_constants.h>
#include "common.cuh"

namespace emp {

constexpr int kMaxKs = 15;

struct PlanePtrs {
    const float* p[kMaxKs];
};

template <int KS>
__device__ __forceinline__ float median_of(float (&v)[KS])
{
    // odd-even transposition network on registers; KS is small (1..15)
#pragma unroll
    for (int pass = 0; pass < KS; ++pass) {
#pragma unroll
        for (int i = (pass & 1); i + 1 < KS; i += 2) {
            const float lo = fminf(v[i], v[i + 1]);
            const float hi = fmaxf(v[i], v[i + 1]);
            v[i] = lo; v[i + 1] = hi;
        }
    }
    return v[(KS - 1) / 2];
}

// torch.median propagates NaN: a window holding one gives NaN
template <int KS>
__device__ __forceinline__ bool any_nan(const float (&v)[KS])
{
    bool nan = false;
#pragma unroll
    for (int i = 0; i < KS; ++i) nan |= (v[i] != v[i]);
    return nan;
}

// One thread per 4 consecutive pixels (VEC) or per pixel; loops over the C channels.
template <int KS, bool VEC, bool SEM_U8>
__global__ void __launch_bounds__(256)
median_harden_kernel(const PlanePtrs planes, int C, size_t hw, float thr, float* __restrict__ median_out,
                     void* __restrict__ sem_out)
{
    constexpr int N = VEC ? 4 : 1;
    const size_t n_items = VEC ? hw / 4 : hw;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
        const size_t px = it * N;
        float bestv[N];
        int bestc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { bestv[j] = 0.f; bestc[j] = 0; }
        for (int c = 0; c < C; ++c) {
            const size_t e = (size_t)c * hw + px;
            float v[N][KS];
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                if (VEC) {
                    const float4 u = __ldcs(reinterpret_cast<const float4*>(planes.p[k] + e));
                    v[0][k] = u.x; v[1 % N][k] = u.y; v[2 % N][k] = u.z; v[3 % N][k] = u.w;
                } else {
                    v[0][k] = __ldcs(planes.p[k] + e);
                }
            }
            float m[N];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const bool nan = any_nan<KS>(v[j]);
                m[j] = median_of<KS>(v[j]);
                if (nan) m[j] = CUDART_NAN_F;
            }
            if (median_out) {
                if (VEC) *reinterpret_cast<float4*>(median_out + e) = make_float4(m[0], m[1 % N], m[2 % N], m[3 % N]);
                else median_out[e] = m[0];
            }
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (c == 0 || m[j] > bestv[j] || (m[j] != m[j] && bestv[j] == bestv[j])) { bestv[j] = m[j]; bestc[j] = c; }   // argmax: NaN is the maximum
        }
        if (sem_out) {
            int cls[N];
#pragma unroll
            for (int j = 0; j < N; ++j) cls[j] = (C == 1) ? (bestv[j] >= thr ? 1 : 0) : bestc[j];
            if (SEM_U8) {
                unsigned char* o = reinterpret_cast<unsigned char*>(sem_out) + px;
                if (VEC) *reinterpret_cast<unsigned*>(o) = (unsigned)cls[0] | ((unsigned)cls[1 % N] << 8) |
                                                           ((unsigned)cls[2 % N] << 16) | ((unsigned)cls[3 % N] << 24);
                else o[0] = (unsigned char)cls[0];
            } else {
                long long* o = reinterpret_cast<long long*>(sem_out) + px;
                if (VEC) {
                    __stcs(reinterpret_cast<longlong2*>(o), make_longlong2(cls[0], cls[1 % N]));
                    __stcs(reinterpret_cast<longlong2*>(o) + 1, make_longlong2(cls[2 % N], cls[3 % N]));
                } else {
                    o[0] = cls[0];
                }
            }
        }
    }
}

template <int KS>
static int launch_median(const PlanePtrs& pp, int C, size_t hw, float thr, float* median_out, void* sem_out,
                         int sem_u8, bool vec, cudaStream_t st)
{
    const size_t items = vec ? hw / 4 : hw;
    size_t blocks = (items + 255) / 256;
    if (blocks > (size_t)device_sm_count() * 16) blocks = (size_t)device_sm_count() * 16;
    if (blocks < 1) blocks = 1;
    const unsigned g = (unsigned)blocks;
    ProfScope ps(ST_MEDIAN, st);
    if (vec) {
        if (sem_u8) median_harden_kernel<KS, true, true><<<g, 256, 0, st>>>(pp, C, hw, thr, median_out, sem_out);
        else median_harden_kernel<KS, true, false><<<g, 256, 0, st>>>(pp, C, hw, thr, median_out, sem_out);
    } else {
        if (sem_u8) median_harden_kernel<KS, false, true><<<g, 256, 0, st>>>(pp, C, hw, thr, median_out, sem_out);
        else median_harden_kernel<KS, false, false><<<g, 256, 0, st>>>(pp, C, hw, thr, median_out, sem_out);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

}  // namespace emp

using namespace emp;

EMP_API int emp_median_harden(const float* const* planes, int ks, int C, int H, int W, float confidence_thr,
                              float* median_out, void* sem_out, int sem_u8, void* stream)
{
    EMP_REQUIRE(planes != nullptr, EMP_ERR_INVALID, "planes is null");
    EMP_REQUIRE(ks >= 1 && ks <= kMaxKs && (ks & 1), EMP_ERR_INVALID, "median kernel size must be odd and <= %d (got %d)", kMaxKs, ks);
    EMP_REQUIRE(C >= 1 && H > 0 && W > 0, EMP_ERR_INVALID, "bad shape C=%d H=%d W=%d", C, H, W);
    EMP_REQUIRE(!sem_u8 || C <= 256, EMP_ERR_INVALID, "uint8 sem needs C <= 256");
    EMP_REQUIRE(median_out || sem_out, EMP_ERR_INVALID, "nothing to compute");
    PlanePtrs pp;
    const size_t hw = (size_t)H * W;
    bool vec = (hw % 4 == 0);
    for (int i = 0; i < kMaxKs; ++i) pp.p[i] = nullptr;
    for (int i = 0; i < ks; ++i) {
        EMP_REQUIRE(planes[i] != nullptr, EMP_ERR_INVALID, "planes[%d] is null", i);
        pp.p[i] = planes[i];
        vec = vec && ((reinterpret_cast<uintptr_t>(planes[i]) & 15u) == 0);
    }
    if (median_out) vec = vec && ((reinterpret_cast<uintptr_t>(median_out) & 15u) == 0);
    if (sem_out) vec = vec && ((reinterpret_cast<uintptr_t>(sem_out) & 15u) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (ks) {
        case 1:  return launch_median<1>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
        case 3:  return launch_median<3>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
        case 5:  return launch_median<5>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
        case 7:  return launch_median<7>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
        case 9:  return launch_median<9>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
        case 11: return launch_median<11>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
        case 13: return launch_median<13>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
        default: return launch_median<15>(pp, C, hw, confidence_thr, median_out, sem_out, sem_u8, vec, st);
    }
}
