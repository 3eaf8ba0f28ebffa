// rle_common.cuh — pieces shared by the two RLE encoders: rle.cu (int64 label map in, emp_rle) and
// stack_block.cu (16-bit code map + label LUT in, emp_stack_block).
#pragma once
#include <string.h>
#include "common.cuh"

namespace emp {

struct RleClasses {
    long long lo[EMP_MAX_LABELS];        // label * L
    long long label[EMP_MAX_LABELS];
    unsigned char ccl[EMP_MAX_LABELS];   // 1: thing class and force_connected
    long long L;
    int n;
};

__device__ __forceinline__ int class_of(long long v, const RleClasses& rc)
{
    int c = -1;
    for (int i = 0; i < rc.n; ++i)
        if (v >= rc.lo[i] && v < rc.lo[i] + rc.L) c = i;
    return (v != 0) ? c : -1;
}

// key-space offset of class ci when n row-runs exist: CCL classes take n keys, others L keys
__device__ __forceinline__ long long key_offset(int ci, int n, const RleClasses& rc)
{
    long long o = 0;
    for (int i = 0; i < ci; ++i) o += rc.ccl[i] ? (long long)n : rc.L;
    return o;
}

__device__ __forceinline__ unsigned run_key(long long v, const RleClasses& rc)
{
    const int c = class_of(v, rc);
    return c >= 0 ? ((unsigned)(c + 1) << 22) | (unsigned)(v - rc.lo[c]) : 0u;
}

__device__ __forceinline__ int uf_find(const int* parent, int x)
{
    int p = parent[x];
    while (p != x) { x = p; p = parent[x]; }
    return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b)
{
    bool done;
    do {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a < b) { const int old = atomicMin(parent + b, a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(parent + a, b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// exclusive in-place scan of data[0..n) by one 1024-thread CTA, 4 items per thread per round
__device__ __forceinline__ int cta_scan_inplace(int* data, long long n, int* s_w)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int carry = 0;
    for (long long base = 0; base < n; base += 4096) {
        const long long i0 = base + (long long)tid * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? data[i0 + k] : 0;
        const int local = v[0] + v[1] + v[2] + v[3];
        int wtot;
        const int wex = warp_excl_scan(local, lane, &wtot);
        if (lane == 0) s_w[warp] = wtot;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) { const int x = s_w[w]; if (w < warp) woff += x; tot += x; }
        int run = carry + woff + wex;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < n) data[i0 + k] = run;
            run += v[k];
        }
        carry += tot;
        __syncthreads();
    }
    return carry;
}

}  // namespace emp

static inline int make_rle_classes(const int64_t* labels, int n_labels, int64_t L, const int64_t* things, int nt,
                            int force_connected, emp::RleClasses* rc)
{
    EMP_REQUIRE(n_labels >= 0 && n_labels <= EMP_MAX_LABELS, EMP_ERR_INVALID, "at most %d labels (got %d)", EMP_MAX_LABELS, n_labels);
    EMP_REQUIRE(L > 0 && L <= (1ll << 22), EMP_ERR_INVALID, "label_divisor must be in (0, 2^22] (got %lld)", (long long)L);
    EMP_REQUIRE(n_labels == 0 || labels, EMP_ERR_INVALID, "labels is null");
    memset(rc, 0, sizeof(*rc));
    rc->n = n_labels;
    rc->L = L;
    for (int i = 0; i < n_labels; ++i) {
        EMP_REQUIRE(labels[i] >= 0 && labels[i] < (1ll << 40), EMP_ERR_INVALID, "label %lld out of range", (long long)labels[i]);
        for (int j = 0; j < i; ++j) EMP_REQUIRE(labels[j] != labels[i], EMP_ERR_INVALID, "duplicate label %lld", (long long)labels[i]);
        rc->label[i] = labels[i];
        rc->lo[i] = labels[i] * L;
        bool thing = false;
        for (int t = 0; t < nt; ++t) thing |= (things[t] == labels[i]);
        rc->ccl[i] = (force_connected && thing) ? 1 : 0;
    }
    return EMP_OK;
}
