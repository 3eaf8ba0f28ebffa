// store_patterns.cu — what a pure label-store stream reaches on this GPU for the store geometries apply_lut can use.
// Build + run on the B200 box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sp profiles/store_patterns.cu && /tmp/sp
// Every kernel writes the same 16 x 4096 x 4096 int64 planes (2.15 GB); time = median of 20 launches (CUDA events).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int H = 4096, W = 4096, B = 16;

// A: linear, one STG.128 per thread per iteration, grid-stride
__global__ void __launch_bounds__(256) k_linear(longlong2* out, size_t n2, long long v)
{
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n2; i += (size_t)gridDim.x * 256) __stcs(out + i, make_longlong2(v, v));
}

// B: block walk — a warp owns a 64-column x (4*items)-row block; per row one 512-byte warp store
template <int LOADS>   // 0: pure stores; 1: one dependent 16-byte flag load per block first; 2: flag load + the value depends on it
__global__ void __launch_bounds__(256) k_blocks(long long* out, const uint4* flags, int items, long long v)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blocks_x = W / 64, blocks_y = H / (4 * items);
    const int blk = blockIdx.x * 8 + warp;
    if (blk >= blocks_x * blocks_y) return;
    long long* pan = out + (size_t)blockIdx.z * H * W;
    long long val = v;
    if (LOADS) {
        const uint4 f = __ldg(flags + (size_t)blockIdx.z * blocks_x * blocks_y + blk);
        if (LOADS == 2) val += f.x;
        else if (f.x == 0xdeadbeefu) return;
    }
    const int by = blk / blocks_x, bx = blk - by * blocks_x;
    long long* p = pan + (size_t)by * 4 * items * W + bx * 64 + 2 * lane;
    for (int r = 0; r < 4 * items; ++r) __stcs(reinterpret_cast<longlong2*>(p + (size_t)r * W), make_longlong2(val, val));
}

// C: 128-column blocks — per row two STG.128 per lane (1 KB contiguous per warp per row)
__global__ void __launch_bounds__(256) k_blocks128(long long* out, int items, long long v)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blocks_x = W / 128, blocks_y = H / (4 * items);
    const int blk = blockIdx.x * 8 + warp;
    if (blk >= blocks_x * blocks_y) return;
    long long* pan = out + (size_t)blockIdx.z * H * W;
    const int by = blk / blocks_x, bx = blk - by * blocks_x;
    long long* p = pan + (size_t)by * 4 * items * W + bx * 128 + 2 * lane;
    for (int r = 0; r < 4 * items; ++r) {
        __stcs(reinterpret_cast<longlong2*>(p + (size_t)r * W), make_longlong2(v, v));
        __stcs(reinterpret_cast<longlong2*>(p + (size_t)r * W + 64), make_longlong2(v, v));
    }
}

// D: row-linear — a CTA owns 4 rows, warp w writes 64-column segments w, w+8, ...
__global__ void __launch_bounds__(256) k_rows(long long* out, long long v)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* pan = out + (size_t)blockIdx.z * H * W + (size_t)blockIdx.x * 4 * W;
    for (int r = 0; r < 4; ++r)
        for (int seg = warp; seg < W / 64; seg += 8)
            __stcs(reinterpret_cast<longlong2*>(pan + (size_t)r * W + seg * 64 + 2 * lane), make_longlong2(v, v));
}

// E: CTA-contiguous — a CTA owns 4 KB-aligned 32 KB chunks (one whole row), every thread 8 x STG.128
__global__ void __launch_bounds__(256) k_cta_rows(longlong2* out, long long v)
{
    longlong2* p = out + ((size_t)blockIdx.z * H + blockIdx.x) * (W / 2);
#pragma unroll
    for (int k = 0; k < 8; ++k) __stcs(p + k * 256 + threadIdx.x, make_longlong2(v, v));
}

template <class F> static float timed(F launch)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    std::vector<float> ts;
    for (int i = 0; i < 20; ++i) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ts.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

int main()
{
    const size_t n = (size_t)B * H * W;
    long long* out; CK(cudaMalloc(&out, n * 8));
    uint4* flags; CK(cudaMalloc(&flags, (size_t)B * 64 * 1024 * 16)); CK(cudaMemset(flags, 0, (size_t)B * 64 * 1024 * 16));
    const double gb = n * 8 / 1e9;
    auto report = [&](const char* name, float ms) { printf("%-58s %.4f ms  %.0f GB/s\n", name, ms, gb / (ms * 1e-3)); };
    report("memset", timed([&] { CK(cudaMemsetAsync(out, 1, n * 8)); }));
    for (int per_sm : {8, 16, 32})
        { char nm[96]; snprintf(nm, 96, "A linear grid-stride, %d CTAs/SM", per_sm); report(nm, timed([&] { k_linear<<<148 * per_sm, 256>>>((longlong2*)out, n / 2, 5); })); }
    report("A linear, one store per thread", timed([&] { k_linear<<<(unsigned)(n / 2 / 256), 256>>>((longlong2*)out, n / 2, 5); }));
    for (int items : {4, 16}) {
        const int nb = (W / 64) * (H / (4 * items));
        dim3 g((nb + 7) / 8, 1, B);
        char nm[96];
        snprintf(nm, 96, "B 64-col blocks x %d strips, pure stores", items); report(nm, timed([&] { k_blocks<0><<<g, 256>>>(out, flags, items, 5); }));
        snprintf(nm, 96, "B 64-col blocks x %d strips, flag load (control dep)", items); report(nm, timed([&] { k_blocks<1><<<g, 256>>>(out, flags, items, 5); }));
        snprintf(nm, 96, "B 64-col blocks x %d strips, flag load (data dep)", items); report(nm, timed([&] { k_blocks<2><<<g, 256>>>(out, flags, items, 5); }));
        const int nb2 = (W / 128) * (H / (4 * items));
        dim3 g2((nb2 + 7) / 8, 1, B);
        snprintf(nm, 96, "C 128-col blocks x %d strips, pure stores", items); report(nm, timed([&] { k_blocks128<<<g2, 256>>>(out, items, 5); }));
    }
    report("D row-linear (CTA = 4 rows)", timed([&] { k_rows<<<dim3(H / 4, 1, B), 256>>>(out, 5); }));
    report("E CTA = one 32 KB row, 8 stores per thread", timed([&] { k_cta_rows<<<dim3(H, 1, B), 256>>>((longlong2*)out, 5); }));
    return 0;
}
