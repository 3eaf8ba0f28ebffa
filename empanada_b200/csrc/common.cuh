// common.cuh — shared helpers for libempanada_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/empanada_b200.h"

#define EMP_API extern "C" __attribute__((visibility("default")))

namespace emp {

void set_error(const char* fmt, ...);

#define EMP_CUDA_CHECK(expr)                                                          \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            emp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                       \
            return EMP_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

#define EMP_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            emp::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// code-map conventions (see DESIGN.md "code map"):
//   0                      -> void
//   1 .. CLS_BASE-1        -> instance id (index into the label LUT)
//   CLS_BASE + c           -> stuff pixel of semantic class c (index into the class LUT)
constexpr uint32_t kClsBase16 = 0xF000u;
constexpr uint32_t kClsBase32 = 0xFFFFF000u;
constexpr int kNumClasses = EMP_MAX_CLASSES;

// Per-tile workspace layout.  Everything in [0, zero_bytes) is cleared by one memset per call.
// Center index: centers binned into a uniform grid of 2^gshift-pixel cells (gshift chosen on the
// device from K, so that a cell holds about one center and there are at most kMaxCells cells).
constexpr int kMaxCells = 16384;

struct WsLayout {
    size_t status, rowcnt, areas, votes, zero_bytes;
    size_t mask, centers, ctr_i, cell_start, cell_fill, sorted, lut, sflags, codes, total;
    int wd;          // mask words per row
    int cls_off;     // label LUT: entries [0, k_cap] map instance ids, [cls_off, cls_off + kNumClasses) map stuff classes
    bool code16;
};

static inline WsLayout ws_layout(int H, int W, int k_cap, int n_things)
{
    WsLayout L;
    if (n_things < 1) n_things = 1;
    L.wd = (W + 31) / 32;
    L.code16 = (uint32_t)k_cap < kClsBase16;  // ids must stay below the class codes
    size_t o = 0;
    L.status = o; o = align_up(o + sizeof(int32_t) * EMP_ST_WORDS, 256);
    L.rowcnt = o; o = align_up(o + sizeof(uint32_t) * (size_t)H, 256);
    L.areas = o;  o = align_up(o + sizeof(uint32_t) * (kNumClasses + 1), 256);   // [kNumClasses]: pixels that are not class-0 stuff
    L.votes = o;  o = align_up(o + sizeof(uint32_t) * ((size_t)k_cap + 1) * n_things, 256);
    L.zero_bytes = o;
    L.mask = o;    o = align_up(o + sizeof(uint32_t) * (size_t)H * L.wd, 256);
    L.centers = o; o = align_up(o + sizeof(float2) * ((size_t)k_cap + 1), 256);     // (cy, cx) = step * (y, x), row-major order
    L.ctr_i = o;   o = align_up(o + sizeof(int2) * ((size_t)k_cap + 1), 256);       // (y, x) clamped to int32, for binning
    L.cell_start = o; o = align_up(o + sizeof(int) * (kMaxCells + 2), 256);
    L.cell_fill = o;  o = align_up(o + sizeof(int) * (kMaxCells + 2), 256);
    L.sorted = o;  o = align_up(o + sizeof(float4) * ((size_t)k_cap + 1), 256);     // (cy, cx, bits of k, -) grouped by cell
    L.cls_off = k_cap + 1;
    L.lut = o;     o = align_up(o + sizeof(int64_t) * ((size_t)k_cap + 1 + kNumClasses), 256);
    // one flag byte per 4 x 64 strip in slots of 16 per block (blocks row-major; a block is 1..16 strips tall)
    L.sflags = o;  o = align_up(o + (size_t)16 * ((W + 63) / 64) * ((H + 3) / 4), 256);
    L.codes = o;   o = align_up(o + (L.code16 ? 2 : 4) * (size_t)H * W, 256);
    L.total = o;
    return L;
}

struct Things {
    long long v[EMP_MAX_THINGS];   // ascending, unique
    int n;
};

int make_things(const int64_t* list, int n, Things* out);

// Optional per-stage device timing (emp_profile_enable / emp_profile_read): one CUDA event pair
// around each kernel launch on the launching stream.  Off by default (no events recorded).
enum { ST_NMS = 0, ST_EMIT = 1, ST_ASSIGN = 2, ST_LUT = 3, ST_APPLY = 4, ST_MEDIAN = 5, ST_RLE_MARK = 6,
       ST_RLE_RUNS = 7, ST_BIN = 8, ST_CHAIN = 9, ST_BLK_KEYS = 10, ST_BLK_MARK = 11, ST_BLK_EMIT = 12, ST_BLK_RUNS = 13,
       ST_BLK_PACK = 14, ST_MEMSET = 15, ST_COUNT = EMP_PROFILE_STAGES };
struct ProfScope {
    ProfScope(int stage, cudaStream_t st);
    ~ProfScope();
    int idx;
    cudaStream_t st;
};   // sorts + dedups; EMP_ERR_INVALID if > max

// ---- batched building blocks shared with stack_block.cu (defined in panoptic.cu) ---------------
int assign_block_items(int H, int W);       // strips per 64-column block the assign kernel uses for an H x W plane
int device_sm_count();
int coarse_ids_batched(int B, const float* hm, size_t hm_stride, const float* off, size_t off_stride, int h, int w,
                       float threshold, int nms_kernel, float step, int32_t* ids_out, size_t ids_stride, int k_cap,
                       char* ws, size_t ws_stride, cudaStream_t st, const unsigned char* need = nullptr, size_t need_stride = 0);
int merge_codes_batched(int B, const unsigned char* sem8, size_t sem_stride, const int32_t* ids, size_t ids_stride, int hc,
                        int wc, int shift, int H, int W, const Things& th, long long label_divisor, long long stuff_area,
                        long long void_label, int k_cap, const int32_t* k_dev, size_t k_dev_stride, char* ws,
                        size_t ws_stride, cudaStream_t st);

int build_luts_batched(int B, int H, int W, const Things& th, long long label_divisor, long long stuff_area, long long void_label,
                       int k_cap, const int32_t* k_dev, size_t k_dev_stride, char* ws, size_t ws_stride, cudaStream_t st);

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ int thing_index(long long c, const Things& th)
{
    int t = -1;
#pragma unroll
    for (int i = 0; i < EMP_MAX_THINGS; ++i)
        if (i < th.n && th.v[i] == c) t = i;
    return t;
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total)
{
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// ---- TMA bulk copy + mbarrier (sm_90+ PTX; SASS: UBLKCP / SYNCS) -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// 1-D bulk copy global -> shared; src, dst 16-byte aligned, bytes a multiple of 16.  Completion is
// signalled on `bar` (complete_tx).  L2 evict-first: the streamed planes are read exactly once.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// One box of a tiled tensor map (cuTensorMapEncodeTiled) global -> shared, rank 3; elements outside
// the tensor are zero-filled and still count towards the barrier's byte total.  SASS: UTMALDG.
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

}  // namespace emp
