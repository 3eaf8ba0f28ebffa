// panoptic.cu — find_instance_center / group_pixels / merge_semantic_and_instance /
// get_panoptic_segmentation (reference empanada/inference/postprocess.py) as sm_100a kernels.
//
// Kernel chain for one tile (all HBM-bound integer / fp32-compare work, no tensor cores):
//   nms_peaks      hm (4 B/px)                 -> peak bitmask (1 bit/px) + per-row counts
//   emit_centers   bitmask                     -> centers in row-major order, K
//   bin_centers    centers                     -> uniform-grid cell index (counting sort by cell)
//   assign         sem (8 B/px, TMA-staged) + off (8 B/px, thing sectors only)
//                                              -> code map (2 B/px) + votes + stuff areas
//   build_lut      votes                       -> label LUT (K+1)
//   apply_lut      code map (2 B/px)           -> pan (8 B/px)
// DESIGN.md has the data layout, the exactness argument for the culled argmin and the roofline.
#include <cuda.h>       // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved through the runtime
#include <limits.h>
#include <math_constants.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace emp {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// per-stage timing
// ---------------------------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int stage; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mutex;                 // the record list is shared by every thread that launches kernels
static std::vector<ProfRec> g_prof;
static size_t g_prof_used = 0;

ProfScope::ProfScope(int stage, cudaStream_t s) : idx(-1), st(s)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (g_prof_used == g_prof.size()) {
        ProfRec r;
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
        r.stage = 0;
        g_prof.push_back(r);
    }
    idx = (int)g_prof_used++;
    g_prof[idx].stage = stage;
    cudaEventRecord(g_prof[idx].a, st);
}

ProfScope::~ProfScope()
{
    if (idx < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if ((size_t)idx < g_prof.size()) cudaEventRecord(g_prof[idx].b, st);
}

int make_things(const int64_t* list, int n, Things* out)
{
    memset(out, 0, sizeof(*out));
    if (n < 0 || (n > 0 && !list)) { set_error("thing_list is null"); return EMP_ERR_INVALID; }
    long long tmp[1024];
    if (n > 1024) { set_error("too many thing classes (%d)", n); return EMP_ERR_INVALID; }
    for (int i = 0; i < n; ++i) tmp[i] = list[i];
    std::sort(tmp, tmp + n);
    int m = (int)(std::unique(tmp, tmp + n) - tmp);
    if (m > EMP_MAX_THINGS) {
        set_error("at most %d distinct thing classes are supported (got %d)", EMP_MAX_THINGS, m);
        return EMP_ERR_INVALID;
    }
    for (int i = 0; i < m; ++i) out->v[i] = tmp[i];
    out->n = m;
    return EMP_OK;
}

// ---------------------------------------------------------------------------------------------
// K1  nms_peaks — postprocess.py:55-68.
//   peak(y,x)  <=>  v > thr  and  v > 0  and  v >= every value in rows y-lo..y+hi, cols x-lo..x+hi
// (clipped; lo = k/2, hi = k-1-lo).  Thresholding neighbours to -1 first (F.threshold) cannot
// change the comparison because v itself is > thr.  NaNs compare false both ways, like -1.
// A streaming pass flags pixels above threshold (a few % of an EM heat-map); only words holding a
// flagged pixel check neighbours, 3x3 first (that alone rejects every slope pixel of a smooth
// heat-map), the full k x k window only for 3x3 maxima.  Output: one ballot word per 32 pixels
// and a per-row popcount.
// ---------------------------------------------------------------------------------------------
// Full k x k window test of ONE 3x3 maximum at (y, x) with value v, by the whole (converged) warp:
// one neighbour per lane through L1/L2, so a peak costs a couple of load round trips instead of a
// 48-load chain.  Only reached for k >= 4 and only by 3x3 maxima (about one pixel per instance).
__device__ __noinline__ bool nms_window_has_bigger(const float* __restrict__ hm, int H, int W, int y, int x, float v,
                                                   int lo, int hi)
{
    const int side = lo + hi + 1, n = side * side;
    bool bigger = false;
    for (int idx = threadIdx.x & 31; idx < n; idx += 32) {
        const int yy = y + idx / side - lo, xx = x + idx % side - lo;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W && __ldg(hm + (size_t)yy * W + xx) > v) bigger = true;
    }
    return __any_sync(0xffffffffu, bigger);
}

// Persistent, warp-autonomous.  The work item is 4 rows x 256 pixels = 4 full 32-byte sectors of the
// peak mask; lane l takes pixels 32j + l (j = 0..7) of each row, so a ballot over the warp IS the mask
// word.  Items are walked in column strips of kNmsBlkItems items (64 rows), strips dealt round-robin to
// the warps of the grid.
// FAST: each item is one TMA tensor copy — box 1 x 6 x 256 floats of the (B,H,W) view: the 4 rows plus
// one halo row above and below; rows outside the image arrive as zeros, which can neither be a
// candidate nor beat one — into a per-warp ring of kNmsStages shared-memory buffers, issued
// kNmsStages items ahead.  Pixels above threshold (a few %) take the 3x3 test straight from shared
// memory (the two halo columns of a strip come through L1/L2); that alone rejects every slope pixel of
// a smooth heat-map, and only 3x3 maxima go on to the full k x k window.  Without FAST (unaligned or
// tiny planes) the same code reads global memory.
#ifndef EMP_NMS_STAGES
#define EMP_NMS_STAGES 3
#endif
#ifndef EMP_NMS_CTAS_PER_SM
#define EMP_NMS_CTAS_PER_SM 3
#endif
#ifndef EMP_NMS_WORDS
#define EMP_NMS_WORDS 8
#endif
constexpr int kNmsRows = 4, kNmsWords = EMP_NMS_WORDS, kNmsStages = EMP_NMS_STAGES, kNmsWarps = 4, kNmsBlkItems = 16, kNmsCtasPerSm = EMP_NMS_CTAS_PER_SM;
#ifndef EMP_NMS_HALO
#define EMP_NMS_HALO 1
#endif
constexpr int kNmsHalo = EMP_NMS_HALO;          // 1: stage one halo row above and below each item
constexpr int kNmsItemW = kNmsWords * 32, kNmsBoxRows = kNmsRows + 2 * kNmsHalo;
constexpr unsigned kNmsStageFloats = kNmsBoxRows * kNmsItemW;
constexpr unsigned kNmsStageBytes = kNmsStageFloats * sizeof(float);

struct NmsArgs {
    CUtensorMap tmap;                           // FAST: (B, H, W) fp32 view of the heat-maps, box 1 x 6 x 256
    const float* hm; size_t hm_stride;
    char* ws; size_t ws_stride;
    size_t o_mask, o_rowcnt;
    int B, H, W, wd, lo, hi;
    int blocks_x, blocks_y, blk_items;          // column strips per tile row / per tile column, items per strip
    float thr;
};

template <bool FAST>
__global__ void __launch_bounds__(kNmsWarps * 32, kNmsCtasPerSm)
nms_peaks_kernel(const __grid_constant__ NmsArgs a)
{
    extern __shared__ __align__(128) unsigned char dsm[];            // [warp][stage][6 rows][256] fp32
    __shared__ uint64_t s_bar[kNmsWarps][kNmsStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, W = a.W, wd = a.wd, lo = a.lo, hi = a.hi;
    const int per_img = a.blocks_x * a.blocks_y;
    const int n_blocks = per_img * a.B;
    const int total_warps = (int)gridDim.x * kNmsWarps;
    const int gw = (int)blockIdx.x * kNmsWarps + warp;
    const float t0 = fmaxf(a.thr, 0.0f);                            // v > thr and v > 0  <=>  v > max(thr, 0)
    const bool before = lo >= 1, after = hi >= 1;                   // does the window reach to -1 / +1 at all?

    struct Strip { int b, row0, colb, nitems; };
    auto decode = [&](int blk) {
        Strip k;
        k.b = 0; k.row0 = 0; k.colb = 0; k.nitems = 0;
        if (blk < n_blocks) {
            k.b = blk / per_img;
            const int r = blk - k.b * per_img;
            const int by = r / a.blocks_x;
            k.colb = (r - by * a.blocks_x) * kNmsItemW;
            k.row0 = by * (a.blk_items * kNmsRows);
            k.nitems = min(a.blk_items, (H - k.row0 + kNmsRows - 1) / kNmsRows);
        }
        return k;
    };

    float* ring = reinterpret_cast<float*>(dsm) + (size_t)warp * kNmsStages * kNmsStageFloats;
    uint64_t policy = 0;
    int p_blk = gw, p_i = 0, inflight = 0, st_issue = 0;
    Strip kp = decode(p_blk);
    auto pump = [&]() {
        while (inflight < kNmsStages && kp.nitems > 0) {
            if (p_i < kp.nitems) {
                if (lane == 0) {
                    mbar_expect_tx(&s_bar[warp][st_issue], kNmsStageBytes);
                    tma_load_3d(ring + (size_t)st_issue * kNmsStageFloats, &a.tmap, kp.colb, kp.row0 + p_i * kNmsRows - kNmsHalo,
                                kp.b, &s_bar[warp][st_issue], policy);
                }
                ++p_i; ++inflight;
                st_issue = st_issue + 1 == kNmsStages ? 0 : st_issue + 1;
            } else {
                p_blk += total_warps; p_i = 0;
                kp = decode(p_blk);
            }
        }
    };
    if (FAST) {
        policy = l2_policy_evict_normal();              // halo rows are read again by the next item
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kNmsStages; ++s) mbar_init(&s_bar[warp][s], 1);
            mbar_fence_init();
        }
        __syncwarp();
        pump();
    }

    int st_cons = 0;
    unsigned parity = 0;
    for (int blk = gw; blk < n_blocks; blk += total_warps) {
        const Strip kc = decode(blk);
        const float* hm = a.hm + (size_t)kc.b * a.hm_stride;
        char* ws = a.ws + (size_t)kc.b * a.ws_stride;
        uint32_t* mask = reinterpret_cast<uint32_t*>(ws + a.o_mask);
        uint32_t* rowcnt = reinterpret_cast<uint32_t*>(ws + a.o_rowcnt);
        const int w0 = kc.colb >> 5;                                    // first mask word of the strip
        const int x0 = kc.colb + lane;

        for (int it = 0; it < kc.nitems; ++it) {
            const int y0 = kc.row0 + it * kNmsRows;
            // FAST: (row 0 of the item, this lane) inside the stage; row -1 / row 4 are the halo rows
            const float* sp = ring + (size_t)st_cons * kNmsStageFloats + kNmsHalo * kNmsItemW + lane;
            unsigned cm = 0;                                            // bit 8*r + j: pixel (y0+r, x0+32j) is a candidate
            if (FAST) {
                mbar_wait(&s_bar[warp][st_cons], parity);
                // most items hold no pixel above threshold: 8 x LDS.128 and a max tree decide that
                const float4* vp = reinterpret_cast<const float4*>(ring + (size_t)st_cons * kNmsStageFloats + kNmsHalo * kNmsItemW) + lane;
                float mx = -CUDART_INF_F;
#pragma unroll
                for (int r = 0; r < kNmsRows; ++r) {
#pragma unroll
                    for (int h = 0; h < kNmsItemW / 128; ++h) {
                        const float4 v = vp[r * (kNmsItemW / 4) + h * 32];
                        mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));       // fmaxf skips NaNs, like "v > t0" does
                    }
                }
                if (!__any_sync(0xffffffffu, mx > t0)) {                // warp-uniform
                    __syncwarp();
                    --inflight;
                    if (++st_cons == kNmsStages) { st_cons = 0; parity ^= 1u; }
                    pump();
#pragma unroll
                    for (int r = 0; r < kNmsRows; ++r)
                        if (y0 + r < H && lane < kNmsWords && w0 + lane < wd) mask[(size_t)(y0 + r) * wd + w0 + lane] = 0u;
                    continue;
                }
#pragma unroll
                for (int r = 0; r < kNmsRows; ++r) {
#pragma unroll
                    for (int j = 0; j < kNmsWords; ++j) cm |= (sp[r * kNmsItemW + 32 * j] > t0 ? 1u : 0u) << (8 * r + j);
                }
            } else {
                const float* p = hm + (size_t)y0 * W + x0;
#pragma unroll
                for (int r = 0; r < kNmsRows; ++r) {
                    const bool rin = y0 + r < H;
#pragma unroll
                    for (int j = 0; j < kNmsWords; ++j) {
                        float v = -CUDART_INF_F;
                        if (rin && x0 + 32 * j < W) v = __ldg(p + 32 * j);
                        cm |= (v > t0 ? 1u : 0u) << (8 * r + j);
                    }
                    p += W;
                }
            }

            // 3x3 test of this lane's candidates (lanes proceed independently)
            unsigned pk = 0;                                            // bit 8*r + j: 3x3 maximum
            for (unsigned rest = cm; rest; rest &= rest - 1) {
                const int bit = __ffs(rest) - 1;
                const int r = bit >> 3, j = bit & 7;
                const int y = y0 + r, x = x0 + 32 * j, cs = 32 * j + lane;      // cs: column within the item
                const float* q = sp + r * kNmsItemW + 32 * j;
                const float c = FAST ? q[0] : __ldg(hm + (size_t)y * W + x);
                auto nbr = [&](int dy, int dx) -> float {                    // neighbour (y + dy, x + dx); -inf outside the window / image
                    if (((dy < 0 || dx < 0) && !before) || ((dy > 0 || dx > 0) && !after)) return -CUDART_INF_F;
                    if (FAST && cs + dx >= 0 && cs + dx < kNmsItemW && (kNmsHalo || (r + dy >= 0 && r + dy < kNmsRows)))
                        return q[dy * kNmsItemW + dx];
                    const int yy = y + dy, xx = x + dx;                    // no staging, or a halo column of the strip
                    return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(hm + (size_t)yy * W + xx) : -CUDART_INF_F;
                };
                // left / right first: on a smooth heat-map they alone reject every pixel off the vertical ridge line
                float m = fmaxf(nbr(0, -1), nbr(0, 1));
                if (!(m > c)) {
                    m = fmaxf(fmaxf(nbr(-1, 0), nbr(1, 0)), m);
                    if (!(m > c)) m = fmaxf(fmaxf(nbr(-1, -1), nbr(-1, 1)), fmaxf(nbr(1, -1), nbr(1, 1)));
                }
                if (!(m > c)) pk |= 1u << bit;
            }
            __syncwarp();

            // k x k window of the 3x3 maxima, one at a time, by the whole warp (k >= 4 only)
            unsigned fin = pk;                                          // verified peaks
            if (lo >= 2) {
                unsigned pend = pk;
                fin = 0;
                for (;;) {
                    const unsigned bal = __ballot_sync(0xffffffffu, pend != 0u);
                    if (!bal) break;                                    // warp-uniform
                    const int src = __ffs(bal) - 1;
                    const int bit = __ffs(__shfl_sync(0xffffffffu, pend, src)) - 1;
                    const int r = bit >> 3, j = bit & 7;
                    const int y = y0 + r, x = kc.colb + src + 32 * j;
                    float v = 0.f;
                    if (lane == src) v = FAST ? sp[r * kNmsItemW + 32 * j] : __ldg(hm + (size_t)y * W + x);
                    v = __shfl_sync(0xffffffffu, v, src);
                    const bool bigger = nms_window_has_bigger(hm, H, W, y, x, v, lo, hi);
                    if (lane == src) {
                        pend &= pend - 1;
                        if (!bigger) fin |= 1u << bit;
                    }
                }
            }
            if (FAST) {
                __syncwarp();                                           // every lane is done with the stage: the slot is free
                --inflight;
                if (++st_cons == kNmsStages) { st_cons = 0; parity ^= 1u; }
                pump();
            }

            unsigned mine[kNmsRows] = {0, 0, 0, 0};                     // lane j (< 8) ends up with word j of each row
            unsigned todo = __reduce_or_sync(0xffffffffu, fin);         // (row, word) pairs holding a peak
            while (todo) {                                              // warp-uniform
                const int bit = __ffs(todo) - 1;
                todo &= todo - 1;
                const int r = bit >> 3, j = bit & 7;
                const unsigned word = __ballot_sync(0xffffffffu, (fin >> bit) & 1u);
                if (lane == j) {
                    if (r == 0) mine[0] = word; else if (r == 1) mine[1] = word; else if (r == 2) mine[2] = word; else mine[3] = word;
                }
            }
#pragma unroll
            for (int r = 0; r < kNmsRows; ++r) {
                const int y = y0 + r;
                if (y < H) {                                            // warp-uniform
                    const bool wl = lane < kNmsWords && w0 + lane < wd;
                    if (wl) mask[(size_t)y * wd + w0 + lane] = mine[r];
                    const unsigned cnt = __reduce_add_sync(0xffffffffu, wl ? (unsigned)__popc(mine[r]) : 0u);
                    if (cnt && lane == 0) atomicAdd(rowcnt + y, cnt);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2  emit_centers — torch.nonzero(ctr_hmp > 0) order (postprocess.py:75): row-major.
// Each CTA owns 32 rows: prefix = sum of the row counts above it, then a warp per row expands
// the row's ballot words into (y,x) pairs at their exact rank.  No atomics, deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
emit_centers_kernel(char* __restrict__ ws_base, size_t ws_stride, size_t o_mask, size_t o_rowcnt,
                    size_t o_centers, size_t o_ctr_i, size_t o_status, int H, int wd, int k_cap, float step,
                    int64_t* __restrict__ ctr_out_base, size_t ctr_out_stride, int cap)
{
    char* ws = ws_base + (size_t)blockIdx.z * ws_stride;
    const uint32_t* mask = reinterpret_cast<const uint32_t*>(ws + o_mask);
    const uint32_t* rowcnt = reinterpret_cast<const uint32_t*>(ws + o_rowcnt);
    float2* centers = reinterpret_cast<float2*>(ws + o_centers);     // (cy, cx) = step * (y, x)
    int2* ctr_i = reinterpret_cast<int2*>(ws + o_ctr_i);
    int32_t* status = reinterpret_cast<int32_t*>(ws + o_status);
    int64_t* ctr_out = ctr_out_base ? ctr_out_base + (size_t)blockIdx.z * ctr_out_stride : nullptr;

    __shared__ int s_part[8];
    __shared__ int s_off[33];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * 32;

    int part = 0;
    for (int i = tid; i < r0; i += 256) part += (int)rowcnt[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (warp == 0) {
        int prefix = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) prefix += s_part[w];
        const int c = (r0 + lane < H) ? (int)rowcnt[r0 + lane] : 0;
        int tot;
        const int ex = warp_excl_scan(c, lane, &tot);
        s_off[lane] = prefix + ex;
        if (lane == 31) s_off[32] = prefix + tot;
    }
    __syncthreads();

    for (int j = 0; j < 4; ++j) {
        const int rr = warp * 4 + j;
        const int y = r0 + rr;
        if (y >= H) break;
        const int base = s_off[rr];
        const int cnt = s_off[rr + 1] - base;
        if (cnt == 0) continue;
        int running = 0;
        for (int wb = 0; wb < wd; wb += 32) {
            const int wi = wb + lane;
            unsigned word = wi < wd ? mask[(size_t)y * wd + wi] : 0u;
            int tot;
            const int ex = warp_excl_scan(__popc(word), lane, &tot);
            int pos = base + running + ex;
            while (word) {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                const int x = wi * 32 + b;
                if (pos < k_cap) {
                    centers[pos] = make_float2(__fmul_rn(step, (float)y), __fmul_rn(step, (float)x));
                    ctr_i[pos] = make_int2(y, x);
                }
                if (ctr_out && pos < cap) { ctr_out[2 * (size_t)pos] = y; ctr_out[2 * (size_t)pos + 1] = x; }
                ++pos;
            }
            running += tot;
            if (running >= cnt) break;          // warp-uniform
        }
    }
    if (r0 + 32 >= H && tid == 0) {
        const int K = s_off[32];
        status[EMP_ST_K] = K;
        if (K > k_cap) atomicOr(status + EMP_ST_FLAGS, EMP_FLAG_K_OVERFLOW);
    }
}

// int64 (K,2) centers supplied by the caller (standalone group_pixels) -> center tables
__global__ void load_centers_kernel(const int64_t* __restrict__ ctr, int K, float step, float2* __restrict__ centers,
                                    int2* __restrict__ ctr_i)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) {        // ctr = step * ctr: int64 -> float32, then one rounded product (postprocess.py:152)
        const long long y = ctr[2 * (size_t)i], x = ctr[2 * (size_t)i + 1];
        centers[i] = make_float2(__fmul_rn(step, (float)y), __fmul_rn(step, (float)x));
        ctr_i[i] = make_int2((int)max(min(y, (long long)INT_MAX), (long long)INT_MIN),
                             (int)max(min(x, (long long)INT_MAX), (long long)INT_MIN));
    }
}

// ---------------------------------------------------------------------------------------------
// K2b  bin_centers — a uniform-grid index over the centers of one tile, built once per tile by one
// CTA: cell size 2^gshift pixels with gshift chosen from K (about one center per cell, at most
// kMaxCells cells), counting sort of the centers by cell.  Within a cell the order is whatever the
// atomics give; the search below compares (distance, index) lexicographically, so the result does
// not depend on it.
//   cell of a center = (clamp(y, 0, H-1) >> gshift, clamp(x, 0, W-1) >> gshift)   (integer pixels)
//   sorted[j]        = (step*y, step*x, bits(k), -)    with the reference's rounding (postprocess.py:152)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int grid_cells(int n, int gs) { return ((n - 1) >> gs) + 1; }

__global__ void __launch_bounds__(1024)
bin_centers_kernel(char* __restrict__ ws_base, size_t ws_stride, size_t o_status, size_t o_centers, size_t o_ctr_i,
                   size_t o_cell_start, size_t o_cell_fill, size_t o_sorted, int H, int W, int k_cap, int k_fixed,
                   float step)
{
    char* ws = ws_base + (size_t)blockIdx.z * ws_stride;
    int32_t* status = reinterpret_cast<int32_t*>(ws + o_status);
    const float2* centers = reinterpret_cast<const float2*>(ws + o_centers);
    const int2* ctr_i = reinterpret_cast<const int2*>(ws + o_ctr_i);
    int* cell_start = reinterpret_cast<int*>(ws + o_cell_start);
    int* cell_fill = reinterpret_cast<int*>(ws + o_cell_fill);
    float4* sorted = reinterpret_cast<float4*>(ws + o_sorted);

    __shared__ int s_warp[32];
    __shared__ int s_k, s_gs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const int K = k_fixed >= 0 ? k_fixed : min(__ldcg(status + EMP_ST_K), k_cap);
        int gs = 30;                                        // one cell: the search degenerates to all K centers
        if (K > 0 && step > 0.0f && step < CUDART_INF_F) {
            const float target = (float)H * (float)W / (float)K;          // pixels per center
            gs = max(3, (int)floorf(0.5f * log2f(target)));
            if (gs > 30) gs = 30;
            while (gs < 30 && (long long)grid_cells(H, gs) * grid_cells(W, gs) > kMaxCells) ++gs;
        }
        s_k = K; s_gs = gs;
        status[EMP_ST_GSHIFT] = gs;
    }
    __syncthreads();
    const int K = s_k, gs = s_gs;
    const int ncy = grid_cells(H, gs), ncx = grid_cells(W, gs), n = ncy * ncx;

    for (int i = tid; i <= n; i += 1024) cell_fill[i] = 0;
    __syncthreads();
    for (int k = tid; k < K; k += 1024) {
        const int2 c = ctr_i[k];
        const int cell = (min(max(c.x, 0), H - 1) >> gs) * ncx + (min(max(c.y, 0), W - 1) >> gs);
        atomicAdd(cell_fill + cell, 1);
    }
    __syncthreads();
    // exclusive scan of the n counts: each thread owns a contiguous chunk
    const int per = (n + 1023) / 1024;
    const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += cell_fill[i];
    int tot;
    int ex = warp_excl_scan(sum, lane, &tot);
    if (lane == 31) s_warp[warp] = tot;
    __syncthreads();
    if (warp == 0) {
        int wt;
        const int wex = warp_excl_scan(s_warp[lane], lane, &wt);
        s_warp[lane] = wex;
    }
    __syncthreads();
    int run = s_warp[warp] + ex;
    for (int i = i0; i < i1; ++i) {
        const int c = cell_fill[i];
        cell_start[i] = run;
        cell_fill[i] = run;
        run += c;
    }
    if (tid == 1023) cell_start[n] = run;                   // == K
    __syncthreads();
    for (int k = tid; k < K; k += 1024) {
        const int2 c = ctr_i[k];
        const int cell = (min(max(c.x, 0), H - 1) >> gs) * ncx + (min(max(c.y, 0), W - 1) >> gs);
        const int pos = atomicAdd(cell_fill + cell, 1);
        const float2 f = centers[k];
        sorted[pos] = make_float4(f.x, f.y, __int_as_float(k), 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// K3  assign — group_pixels (postprocess.py:146-167, :97-116) fused with the thing mask of
// get_instance_segmentation (:207-221) and the vote / stuff-area pass of
// merge_semantic_and_instance (:253-294).
//
// Persistent kernel, warp-autonomous.  Warps draw 64 x 64-pixel blocks (the first statically, the
// rest from a device-wide counter, so instance-heavy blocks balance out) and walk each block as 16
// strips of 4 rows x 64 columns; lane l owns columns 2l, 2l+1 of each row, so every warp-wide access
// covers one whole 64-pixel row segment: 512 B of int64 sem, 256 B of an offset plane (LDG.64),
// 128 B of uint16 codes (STG.32).  The sem plane — the only stream every pixel needs — is staged
// through a per-warp ring of kStages shared-memory buffers, each filled by ONE TMA tensor copy
// (cp.async.bulk.tensor.3d, box 1 x 4 x 64 of the (B,H,W) view, out-of-image elements zero-filled by
// the hardware, completion on an mbarrier): lane 0 issues the strip kStages iterations ahead — across
// block boundaries — so loads stay in flight while the warp classifies, searches and stores.  There
// is no block-wide barrier anywhere in the loop.
//
// Exact nearest center.  For each pixel the reference takes, over ALL K centers,
//       d_k = sqrt_rn(fma(dx, dx, rn(dy*dy))),  dy = cy_k - ly,  dx = cx_k - lx   (fp32)
// and keeps the first minimum.  The warp bounds the cells of its thing pixels' shifted locations
// (ly,lx) by a box, widens it by r cells (r = 1, 2, 4, ...) and evaluates d_k — in exactly that
// arithmetic — for the centers binned in the block, keeping the lexicographic minimum of (d_k, k).
// A pixel is settled when its best d is below 0.9999 x its distance to the block's outer edge: every
// center outside the block is farther than that edge in y or in x, far beyond the few-ulp rounding
// of d_k, so it can neither win nor tie.  (Block edges are computed as step * (cell << gshift) with
// the rounding the centers themselves got, which is monotone; a block side on the grid border is
// +-inf because centers beyond the image are binned into border cells.)  If any pixel of the warp is
// not settled the ring doubles; a block covering the whole grid is the reference's full search.
// ---------------------------------------------------------------------------------------------
enum { SEM_NONE = 0, SEM_I64 = 1, SEM_U8 = 2 };
enum { ID_ARGMIN = 0, ID_DENSE = 1, ID_COARSE = 2 };
enum { OUT_CODE16 = 0, OUT_CODE32 = 1, OUT_IDS64 = 2, OUT_IDS32 = 3 };

struct AssignArgs {
    CUtensorMap tmap;                          // FAST: (B, H, W) view of the sem planes, box 1 x 4 x 64
    const void* sem;    size_t sem_stride;     // elements per tile
    const float* off;   size_t off_stride;     // floats per tile (2*H*W)
    const void* ids_in; size_t ids_stride;     // ID_DENSE: int64 H*W, ID_COARSE: int32 hc*wc
    void* out;          size_t out_stride;     // elements per tile
    char* ws;           size_t ws_stride;
    size_t o_status, o_votes, o_areas, o_cell_start, o_sorted, o_sflags;
    int32_t* counter;                          // zeroed device word: dynamic block hand-out
    int B, H, W, wc, shift;
    int blocks_x, blocks_y, blk_items;         // blocks (blk_items strips of 4 x 64) per tile row / column
    float step;
    int chunksize, k_cap, k_fixed;             // k_fixed >= 0: K known on the host
    long long max_id;
    int vec;                                   // host: W % 4 == 0 and all planes 16-byte aligned
    int fast;                                  // host: launch the FAST instantiation (vector accesses + TMA staging)
    unsigned long long thing_bits;             // bit c set <=> class c (< 64) is a thing class
    int things_small;                          // every thing class is < 64 (bit test suffices)
    Things things;
};

#ifndef EMP_ASSIGN_CTAS_PER_SM
#define EMP_ASSIGN_CTAS_PER_SM 4
#endif
constexpr int kItemW = 64, kItemH = 4, kAssignThreads = 128, kAssignWarps = 4, kAssignCtasPerSm = EMP_ASSIGN_CTAS_PER_SM, kPx = 8;
#ifndef EMP_ASSIGN_STAGES
#define EMP_ASSIGN_STAGES 3
#endif
constexpr int kStages = EMP_ASSIGN_STAGES;
constexpr int kBlkItems = 16;                  // most strips per block (64 rows); small images use shorter blocks
#ifndef EMP_ASSIGN_STAGE_ITEMS
#define EMP_ASSIGN_STAGE_ITEMS 1
#endif
// strips per TMA box / ring stage.  Measured on config 2: 1 -> assign 0.58 ms, 2 -> 0.67 ms, 4 -> 0.84 ms: halving
// the barrier waits and issues does not pay for the coarser prefetch (and the larger ring), so one strip per box.
constexpr int kStageItems = EMP_ASSIGN_STAGE_ITEMS;
constexpr unsigned kInfoThing = 0x8000u, kInfoBad = 0x4000u;   // per-pixel 16-bit info word
constexpr unsigned kNoKey = 0xFFFFFFFFu;

struct LutScratch {
    int run[EMP_MAX_THINGS];
    int wcnt[8];
};

// 16-bit info word of one pixel: thing -> kInfoThing | thing index, stuff -> class id,
// class outside [0, 4096) -> kInfoBad (reported through EMP_FLAG_CLASS_RANGE).
__device__ __noinline__ unsigned classify_slow(long long v, unsigned long long thing_bits, int things_small,
                                               const Things& things)
{
    if ((unsigned long long)v < 64ull) {
        const unsigned c = (unsigned)v;
        if ((thing_bits >> c) & 1ull)
            return kInfoThing | (unsigned)__popcll(thing_bits & ((1ull << c) - 1ull));
        return c;
    }
    if (!things_small) {
        const int t = thing_index(v, things);
        if (t >= 0) return kInfoThing | (unsigned)t;
    }
    if (v < 0 || v >= kNumClasses) return kInfoBad;
    return (unsigned)v;
}

// class ids known to be < 64: one shift / mask, no branches
__device__ __forceinline__ unsigned classify_small(unsigned c, unsigned long long tbits, bool multi)
{
    const unsigned th = (unsigned)(tbits >> c) & 1u;
    const unsigned t = multi ? (unsigned)__popcll(tbits & ((1ull << c) - 1ull)) : 0u;
    return th ? (kInfoThing | t) : c;
}

// rare paths are kept out of line so that the per-pixel code stays small
__device__ __noinline__ void count_slow(uint32_t* p, unsigned v) { atomicAdd(p, v); }

__device__ __forceinline__ unsigned classify(long long v, unsigned long long tbits, bool multi, int things_small,
                                             const Things& things)
{
    if ((unsigned long long)v < 64ull) return classify_small((unsigned)v, tbits, multi);
    return classify_slow(v, tbits, things_small, things);
}

// merge_semantic_and_instance's bookkeeping (postprocess.py:263-281), by one 256-thread CTA:
//   id -> majority thing class (ties -> smallest class, torch.mode) * L + 1-based rank among voted
//   ids of that class in ascending id order.
__device__ void build_label_lut(long long K, const uint32_t* votes, const Things& things, long long label_divisor,
                                long long void_label, long long* lut, LutScratch& sc)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = things.n;
    const int T = nt > 0 ? nt : 1;
    if (tid < EMP_MAX_THINGS) sc.run[tid] = 0;
    if (tid == 0) lut[0] = void_label;
    __syncthreads();
    for (long long base = 1; base <= K; base += 256) {
        const long long id = base + tid;
        int t = -1;
        if (id <= K) {
            uint32_t best = 0;
            for (int c = 0; c < T; ++c) {
                const uint32_t v = __ldcg(votes + (size_t)id * T + c);
                if (v > best) { best = v; t = c; }
            }
            if (t < 0) lut[id] = void_label;
        }
        for (int c = 0; c < nt; ++c) {
            const bool f = (t == c);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) sc.wcnt[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) { const int x = sc.wcnt[w8]; if (w8 < warp) woff += x; tot += x; }
            if (f) lut[id] = things.v[c] * label_divisor + (long long)(sc.run[c] + woff + __popc(bal & lanemask_lt()) + 1);
            __syncthreads();
            if (tid == 0) sc.run[c] += tot;
        }
        __syncthreads();
    }
}

// number of instance ids the label LUT must cover (device-side counts resolved here)
__device__ __forceinline__ long long lut_extent(int k_fixed, int k_cap, const int32_t* status, const int32_t* k_dev)
{
    long long K = k_fixed >= 0 ? (long long)k_fixed : (long long)min(__ldcg(status + EMP_ST_K), k_cap);
    if (k_dev) K = min(K, (long long)max(__ldcg(k_dev), 0));
    return K;
}

struct GridView {
    const int* cell_start;
    const float4* sorted;
    int gs, ncy, ncx;
    float inv_cell;         // 1 / (step * 2^gs)
    float step;
};

// The reference's own comparison for ONE pixel over the centers binned in cell rows ya..yb, columns
// xa..xb: minimum of the rounded distances sqrt_rn(s), ties to the lower index.  Out of line and
// scalar: it only runs for pixels that met a near-tie in the sqrt-free pass below (exact ties need
// integer-like offsets).  Returns (bits of s of the winner) << 32 | index.
__device__ __noinline__ unsigned long long precise_pixel(const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                                                         int ncx, int ya, int yb, int xa, int xb, float ly, float lx)
{
    float bd = CUDART_INF_F, bs = CUDART_INF_F;
    int bk = INT_MAX;
    for (int row = ya; row <= yb; ++row) {
        const int s = __ldg(cell_start + row * ncx + xa);
        const int e = __ldg(cell_start + row * ncx + xb + 1);
        for (int j = s; j < e; ++j) {
            const float4 c = __ldg(sorted + j);
            const int ck = __float_as_int(c.z);
            const float dy = __fsub_rn(c.x, ly);
            const float dx = __fsub_rn(c.y, lx);
            const float s2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
            const float d = __fsqrt_rn(s2);
            if (d < bd || (d == bd && ck < bk)) { bd = d; bs = s2; bk = ck; }
        }
    }
    return ((unsigned long long)__float_as_uint(bs) << 32) | (unsigned)bk;
}

// Exact nearest center for the warp's thing pixels (8 per lane), see the comment above.  Called by
// all 32 lanes (converged).  bs / bk come back as (d_k^2, k) of the lexicographic minimum of
// (d_k, k) over all centers; bk stays INT_MAX when no distance compared below +inf (non-finite
// locations).
//
// The pass over a block keeps s = d^2 and evaluates no sqrt: sqrt_rn is monotone, and two s that
// differ by more than 1e-6 relative have sqrt_rn values several ulps apart, so
// "s2 < bs * (1 - 1e-6)" is a strict win and "s2 > bs * (1 + 1e-6)" a strict loss.  A pixel that
// meets a candidate inside that sliver is redone by precise_pixel().
__device__ __forceinline__ void grid_nearest_8px(const GridView& g, unsigned thing, const float (&ly)[kPx],
                                                 const float (&lx)[kPx], float (&bs)[kPx], int (&bk)[kPx])
{
    // box of the shifted locations (fminf / fmaxf skip NaNs), then its cell range over the warp
    float fy0 = CUDART_INF_F, fy1 = -CUDART_INF_F, fx0 = CUDART_INF_F, fx1 = -CUDART_INF_F;
    float nanacc = 0.f;                                 // stays 0 unless a thing pixel's location is NaN / inf
#pragma unroll
    for (int p = 0; p < kPx; ++p) {
        bs[p] = CUDART_INF_F; bk[p] = INT_MAX;
        if ((thing >> p) & 1u) {
            fy0 = fminf(fy0, ly[p]); fy1 = fmaxf(fy1, ly[p]);
            fx0 = fminf(fx0, lx[p]); fx1 = fmaxf(fx1, lx[p]);
            nanacc = __fmaf_rn(ly[p], 0.f, __fmaf_rn(lx[p], 0.f, nanacc));
        }
    }
    unsigned fin = thing;                               // thing pixels with a finite location
    if (__any_sync(0xffffffffu, !(nanacc == 0.f))) {    // warp-uniform, rare: sort out which ones
        fin = 0;
#pragma unroll
        for (int p = 0; p < kPx; ++p)
            if (((thing >> p) & 1u) && fabsf(ly[p]) < CUDART_INF_F && fabsf(lx[p]) < CUDART_INF_F) fin |= 1u << p;
        fy0 = fx0 = CUDART_INF_F; fy1 = fx1 = -CUDART_INF_F;
#pragma unroll
        for (int p = 0; p < kPx; ++p) {
            if ((fin >> p) & 1u) {
                fy0 = fminf(fy0, ly[p]); fy1 = fmaxf(fy1, ly[p]);
                fx0 = fminf(fx0, lx[p]); fx1 = fmaxf(fx1, lx[p]);
            }
        }
    }
    int y0 = INT_MAX, y1 = -1, x0 = INT_MAX, x1 = -1;
    if (fin) {
        y0 = min(max(__float2int_rd(fy0 * g.inv_cell), 0), g.ncy - 1);
        y1 = min(max(__float2int_rd(fy1 * g.inv_cell), 0), g.ncy - 1);
        x0 = min(max(__float2int_rd(fx0 * g.inv_cell), 0), g.ncx - 1);
        x1 = min(max(__float2int_rd(fx1 * g.inv_cell), 0), g.ncx - 1);
    }
    y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
    x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
    if (y1 < 0) return;                                 // warp-uniform: nothing finite to search for

    // ring 0 is the box's own cells: a location within a pixel or two of its center settles there
    // unless it sits right at a cell border
#pragma unroll 1
    for (int r = 0;; r = r ? 2 * r : 1) {
        const int ya = max(y0 - r, 0), yb = min(y1 + r, g.ncy - 1);
        const int xa = max(x0 - r, 0), xb = min(x1 + r, g.ncx - 1);
        unsigned near = 0;
#pragma unroll
        for (int p = 0; p < kPx; ++p) { bs[p] = CUDART_INF_F; bk[p] = INT_MAX; }
#pragma unroll 1
        for (int row = ya; row <= yb; ++row) {
            const int s = __ldg(g.cell_start + row * g.ncx + xa);
            const int e = __ldg(g.cell_start + row * g.ncx + xb + 1);
#pragma unroll 1
            for (int j = s; j < e; ++j) {
                const float4 c = __ldg(g.sorted + j);
                const int ck = __float_as_int(c.z);
#pragma unroll
                for (int p = 0; p < kPx; ++p) {
                    const float dy = __fsub_rn(c.x, ly[p]);
                    const float dx = __fsub_rn(c.y, lx[p]);
                    const float s2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
                    const bool better = s2 < __fmul_rn(bs[p], 0.999999f);
                    near |= ((!better && s2 <= __fmul_rn(bs[p], 1.000001f)) ? 1u : 0u) << p;
                    bs[p] = better ? s2 : bs[p];
                    bk[p] = better ? ck : bk[p];
                }
            }
        }
        near &= fin;
        if (near) {                                     // rare, divergent
#pragma unroll
            for (int p = 0; p < kPx; ++p) {
                if ((near >> p) & 1u) {
                    const unsigned long long v = precise_pixel(g.cell_start, g.sorted, g.ncx, ya, yb, xa, xb, ly[p], lx[p]);
                    bs[p] = __uint_as_float((unsigned)(v >> 32));
                    bk[p] = (int)(unsigned)v;
                }
            }
        }
        __syncwarp();
        if (ya == 0 && yb == g.ncy - 1 && xa == 0 && xb == g.ncx - 1) break;     // every center seen
        const float Ylo = ya == 0 ? -CUDART_INF_F : __fmul_rn(g.step, (float)(ya << g.gs));
        const float Yhi = yb == g.ncy - 1 ? CUDART_INF_F : __fmul_rn(g.step, (float)((yb + 1) << g.gs));
        const float Xlo = xa == 0 ? -CUDART_INF_F : __fmul_rn(g.step, (float)(xa << g.gs));
        const float Xhi = xb == g.ncx - 1 ? CUDART_INF_F : __fmul_rn(g.step, (float)((xb + 1) << g.gs));
        // settled iff d < 0.9999 * m, m = distance to the block's outer edge; tested on squares with the
        // slack rounded in our favour.  First per lane — its pixels' box against its largest s — which
        // almost always decides; pixel by pixel only if some lane fails that.
        float smax = 0.f;
#pragma unroll
        for (int p = 0; p < kPx; ++p) smax = ((fin >> p) & 1u) ? fmaxf(smax, bs[p]) : smax;
        const float ml = fminf(fminf(fy0 - Ylo, Yhi - fy1), fminf(fx0 - Xlo, Xhi - fx1));
        bool ok = fin == 0u || (ml > 0.f && smax < 0.9997f * (ml * ml));
        if (!__all_sync(0xffffffffu, ok)) {             // warp-uniform
            ok = true;
#pragma unroll
            for (int p = 0; p < kPx; ++p) {
                const float m = fminf(fminf(ly[p] - Ylo, Yhi - ly[p]), fminf(lx[p] - Xlo, Xhi - lx[p]));
                if (((fin >> p) & 1u) && !(m > 0.f && bs[p] < 0.9997f * (m * m))) ok = false;
            }
            if (!__all_sync(0xffffffffu, ok)) continue;
        }
        break;
    }
    // non-finite locations: every d_k is +inf or NaN; the caller maps "no index" to the reference's answer
#pragma unroll
    for (int p = 0; p < kPx; ++p)
        if (!((fin >> p) & 1u)) { bs[p] = CUDART_INF_F; bk[p] = INT_MAX; }
}

// a block of kBlkItems strips (64 rows x 64 columns) of one tile: the unit handed out to warps
struct Blk {
    int b, row0, colb, nitems;      // nitems == 0: no such block (past the end)
};

template <int SEM, int IDM, int OUT, bool FAST>
__global__ void __launch_bounds__(kAssignThreads, kAssignCtasPerSm)
assign_kernel(const __grid_constant__ AssignArgs a)
{
    constexpr bool kCodes = (OUT == OUT_CODE16 || OUT == OUT_CODE32);
    constexpr bool kTma = FAST && SEM != SEM_NONE;
    constexpr uint32_t kClsBase = (OUT == OUT_CODE16) ? kClsBase16 : kClsBase32;
    constexpr int kRowBytes = (SEM == SEM_I64) ? kItemW * 8 : kItemW;       // one staged row of sem
    constexpr unsigned kItemBytes = kItemH * kRowBytes;
    constexpr unsigned kStageBytes = kStageItems * kItemBytes;
    extern __shared__ __align__(128) unsigned char dsm[];                   // [warp][stage][row][kRowBytes]
    __shared__ uint64_t s_bar[kAssignWarps][kStages];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;
    const size_t HW = (size_t)H * W;
    const int T = a.things.n > 0 ? a.things.n : 1;
    const bool multi = a.things.n > 1;
    const bool class0_stuff = !(a.thing_bits & 1ull);
    const int per_img = a.blocks_x * a.blocks_y;           // host guarantees per_img * B + warps < 2^31
    const int n_blocks = per_img * a.B;
    const int total_warps = (int)gridDim.x * kAssignWarps;

    auto decode = [&](int blk) {
        Blk k;
        k.b = a.B; k.row0 = 0; k.colb = 0; k.nitems = 0;
        if (blk < n_blocks) {
            k.b = blk / per_img;
            const int r = blk - k.b * per_img;
            const int by = r / a.blocks_x;
            k.colb = (r - by * a.blocks_x) * kItemW;
            k.row0 = by * (a.blk_items * kItemH);
            k.nitems = min(a.blk_items, (H - k.row0 + kItemH - 1) / kItemH);
        }
        return k;
    };
    // the first block of every warp is static; the rest come from a device-wide counter, so warps
    // that drew instance-heavy blocks simply take fewer of them
    auto grab = [&]() {
        int v = 0;
        if (lane == 0) v = atomicAdd(a.counter, 1);
        return total_warps + __shfl_sync(0xffffffffu, v, 0);
    };
    int cur = (int)blockIdx.x * kAssignWarps + warp;
    int nxt = cur < n_blocks ? grab() : n_blocks;
    Blk kc = decode(cur), kn = decode(nxt);

    // ---- TMA ring: one box of kStageItems strips (64 x 8 pixels) of the sem plane per stage, kStages boxes ahead ----
    unsigned char* ring = dsm + (size_t)warp * kStages * kStageBytes;
    uint64_t policy = 0;
    int p_which = 0, p_i = 0, inflight = 0, st_issue = 0;   // prefetch position: block (0 cur, 1 nxt, 2 beyond), box within it
    auto pump = [&]() {
        while (inflight < kStages && p_which < 2) {
            const int nit = ((p_which == 0 ? kc.nitems : kn.nitems) + kStageItems - 1) / kStageItems;     // boxes of the block
            if (p_i < nit) {
                if (lane == 0) {
                    const int pb = p_which == 0 ? kc.b : kn.b;
                    const int prow = (p_which == 0 ? kc.row0 : kn.row0) + p_i * (kStageItems * kItemH);
                    const int pcol = p_which == 0 ? kc.colb : kn.colb;
                    mbar_expect_tx(&s_bar[warp][st_issue], kStageBytes);
                    tma_load_3d(ring + (size_t)st_issue * kStageBytes, &a.tmap, pcol, prow, pb, &s_bar[warp][st_issue], policy);
                }
                ++p_i; ++inflight;
                st_issue = st_issue + 1 == kStages ? 0 : st_issue + 1;
            } else if (nit == 0) {
                break;                                      // no such block: nothing further is known yet
            } else {
                ++p_which; p_i = 0;
            }
        }
    };
    if (kTma) {
        policy = l2_policy_evict_first();
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) mbar_init(&s_bar[warp][s], 1);
            mbar_fence_init();
        }
        __syncwarp();
        pump();
    }

    // per-image state
    int cur_b = -1;
    int K = 0;
    bool chunked = false;
    GridView g;
    g.cell_start = nullptr; g.sorted = nullptr; g.gs = 30; g.ncy = 1; g.ncx = 1; g.inv_cell = 0.f; g.step = a.step;
    int32_t* status = nullptr;
    uint32_t* votes = nullptr;
    uint32_t* areas = nullptr;
    unsigned char* sflags = nullptr;    // one byte per strip: 1 = all class-0 stuff (no codes stored), 0 = see the code map
    unsigned deficit_acc = 0;       // in-image pixels of the current image that are NOT class-0 stuff
    unsigned akey_acc = kNoKey, acnt_acc = 0;    // pending stuff-area count of one non-zero class
    int flags = 0;

    auto flush_image = [&]() {      // warp-uniform call sites only
        if (cur_b < 0) return;
        if (kCodes) {
            const unsigned dsum = __reduce_add_sync(0xffffffffu, deficit_acc);
            if (dsum && lane == 0) atomicAdd(areas + kNumClasses, dsum);
            const unsigned peers = __match_any_sync(0xffffffffu, akey_acc);
            const unsigned asum = __reduce_add_sync(peers, acnt_acc);
            if (akey_acc != kNoKey && asum && lane == __ffs(peers) - 1) atomicAdd(areas + akey_acc, asum);
        }
        const int f = __reduce_or_sync(0xffffffffu, flags);
        if (f && lane == 0) atomicOr(status + EMP_ST_FLAGS, f);
        deficit_acc = 0; akey_acc = kNoKey; acnt_acc = 0; flags = 0;
    };

    int st_cons = 0;
    unsigned parity = 0;
    while (kc.nitems > 0) {                                 // warp-uniform
        const int b = kc.b;
        if (b != cur_b) {
            flush_image();
            cur_b = b;
            char* ws = a.ws + (size_t)b * a.ws_stride;
            status = reinterpret_cast<int32_t*>(ws + a.o_status);
            votes = reinterpret_cast<uint32_t*>(ws + a.o_votes);
            areas = reinterpret_cast<uint32_t*>(ws + a.o_areas);
            sflags = reinterpret_cast<unsigned char*>(ws + a.o_sflags);
            if (IDM == ID_ARGMIN) {
                K = a.k_fixed >= 0 ? a.k_fixed : min(__ldcg(status + EMP_ST_K), a.k_cap);
                chunked = K > a.chunksize;
                g.gs = __ldcg(status + EMP_ST_GSHIFT);
                g.ncy = grid_cells(H, g.gs); g.ncx = grid_cells(W, g.gs);
                g.inv_cell = __frcp_rn(__fmul_rn(a.step, (float)(1 << g.gs)));
                g.cell_start = reinterpret_cast<const int*>(ws + a.o_cell_start);
                g.sorted = reinterpret_cast<const float4*>(ws + a.o_sorted);
            }
        }
        const int col0 = kc.colb + 2 * lane;
        const bool cols_full = kc.colb + kItemW <= W;
        // strip flags of a block are contiguous: [(block row * blocks_x + block column) * kBlkItems + strip]
        unsigned char* bflags = sflags + ((size_t)(kc.row0 / (a.blk_items * kItemH)) * a.blocks_x + kc.colb / kItemW) * kBlkItems;
        const float xc0 = __fmul_rn((float)col0, a.step), xc1 = __fmul_rn((float)(col0 + 1), a.step);

        for (int it = 0; it < kc.nitems; ++it) {
            const int row0 = kc.row0 + it * kItemH;
            const bool full = cols_full && row0 + kItemH <= H;              // warp-uniform: no edge masking needed
            const size_t px0 = (size_t)row0 * W + col0;                     // this lane's first pixel within the tile

            // ---- sem: staged by TMA (FAST) or loaded directly -----------------------------------------
            long long sv[kPx];
#pragma unroll
            for (int p = 0; p < kPx; ++p) sv[p] = 0;
            if (kTma) {
                const int sub = it % kStageItems;           // strip within the staged box
                if (sub == 0) mbar_wait(&s_bar[warp][st_cons], parity);
                const unsigned char* sp = ring + (size_t)st_cons * kStageBytes + sub * kItemBytes;
#pragma unroll
                for (int i = 0; i < kItemH; ++i) {
                    if (SEM == SEM_I64) {
                        const longlong2 u = *reinterpret_cast<const longlong2*>(sp + i * kRowBytes + lane * 16);
                        sv[2 * i] = u.x; sv[2 * i + 1] = u.y;
                    } else {
                        const unsigned u = *reinterpret_cast<const unsigned short*>(sp + i * kRowBytes + lane * 2);
                        sv[2 * i] = u & 255u; sv[2 * i + 1] = u >> 8;
                    }
                }
                if (sub == kStageItems - 1 || it + 1 == kc.nitems) {        // last strip of the box: the slot is free
                    __syncwarp();                           // every lane has its values
                    --inflight;
                    if (++st_cons == kStages) { st_cons = 0; parity ^= 1u; }
                    if (p_which == 0 && p_i * kStageItems < kc.nitems) {    // common case: the next box of this block
                        if (lane == 0) {
                            mbar_expect_tx(&s_bar[warp][st_issue], kStageBytes);
                            tma_load_3d(ring + (size_t)st_issue * kStageBytes, &a.tmap, kc.colb, kc.row0 + p_i * (kStageItems * kItemH),
                                        kc.b, &s_bar[warp][st_issue], policy);
                        }
                        ++p_i; ++inflight;
                        st_issue = st_issue + 1 == kStages ? 0 : st_issue + 1;
                    } else {
                        pump();                             // block boundary: the general state machine
                    }
                }
            } else if (SEM != SEM_NONE) {
#pragma unroll
                for (int i = 0; i < kItemH; ++i) {
                    const bool rin = row0 + i < H && col0 < W;
                    const bool c1in = col0 + 1 < W;
                    if (SEM == SEM_I64) {
                        const long long* sp = reinterpret_cast<const long long*>(a.sem) + (size_t)b * a.sem_stride + px0 + (size_t)i * W;
                        if (rin) sv[2 * i] = __ldcs(sp);
                        if (rin && c1in) sv[2 * i + 1] = __ldcs(sp + 1);
                    } else {
                        const unsigned char* sp = reinterpret_cast<const unsigned char*>(a.sem) + (size_t)b * a.sem_stride + px0 + (size_t)i * W;
                        if (rin) sv[2 * i] = sp[0];
                        if (rin && c1in) sv[2 * i + 1] = sp[1];
                    }
                }
            }

            // ---- fast path: a full strip of class-0 background (the bulk of an EM tile) ----------------
            if (SEM != SEM_NONE && IDM != ID_DENSE && kCodes) {
                unsigned long long orall = 0ull;
#pragma unroll
                for (int p = 0; p < kPx; ++p) orall |= (unsigned long long)sv[p];
                if (full && class0_stuff && __all_sync(0xffffffffu, orall == 0ull)) {      // warp-uniform
                    // nothing to vote or count (class 0's area is taken by complement), and no codes:
                    // the strip flag tells apply_lut that every pixel is class-0 stuff
                    if (lane == 0) bflags[it] = 1;
                    continue;
                }
            }
            if (kCodes && lane == 0) bflags[it] = 0;

            // ---- general path ---------------------------------------------------------------------------
            unsigned inb = 0xFFu;       // bit p: pixel p = 2*i + j is inside the image
            if (!full) {
                inb = 0;
#pragma unroll
                for (int i = 0; i < kItemH; ++i) {
                    if (row0 + i < H && col0 < W) inb |= 1u << (2 * i);
                    if (row0 + i < H && col0 + 1 < W) inb |= 2u << (2 * i);
                }
            }
            unsigned w[kPx];            // info word: thing -> kInfoThing | thing index, stuff -> class, kInfoBad
            int idv[kPx];               // instance id (ID_DENSE / ID_COARSE) or argmin result
            unsigned thing = 0;         // bit p: pixel p takes an instance id
            unsigned bad = 0;           // bit p: class out of range
            unsigned cls0 = 0;          // bit p: class 0 and not a thing
#pragma unroll
            for (int p = 0; p < kPx; ++p) {
                w[p] = (SEM == SEM_NONE) ? kInfoThing : classify(sv[p], a.thing_bits, multi, a.things_small, a.things);
                thing |= ((w[p] >> 15) & 1u) << p;
                bad |= ((w[p] >> 14) & 1u) << p;
                cls0 |= (w[p] == 0u ? 1u : 0u) << p;
                idv[p] = 0;
            }
            thing &= inb; bad &= inb; cls0 &= inb;          // rows / columns outside the image count for nothing
            if (bad) flags |= EMP_FLAG_CLASS_RANGE;

            // instance ids supplied by the caller (merge entry points)
            unsigned idpos = 0;         // ID_DENSE: bit p: the pixel carries an instance id > 0
            if (IDM == ID_DENSE) {
#pragma unroll
                for (int i = 0; i < kItemH; ++i) {
                    const long long* ip = reinterpret_cast<const long long*>(a.ids_in) + (size_t)b * a.ids_stride + px0 + (size_t)i * W;
                    long long v0 = 0, v1 = 0;
                    if (FAST) {
                        if ((inb >> (2 * i)) & 1u) { const longlong2 u = __ldcs(reinterpret_cast<const longlong2*>(ip)); v0 = u.x; v1 = u.y; }
                    } else {
                        if ((inb >> (2 * i)) & 1u) v0 = __ldcs(ip);
                        if ((inb >> (2 * i)) & 2u) v1 = __ldcs(ip + 1);
                    }
                    if (v0 < 0 || v0 > a.max_id) { flags |= EMP_FLAG_ID_RANGE; v0 = 0; }
                    if (v1 < 0 || v1 > a.max_id) { flags |= EMP_FLAG_ID_RANGE; v1 = 0; }
                    idv[2 * i] = (int)v0; idv[2 * i + 1] = (int)v1;
                    idpos |= (v0 > 0 ? 1u : 0u) << (2 * i);
                    idpos |= (v1 > 0 ? 2u : 0u) << (2 * i);
                }
            } else if (IDM == ID_COARSE) {
#pragma unroll
                for (int i = 0; i < kItemH; ++i) {
                    const int* ip = reinterpret_cast<const int*>(a.ids_in) + (size_t)b * a.ids_stride;
                    const size_t crow = (size_t)((row0 + i) >> a.shift) * a.wc;
                    int v0 = 0, v1 = 0;
                    if ((inb >> (2 * i)) & 1u) v0 = __ldg(ip + crow + (col0 >> a.shift));
                    if ((inb >> (2 * i)) & 2u) v1 = __ldg(ip + crow + ((col0 + 1) >> a.shift));
                    if (v0 < 0 || v0 > a.max_id) { flags |= EMP_FLAG_ID_RANGE; v0 = 0; }
                    if (v1 < 0 || v1 > a.max_id) { flags |= EMP_FLAG_ID_RANGE; v1 = 0; }
                    idv[2 * i] = v0; idv[2 * i + 1] = v1;
                }
            }

            // nearest center over the cell index
            if (IDM == ID_ARGMIN && K > 0 && __any_sync(0xffffffffu, thing != 0)) {       // warp-uniform
                float ly[kPx], lx[kPx];
#pragma unroll
                for (int i = 0; i < kItemH; ++i) {
                    float2 fy = make_float2(0.f, 0.f), fx = make_float2(0.f, 0.f);
                    const float* oy = a.off + (size_t)b * a.off_stride + px0 + (size_t)i * W;
                    const float* ox = oy + HW;
                    const unsigned t2 = (thing >> (2 * i)) & 3u;
                    if (FAST) {
                        if (t2) { fy = __ldcs(reinterpret_cast<const float2*>(oy)); fx = __ldcs(reinterpret_cast<const float2*>(ox)); }
                    } else {
                        if (t2 & 1u) { fy.x = __ldcs(oy); fx.x = __ldcs(ox); }
                        if (t2 & 2u) { fy.y = __ldcs(oy + 1); fx.y = __ldcs(ox + 1); }
                    }
                    const float ycoord = __fmul_rn((float)(row0 + i), a.step);         // arange(0, H*step, step)
                    ly[2 * i] = __fadd_rn(ycoord, fy.x);
                    ly[2 * i + 1] = __fadd_rn(ycoord, fy.y);
                    lx[2 * i] = __fadd_rn(xc0, fx.x);
                    lx[2 * i + 1] = __fadd_rn(xc1, fx.y);
                }
                float bs[kPx];
                grid_nearest_8px(g, thing, ly, lx, bs, idv);
                // index -> 1-based id; "no index" (non-finite location) is id 1 without the sentinel, 0 with it
                // (values of non-thing pixels are never used)
                unsigned far = 0;
#pragma unroll
                for (int p = 0; p < kPx; ++p) {
                    idv[p] = idv[p] == INT_MAX ? (chunked ? 0 : 1) : idv[p] + 1;
                    far |= (bs[p] >= 9.99e9f ? 1u : 0u) << p;
                }
                if (chunked && (far & thing)) {         // sentinel (postprocess.py:98,:110): d < 1e5 on the rounded sqrt
#pragma unroll
                    for (int p = 0; p < kPx; ++p)
                        if (((far >> p) & 1u) && !(bs[p] < 1.001e10f && __fsqrt_rn(bs[p]) < 1e5f)) idv[p] = 0;
                }
            }

            unsigned voted = 0;         // bit p: a thing pixel with an instance id
#pragma unroll
            for (int p = 0; p < kPx; ++p) voted |= (idv[p] != 0 ? 1u : 0u) << p;
            voted &= thing;
            // stuff pixels (pasted by class if the class is large enough): in the image, not a thing,
            // class in range, and — with caller-supplied ids — not covered by an instance
            const unsigned st0 = inb & ~thing & ~bad & ~idpos;

            // votes (postprocess.py:263-273), stuff areas (:284-291)
            if (kCodes) {
                deficit_acc += (unsigned)__popc(inb & ~(st0 & cls0));       // class-0 stuff is counted by complement
                // votes: one warp-aggregated atomic per distinct (id, class) key
                unsigned vk[kPx];
#pragma unroll
                for (int p = 0; p < kPx; ++p)
                    vk[p] = ((voted >> p) & 1u) ? (unsigned)idv[p] * (unsigned)T + (w[p] & 15u) : kNoKey;
                for (;;) {
                    unsigned mine = kNoKey;
#pragma unroll
                    for (int p = kPx - 1; p >= 0; --p) mine = vk[p] != kNoKey ? vk[p] : mine;
                    const unsigned bal = __ballot_sync(0xffffffffu, mine != kNoKey);
                    if (!bal) break;                        // warp-uniform
                    const int leader = __ffs(bal) - 1;
                    const unsigned key = __shfl_sync(0xffffffffu, mine, leader);
                    int cnt = 0;
#pragma unroll
                    for (int p = 0; p < kPx; ++p) {
                        const bool m = vk[p] == key;
                        cnt += m ? 1 : 0;
                        vk[p] = m ? kNoKey : vk[p];
                    }
                    const int sum = __reduce_add_sync(0xffffffffu, cnt);
                    if (lane == leader) atomicAdd(votes + key, (uint32_t)sum);
                }
                // stuff areas of non-zero classes: accumulated per lane while the class stays the same
                const unsigned stuff = st0 & ~cls0;
                if (stuff) {
                    unsigned key = kNoKey, cnt = 0, rest = 0;
#pragma unroll
                    for (int p = kPx - 1; p >= 0; --p)
                        if ((stuff >> p) & 1u) key = w[p];
#pragma unroll
                    for (int p = 0; p < kPx; ++p) {
                        const bool isa = (stuff >> p) & 1u;
                        cnt += (isa && w[p] == key) ? 1u : 0u;
                        rest |= ((isa && w[p] != key) ? 1u : 0u) << p;
                    }
                    if (key != akey_acc) {
                        if (akey_acc != kNoKey && acnt_acc) count_slow(areas + akey_acc, acnt_acc);
                        akey_acc = key; acnt_acc = 0;
                    }
                    acnt_acc += cnt;
                    if (rest) {
#pragma unroll
                        for (int p = 0; p < kPx; ++p)
                            if ((rest >> p) & 1u) count_slow(areas + w[p], 1u);
                    }
                }
            }

            // codes / ids out
#pragma unroll
            for (int i = 0; i < kItemH; ++i) {
                if ((inb >> (2 * i)) & 1u) {
                    unsigned c[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int p = 2 * i + j;
                        if (kCodes) c[j] = ((voted >> p) & 1u) ? (unsigned)idv[p] : (((st0 >> p) & 1u) ? kClsBase + w[p] : 0u);
                        else c[j] = ((thing >> p) & 1u) ? (unsigned)idv[p] : 0u;
                    }
                    const bool c1in = (inb >> (2 * i)) & 2u;
                    const size_t o = (size_t)b * a.out_stride + px0 + (size_t)i * W;
                    if (OUT == OUT_CODE16) {
                        unsigned short* op = reinterpret_cast<unsigned short*>(a.out) + o;
                        if (FAST) *reinterpret_cast<unsigned*>(op) = c[0] | (c[1] << 16);
                        else { op[0] = (unsigned short)c[0]; if (c1in) op[1] = (unsigned short)c[1]; }
                    } else if (OUT == OUT_CODE32) {
                        unsigned* op = reinterpret_cast<unsigned*>(a.out) + o;
                        if (FAST) *reinterpret_cast<uint2*>(op) = make_uint2(c[0], c[1]);
                        else { op[0] = c[0]; if (c1in) op[1] = c[1]; }
                    } else if (OUT == OUT_IDS64) {
                        long long* op = reinterpret_cast<long long*>(a.out) + o;
                        if (FAST) __stcs(reinterpret_cast<longlong2*>(op), make_longlong2((long long)c[0], (long long)c[1]));
                        else { op[0] = (long long)c[0]; if (c1in) op[1] = (long long)c[1]; }
                    } else {
                        int* op = reinterpret_cast<int*>(a.out) + o;
                        if (FAST) *reinterpret_cast<int2*>(op) = make_int2((int)c[0], (int)c[1]);
                        else { op[0] = (int)c[0]; if (c1in) op[1] = (int)c[1]; }
                    }
                }
            }
        }

        // next block; the prefetcher is already inside it (or beyond it)
        kc = kn;
        nxt = kc.nitems > 0 ? grab() : n_blocks;
        kn = decode(nxt);
        if (kTma) {
            if (p_which == 0) p_i = 0; else --p_which;
            pump();
        }
    }
    flush_image();
}

// label LUT: one CTA per tile
__global__ void __launch_bounds__(256)
build_lut_kernel(char* ws_base, size_t ws_stride, size_t o_status, size_t o_votes, size_t o_lut, size_t o_areas, int k_cap,
                 int k_fixed, const int32_t* k_dev, size_t k_dev_stride, const __grid_constant__ Things things,
                 long long label_divisor, long long void_label, long long stuff_area, long long n_px)
{
    if (k_dev) k_dev += (size_t)blockIdx.z * k_dev_stride;         // per-tile center count (int32 words apart)
    char* ws = ws_base + (size_t)blockIdx.z * ws_stride;
    // stuff classes (postprocess.py:287-294): class c keeps its label c * L if it covers at least stuff_area pixels, else
    // void; class 0's area is the complement of everything assign counted.  Entries of thing classes are never looked up.
    {
        const uint32_t* areas = reinterpret_cast<const uint32_t*>(ws + o_areas);
        long long* cls = reinterpret_cast<long long*>(ws + o_lut) + (k_cap + 1);
        static_assert(kNumClasses % 1024 == 0, "class table is read as uint4 by 256 threads");
        const long long rest = (long long)__ldcg(areas + kNumClasses);
#pragma unroll
        for (int r = 0; r < kNumClasses / 1024; ++r) {              // all loads of a thread are in flight together
            const int c0 = (r * 256 + threadIdx.x) * 4;
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(areas) + r * 256 + threadIdx.x);
            const uint32_t ar[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                const long long area = c ? (long long)ar[j] : n_px - rest;
                cls[c] = area >= stuff_area ? (long long)c * label_divisor : void_label;
            }
        }
    }
    __shared__ LutScratch sc;
    __shared__ long long s_k;
    if (threadIdx.x == 0) s_k = lut_extent(k_fixed, k_cap, reinterpret_cast<const int32_t*>(ws + o_status), k_dev);
    __syncthreads();
    build_label_lut(s_k, reinterpret_cast<const uint32_t*>(ws + o_votes), things, label_divisor, void_label,
                    reinterpret_cast<long long*>(ws + o_lut), sc);
}

// ---------------------------------------------------------------------------------------------
// K5  apply_lut — code map -> int64 panoptic labels (postprocess.py:281, :287-294).
//   code 0 -> void; 1..CLS_BASE-1 -> lut[id]; CLS_BASE + c -> c*L if area[c] >= stuff_area else void
// (a thing-class pixel never carries a class code, so no thing test is needed here).
// Same geometry as assign: a warp takes a 64 x 64 block and walks its 16 strips of 4 rows x 64
// columns, lane l owning columns 2l, 2l+1 (128 B of codes in, 512 B of labels out per row).  The
// block's 16 strip flags arrive in one 16-byte load; a flagged strip (all class-0 stuff, the bulk of
// an EM tile) has no codes to read: the warp streams out the one label.
// ---------------------------------------------------------------------------------------------
struct ApplyArgs {
    char* ws; size_t ws_stride;
    size_t o_codes, o_lut, o_sflags;
    long long* pan; size_t n_px;
    int H, W, blocks_x, blocks_y, blk_items;
    unsigned cls_off;                                               // lut[cls_off + c]: label of stuff class c
};

// one table look-up per pixel: build_lut resolved ids AND stuff classes (area test included) into labels
template <bool C16>
__device__ __forceinline__ long long decode(unsigned code, const long long* __restrict__ lut, unsigned cls_off)
{
    constexpr uint32_t base = C16 ? kClsBase16 : kClsBase32;
    return __ldg(lut + (code >= base ? code - base + cls_off : code));
}

// General kernel (32-bit codes, unaligned planes, odd widths): one 64 x 64 block per warp; with FAST the code loads
// of up to kBatch unflagged strips are issued together from registers.
template <bool C16, bool FAST>
__global__ void __launch_bounds__(256)
apply_lut_kernel(const __grid_constant__ ApplyArgs a)
{
    constexpr uint32_t kClsBase = C16 ? kClsBase16 : kClsBase32;
    constexpr int kBatch = 4;                         // unflagged strips whose code loads are issued together
    char* ws = a.ws + (size_t)blockIdx.z * a.ws_stride;
    const long long* lut = reinterpret_cast<const long long*>(ws + a.o_lut);
    const uint4* flags = reinterpret_cast<const uint4*>(ws + a.o_sflags);
    long long* pan = a.pan + (size_t)blockIdx.z * a.n_px;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, W = a.W;
    const int blk = (int)blockIdx.x * 8 + warp;
    if (blk >= a.blocks_x * a.blocks_y) return;                     // warp-uniform
    const uint4 fl = __ldg(flags + blk);
    const long long bg = decode<C16>(kClsBase, lut, a.cls_off);     // label of class-0 stuff, once per warp
    {                                                               // the warp's block
        const int by = blk / a.blocks_x, bx = blk - by * a.blocks_x;
        const int colb = bx * kItemW, rowb = by * (a.blk_items * kItemH);
        const int col0 = colb + 2 * lane;
        const int nitems = min(a.blk_items, (H - rowb + kItemH - 1) / kItemH);
        const bool cols_full = colb + kItemW <= W;
        unsigned fmask = 0;                                         // bit it: strip `it` is flagged (all class-0 stuff)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned wq = q == 0 ? fl.x : q == 1 ? fl.y : q == 2 ? fl.z : fl.w;
#pragma unroll
            for (int k = 0; k < 4; ++k) fmask |= (((wq >> (8 * k)) & 0xFFu) != 0u ? 1u : 0u) << (4 * q + k);
        }
        const unsigned live = (1u << nitems) - 1u;                  // nitems <= 16
        fmask &= live;

        if (FAST) {
            // unflagged full strips: issue the code loads of up to kBatch strips together ...
            unsigned todo = live & ~fmask;
            if (!cols_full) todo = 0;
            else if (rowb + nitems * kItemH > H) todo &= ~(1u << (nitems - 1));      // a bottom strip cut by the image edge
            const unsigned edge = (live & ~fmask) & ~todo;          // partial strips: scalar path below
            // ... but first the flagged strips (always full strips): pure stores, nothing to wait for
            for (unsigned m = fmask; m; m &= m - 1) {
                const size_t px0 = (size_t)(rowb + (__ffs(m) - 1) * kItemH) * W + col0;
#pragma unroll
                for (int i = 0; i < kItemH; ++i) __stcs(reinterpret_cast<longlong2*>(pan + px0 + (size_t)i * W), make_longlong2(bg, bg));
            }
            while (todo) {
                int its[kBatch];
                unsigned cw[kBatch][kItemH], ch[kBatch][kItemH];
#pragma unroll
                for (int s2 = 0; s2 < kBatch; ++s2) {
                    its[s2] = todo ? __ffs(todo) - 1 : -1;
                    todo &= todo - 1;                               // (0 & -1 stays 0)
#pragma unroll
                    for (int i = 0; i < kItemH; ++i) {
                        cw[s2][i] = 0; ch[s2][i] = 0;
                        if (its[s2] >= 0) {
                            const size_t e = (size_t)(rowb + its[s2] * kItemH + i) * W + col0;
                            if (C16) {
                                cw[s2][i] = __ldcs(reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned short*>(ws + a.o_codes) + e));
                            } else {
                                const uint2 u = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned*>(ws + a.o_codes) + e));
                                cw[s2][i] = u.x; ch[s2][i] = u.y;
                            }
                        }
                    }
                }
#pragma unroll
                for (int s2 = 0; s2 < kBatch; ++s2) {
                    if (its[s2] < 0) continue;
#pragma unroll
                    for (int i = 0; i < kItemH; ++i) {
                        const size_t e = (size_t)(rowb + its[s2] * kItemH + i) * W + col0;
                        const unsigned c0 = C16 ? (cw[s2][i] & 0xFFFFu) : cw[s2][i];
                        const unsigned c1 = C16 ? (cw[s2][i] >> 16) : ch[s2][i];
                        __stcs(reinterpret_cast<longlong2*>(pan + e), make_longlong2(decode<C16>(c0, lut, a.cls_off), decode<C16>(c1, lut, a.cls_off)));
                    }
                }
            }
            fmask = ~edge;                                          // what is left for the scalar loop: the partial strips
        }

        for (int it = 0; it < nitems; ++it) {
            const bool flagged = (fmask >> it) & 1u;
            if (FAST && flagged) continue;                          // done above
            const int row0 = rowb + it * kItemH;
            const size_t px0 = (size_t)row0 * W + col0;
#pragma unroll
            for (int i = 0; i < kItemH; ++i) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (row0 + i < H && col0 + j < W) {
                        const size_t e = px0 + (size_t)i * W + j;
                        const unsigned code = (!FAST && flagged) ? kClsBase
                                            : C16 ? (unsigned)reinterpret_cast<const unsigned short*>(ws + a.o_codes)[e]
                                                  : reinterpret_cast<const unsigned*>(ws + a.o_codes)[e];
                        pan[e] = decode<C16>(code, lut, a.cls_off);
                    }
                }
            }
        }
    }
}

// apply_lut for the common case (16-bit codes, aligned planes): the same block walk, but ALL code loads of a
// block are started up front as 4-byte cp.async copies into a per-warp staging area in shared memory (no
// registers held, up to 64 in flight per lane), the flagged strips are streamed out while they fly, and the
// unflagged strips are decoded from shared memory after one wait.  The register-batched kernel above needs a
// DRAM round trip per kBatch strips, under a write stream that makes each one long.  Each 64 x 64 block is shared
// by kApplySplit warps (a pure store stream runs best with little work per warp: profiles/store_patterns.cu).
// Measured on config 2 (ms per batch): register-batched 0.384, staged split 1 / 2 / 4: 0.3765 / 0.3785 / 0.3749; an
// L2 evict-first hint on the code copies changes nothing.  With every strip forced "flagged" the kernel takes 0.321
// and a bare store kernel of this geometry 0.304: what is left above that comes with the code reads of thing strips.
constexpr int kApplySplit = 4;                                      // warps sharing one block (each takes 16 / split strips)
constexpr int kApplyPer = kBlkItems / kApplySplit;                  // strips per warp
__global__ void __launch_bounds__(128)
apply_lut_staged_kernel(const __grid_constant__ ApplyArgs a)
{
    constexpr uint32_t kClsBase = kClsBase16;
    __shared__ unsigned stage[4][kApplyPer * kItemH][32];          // per warp: its strips' codes (two per word)
    char* ws = a.ws + (size_t)blockIdx.z * a.ws_stride;
    const long long* lut = reinterpret_cast<const long long*>(ws + a.o_lut);
    const unsigned short* codes = reinterpret_cast<const unsigned short*>(ws + a.o_codes);
    long long* pan = a.pan + (size_t)blockIdx.z * a.n_px;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, W = a.W;
    const int blk = (int)blockIdx.x * (4 / kApplySplit) + warp / kApplySplit;
    const int part = warp % kApplySplit;
    if (blk >= a.blocks_x * a.blocks_y) return;                     // warp-uniform
    const uint4 fl = __ldg(reinterpret_cast<const uint4*>(ws + a.o_sflags) + blk);
    const int by = blk / a.blocks_x, bx = blk - by * a.blocks_x;
    const int colb = bx * kItemW, rowb = by * (a.blk_items * kItemH);
    const int col0 = colb + 2 * lane;
    const int nitems = min(a.blk_items, (H - rowb + kItemH - 1) / kItemH);
    unsigned fmask = 0;                                             // bit it: strip `it` is flagged (all class-0 stuff)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned wq = q == 0 ? fl.x : q == 1 ? fl.y : q == 2 ? fl.z : fl.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) fmask |= (((wq >> (8 * k)) & 0xFFu) != 0u ? 1u : 0u) << (4 * q + k);
    }
    const unsigned mine = ((1u << kApplyPer) - 1u) << (part * kApplyPer);        // this warp's strips of the block
    const unsigned live = ((1u << nitems) - 1u) & mine;             // nitems <= 16
    if (!live) return;
    fmask &= live;
    unsigned todo = live & ~fmask;                                  // unflagged full strips
    if (colb + kItemW > W) todo = 0;
    else if (rowb + nitems * kItemH > H) todo &= ~(1u << (nitems - 1));          // a bottom strip cut by the image edge
    const unsigned edge = (live & ~fmask) & ~todo;                  // partial strips: scalar path at the end

    for (unsigned m = todo; m; m &= m - 1) {                        // 1. start every code load of the block
        const int it = __ffs(m) - 1;
#pragma unroll
        for (int i = 0; i < kItemH; ++i) {
            const unsigned short* src = codes + (size_t)(rowb + it * kItemH + i) * W + col0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&stage[warp][(it - part * kApplyPer) * kItemH + i][lane])), "l"(src) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const long long bg = decode<true>(kClsBase, lut, a.cls_off);    // label of class-0 stuff
    for (unsigned m = fmask; m; m &= m - 1) {                       // 2. flagged strips: pure stores
        const size_t px0 = (size_t)(rowb + (__ffs(m) - 1) * kItemH) * W + col0;
#pragma unroll
        for (int i = 0; i < kItemH; ++i) __stcs(reinterpret_cast<longlong2*>(pan + px0 + (size_t)i * W), make_longlong2(bg, bg));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");            // 3. every lane reads back only what it copied itself
    for (unsigned m = todo; m; m &= m - 1) {
        const int it = __ffs(m) - 1;
        const size_t px0 = (size_t)(rowb + it * kItemH) * W + col0;
        unsigned w[kItemH];
#pragma unroll
        for (int i = 0; i < kItemH; ++i) w[i] = stage[warp][(it - part * kApplyPer) * kItemH + i][lane];
#pragma unroll
        for (int i = 0; i < kItemH; ++i)
            __stcs(reinterpret_cast<longlong2*>(pan + px0 + (size_t)i * W),
                   make_longlong2(decode<true>(w[i] & 0xFFFFu, lut, a.cls_off), decode<true>(w[i] >> 16, lut, a.cls_off)));
    }
    for (unsigned m = edge; m; m &= m - 1) {                        // 4. strips cut by the right / bottom image edge
        const int row0 = rowb + (__ffs(m) - 1) * kItemH;
#pragma unroll
        for (int i = 0; i < kItemH; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (row0 + i < H && col0 + j < W) {
                    const size_t e = (size_t)(row0 + i) * W + col0 + j;
                    pan[e] = decode<true>(codes[e], lut, a.cls_off);
                }
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------------------------
static int sm_count()
{
    static std::atomic<int> cached[64];         // per device ordinal; 0 = not asked yet
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && (v = cached[dev].load(std::memory_order_relaxed)) > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    if (dev >= 0 && dev < 64) cached[dev].store(v, std::memory_order_relaxed);
    return v;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_ws(const void* ws, size_t ws_bytes, size_t need)
{
    EMP_REQUIRE(ws != nullptr, EMP_ERR_WORKSPACE, "workspace is null");
    EMP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, EMP_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    EMP_REQUIRE(ws_bytes >= need, EMP_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, need);
    return EMP_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (B, H, W) tiled tensor map over B planes `plane_stride` elements apart, box 1 x box_h x box_w
static int make_plane_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, size_t elt, const void* base, int B, int H, int W,
                                 size_t plane_stride, int box_h, int box_w)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        EMP_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        EMP_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, EMP_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)W * elt, (cuuint64_t)plane_stride * elt};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(out, dtype, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EMP_REQUIRE(r == CUDA_SUCCESS, EMP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EMP_OK;
}

int launch_centers(int B, const float* hm, int H, int W, float thr, int k, float step, const WsLayout& L,
                   char* ws, size_t ws_stride, int k_cap, int64_t* ctr_out, int cap, cudaStream_t st, size_t hm_stride = 0)
{
    NmsArgs n;
    memset(&n, 0, sizeof(n));
    n.hm = hm; n.hm_stride = hm_stride ? hm_stride : (size_t)H * W;
    n.ws = ws; n.ws_stride = ws_stride; n.o_mask = L.mask; n.o_rowcnt = L.rowcnt;
    n.B = B; n.H = H; n.W = W; n.wd = L.wd; n.lo = k / 2; n.hi = k - 1 - n.lo; n.thr = thr;
    n.blocks_x = (W + kNmsItemW - 1) / kNmsItemW;
    n.blk_items = kNmsBlkItems;                 // shorter strips for small batches: at least ~4 strips per resident warp, so the
                                                // warps that draw one strip more than the others do not set the launch's time
    const long long want_strips = 4ll * sm_count() * kNmsCtasPerSm * kNmsWarps;
    while (n.blk_items > 1 && (long long)((H + n.blk_items * kNmsRows - 1) / (n.blk_items * kNmsRows)) * n.blocks_x * B < want_strips) n.blk_items >>= 1;
    n.blocks_y = (H + n.blk_items * kNmsRows - 1) / (n.blk_items * kNmsRows);
    const long long n_blocks = (long long)n.blocks_x * n.blocks_y * B;
    EMP_REQUIRE(n_blocks < (1ll << 30), EMP_ERR_INVALID, "batch too large for one launch");
    // TMA staging needs a 16-byte aligned base and row pitch, and the box must fit the tensor
    const bool fast = aligned16(hm) && W % 4 == 0 && W >= kNmsItemW && H >= kNmsBoxRows && (n.hm_stride % 4 == 0);
    if (fast) {
        const int rc = make_plane_tensor_map(&n.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, hm, B, H, W, n.hm_stride, kNmsBoxRows, kNmsItemW);
        if (rc) return rc;
    }
    long long blocks = (n_blocks + kNmsWarps - 1) / kNmsWarps;
    const long long resident = (long long)sm_count() * kNmsCtasPerSm;
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    {
        ProfScope ps(ST_NMS, st);
        if (fast) {
            const int ring_bytes = kNmsWarps * kNmsStages * (int)kNmsStageBytes;
            EMP_CUDA_CHECK(cudaFuncSetAttribute(nms_peaks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes));
            EMP_CUDA_CHECK(cudaFuncSetAttribute(nms_peaks_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            nms_peaks_kernel<true><<<(unsigned)blocks, kNmsWarps * 32, ring_bytes, st>>>(n);
        } else {
            nms_peaks_kernel<false><<<(unsigned)blocks, kNmsWarps * 32, 0, st>>>(n);
        }
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    dim3 g2((H + 31) / 32, 1, B);
    {
        ProfScope ps(ST_EMIT, st);
        emit_centers_kernel<<<g2, 256, 0, st>>>(ws, ws_stride, L.mask, L.rowcnt, L.centers, L.ctr_i, L.status, H, L.wd, k_cap,
                                                step, ctr_out, (size_t)cap * 2, cap);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

// Strips per block for an H x W tile: as tall as possible (fewer hand-outs, longer prefetch runs)
// while a tile still yields a few thousand blocks, so that small images (the 512 x 512 coarse maps of
// the stack path) keep every warp of the persistent grid busy.  assign and apply_lut must agree.
static int block_items(int H, int W)
{
    const long long strips_x = (W + kItemW - 1) / kItemW, strips_y = (H + kItemH - 1) / kItemH;
    int b = kBlkItems;
    while (b > 1 && ((strips_y + b - 1) / b) * strips_x < 4096) b >>= 1;
    return b;
}

template <int SEM, int IDM, int OUT, bool FAST>
static int launch_assign_f(const AssignArgs& a, cudaStream_t st)
{
    constexpr int row_bytes = (SEM == SEM_I64) ? kItemW * 8 : kItemW;
    constexpr int ring_bytes = kAssignWarps * kStages * kStageItems * kItemH * row_bytes;
    const size_t smem = (FAST && SEM != SEM_NONE) ? ring_bytes : 0;
    EMP_CUDA_CHECK(cudaFuncSetAttribute(assign_kernel<SEM, IDM, OUT, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes));
    const long long n_blocks = (long long)a.blocks_x * a.blocks_y * a.B;
    long long blocks = (n_blocks + kAssignWarps - 1) / kAssignWarps;
    const long long resident = (long long)sm_count() * kAssignCtasPerSm;
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    ProfScope ps(ST_ASSIGN, st);
    assign_kernel<SEM, IDM, OUT, FAST><<<(unsigned)blocks, kAssignThreads, smem, st>>>(a);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

// (B, H, W) tensor map over the sem planes for the assign kernel's TMA ring: box 1 x 4 x 64
static int make_sem_tensor_map(AssignArgs& a, int sem_mode)
{
    return make_plane_tensor_map(&a.tmap, sem_mode == SEM_I64 ? CU_TENSOR_MAP_DATA_TYPE_INT64 : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                                 sem_mode == SEM_I64 ? 8 : 1, a.sem, a.B, a.H, a.W, a.sem_stride, kStageItems * kItemH, kItemW);
}

template <int SEM, int IDM, int OUT>
static int launch_assign_t(const AssignArgs& a, cudaStream_t st)
{
    return a.fast ? launch_assign_f<SEM, IDM, OUT, true>(a, st) : launch_assign_f<SEM, IDM, OUT, false>(a, st);
}

template <int SEM, int IDM>
static int launch_assign_out(int out_mode, const AssignArgs& a, cudaStream_t st)
{
    switch (out_mode) {
        case OUT_CODE16: return launch_assign_t<SEM, IDM, OUT_CODE16>(a, st);
        case OUT_CODE32: return launch_assign_t<SEM, IDM, OUT_CODE32>(a, st);
        case OUT_IDS64:  return launch_assign_t<SEM, IDM, OUT_IDS64>(a, st);
        default:         return launch_assign_t<SEM, IDM, OUT_IDS32>(a, st);
    }
}

// a.B tiles in one launch; a.sem / a.off / a.ids_in / a.out / a.ws point at the first of them
int launch_assign(int sem_mode, int id_mode, int out_mode, AssignArgs& a, cudaStream_t st)
{
    a.blk_items = block_items(a.H, a.W);
    a.blocks_x = (a.W + kItemW - 1) / kItemW;
    a.blocks_y = (a.H + a.blk_items * kItemH - 1) / (a.blk_items * kItemH);
    EMP_REQUIRE((long long)a.blocks_x * a.blocks_y * a.B < (1ll << 30), EMP_ERR_INVALID, "batch too large for one launch");
    a.counter = reinterpret_cast<int32_t*>(a.ws + a.o_status) + EMP_ST_TICKET;      // zeroed with the status block
    // FAST: every plane moves as 8/16-byte vectors and the sem plane is staged by TMA tensor copies
    // (16-byte aligned base and row pitch); otherwise scalar accesses throughout.
    a.fast = a.vec;
    if (sem_mode == SEM_I64) a.fast = a.vec && aligned16(a.sem) && (a.sem_stride % 2 == 0);
    else if (sem_mode == SEM_U8) a.fast = a.vec && (a.W % 16 == 0) && aligned16(a.sem) && (a.sem_stride % 16 == 0);
    if (sem_mode != SEM_NONE && (a.W < kItemW || a.H < kStageItems * kItemH)) a.fast = 0;       // the TMA box must fit the tensor
    if (a.fast && sem_mode != SEM_NONE) {
        const int rc = make_sem_tensor_map(a, sem_mode);
        if (rc) return rc;
    }
    if (id_mode == ID_ARGMIN) {
        if (sem_mode == SEM_NONE) return launch_assign_out<SEM_NONE, ID_ARGMIN>(out_mode, a, st);
        if (sem_mode == SEM_I64) return launch_assign_out<SEM_I64, ID_ARGMIN>(out_mode, a, st);
        return launch_assign_out<SEM_U8, ID_ARGMIN>(out_mode, a, st);
    }
    if (id_mode == ID_DENSE) {
        EMP_REQUIRE(sem_mode == SEM_I64, EMP_ERR_INVALID, "dense-id merge needs int64 sem");
        return launch_assign_out<SEM_I64, ID_DENSE>(out_mode, a, st);
    }
    if (sem_mode == SEM_I64) return launch_assign_out<SEM_I64, ID_COARSE>(out_mode, a, st);
    EMP_REQUIRE(sem_mode == SEM_U8, EMP_ERR_INVALID, "coarse-id merge needs int64 or uint8 sem");
    return launch_assign_out<SEM_U8, ID_COARSE>(out_mode, a, st);
}

void fill_assign_common(AssignArgs& a, const WsLayout& L, const Things& th)
{
    a.o_status = L.status; a.o_votes = L.votes; a.o_areas = L.areas;
    a.o_cell_start = L.cell_start; a.o_sorted = L.sorted; a.o_sflags = L.sflags;
    a.things = th;
    a.thing_bits = 0ull;
    a.things_small = 1;
    for (int i = 0; i < th.n; ++i) {
        if (th.v[i] >= 0 && th.v[i] < 64) a.thing_bits |= 1ull << th.v[i];
        else a.things_small = 0;
    }
}

int launch_apply(int B, const WsLayout& L, char* ws, size_t ws_stride, int64_t* pan_out, int H, int W, cudaStream_t st)
{
    ApplyArgs a;
    memset(&a, 0, sizeof(a));
    a.ws = ws; a.ws_stride = ws_stride;
    a.o_codes = L.codes; a.o_lut = L.lut; a.o_sflags = L.sflags; a.cls_off = (unsigned)L.cls_off;
    a.pan = reinterpret_cast<long long*>(pan_out); a.n_px = (size_t)H * W;
    a.H = H; a.W = W;
    a.blk_items = block_items(H, W);
    a.blocks_x = (W + kItemW - 1) / kItemW;
    a.blocks_y = (H + a.blk_items * kItemH - 1) / (a.blk_items * kItemH);
    const bool fast = aligned16(pan_out) && W % 4 == 0;             // 16-byte label stores, 4-byte code loads
    const long long n_blocks = (long long)a.blocks_x * a.blocks_y;
    ProfScope ps(ST_APPLY, st);
    // Non-persistent grids on purpose: a store stream runs best with little work per warp (profiles/README.md:
    // grids of 3 ... 12 resident CTAs per SM walking several blocks each were 10 - 25 % slower, a row-linear
    // variant slower on dense tiles).
    if (L.code16 && fast) {
        constexpr int per_cta = 4 / kApplySplit;                    // blocks per 4-warp CTA
        apply_lut_staged_kernel<<<dim3((unsigned)((n_blocks + per_cta - 1) / per_cta), 1, B), 128, 0, st>>>(a);
    } else {
        const dim3 grid((unsigned)((n_blocks + 7) / 8), 1, B);
        if (L.code16) apply_lut_kernel<true, false><<<grid, 256, 0, st>>>(a);
        else if (fast) apply_lut_kernel<false, true><<<grid, 256, 0, st>>>(a);
        else apply_lut_kernel<false, false><<<grid, 256, 0, st>>>(a);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

int launch_build_lut(int B, const WsLayout& L, char* ws, size_t ws_stride, int k_cap, int k_fixed, const int32_t* k_dev,
                     const Things& th, long long label_divisor, long long void_label, long long stuff_area, int H, int W,
                     cudaStream_t st, size_t k_dev_stride = 0)
{
    ProfScope ps(ST_LUT, st);
    build_lut_kernel<<<dim3(1, 1, B), 256, 0, st>>>(ws, ws_stride, L.status, L.votes, L.lut, L.areas, k_cap, k_fixed, k_dev,
                                                    k_dev_stride, th, label_divisor, void_label, stuff_area, (long long)H * W);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

// cell index over the centers of B tiles (k_fixed < 0: K comes from the status block)
int launch_bin(int B, const WsLayout& L, char* ws, size_t ws_stride, int H, int W, int k_cap, int k_fixed, float step,
               cudaStream_t st)
{
    ProfScope ps(ST_BIN, st);
    bin_centers_kernel<<<dim3(1, 1, B), 1024, 0, st>>>(ws, ws_stride, L.status, L.centers, L.ctr_i, L.cell_start,
                                                       L.cell_fill, L.sorted, H, W, k_cap, k_fixed, step);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

int load_centers(const int64_t* ctr, int K, float step, const WsLayout& L, char* ws, cudaStream_t st)
{
    if (K > 0) {
        load_centers_kernel<<<(K + 255) / 256, 256, 0, st>>>(ctr, K, step, reinterpret_cast<float2*>(ws + L.centers),
                                                             reinterpret_cast<int2*>(ws + L.ctr_i));
        EMP_CUDA_CHECK(cudaGetLastError());
    }
    return EMP_OK;
}

// Tiles per assign -> build_lut -> apply launch group of the fused path.  Measured on B200
// (profiles/): one group holding the whole batch is fastest — the persistent assign kernel wants
// many blocks per warp to balance, and strip flags already keep most of the code map out of DRAM.
// EMP_TILE_GROUP=g (tuning knob) forces groups of g tiles.
static int tile_group(int B)
{
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("EMP_TILE_GROUP");
        env = e ? atoi(e) : 0;
    }
    return env > 0 ? env : B;
}

// ---- batched building blocks of the stack path (stack_block.cu) --------------------------------
int assign_block_items(int H, int W) { return block_items(H, W); }

int device_sm_count() { return sm_count(); }

// B coarse maps -> nearest-center ids (engines.py:257-272 without the upsample): centers, cell index and the exact
// argmin, one launch per kernel for the whole batch.  ws: B workspaces of ws_layout(h, w, k_cap, 1), ws_stride apart.
int coarse_ids_batched(int B, const float* hm, size_t hm_stride, const float* off, size_t off_stride, int h, int w,
                       float threshold, int nms_kernel, float step, int32_t* ids_out, size_t ids_stride, int k_cap,
                       char* ws, size_t ws_stride, cudaStream_t st, const unsigned char* need, size_t need_stride)
{
    const WsLayout L = ws_layout(h, w, k_cap, 1);
    EMP_REQUIRE(ws_stride >= L.total && ws_stride % 256 == 0, EMP_ERR_WORKSPACE, "coarse workspace stride too small");
    {
        ProfScope ps(ST_MEMSET, st);
        EMP_CUDA_CHECK(cudaMemset2DAsync(ws, ws_stride, 0, L.zero_bytes, (size_t)B, st));
    }
    int rc;
    if ((rc = launch_centers(B, hm, h, w, threshold, nms_kernel, step, L, ws, ws_stride, k_cap, nullptr, 0, st, hm_stride))) return rc;
    if ((rc = launch_bin(B, L, ws, ws_stride, h, w, k_cap, -1, step, st))) return rc;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.off = off; a.off_stride = off_stride;
    a.out = ids_out; a.out_stride = ids_stride;
    a.ws = ws; a.ws_stride = ws_stride;
    a.B = B; a.H = h; a.W = w; a.step = step; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
    a.vec = (w % 4 == 0) && aligned16(off) && aligned16(ids_out) && (off_stride % 4 == 0) && (ids_stride % 4 == 0);
    if (need) {
        // only cells that hold a thing pixel need their nearest center (everywhere else the id is multiplied by 0,
        // engines.py:283-285): the need map plays the class map of get_instance_segmentation with thing class 1
        Things one; memset(&one, 0, sizeof(one)); one.n = 1; one.v[0] = 1;
        fill_assign_common(a, L, one);
        a.sem = need; a.sem_stride = need_stride;
        return launch_assign(SEM_U8, ID_ARGMIN, OUT_IDS32, a, st);
    }
    { Things none; memset(&none, 0, sizeof(none)); fill_assign_common(a, L, none); }
    return launch_assign(SEM_NONE, ID_ARGMIN, OUT_IDS32, a, st);
}

// B hardened class maps (uint8) + their coarse id maps -> code maps, strip flags, votes and label LUTs
// (postprocess.py:253-294 up to, but without, the write of the int64 label map).  ws: B workspaces of
// ws_layout(H, W, k_cap, n_things); k_dev: the slices' center counts, k_dev_stride int32 words apart.
int merge_codes_batched(int B, const unsigned char* sem8, size_t sem_stride, const int32_t* ids, size_t ids_stride, int hc,
                        int wc, int shift, int H, int W, const Things& th, long long label_divisor, long long stuff_area,
                        long long void_label, int k_cap, const int32_t* k_dev, size_t k_dev_stride, char* ws,
                        size_t ws_stride, cudaStream_t st)
{
    const WsLayout L = ws_layout(H, W, k_cap, th.n);
    EMP_REQUIRE(L.code16, EMP_ERR_INVALID, "the stack block needs 16-bit codes (k_cap < %u)", kClsBase16);
    EMP_REQUIRE(ws_stride >= L.total && ws_stride % 256 == 0, EMP_ERR_WORKSPACE, "merge workspace stride too small");
    {
        ProfScope ps(ST_MEMSET, st);
        EMP_CUDA_CHECK(cudaMemset2DAsync(ws, ws_stride, 0, L.zero_bytes, (size_t)B, st));
    }
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.sem = sem8; a.sem_stride = sem_stride;
    a.ids_in = ids; a.ids_stride = ids_stride; a.wc = wc; a.shift = shift;
    a.out = ws + L.codes; a.out_stride = ws_stride / 2;
    a.ws = ws; a.ws_stride = ws_stride;
    fill_assign_common(a, L, th);
    a.B = B; a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = k_cap;
    a.max_id = k_cap;
    a.vec = (W % 4 == 0) && aligned16(sem8);
    int rc;
    if ((rc = launch_assign(SEM_U8, ID_COARSE, OUT_CODE16, a, st))) return rc;
    return launch_build_lut(B, L, ws, ws_stride, k_cap, k_cap, k_dev, th, label_divisor, void_label, stuff_area, H, W, st, k_dev_stride);
}

// label LUTs of B tiles from their votes / areas (build_lut_kernel), for merge kernels defined elsewhere
int build_luts_batched(int B, int H, int W, const Things& th, long long label_divisor, long long stuff_area, long long void_label,
                       int k_cap, const int32_t* k_dev, size_t k_dev_stride, char* ws, size_t ws_stride, cudaStream_t st)
{
    const WsLayout L = ws_layout(H, W, k_cap, th.n);
    return launch_build_lut(B, L, ws, ws_stride, k_cap, k_cap, k_dev, th, label_divisor, void_label, stuff_area, H, W, st, k_dev_stride);
}

}  // namespace emp

// =============================================================================================
// C ABI
// =============================================================================================
using namespace emp;

EMP_API int emp_version(void) { return 100; }

EMP_API int emp_profile_enable(int on)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_on.store(on != 0);
    g_prof_used = 0;
    return EMP_OK;
}

EMP_API int emp_profile_read(double* ms_per_stage, int* launches_per_stage)
{
    for (int i = 0; i < ST_COUNT; ++i) { ms_per_stage[i] = 0.0; launches_per_stage[i] = 0; }   // ST_COUNT == EMP_PROFILE_STAGES
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (size_t i = 0; i < g_prof_used; ++i) {
        EMP_CUDA_CHECK(cudaEventSynchronize(g_prof[i].b));
        float ms = 0.f;
        EMP_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b));
        ms_per_stage[g_prof[i].stage] += ms;
        launches_per_stage[g_prof[i].stage] += 1;
    }
    g_prof_used = 0;
    return EMP_OK;
}
EMP_API const char* emp_last_error(void) { return g_err; }

EMP_API size_t emp_workspace_bytes(int H, int W, int k_cap, int n_things)
{
    if (H <= 0 || W <= 0 || k_cap < 0) return 0;
    return ws_layout(H, W, k_cap, n_things).total;
}

static int check_image(int H, int W)
{
    EMP_REQUIRE(H > 0 && W > 0, EMP_ERR_INVALID, "bad image size %d x %d", H, W);
    EMP_REQUIRE((long long)H * W < (1ll << 31), EMP_ERR_INVALID, "image too large (%d x %d)", H, W);
    return EMP_OK;
}

EMP_API int emp_find_centers(const float* hm, int H, int W, float threshold, int nms_kernel,
                             int64_t* ctr_out, int cap, void* ws, size_t ws_bytes, void* stream)
{
    int rc = check_image(H, W);
    if (rc) return rc;
    EMP_REQUIRE(hm != nullptr, EMP_ERR_INVALID, "hm is null");
    EMP_REQUIRE(nms_kernel >= 1, EMP_ERR_INVALID, "nms_kernel must be >= 1");
    EMP_REQUIRE(cap >= 0, EMP_ERR_INVALID, "cap must be >= 0");
    const WsLayout L = ws_layout(H, W, cap, 1);
    if ((rc = check_ws(ws, ws_bytes, L.total))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EMP_CUDA_CHECK(cudaMemsetAsync(ws, 0, L.zero_bytes, st));
    return launch_centers(1, hm, H, W, threshold, nms_kernel, 1.0f, L, static_cast<char*>(ws), L.total, cap, ctr_out, cap, st);
}

EMP_API int emp_group_pixels(const int64_t* ctr, int K, const float* off, int H, int W, float step,
                             int chunksize, void* ids_out, int ids_i32, void* ws, size_t ws_bytes,
                             void* stream)
{
    int rc = check_image(H, W);
    if (rc) return rc;
    EMP_REQUIRE(ctr && off && ids_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(K > 0, EMP_ERR_INVALID, "group_pixels needs at least one center");
    const WsLayout L = ws_layout(H, W, K, 1);
    if ((rc = check_ws(ws, ws_bytes, L.total))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EMP_CUDA_CHECK(cudaMemsetAsync(ws, 0, L.zero_bytes, st));
    if ((rc = load_centers(ctr, K, step, L, static_cast<char*>(ws), st))) return rc;
    if ((rc = launch_bin(1, L, static_cast<char*>(ws), L.total, H, W, K, K, step, st))) return rc;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.off = off; a.off_stride = (size_t)2 * H * W;
    a.out = ids_out; a.out_stride = (size_t)H * W;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    { Things none; memset(&none, 0, sizeof(none)); fill_assign_common(a, L, none); }
    a.B = 1; a.H = H; a.W = W; a.step = step; a.chunksize = chunksize; a.k_cap = K; a.k_fixed = K;
    a.vec = (W % 4 == 0) && aligned16(off) && aligned16(ids_out);
    return launch_assign(SEM_NONE, ID_ARGMIN, ids_i32 ? OUT_IDS32 : OUT_IDS64, a, st);
}

EMP_API int emp_coarse_ids(const float* hm, const float* off, int h, int w, float threshold, int nms_kernel,
                           float step, int32_t* ids_out, int k_cap, void* ws, size_t ws_bytes, void* stream)
{
    int rc = check_image(h, w);
    if (rc) return rc;
    EMP_REQUIRE(hm && off && ids_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(nms_kernel >= 1 && k_cap >= 1, EMP_ERR_INVALID, "bad nms_kernel / k_cap");
    const WsLayout L = ws_layout(h, w, k_cap, 1);
    if ((rc = check_ws(ws, ws_bytes, L.total))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EMP_CUDA_CHECK(cudaMemsetAsync(ws, 0, L.zero_bytes, st));
    if ((rc = launch_centers(1, hm, h, w, threshold, nms_kernel, step, L, static_cast<char*>(ws), L.total, k_cap, nullptr, 0, st))) return rc;
    if ((rc = launch_bin(1, L, static_cast<char*>(ws), L.total, h, w, k_cap, -1, step, st))) return rc;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.off = off; a.off_stride = (size_t)2 * h * w;
    a.out = ids_out; a.out_stride = (size_t)h * w;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    { Things none; memset(&none, 0, sizeof(none)); fill_assign_common(a, L, none); }
    a.B = 1; a.H = h; a.W = w; a.step = step; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
    a.vec = (w % 4 == 0) && aligned16(off) && aligned16(ids_out);
    return launch_assign(SEM_NONE, ID_ARGMIN, OUT_IDS32, a, st);
}

EMP_API int emp_instance_segmentation(const int64_t* sem, const float* hm, const float* off, int H, int W,
                                      const int64_t* thing_list, int n_things, float threshold,
                                      int nms_kernel, int64_t* ins_out, int64_t* ctr_out, int cap,
                                      int k_cap, void* ws, size_t ws_bytes, void* stream)
{
    int rc = check_image(H, W);
    if (rc) return rc;
    EMP_REQUIRE(sem && hm && off && ins_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(nms_kernel >= 1 && k_cap >= 1, EMP_ERR_INVALID, "bad nms_kernel / k_cap");
    Things th;
    if ((rc = make_things(thing_list, n_things, &th))) return rc;
    const WsLayout L = ws_layout(H, W, k_cap, th.n);
    if ((rc = check_ws(ws, ws_bytes, L.total))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EMP_CUDA_CHECK(cudaMemsetAsync(ws, 0, L.zero_bytes, st));
    if ((rc = launch_centers(1, hm, H, W, threshold, nms_kernel, 1.0f, L, static_cast<char*>(ws), L.total, k_cap, ctr_out, cap, st))) return rc;
    if ((rc = launch_bin(1, L, static_cast<char*>(ws), L.total, H, W, k_cap, -1, 1.0f, st))) return rc;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.sem = sem; a.sem_stride = (size_t)H * W;
    a.off = off; a.off_stride = (size_t)2 * H * W;
    a.out = ins_out; a.out_stride = (size_t)H * W;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    fill_assign_common(a, L, th);
    a.B = 1; a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
    a.vec = (W % 4 == 0) && aligned16(sem) && aligned16(off) && aligned16(ins_out);
    return launch_assign(SEM_I64, ID_ARGMIN, OUT_IDS64, a, st);
}

static int merge_common(const void* sem, int sem_mode, int id_mode, const void* ids_in, int wc, int shift,
                        int H, int W, int64_t label_divisor, const int64_t* thing_list, int n_things,
                        int64_t stuff_area, int64_t void_label, int64_t max_id, const int32_t* k_dev,
                        int64_t* pan_out, void* ws, size_t ws_bytes, void* stream)
{
    int rc = check_image(H, W);
    if (rc) return rc;
    EMP_REQUIRE(sem && ids_in && pan_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(max_id >= 0 && max_id < (1ll << 30), EMP_ERR_INVALID, "max_id %lld out of range", (long long)max_id);
    Things th;
    if ((rc = make_things(thing_list, n_things, &th))) return rc;
    const int k_cap = (int)max_id;
    const WsLayout L = ws_layout(H, W, k_cap, th.n);
    if ((rc = check_ws(ws, ws_bytes, L.total))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EMP_CUDA_CHECK(cudaMemsetAsync(ws, 0, L.zero_bytes, st));
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.sem = sem; a.sem_stride = (size_t)H * W;
    a.ids_in = ids_in; a.ids_stride = 0; a.wc = wc; a.shift = shift;
    a.out = static_cast<char*>(ws) + L.codes; a.out_stride = 0;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    fill_assign_common(a, L, th);
    a.B = 1; a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = k_cap;
    a.max_id = max_id;
    a.vec = (W % 4 == 0) && aligned16(sem) && (id_mode != ID_DENSE || aligned16(ids_in));
    if ((rc = launch_assign(sem_mode, id_mode, L.code16 ? OUT_CODE16 : OUT_CODE32, a, st))) return rc;
    if ((rc = launch_build_lut(1, L, static_cast<char*>(ws), L.total, k_cap, k_cap, k_dev, th, label_divisor, void_label, stuff_area, H, W, st))) return rc;
    return launch_apply(1, L, static_cast<char*>(ws), L.total, pan_out, H, W, st);
}

EMP_API int emp_merge(const int64_t* sem, const int64_t* ins, int H, int W, int64_t label_divisor,
                      const int64_t* thing_list, int n_things, int64_t stuff_area, int64_t void_label,
                      int64_t max_id, int64_t* pan_out, void* ws, size_t ws_bytes, void* stream)
{
    return merge_common(sem, SEM_I64, ID_DENSE, ins, 0, 0, H, W, label_divisor, thing_list, n_things,
                        stuff_area, void_label, max_id, nullptr, pan_out, ws, ws_bytes, stream);
}

EMP_API int emp_merge_coarse(const void* sem, int sem_u8, const int32_t* coarse_ids, int hc, int wc,
                             int shift, int H, int W, int64_t label_divisor, const int64_t* thing_list,
                             int n_things, int64_t stuff_area, int64_t void_label, int64_t max_id,
                             const int32_t* k_dev, int64_t* pan_out, void* ws, size_t ws_bytes, void* stream)
{
    EMP_REQUIRE(shift >= 0 && shift < 16, EMP_ERR_INVALID, "bad shift %d", shift);
    EMP_REQUIRE(hc > 0 && wc > 0 && ((H - 1) >> shift) < hc && ((W - 1) >> shift) < wc, EMP_ERR_INVALID,
                "coarse map %d x %d << %d does not cover %d x %d", hc, wc, shift, H, W);
    return merge_common(sem, sem_u8 ? SEM_U8 : SEM_I64, ID_COARSE, coarse_ids, wc, shift, H, W, label_divisor,
                        thing_list, n_things, stuff_area, void_label, max_id, k_dev, pan_out, ws, ws_bytes, stream);
}

EMP_API int emp_panoptic_batched(int B, const void* sem, int sem_u8, const float* hm, const float* off,
                                 int H, int W, const int64_t* thing_list, int n_things,
                                 int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                                 float threshold, int nms_kernel, int64_t* pan_out, int64_t* ctr_out,
                                 int cap, int k_cap, void* ws, size_t ws_bytes_per_tile, void* stream)
{
    int rc = check_image(H, W);
    if (rc) return rc;
    EMP_REQUIRE(B >= 1 && B <= 65535, EMP_ERR_INVALID, "bad batch %d", B);
    EMP_REQUIRE(sem && hm && off && pan_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(nms_kernel >= 1 && k_cap >= 1, EMP_ERR_INVALID, "bad nms_kernel / k_cap");
    Things th;
    if ((rc = make_things(thing_list, n_things, &th))) return rc;
    const WsLayout L = ws_layout(H, W, k_cap, th.n);
    EMP_REQUIRE(ws_bytes_per_tile % 256 == 0, EMP_ERR_WORKSPACE, "per-tile workspace stride must be a multiple of 256");
    if ((rc = check_ws(ws, ws_bytes_per_tile, L.total))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* wsb = static_cast<char*>(ws);
    const size_t n_px = (size_t)H * W;

    {
        ProfScope ps(ST_MEMSET, st);
        EMP_CUDA_CHECK(cudaMemset2DAsync(wsb, ws_bytes_per_tile, 0, L.zero_bytes, (size_t)B, st));
    }
    if ((rc = launch_centers(B, hm, H, W, threshold, nms_kernel, 1.0f, L, wsb, ws_bytes_per_tile, k_cap, ctr_out, cap, st))) return rc;
    if ((rc = launch_bin(B, L, wsb, ws_bytes_per_tile, H, W, k_cap, -1, 1.0f, st))) return rc;

    const size_t sem_elt = sem_u8 ? 1 : 8;
    // assign -> label LUT -> apply in groups of tiles small enough that a group's code maps
    // (2 B/px) are still in L2 when apply_lut reads them back.
    const int G = tile_group(B);
    for (int b0 = 0; b0 < B; b0 += G) {
        const int nb = std::min(G, B - b0);
        char* wsg = wsb + (size_t)b0 * ws_bytes_per_tile;
        AssignArgs a;
        memset(&a, 0, sizeof(a));
        a.sem = static_cast<const char*>(sem) + (size_t)b0 * n_px * sem_elt; a.sem_stride = n_px;
        a.off = off + (size_t)b0 * 2 * n_px; a.off_stride = 2 * n_px;
        a.out = wsg + L.codes; a.out_stride = ws_bytes_per_tile / (L.code16 ? 2 : 4);
        a.ws = wsg; a.ws_stride = ws_bytes_per_tile;
        fill_assign_common(a, L, th);
        a.B = nb; a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
        a.vec = (W % 4 == 0) && aligned16(a.sem) && aligned16(a.off) && (n_px % 4 == 0 || nb == 1);
        if ((rc = launch_assign(sem_u8 ? SEM_U8 : SEM_I64, ID_ARGMIN, L.code16 ? OUT_CODE16 : OUT_CODE32, a, st))) return rc;
        if ((rc = launch_build_lut(nb, L, wsg, ws_bytes_per_tile, k_cap, -1, nullptr, th, label_divisor, void_label, stuff_area, H, W, st))) return rc;
        if ((rc = launch_apply(nb, L, wsg, ws_bytes_per_tile, pan_out + (size_t)b0 * n_px, H, W, st))) return rc;
    }
    return EMP_OK;
}
