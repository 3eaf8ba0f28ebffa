// panoptic.cu — find_instance_center / group_pixels / merge_semantic_and_instance /
// get_panoptic_segmentation (reference empanada/inference/postprocess.py) as sm_100a kernels.
//
// Kernel chain for one tile (all HBM-bound integer / fp32-compare work, no tensor cores):
//   nms_peaks      hm (4 B/px)                 -> peak bitmask (1 bit/px) + per-row counts
//   emit_centers   bitmask                     -> centers in row-major order, K
//   assign         sem (8 B/px) + off (8 B/px, thing sectors only)
//                                              -> code map (2 B/px) + votes + stuff areas
//   build_lut      votes, areas                -> label LUT (K+1) + class LUT
//   apply_lut      code map (2 B/px)           -> pan (8 B/px)
// DESIGN.md has the data layout, the exactness argument for the culled argmin and the roofline.
#include <math_constants.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
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
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static size_t g_prof_used = 0;

ProfScope::ProfScope(int stage, cudaStream_t s) : idx(-1), st(s)
{
    if (!g_prof_on) return;
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
    if (idx >= 0) cudaEventRecord(g_prof[idx].b, st);
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
__device__ __forceinline__ bool window_is_peak(const float* __restrict__ hm, int H, int W, int y,
                                               int x, float v, int lo, int hi)
{
    const int span = 2 * lo;            // lo >= hi always
    for (int i = 0; i <= span; ++i) {
        const int dy = (i & 1) ? -((i + 1) >> 1) : (i >> 1);      // 0,-1,+1,-2,+2,...
        if (dy < -lo || dy > hi) continue;
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const float* row = hm + (size_t)yy * W;
        for (int j = 0; j <= span; ++j) {
            const int dx = (j & 1) ? -((j + 1) >> 1) : (j >> 1);
            if (dx < -lo || dx > hi) continue;
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            if (__ldg(row + xx) > v) return false;
        }
    }
    return true;
}

// Called by a whole (converged) warp for one 32-pixel word of one row in which some lane is above
// threshold (a few % of all words).  Neighbour values come through L1 here (the streaming loop does
// not keep them): 3x3 neighbourhood first, the rest of the k x k window only for 3x3 maxima.
__device__ __noinline__ unsigned nms_word_check(const float* __restrict__ hm, int H, int W, int y, int x,
                                                bool cand, float thr, int lo, int hi)
{
    bool peak = false;
    if (cand) {
        const float c = __ldg(hm + (size_t)y * W + x);
        const bool before = lo >= 1, after = hi >= 1;   // window reaches to -1 / +1 at all?
        float m = -CUDART_INF_F;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dy == 0 && dx == 0) continue;
                if ((dy < 0 || dx < 0) && !before) continue;
                if ((dy > 0 || dx > 0) && !after) continue;
                const int yy = y + dy, xx = x + dx;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) m = fmaxf(m, __ldg(hm + (size_t)yy * W + xx));
            }
        }
        peak = !(m > c);
        if (peak && lo >= 2) peak = window_is_peak(hm, H, W, y, x, c, lo, hi);
    }
    return __ballot_sync(0xffffffffu, peak);
}

// One warp owns 4 rows x 256 pixels = 4 full 32-byte sectors of the peak mask, so it needs no
// shared memory and no barrier: lane l streams pixels 32j + l (j = 0..7) of each row (128-byte
// coalesced loads), keeps one "above threshold" flag per pixel, and only words with a flagged
// lane take the neighbourhood check.
constexpr int kNmsRows = 4, kNmsWords = 8;

__global__ void __launch_bounds__(256)
nms_peaks_kernel(const float* __restrict__ hm_base, size_t hm_stride, int H, int W, float thr,
                 int lo, int hi, char* __restrict__ ws_base, size_t ws_stride, size_t o_mask,
                 size_t o_rowcnt, int wd)
{
    const float* hm = hm_base + (size_t)blockIdx.z * hm_stride;
    char* ws = ws_base + (size_t)blockIdx.z * ws_stride;
    uint32_t* mask = reinterpret_cast<uint32_t*>(ws + o_mask);
    uint32_t* rowcnt = reinterpret_cast<uint32_t*>(ws + o_rowcnt);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w0 = blockIdx.x * kNmsWords;                          // first mask word of this warp
    const int y0 = (blockIdx.y * 8 + warp) * kNmsRows;
    if (y0 >= H) return;                                            // warp-uniform
    const int x0 = w0 * 32 + lane;

    unsigned cm = 0;                                                // bit 8*r + j: pixel (y0+r, x0+32j) is a candidate
    const long long Wl = W;
    const float* p = hm + (long long)y0 * Wl + x0;
#pragma unroll
    for (int r = 0; r < kNmsRows; ++r) {
        const bool rin = y0 + r < H;
#pragma unroll
        for (int j = 0; j < kNmsWords; ++j) {
            float v = -CUDART_INF_F;
            if (rin && x0 + 32 * j < W) v = __ldcs(p + 32 * j);
            cm |= ((v > thr && v > 0.0f) ? 1u : 0u) << (8 * r + j);
        }
        p += Wl;
    }
    unsigned mine[kNmsRows] = {0, 0, 0, 0};                         // lane j (< 8) ends up with word j of each row
    unsigned todo = __reduce_or_sync(0xffffffffu, cm);              // (row, word) pairs with any candidate
    while (todo) {                                                  // warp-uniform
        const int bit = __ffs(todo) - 1;
        todo &= todo - 1;
        const int r = bit >> 3, j = bit & 7;
        const unsigned word = nms_word_check(hm, H, W, y0 + r, x0 + 32 * j, (cm >> bit) & 1u, thr, lo, hi);
        if (lane == j) {
            if (r == 0) mine[0] = word; else if (r == 1) mine[1] = word; else if (r == 2) mine[2] = word; else mine[3] = word;
        }
    }
#pragma unroll
    for (int r = 0; r < kNmsRows; ++r) {
        const int y = y0 + r;
        if (y < H) {                                                // warp-uniform
            const bool wl = lane < kNmsWords && w0 + lane < wd;
            if (wl) mask[(size_t)y * wd + w0 + lane] = mine[r];
            const unsigned cnt = __reduce_add_sync(0xffffffffu, wl ? (unsigned)__popc(mine[r]) : 0u);
            if (cnt && lane == 0) atomicAdd(rowcnt + y, cnt);
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
                    size_t o_centers, size_t o_status, int H, int wd, int k_cap, float step,
                    int64_t* __restrict__ ctr_out_base, size_t ctr_out_stride, int cap)
{
    char* ws = ws_base + (size_t)blockIdx.z * ws_stride;
    const uint32_t* mask = reinterpret_cast<const uint32_t*>(ws + o_mask);
    const uint32_t* rowcnt = reinterpret_cast<const uint32_t*>(ws + o_rowcnt);
    float2* centers = reinterpret_cast<float2*>(ws + o_centers);     // (cy, cx) = step * (y, x)
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
                if (pos < k_cap) centers[pos] = make_float2(__fmul_rn(step, (float)y), __fmul_rn(step, (float)x));
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

// int64 (K,2) centers supplied by the caller (standalone group_pixels) -> int2 table
__global__ void load_centers_kernel(const int64_t* __restrict__ ctr, int K, float step, float2* __restrict__ centers)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K)          // ctr = step * ctr: int64 -> float32, then one rounded product (postprocess.py:152)
        centers[i] = make_float2(__fmul_rn(step, (float)ctr[2 * (size_t)i]), __fmul_rn(step, (float)ctr[2 * (size_t)i + 1]));
}

// ---------------------------------------------------------------------------------------------
// K3  assign — group_pixels (postprocess.py:146-167, :97-116) fused with the thing mask of
// get_instance_segmentation (:207-221) and the vote / stuff-area pass of
// merge_semantic_and_instance (:253-294); the last CTA to finish also builds the label LUT
// (:263-281), so no separate single-CTA launch sits between assign and apply_lut.
//
// One CTA per 64x32 pixel tile, 8 pixels per thread (2 row groups x 4 consecutive columns, so
// sem / offsets / codes move as 128-bit / 64-bit vectors and every warp touches whole lines).
// Tiles without a thing pixel (most of an EM tile) leave after one barrier: classify, count
// stuff area, write class codes.
//
// Exact culled argmin.  For each pixel the reference takes, over ALL K centers,
//       d_k = sqrt_rn(fma(dx, dx, rn(dy*dy))),  dy = cy_k - ly,  dx = cx_k - lx   (fp32)
// and keeps the first minimum.  The CTA bounds the shifted locations (ly,lx) of its thing
// pixels by a box, takes U2 = min_k maxdist^2(box, c_k) and keeps only centers with
// mindist^2(box, c_k) <= U2 * 1.001 + 1e-6: a dropped center is farther from every point of the
// box than some kept center by far more than the few-ulp rounding of d_k, so it can neither win
// nor tie.  Survivors are compacted in ascending k (ballot + prefix), so "first minimum" is a
// strict < on the rounded sqrt; sqrt is monotone, so it is only evaluated when s = d^2 improves.
// A non-finite location disables the cull for the tile.
// ---------------------------------------------------------------------------------------------
enum { SEM_NONE = 0, SEM_I64 = 1, SEM_U8 = 2 };
enum { ID_ARGMIN = 0, ID_DENSE = 1, ID_COARSE = 2 };
enum { OUT_CODE16 = 0, OUT_CODE32 = 1, OUT_IDS64 = 2, OUT_IDS32 = 3 };

struct AssignArgs {
    const void* sem;    size_t sem_stride;     // elements per tile
    const float* off;   size_t off_stride;     // floats per tile (2*H*W)
    const void* ids_in; size_t ids_stride;     // ID_DENSE: int64 H*W, ID_COARSE: int32 hc*wc
    void* out;          size_t out_stride;     // elements per tile
    char* ws;           size_t ws_stride;
    size_t o_status, o_centers, o_votes, o_areas, o_lut;
    const int32_t* k_dev;                      // optional device count bounding the LUT build
    int H, W, wc, shift;
    float step;
    int chunksize, k_cap, k_fixed;             // k_fixed >= 0: K known on the host
    long long max_id;
    long long label_divisor, void_label;
    int vec;                                   // 1: W % 4 == 0 and all planes 16-byte aligned
    unsigned long long thing_bits;             // bit c set <=> class c (< 64) is a thing class
    int things_small;                          // every thing class is < 64 (bit test suffices)
    Things things;
};

constexpr int kTileW = 64, kTileH = 32, kAssignThreads = 256, kPx = 8;
constexpr int kCandCap = 1024;
constexpr int kAreaBins = 64, kVoteSlots = 64;
constexpr unsigned kEmptyKey = 0xFFFFFFFFu;
constexpr unsigned kInfoThing = 0x8000u, kInfoBad = 0x4000u;   // per-pixel 16-bit info word

struct LutScratch {
    int run[EMP_MAX_THINGS];
    int wcnt[8];
};

struct AssignSmem {
    float cy[kCandCap], cx[kCandCap];
    int ck[kCandCap];
    float red[8][4];
    int redi[8];
    int wcnt[8];
    unsigned area[kAreaBins];
    unsigned vkey[kVoteSlots], vcnt[kVoteSlots];
    LutScratch lut;
    int last;
    int kshared;
};

// Rare paths are kept out of line so that the per-pixel code stays small and branch-light.
__device__ __noinline__ void vote_insert(AssignSmem& sm, uint32_t* votes, unsigned key, int cnt)
{
    unsigned h = (key * 2654435761u) >> 26;
#pragma unroll 1
    for (int probe = 0; probe < kVoteSlots; ++probe) {
        const unsigned slot = (h + probe) & (kVoteSlots - 1);
        const unsigned prev = atomicCAS(&sm.vkey[slot], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) { atomicAdd(&sm.vcnt[slot], (unsigned)cnt); return; }
    }
    atomicAdd(votes + key, (uint32_t)cnt);      // table full: straight to global
}

__device__ __forceinline__ void area_insert(unsigned* s_area, uint32_t* areas, unsigned cls, int cnt)
{
    if (cls < (unsigned)kAreaBins) atomicAdd(&s_area[cls], (unsigned)cnt);
    else atomicAdd(areas + cls, (uint32_t)cnt);
}

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

// merge_semantic_and_instance's bookkeeping (postprocess.py:263-281), by one 256-thread CTA:
//   id -> majority thing class (ties -> smallest class, torch.mode) * L + 1-based rank among voted
//   ids of that class in ascending id order.  `lut` may point to shared or global memory.
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

// Exact nearest center for the CTA's thing pixels (8 per thread): CTA-level cull of all K centers
// against the box of shifted locations, ordered compaction of the survivors into shared memory,
// warp-level cull against the warp's own box, then the reference's fp32 distance on what is left.
// Must be called by all threads of the CTA (it synchronises).  idv[p] receives the 1-based id.
__device__ __forceinline__ void nearest_center_8px(AssignSmem& sm, const float2* __restrict__ centers, int K,
                                                   int chunksize, unsigned thing, const float (&ly)[kPx],
                                                   const float (&lx)[kPx], int (&idv)[kPx])
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float by0 = CUDART_INF_F, by1 = -CUDART_INF_F, bx0 = CUDART_INF_F, bx1 = -CUDART_INF_F;
    int nonfinite = 0;
#pragma unroll
    for (int p = 0; p < kPx; ++p) {
        if (thing & (1u << p)) {
            by0 = fminf(by0, ly[p]); by1 = fmaxf(by1, ly[p]);
            bx0 = fminf(bx0, lx[p]); bx1 = fmaxf(bx1, lx[p]);
            if (!isfinite(ly[p]) || !isfinite(lx[p])) nonfinite = 1;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, d));
        by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, d));
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, d));
        bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, d));
    }
    nonfinite = __any_sync(0xffffffffu, nonfinite) ? 1 : 0;
    const float wy0 = by0, wy1 = by1, wx0 = bx0, wx1 = bx1;    // this warp's own box (4 x 64 px)
    const bool wnf = nonfinite != 0;
    const bool warp_has_thing = __any_sync(0xffffffffu, thing != 0);
    if (lane == 0) {
        sm.red[warp][0] = by0; sm.red[warp][1] = by1; sm.red[warp][2] = bx0; sm.red[warp][3] = bx1;
        sm.redi[warp] = nonfinite;
    }
    __syncthreads();
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) {
        by0 = fminf(by0, sm.red[w8][0]); by1 = fmaxf(by1, sm.red[w8][1]);
        bx0 = fminf(bx0, sm.red[w8][2]); bx1 = fmaxf(bx1, sm.red[w8][3]);
        nonfinite |= sm.redi[w8];
    }

    // sweep 1: U2 = min_k maxdist^2(box, c_k)
    float u2 = CUDART_INF_F;
    for (int k = tid; k < K; k += kAssignThreads) {
        const float2 c = __ldg(centers + k);
        const float my = fmaxf(fabsf(c.x - by0), fabsf(c.x - by1));
        const float mx = fmaxf(fabsf(c.y - bx0), fabsf(c.y - bx1));
        u2 = fminf(u2, my * my + mx * mx);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) u2 = fminf(u2, __shfl_xor_sync(0xffffffffu, u2, d));
    __syncthreads();                        // sm.red reads above are done
    if (lane == 0) sm.red[warp][0] = u2;
    __syncthreads();
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) u2 = fminf(u2, sm.red[w8][0]);
    const float thr2 = nonfinite ? CUDART_INF_F : u2 * 1.001f + 1e-6f;

    float best_s[kPx];
    int best_k[kPx];
#pragma unroll
    for (int p = 0; p < kPx; ++p) { best_s[p] = CUDART_INF_F; best_k[p] = -1; }

    // sweep 2: ordered compaction of survivors, evaluated in batches of <= kCandCap
    int n_list = 0;
    for (int base = 0; base < K; base += kAssignThreads) {
        const int k = base + tid;
        bool keep = false;
        float2 c = make_float2(0.f, 0.f);
        if (k < K) {
            c = __ldg(centers + k);                  // (cy, cx) = step * ctr (postprocess.py:152)
            const float dy = fmaxf(fmaxf(by0 - c.x, c.x - by1), 0.f);
            const float dx = fmaxf(fmaxf(bx0 - c.y, c.y - bx1), 0.f);
            keep = nonfinite || !(dy * dy + dx * dx > thr2);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sm.wcnt[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) { const int cc = sm.wcnt[w8]; if (w8 < warp) woff += cc; tot += cc; }
        if (keep) {
            const int pos = n_list + woff + __popc(bal & lanemask_lt());
            sm.cy[pos] = c.x; sm.cx[pos] = c.y; sm.ck[pos] = k;
        }
        n_list += tot;
        __syncthreads();
        if (n_list > kCandCap - kAssignThreads || base + kAssignThreads >= K) {
            if (warp_has_thing) {                   // warp-uniform
                // second, warp-level cull against this warp's own (much smaller) box: one
                // lane per candidate, then only the survivors are evaluated per pixel
                float wu2 = CUDART_INF_F;
                for (int j0 = 0; j0 < n_list; j0 += 32) {
                    const int j = j0 + lane;
                    if (j < n_list) {
                        const float ccy = sm.cy[j], ccx = sm.cx[j];
                        const float my = fmaxf(fabsf(ccy - wy0), fabsf(ccy - wy1));
                        const float mx = fmaxf(fabsf(ccx - wx0), fabsf(ccx - wx1));
                        wu2 = fminf(wu2, my * my + mx * mx);
                    }
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) wu2 = fminf(wu2, __shfl_xor_sync(0xffffffffu, wu2, d));
                const float wthr2 = wnf ? CUDART_INF_F : wu2 * 1.001f + 1e-6f;
                for (int j0 = 0; j0 < n_list; j0 += 32) {
                    const int j = j0 + lane;
                    bool wkeep = false;
                    if (j < n_list) {
                        const float ccy = sm.cy[j], ccx = sm.cx[j];
                        const float dy = fmaxf(fmaxf(wy0 - ccy, ccy - wy1), 0.f);
                        const float dx = fmaxf(fmaxf(wx0 - ccx, ccx - wx1), 0.f);
                        wkeep = wnf || !(dy * dy + dx * dx > wthr2);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, wkeep);
                    while (m) {                     // ascending j: ascending center index
                        const int jj = j0 + __ffs(m) - 1;
                        m &= m - 1;
                        const float ccy = sm.cy[jj], ccx = sm.cx[jj];
                        const int ck = sm.ck[jj];
#pragma unroll
                        for (int p = 0; p < kPx; ++p) {
                            if (thing & (1u << p)) {
                                const float dy = __fsub_rn(ccy, ly[p]);
                                const float dx = __fsub_rn(ccx, lx[p]);
                                const float s2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
                                if (s2 < best_s[p]) {
                                    if (__fsqrt_rn(s2) < __fsqrt_rn(best_s[p])) best_k[p] = ck;
                                    best_s[p] = s2;
                                }
                            }
                        }
                    }
                }
            }
            n_list = 0;
            __syncthreads();
        }
    }
    const bool chunked = K > chunksize;
#pragma unroll
    for (int p = 0; p < kPx; ++p) {
        int id = 0;
        if (thing & (1u << p)) {
            if (chunked) id = (best_k[p] >= 0 && __fsqrt_rn(best_s[p]) < 1e5f) ? best_k[p] + 1 : 0;
            else id = best_k[p] >= 0 ? best_k[p] + 1 : 1;
        }
        idv[p] = id;
    }
}

template <int SEM, int IDM, int OUT>
__global__ void __launch_bounds__(kAssignThreads, 3)
assign_kernel(const __grid_constant__ AssignArgs a)
{
    constexpr bool kCodes = (OUT == OUT_CODE16 || OUT == OUT_CODE32);
    constexpr uint32_t kClsBase = (OUT == OUT_CODE16) ? kClsBase16 : kClsBase32;
    __shared__ AssignSmem sm;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z;
    const int H = a.H, W = a.W;
    const size_t HW = (size_t)H * W;
    char* ws = a.ws + (size_t)b * a.ws_stride;
    int32_t* status = reinterpret_cast<int32_t*>(ws + a.o_status);
    const float2* centers = reinterpret_cast<const float2*>(ws + a.o_centers);
    uint32_t* votes = reinterpret_cast<uint32_t*>(ws + a.o_votes);
    uint32_t* areas = reinterpret_cast<uint32_t*>(ws + a.o_areas);
    const int T = a.things.n > 0 ? a.things.n : 1;
    const bool multi = a.things.n > 1;

    // Thread <-> pixel map: warp w owns tile rows 4w..4w+3, lane l owns columns 2l, 2l+1 of each,
    // so one warp-wide access covers one whole 64-pixel tile row: 512 B of int64 sem (LDG.128),
    // 256 B of an offset plane (LDG.64), 128 B of uint16 codes (STG.32) — always full sectors.
    const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
    const int col0 = tx0 + 2 * lane;
    const int row0 = ty0 + 4 * warp;
    const bool cin = col0 < W;
    const bool c1in = col0 + 1 < W;

    // ---- phase 1a: issue every load of the tile before anything depends on one --------------
    int Kld = 0;                // thread 0 resolves the device-side id count for the whole CTA
    if ((IDM == ID_ARGMIN || kCodes) && tid == 0) Kld = (int)lut_extent(a.k_fixed, a.k_cap, status, a.k_dev);
    if (kCodes) {
        if (tid < kAreaBins) sm.area[tid] = 0;
        if (tid < kVoteSlots) { sm.vkey[tid] = kEmptyKey; sm.vcnt[tid] = 0; }
    }
    long long sv[kPx];
    long long iv[kPx];
#pragma unroll
    for (int p = 0; p < kPx; ++p) { sv[p] = 0; iv[p] = 0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + i;
        const bool rin = row < H && cin;
        const size_t e = (size_t)row * W + col0;
        if (SEM == SEM_I64) {
            const long long* sp = reinterpret_cast<const long long*>(a.sem) + (size_t)b * a.sem_stride + e;
            if (a.vec) {
                if (rin) { const longlong2 u = __ldcs(reinterpret_cast<const longlong2*>(sp)); sv[2 * i] = u.x; sv[2 * i + 1] = u.y; }
            } else {
                if (rin) sv[2 * i] = __ldcs(sp);
                if (rin && c1in) sv[2 * i + 1] = __ldcs(sp + 1);
            }
        } else if (SEM == SEM_U8) {
            const unsigned char* sp = reinterpret_cast<const unsigned char*>(a.sem) + (size_t)b * a.sem_stride + e;
            if (a.vec) {
                if (rin) { const unsigned u = __ldcs(reinterpret_cast<const unsigned short*>(sp)); sv[2 * i] = u & 255u; sv[2 * i + 1] = u >> 8; }
            } else {
                if (rin) sv[2 * i] = sp[0];
                if (rin && c1in) sv[2 * i + 1] = sp[1];
            }
        }
        if (IDM == ID_DENSE) {
            const long long* ip = reinterpret_cast<const long long*>(a.ids_in) + (size_t)b * a.ids_stride + e;
            if (a.vec) {
                if (rin) { const longlong2 u = __ldcs(reinterpret_cast<const longlong2*>(ip)); iv[2 * i] = u.x; iv[2 * i + 1] = u.y; }
            } else {
                if (rin) iv[2 * i] = __ldcs(ip);
                if (rin && c1in) iv[2 * i + 1] = __ldcs(ip + 1);
            }
        } else if (IDM == ID_COARSE) {
            const int* ip = reinterpret_cast<const int*>(a.ids_in) + (size_t)b * a.ids_stride;
            const size_t crow = (size_t)(row >> a.shift) * a.wc;
            if (rin) iv[2 * i] = __ldg(ip + crow + (col0 >> a.shift));
            if (rin && c1in) iv[2 * i + 1] = __ldg(ip + crow + ((col0 + 1) >> a.shift));
        }
    }

    // ---- phase 1b: classify ------------------------------------------------------------------
    unsigned inb = 0;           // bit p: pixel p = 2*i + j is inside the image
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool rin = (row0 + i) < H && cin;
        if (rin) inb |= 1u << (2 * i);
        if (rin && c1in) inb |= 2u << (2 * i);
    }
    unsigned long long orall = 0ull, orid = 0ull;
#pragma unroll
    for (int p = 0; p < kPx; ++p) { orall |= (unsigned long long)sv[p]; orid |= (unsigned long long)iv[p]; }
    // Fast path: the thread's 8 pixels are all in the image, all background class 0 (not a thing
    // class) and carry no instance id that could void them — the bulk of an EM tile.
    const bool pure_bg = SEM != SEM_NONE && inb == 0xFFu && orall == 0ull && !(a.thing_bits & 1ull) &&
                         (IDM != ID_DENSE || orid == 0ull);
    unsigned w[kPx];            // 16-bit info word per pixel
    unsigned thing = 0;         // bit p: pixel p takes an instance id
    unsigned bad = 0;
    int flags = 0;
    int idv[kPx];               // instance id (ID_DENSE / ID_COARSE) or argmin result
#pragma unroll
    for (int p = 0; p < kPx; ++p) { w[p] = 0; idv[p] = 0; }
    if (!pure_bg) {
        if (SEM == SEM_NONE) {
#pragma unroll
            for (int p = 0; p < kPx; ++p) w[p] = kInfoThing;
        } else if (orall < 64ull) {
#pragma unroll
            for (int p = 0; p < kPx; ++p) w[p] = classify_small((unsigned)sv[p], a.thing_bits, multi);
        } else {
#pragma unroll
            for (int p = 0; p < kPx; ++p) w[p] = classify_slow(sv[p], a.thing_bits, a.things_small, a.things);
        }
#pragma unroll
        for (int p = 0; p < kPx; ++p) {
            w[p] = ((inb >> p) & 1u) ? w[p] : 0u;
            thing |= ((w[p] >> 15) & 1u) << p;
            bad |= ((w[p] >> 14) & 1u) << p;
        }
        if (bad) flags |= EMP_FLAG_CLASS_RANGE;
        if (IDM != ID_ARGMIN) {
#pragma unroll
            for (int p = 0; p < kPx; ++p) {
                long long v = iv[p];
                if (v < 0 || v > a.max_id) { flags |= EMP_FLAG_ID_RANGE; v = 0; }
                idv[p] = (int)v;
            }
        }
    }

    // ---- phase 2: nearest center over the culled candidate list ------------------------------
    if ((IDM == ID_ARGMIN || kCodes) && tid == 0) sm.kshared = Kld;
    // Barrier 1 (also publishes the smem tables): is the whole CTA plain background?  Then its
    // codes are a constant and it has nothing to vote, count or search: store and leave.  (The
    // area of class 0 is derived in apply_lut from what the other CTAs count, see below.)
    if (kCodes && SEM != SEM_NONE) {
        if (!__syncthreads_or(!pure_bg)) {                  // block-uniform
            {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const size_t o = (size_t)b * a.out_stride + (size_t)(row0 + i) * W + col0;
                    if (OUT == OUT_CODE16) {
                        unsigned short* op = reinterpret_cast<unsigned short*>(a.out) + o;
                        if (a.vec) *reinterpret_cast<unsigned*>(op) = kClsBase | (kClsBase << 16);
                        else { op[0] = (unsigned short)kClsBase; op[1] = (unsigned short)kClsBase; }
                    } else {
                        unsigned* op = reinterpret_cast<unsigned*>(a.out) + o;
                        if (a.vec) *reinterpret_cast<uint2*>(op) = make_uint2(kClsBase, kClsBase);
                        else { op[0] = kClsBase; op[1] = kClsBase; }
                    }
                }
                return;
            }
        }
    }
    if (IDM == ID_ARGMIN) {
        const int any = __syncthreads_or(thing != 0);
        const int K = sm.kshared;
        if (any && K > 0) {                                 // block-uniform
            float ly[kPx], lx[kPx];
            {
                float2 fy[4], fx[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    fy[i] = make_float2(0.f, 0.f); fx[i] = make_float2(0.f, 0.f);
                    const float* oy = a.off + (size_t)b * a.off_stride + (size_t)(row0 + i) * W + col0;
                    const float* ox = oy + HW;
                    const unsigned t2 = (thing >> (2 * i)) & 3u;
                    if (a.vec) {
                        if (t2) { fy[i] = __ldcs(reinterpret_cast<const float2*>(oy)); fx[i] = __ldcs(reinterpret_cast<const float2*>(ox)); }
                    } else {
                        if (t2 & 1u) { fy[i].x = __ldcs(oy); fx[i].x = __ldcs(ox); }
                        if (t2 & 2u) { fy[i].y = __ldcs(oy + 1); fx[i].y = __ldcs(ox + 1); }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float ycoord = __fmul_rn((float)(row0 + i), a.step);     // arange(0, H*step, step)
                    ly[2 * i] = __fadd_rn(ycoord, fy[i].x);
                    ly[2 * i + 1] = __fadd_rn(ycoord, fy[i].y);
                    lx[2 * i] = __fadd_rn(__fmul_rn((float)col0, a.step), fx[i].x);
                    lx[2 * i + 1] = __fadd_rn(__fmul_rn((float)(col0 + 1), a.step), fx[i].y);
                }
            }
            nearest_center_8px(sm, centers, K, a.chunksize, thing, ly, lx, idv);
        }
    }

    // ---- phase 3: outputs, votes (postprocess.py:263-273), stuff areas (:284-291) ---------------
    unsigned code[kPx];
    unsigned vkey = kEmptyKey, akey = kEmptyKey;
    int vcnt = 0, acnt = 0;
    int deficit = 0;            // in-image pixels that are NOT class-0 stuff (area[0] = H*W - sum)
    if (pure_bg) {
#pragma unroll
        for (int p = 0; p < kPx; ++p) code[p] = kCodes ? kClsBase : 0u;
    } else {
        unsigned voted = 0, stuff = 0;      // bit p: pixel p votes for its instance / counts as stuff area
#pragma unroll
        for (int p = 0; p < kPx; ++p) {
            const bool in = (inb >> p) & 1u;
            const bool th = (thing >> p) & 1u;
            const bool bd = (bad >> p) & 1u;
            const int id = idv[p];
            const bool v = th && id != 0;
            const bool st0 = in && !th && !bd && !(IDM == ID_DENSE && id > 0);
            const bool st = st0 && w[p] != 0u;              // class-0 stuff is counted by complement
            deficit += (in && !(st0 && w[p] == 0u)) ? 1 : 0;
            voted |= (v ? 1u : 0u) << p;
            stuff |= (st ? 1u : 0u) << p;
            if (kCodes) code[p] = v ? (unsigned)id : (st0 ? kClsBase + w[p] : 0u);
            else code[p] = th ? (unsigned)id : 0u;
        }
        if (kCodes) {
            // key of the thread's first voting / stuff pixel and how many of its pixels share it;
            // stragglers with another key take the out-of-line insert
#pragma unroll
            for (int p = kPx - 1; p >= 0; --p) {
                if ((voted >> p) & 1u) vkey = (unsigned)idv[p] * (unsigned)T + (w[p] & 15u);
                if ((stuff >> p) & 1u) akey = w[p];
            }
            unsigned vrest = 0, arest = 0;
#pragma unroll
            for (int p = 0; p < kPx; ++p) {
                const unsigned kv = (unsigned)idv[p] * (unsigned)T + (w[p] & 15u);
                const bool isv = (voted >> p) & 1u, isa = (stuff >> p) & 1u;
                vcnt += (isv && kv == vkey) ? 1 : 0;
                vrest |= ((isv && kv != vkey) ? 1u : 0u) << p;
                acnt += (isa && w[p] == akey) ? 1 : 0;
                arest |= ((isa && w[p] != akey) ? 1u : 0u) << p;
            }
            if (vrest | arest) {
#pragma unroll
                for (int p = 0; p < kPx; ++p) {
                    if ((vrest >> p) & 1u) vote_insert(sm, votes, (unsigned)idv[p] * (unsigned)T + (w[p] & 15u), 1);
                    if ((arest >> p) & 1u) area_insert(sm.area, areas, w[p], 1);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + i;
        if (row < H && cin) {
            const size_t o = (size_t)b * a.out_stride + (size_t)row * W + col0;
            const unsigned c0 = code[2 * i], c1 = code[2 * i + 1];
            if (OUT == OUT_CODE16) {
                unsigned short* op = reinterpret_cast<unsigned short*>(a.out) + o;
                if (a.vec) *reinterpret_cast<unsigned*>(op) = c0 | (c1 << 16);
                else { op[0] = (unsigned short)c0; if (c1in) op[1] = (unsigned short)c1; }
            } else if (OUT == OUT_CODE32) {
                unsigned* op = reinterpret_cast<unsigned*>(a.out) + o;
                if (a.vec) *reinterpret_cast<uint2*>(op) = make_uint2(c0, c1);
                else { op[0] = c0; if (c1in) op[1] = c1; }
            } else if (OUT == OUT_IDS64) {
                long long* op = reinterpret_cast<long long*>(a.out) + o;
                if (a.vec) __stcs(reinterpret_cast<longlong2*>(op), make_longlong2((long long)c0, (long long)c1));
                else { op[0] = (long long)c0; if (c1in) op[1] = (long long)c1; }
            } else {
                int* op = reinterpret_cast<int*>(a.out) + o;
                if (a.vec) *reinterpret_cast<int2*>(op) = make_int2((int)c0, (int)c1);
                else { op[0] = (int)c0; if (c1in) op[1] = (int)c1; }
            }
        }
    }
    if (flags) atomicOr(status + EMP_ST_FLAGS, flags);

    if (kCodes) {
        // one warp-aggregated insert per distinct key, then one global atomic per live bin per CTA
        {
            const unsigned peers = __match_any_sync(0xffffffffu, vkey);
            const int sum = __reduce_add_sync(peers, vcnt);
            if (vkey != kEmptyKey && lane == __ffs(peers) - 1) vote_insert(sm, votes, vkey, sum);
        }
        {
            const unsigned peers = __match_any_sync(0xffffffffu, akey);
            const int sum = __reduce_add_sync(peers, acnt);
            if (akey != kEmptyKey && lane == __ffs(peers) - 1) area_insert(sm.area, areas, akey, sum);
        }
        {
            const int sum = __reduce_add_sync(0xffffffffu, deficit);
            if (sum && lane == 0) atomicAdd(&sm.area[0], (unsigned)sum);    // bin 0 holds the deficit
        }
        __syncthreads();
        if (tid < kAreaBins && sm.area[tid]) atomicAdd(areas + (tid == 0 ? kNumClasses : tid), sm.area[tid]);
        if (tid < kVoteSlots && sm.vkey[tid] != kEmptyKey && sm.vcnt[tid]) atomicAdd(votes + sm.vkey[tid], sm.vcnt[tid]);

    }
}

// ---------------------------------------------------------------------------------------------
// Fused get_panoptic_segmentation path (emp_panoptic_batched) = three kernels per tile:
//
//   classify      every pixel: sem (8 B/px) -> code map (2 B/px).  Stuff pixels get their class
//                 code, thing pixels a "pending" code (kPend + thing index).  Few registers, no
//                 search: a pure streaming kernel at high occupancy.  CTAs holding thing pixels
//                 append their tile to a worklist; stuff areas are counted here (class 0 by
//                 complement, so plain-background CTAs count nothing and leave after one barrier).
//   argmin_tiles  persistent CTAs over the worklist (thing tiles only, ~1/4 of an EM tile):
//                 offsets (8 B/px, thing rows only) -> exact nearest center -> id codes + votes.
//                 The last CTA to finish builds the label LUT.
//   apply_lut     code map -> int64 labels.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPend16 = 0xEFF0u, kPend32 = 0xFFFFEFF0u;    // + thing index: thing pixel awaiting its id

struct ClassifyArgs {
    const void* sem;
    void* codes;
    char* ws;
    size_t o_status, o_areas, o_worklist;
    int H, W, vec;
    unsigned long long thing_bits;
    int things_small;
    Things things;
};

template <int SEM, bool C16>
__global__ void __launch_bounds__(256, 5)
classify_kernel(const __grid_constant__ ClassifyArgs a)
{
    constexpr uint32_t kClsBase = C16 ? kClsBase16 : kClsBase32;
    constexpr uint32_t kPend = C16 ? kPend16 : kPend32;
    __shared__ unsigned s_area[kAreaBins];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;
    int32_t* status = reinterpret_cast<int32_t*>(a.ws + a.o_status);
    uint32_t* areas = reinterpret_cast<uint32_t*>(a.ws + a.o_areas);
    const bool multi = a.things.n > 1;

    const int col0 = blockIdx.x * kTileW + 2 * lane;
    const int row0 = blockIdx.y * kTileH + 4 * warp;
    const bool cin = col0 < W, c1in = col0 + 1 < W;
    if (tid < kAreaBins) s_area[tid] = 0;

    long long sv[kPx];
#pragma unroll
    for (int p = 0; p < kPx; ++p) sv[p] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool rin = row0 + i < H && cin;
        const size_t e = (size_t)(row0 + i) * W + col0;
        if (SEM == SEM_I64) {
            const long long* sp = reinterpret_cast<const long long*>(a.sem) + e;
            if (a.vec) {
                if (rin) { const longlong2 u = __ldcs(reinterpret_cast<const longlong2*>(sp)); sv[2 * i] = u.x; sv[2 * i + 1] = u.y; }
            } else {
                if (rin) sv[2 * i] = __ldcs(sp);
                if (rin && c1in) sv[2 * i + 1] = __ldcs(sp + 1);
            }
        } else {
            const unsigned char* sp = reinterpret_cast<const unsigned char*>(a.sem) + e;
            if (a.vec) {
                if (rin) { const unsigned u = __ldcs(reinterpret_cast<const unsigned short*>(sp)); sv[2 * i] = u & 255u; sv[2 * i + 1] = u >> 8; }
            } else {
                if (rin) sv[2 * i] = sp[0];
                if (rin && c1in) sv[2 * i + 1] = sp[1];
            }
        }
    }
    unsigned inb = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool rin = row0 + i < H && cin;
        if (rin) inb |= 1u << (2 * i);
        if (rin && c1in) inb |= 2u << (2 * i);
    }
    unsigned long long orall = 0ull;
#pragma unroll
    for (int p = 0; p < kPx; ++p) orall |= (unsigned long long)sv[p];
    const bool pure_bg = inb == 0xFFu && orall == 0ull && !(a.thing_bits & 1ull);

    unsigned code[kPx];
    unsigned thing = 0, akey = kEmptyKey;
    int acnt = 0, deficit = 0, flags = 0;
    if (pure_bg) {
#pragma unroll
        for (int p = 0; p < kPx; ++p) code[p] = kClsBase;
    } else {
        unsigned w[kPx];
        if (orall < 64ull) {
#pragma unroll
            for (int p = 0; p < kPx; ++p) w[p] = classify_small((unsigned)sv[p], a.thing_bits, multi);
        } else {
#pragma unroll
            for (int p = 0; p < kPx; ++p) w[p] = classify_slow(sv[p], a.thing_bits, a.things_small, a.things);
        }
        unsigned stuff = 0;
#pragma unroll
        for (int p = 0; p < kPx; ++p) {
            const bool in = (inb >> p) & 1u;
            const bool th = in && (w[p] & kInfoThing);
            const bool bd = in && (w[p] & kInfoBad);
            const bool st0 = in && !th && !bd;
            thing |= (th ? 1u : 0u) << p;
            if (bd) flags |= EMP_FLAG_CLASS_RANGE;
            code[p] = th ? kPend + (w[p] & 15u) : (st0 ? kClsBase + w[p] : 0u);
            deficit += (in && !(st0 && w[p] == 0u)) ? 1 : 0;    // class-0 stuff is counted by complement
            stuff |= ((st0 && w[p] != 0u) ? 1u : 0u) << p;
        }
#pragma unroll
        for (int p = kPx - 1; p >= 0; --p) if ((stuff >> p) & 1u) akey = w[p];
        unsigned arest = 0;
#pragma unroll
        for (int p = 0; p < kPx; ++p) {
            const bool isa = (stuff >> p) & 1u;
            acnt += (isa && w[p] == akey) ? 1 : 0;
            arest |= ((isa && w[p] != akey) ? 1u : 0u) << p;
        }
        if (arest) {
#pragma unroll
            for (int p = 0; p < kPx; ++p) if ((arest >> p) & 1u) area_insert(s_area, areas, w[p], 1);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (row0 + i < H && cin) {
            const size_t o = (size_t)(row0 + i) * W + col0;
            const unsigned c0 = code[2 * i], c1 = code[2 * i + 1];
            if (C16) {
                unsigned short* op = reinterpret_cast<unsigned short*>(a.codes) + o;
                if (a.vec) *reinterpret_cast<unsigned*>(op) = c0 | (c1 << 16);
                else { op[0] = (unsigned short)c0; if (c1in) op[1] = (unsigned short)c1; }
            } else {
                unsigned* op = reinterpret_cast<unsigned*>(a.codes) + o;
                if (a.vec) *reinterpret_cast<uint2*>(op) = make_uint2(c0, c1);
                else { op[0] = c0; if (c1in) op[1] = c1; }
            }
        }
    }
    // Barrier 1 (also publishes s_area): plain-background CTAs are done.
    if (!__syncthreads_or(!pure_bg)) return;                        // block-uniform

    if (flags) atomicOr(status + EMP_ST_FLAGS, flags);
    {
        const unsigned peers = __match_any_sync(0xffffffffu, akey);
        const int sum = __reduce_add_sync(peers, acnt);
        if (akey != kEmptyKey && lane == __ffs(peers) - 1) area_insert(s_area, areas, akey, sum);
    }
    {
        const int sum = __reduce_add_sync(0xffffffffu, deficit);
        if (sum && lane == 0) atomicAdd(&s_area[0], (unsigned)sum);             // bin 0 holds the deficit
    }
    const int any_thing = __syncthreads_or(thing != 0);
    if (tid < kAreaBins && s_area[tid]) atomicAdd(areas + (tid == 0 ? kNumClasses : tid), s_area[tid]);
    if (any_thing && tid == 0) {
        const int idx = atomicAdd(status + EMP_ST_NTILES, 1);
        reinterpret_cast<uint32_t*>(a.ws + a.o_worklist)[idx] = blockIdx.y * gridDim.x + blockIdx.x;
    }
}

struct ArgminArgs {
    const float* off;
    void* codes;
    char* ws;
    size_t o_status, o_centers, o_votes, o_lut, o_worklist;
    int H, W, tiles_x, vec, k_cap, chunksize;
    float step;
    long long label_divisor, void_label;
    Things things;
};

template <bool C16>
__global__ void __launch_bounds__(kAssignThreads, 3)
argmin_tiles_kernel(const __grid_constant__ ArgminArgs a)
{
    constexpr uint32_t kPend = C16 ? kPend16 : kPend32;
    __shared__ AssignSmem sm;
    __shared__ int s_ntiles;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, W = a.W;
    const size_t HW = (size_t)H * W;
    int32_t* status = reinterpret_cast<int32_t*>(a.ws + a.o_status);
    const float2* centers = reinterpret_cast<const float2*>(a.ws + a.o_centers);
    uint32_t* votes = reinterpret_cast<uint32_t*>(a.ws + a.o_votes);
    const uint32_t* worklist = reinterpret_cast<const uint32_t*>(a.ws + a.o_worklist);
    const int T = a.things.n > 0 ? a.things.n : 1;

    if (tid == 0) {
        sm.kshared = min(__ldcg(status + EMP_ST_K), a.k_cap);
        s_ntiles = __ldcg(status + EMP_ST_NTILES);
    }
    if (tid < kVoteSlots) { sm.vkey[tid] = kEmptyKey; sm.vcnt[tid] = 0; }
    __syncthreads();
    const int K = sm.kshared;
    const int ntiles = s_ntiles;

    for (int it = blockIdx.x; it < ntiles; it += gridDim.x) {          // block-uniform
        const unsigned tile = __ldcg(worklist + it);
        const int col0 = (int)(tile % (unsigned)a.tiles_x) * kTileW + 2 * lane;
        const int row0 = (int)(tile / (unsigned)a.tiles_x) * kTileH + 4 * warp;
        const bool cin = col0 < W, c1in = col0 + 1 < W;

        // codes written by classify (L2-resident): which of my pixels await an id?
        unsigned code[kPx];
#pragma unroll
        for (int p = 0; p < kPx; ++p) code[p] = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool rin = row0 + i < H && cin;
            const size_t o = (size_t)(row0 + i) * W + col0;
            if (C16) {
                const unsigned short* cp = reinterpret_cast<const unsigned short*>(a.codes) + o;
                if (a.vec) {
                    if (rin) { const unsigned u = __ldcg(reinterpret_cast<const unsigned*>(cp)); code[2 * i] = u & 0xFFFFu; code[2 * i + 1] = u >> 16; }
                } else {
                    if (rin) code[2 * i] = __ldcg(cp);
                    if (rin && c1in) code[2 * i + 1] = __ldcg(cp + 1);
                }
            } else {
                const unsigned* cp = reinterpret_cast<const unsigned*>(a.codes) + o;
                if (a.vec) {
                    if (rin) { const uint2 u = __ldcg(reinterpret_cast<const uint2*>(cp)); code[2 * i] = u.x; code[2 * i + 1] = u.y; }
                } else {
                    if (rin) code[2 * i] = __ldcg(cp);
                    if (rin && c1in) code[2 * i + 1] = __ldcg(cp + 1);
                }
            }
        }
        unsigned thing = 0;
#pragma unroll
        for (int p = 0; p < kPx; ++p) thing |= ((code[p] - kPend < 16u) ? 1u : 0u) << p;

        int idv[kPx];
#pragma unroll
        for (int p = 0; p < kPx; ++p) idv[p] = 0;
        if (K > 0) {                                                    // block-uniform
            float ly[kPx], lx[kPx];
            float2 fy[4], fx[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fy[i] = make_float2(0.f, 0.f); fx[i] = make_float2(0.f, 0.f);
                const float* oy = a.off + (size_t)(row0 + i) * W + col0;
                const float* ox = oy + HW;
                const unsigned t2 = (thing >> (2 * i)) & 3u;
                if (a.vec) {
                    if (t2) { fy[i] = __ldcs(reinterpret_cast<const float2*>(oy)); fx[i] = __ldcs(reinterpret_cast<const float2*>(ox)); }
                } else {
                    if (t2 & 1u) { fy[i].x = __ldcs(oy); fx[i].x = __ldcs(ox); }
                    if (t2 & 2u) { fy[i].y = __ldcs(oy + 1); fx[i].y = __ldcs(ox + 1); }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float ycoord = __fmul_rn((float)(row0 + i), a.step);         // arange(0, H*step, step)
                ly[2 * i] = __fadd_rn(ycoord, fy[i].x);
                ly[2 * i + 1] = __fadd_rn(ycoord, fy[i].y);
                lx[2 * i] = __fadd_rn(__fmul_rn((float)col0, a.step), fx[i].x);
                lx[2 * i + 1] = __fadd_rn(__fmul_rn((float)(col0 + 1), a.step), fx[i].y);
            }
            nearest_center_8px(sm, centers, K, a.chunksize, thing, ly, lx, idv);
        }

        // id codes back into the code map (only rows that held pending pixels), votes
        unsigned vkey = kEmptyKey;
        int vcnt = 0;
        unsigned vrest = 0;
#pragma unroll
        for (int p = kPx - 1; p >= 0; --p)
            if (((thing >> p) & 1u) && idv[p] != 0) vkey = (unsigned)idv[p] * (unsigned)T + (code[p] - kPend);
#pragma unroll
        for (int p = 0; p < kPx; ++p) {
            const bool isv = ((thing >> p) & 1u) && idv[p] != 0;
            const unsigned kv = (unsigned)idv[p] * (unsigned)T + (code[p] - kPend);
            vcnt += (isv && kv == vkey) ? 1 : 0;
            vrest |= ((isv && kv != vkey) ? 1u : 0u) << p;
        }
        if (vrest) {
#pragma unroll
            for (int p = 0; p < kPx; ++p)
                if ((vrest >> p) & 1u) vote_insert(sm, votes, (unsigned)idv[p] * (unsigned)T + (code[p] - kPend), 1);
        }
        {
            const unsigned peers = __match_any_sync(0xffffffffu, vkey);
            const int sum = __reduce_add_sync(peers, vcnt);
            if (vkey != kEmptyKey && lane == __ffs(peers) - 1) vote_insert(sm, votes, vkey, sum);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned t2 = (thing >> (2 * i)) & 3u;
            if (t2) {
                const unsigned c0 = (t2 & 1u) ? (unsigned)idv[2 * i] : code[2 * i];
                const unsigned c1 = (t2 & 2u) ? (unsigned)idv[2 * i + 1] : code[2 * i + 1];
                const size_t o = (size_t)(row0 + i) * W + col0;
                if (C16) {
                    unsigned short* op = reinterpret_cast<unsigned short*>(a.codes) + o;
                    if (a.vec) *reinterpret_cast<unsigned*>(op) = c0 | (c1 << 16);
                    else { if (t2 & 1u) op[0] = (unsigned short)c0; if (t2 & 2u) op[1] = (unsigned short)c1; }
                } else {
                    unsigned* op = reinterpret_cast<unsigned*>(a.codes) + o;
                    if (a.vec) *reinterpret_cast<uint2*>(op) = make_uint2(c0, c1);
                    else { if (t2 & 1u) op[0] = c0; if (t2 & 2u) op[1] = c1; }
                }
            }
        }
        __syncthreads();
        if (tid < kVoteSlots) {
            if (sm.vkey[tid] != kEmptyKey && sm.vcnt[tid]) atomicAdd(votes + sm.vkey[tid], sm.vcnt[tid]);
            sm.vkey[tid] = kEmptyKey; sm.vcnt[tid] = 0;
        }
        __syncthreads();
    }

    // Last CTA builds the label LUT.  The barrier orders every thread's atomics before thread 0's
    // device-scope fence (fences are cumulative), which orders them before the ticket; the last CTA
    // reads the votes straight from L2 (__ldcg).
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        sm.last = (atomicAdd(status + EMP_ST_TICKET, 1) == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sm.last) {
        __threadfence();
        build_label_lut((long long)K, votes, a.things, a.label_divisor, a.void_label,
                        reinterpret_cast<long long*>(a.ws + a.o_lut), sm.lut);
    }
}

// label LUT for the merge entry points (emp_merge / emp_merge_coarse): one CTA
__global__ void __launch_bounds__(256)
build_lut_kernel(char* ws, size_t o_status, size_t o_votes, size_t o_lut, int k_cap, int k_fixed,
                 const int32_t* k_dev, const __grid_constant__ Things things, long long label_divisor, long long void_label)
{
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
// 16 codes per thread per iteration; a thread whose codes are all equal (background) decodes once.
// ---------------------------------------------------------------------------------------------
struct ApplyArgs {
    char* ws; size_t ws_stride;
    size_t o_codes, o_lut, o_areas;
    long long* pan; size_t n_px;
    long long label_divisor, stuff_area, void_label;
    int vec;
};

template <bool C16>
__device__ __forceinline__ long long decode(unsigned code, const long long* __restrict__ lut,
                                            const uint32_t* __restrict__ areas, const ApplyArgs& a)
{
    constexpr uint32_t base = C16 ? kClsBase16 : kClsBase32;
    if (code >= base) {
        const unsigned c = code - base;
        const long long area = c ? (long long)__ldg(areas + c) : (long long)a.n_px - (long long)__ldg(areas + kNumClasses);
        return (area >= a.stuff_area) ? (long long)c * a.label_divisor : a.void_label;
    }
    return __ldg(lut + code);
}

// A warp handles 512 consecutive pixels: in step q (0..7) lane l owns pixels 64q + 2l, 64q + 2l + 1,
// so each warp-wide load (128 B of uint16 codes) and store (512 B of int64 labels) is one
// contiguous run of full sectors.  One group per warp: short CTAs, cheap last wave.
template <bool C16>
__global__ void __launch_bounds__(256)
apply_lut_kernel(const __grid_constant__ ApplyArgs a)
{
    char* ws = a.ws + (size_t)blockIdx.z * a.ws_stride;
    const long long* lut = reinterpret_cast<const long long*>(ws + a.o_lut);
    const uint32_t* areas = reinterpret_cast<const uint32_t*>(ws + a.o_areas);
    long long* pan = a.pan + (size_t)blockIdx.z * a.n_px;
    const int lane = threadIdx.x & 31;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t n512 = a.vec ? a.n_px / 512 : 0;

    for (size_t g = warp_global; g < n512; g += n_warps) {
        const size_t base = g * 512;
        unsigned c0[8], c1[8];
        if (C16) {
            const unsigned* cp = reinterpret_cast<const unsigned*>(ws + a.o_codes) + base / 2 + lane;
#pragma unroll
            for (int q = 0; q < 8; ++q) { const unsigned u = __ldcs(cp + q * 32); c0[q] = u & 0xFFFFu; c1[q] = u >> 16; }
        } else {
            const uint2* cp = reinterpret_cast<const uint2*>(ws + a.o_codes) + base / 2 + lane;
#pragma unroll
            for (int q = 0; q < 8; ++q) { const uint2 u = __ldcs(cp + q * 32); c0[q] = u.x; c1[q] = u.y; }
        }
        longlong2* op = reinterpret_cast<longlong2*>(pan + base) + lane;
        bool same = true;
#pragma unroll
        for (int q = 0; q < 8; ++q) same = same && (c0[q] == c0[0]) && (c1[q] == c0[0]);
        if (same) {
            const long long v = decode<C16>(c0[0], lut, areas, a);
#pragma unroll
            for (int q = 0; q < 8; ++q) __stcs(op + q * 32, make_longlong2(v, v));
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                __stcs(op + q * 32, make_longlong2(decode<C16>(c0[q], lut, areas, a), decode<C16>(c1[q], lut, areas, a)));
        }
    }
    // tail (and the whole image when it is not 512-divisible / aligned): one pixel per thread
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = n512 * 512 + t0; i < a.n_px; i += stride) {
        const unsigned code = C16 ? (unsigned)reinterpret_cast<const unsigned short*>(ws + a.o_codes)[i]
                                  : reinterpret_cast<const unsigned*>(ws + a.o_codes)[i];
        pan[i] = decode<C16>(code, lut, areas, a);
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------------------------
static int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n = v;
    }
    return n;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_ws(const void* ws, size_t ws_bytes, size_t need)
{
    EMP_REQUIRE(ws != nullptr, EMP_ERR_WORKSPACE, "workspace is null");
    EMP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, EMP_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    EMP_REQUIRE(ws_bytes >= need, EMP_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, need);
    return EMP_OK;
}

int launch_centers(int B, const float* hm, int H, int W, float thr, int k, float step, const WsLayout& L,
                   char* ws, size_t ws_stride, int k_cap, int64_t* ctr_out, int cap, cudaStream_t st)
{
    const int lo = k / 2, hi = k - 1 - lo;
    dim3 g1((L.wd + kNmsWords - 1) / kNmsWords, (H + 8 * kNmsRows - 1) / (8 * kNmsRows), B);
    {
        ProfScope ps(ST_NMS, st);
        nms_peaks_kernel<<<g1, 256, 0, st>>>(hm, (size_t)H * W, H, W, thr, lo, hi, ws, ws_stride, L.mask, L.rowcnt, L.wd);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    dim3 g2((H + 31) / 32, 1, B);
    {
        ProfScope ps(ST_EMIT, st);
        emit_centers_kernel<<<g2, 256, 0, st>>>(ws, ws_stride, L.mask, L.rowcnt, L.centers, L.status, H, L.wd, k_cap,
                                                step, ctr_out, (size_t)cap * 2, cap);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

template <int SEM, int IDM>
static int launch_assign_out(int out_mode, const AssignArgs& a, dim3 grid, cudaStream_t st)
{
    ProfScope ps(ST_ASSIGN, st);
    switch (out_mode) {
        case OUT_CODE16: assign_kernel<SEM, IDM, OUT_CODE16><<<grid, kAssignThreads, 0, st>>>(a); break;
        case OUT_CODE32: assign_kernel<SEM, IDM, OUT_CODE32><<<grid, kAssignThreads, 0, st>>>(a); break;
        case OUT_IDS64:  assign_kernel<SEM, IDM, OUT_IDS64><<<grid, kAssignThreads, 0, st>>>(a); break;
        default:         assign_kernel<SEM, IDM, OUT_IDS32><<<grid, kAssignThreads, 0, st>>>(a); break;
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

int launch_assign(int B, int sem_mode, int id_mode, int out_mode, const AssignArgs& a, cudaStream_t st)
{
    dim3 grid((a.W + kTileW - 1) / kTileW, (a.H + kTileH - 1) / kTileH, B);
    if (id_mode == ID_ARGMIN) {
        if (sem_mode == SEM_NONE) return launch_assign_out<SEM_NONE, ID_ARGMIN>(out_mode, a, grid, st);
        if (sem_mode == SEM_I64) return launch_assign_out<SEM_I64, ID_ARGMIN>(out_mode, a, grid, st);
        return launch_assign_out<SEM_U8, ID_ARGMIN>(out_mode, a, grid, st);
    }
    if (id_mode == ID_DENSE) {
        EMP_REQUIRE(sem_mode == SEM_I64, EMP_ERR_INVALID, "dense-id merge needs int64 sem");
        return launch_assign_out<SEM_I64, ID_DENSE>(out_mode, a, grid, st);
    }
    if (sem_mode == SEM_I64) return launch_assign_out<SEM_I64, ID_COARSE>(out_mode, a, grid, st);
    EMP_REQUIRE(sem_mode == SEM_U8, EMP_ERR_INVALID, "coarse-id merge needs int64 or uint8 sem");
    return launch_assign_out<SEM_U8, ID_COARSE>(out_mode, a, grid, st);
}

void fill_assign_common(AssignArgs& a, const WsLayout& L, const Things& th, long long label_divisor, long long void_label)
{
    a.o_status = L.status; a.o_centers = L.centers; a.o_votes = L.votes; a.o_areas = L.areas; a.o_lut = L.lut;
    a.things = th;
    a.thing_bits = 0ull;
    a.things_small = 1;
    for (int i = 0; i < th.n; ++i) {
        if (th.v[i] >= 0 && th.v[i] < 64) a.thing_bits |= 1ull << th.v[i];
        else a.things_small = 0;
    }
    a.label_divisor = label_divisor;
    a.void_label = void_label;
}

int launch_apply(int B, const WsLayout& L, char* ws, size_t ws_stride, long long label_divisor, long long stuff_area,
                 long long void_label, int64_t* pan_out, size_t n_px, cudaStream_t st)
{
    ApplyArgs a;
    memset(&a, 0, sizeof(a));
    a.ws = ws; a.ws_stride = ws_stride;
    a.o_codes = L.codes; a.o_lut = L.lut; a.o_areas = L.areas;
    a.pan = reinterpret_cast<long long*>(pan_out); a.n_px = n_px;
    a.label_divisor = label_divisor; a.stuff_area = stuff_area; a.void_label = void_label;
    a.vec = aligned16(pan_out);
    const size_t groups = a.vec ? n_px / 512 : 0;
    size_t blocks = groups ? (groups + 7) / 8 : (n_px + 255) / 256;
    if (blocks > (1u << 30)) blocks = 1u << 30;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, 1, B);
    ProfScope ps(ST_APPLY, st);
    if (L.code16) apply_lut_kernel<true><<<grid, 256, 0, st>>>(a);
    else apply_lut_kernel<false><<<grid, 256, 0, st>>>(a);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

int launch_build_lut(const WsLayout& L, char* ws, int k_cap, int k_fixed, const int32_t* k_dev, const Things& th,
                     long long label_divisor, long long void_label, cudaStream_t st)
{
    ProfScope ps(ST_LUT, st);
    build_lut_kernel<<<1, 256, 0, st>>>(ws, L.status, L.votes, L.lut, k_cap, k_fixed, k_dev, th, label_divisor, void_label);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

// classify -> argmin_tiles (+ LUT) for one tile of the fused path
int launch_classify_argmin(const void* sem, int sem_u8, const float* off, int H, int W, const WsLayout& L, char* ws,
                           int k_cap, const Things& th, long long label_divisor, long long void_label, cudaStream_t st)
{
    const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
    ClassifyArgs c;
    memset(&c, 0, sizeof(c));
    c.sem = sem; c.codes = ws + L.codes; c.ws = ws;
    c.o_status = L.status; c.o_areas = L.areas; c.o_worklist = L.worklist;
    c.H = H; c.W = W;
    c.vec = (W % 4 == 0) && aligned16(sem);
    c.things = th; c.thing_bits = 0ull; c.things_small = 1;
    for (int i = 0; i < th.n; ++i) {
        if (th.v[i] >= 0 && th.v[i] < 64) c.thing_bits |= 1ull << th.v[i];
        else c.things_small = 0;
    }
    dim3 grid(tiles_x, tiles_y, 1);
    {
        ProfScope ps(ST_ASSIGN, st);
        if (L.code16) {
            if (sem_u8) classify_kernel<SEM_U8, true><<<grid, 256, 0, st>>>(c);
            else classify_kernel<SEM_I64, true><<<grid, 256, 0, st>>>(c);
        } else {
            if (sem_u8) classify_kernel<SEM_U8, false><<<grid, 256, 0, st>>>(c);
            else classify_kernel<SEM_I64, false><<<grid, 256, 0, st>>>(c);
        }
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    ArgminArgs g;
    memset(&g, 0, sizeof(g));
    g.off = off; g.codes = ws + L.codes; g.ws = ws;
    g.o_status = L.status; g.o_centers = L.centers; g.o_votes = L.votes; g.o_lut = L.lut; g.o_worklist = L.worklist;
    g.H = H; g.W = W; g.tiles_x = tiles_x; g.k_cap = k_cap; g.chunksize = 20; g.step = 1.0f;
    g.vec = (W % 4 == 0) && aligned16(off);
    g.label_divisor = label_divisor; g.void_label = void_label; g.things = th;
    int blocks = sm_count() * 3;
    if (blocks > tiles_x * tiles_y) blocks = tiles_x * tiles_y;
    {
        ProfScope ps(ST_LUT, st);       // profiling slot 3: argmin over thing tiles + label LUT
        if (L.code16) argmin_tiles_kernel<true><<<blocks, kAssignThreads, 0, st>>>(g);
        else argmin_tiles_kernel<false><<<blocks, kAssignThreads, 0, st>>>(g);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

int load_centers(const int64_t* ctr, int K, float step, const WsLayout& L, char* ws, cudaStream_t st)
{
    if (K > 0) {
        load_centers_kernel<<<(K + 255) / 256, 256, 0, st>>>(ctr, K, step, reinterpret_cast<float2*>(ws + L.centers));
        EMP_CUDA_CHECK(cudaGetLastError());
    }
    return EMP_OK;
}

}  // namespace emp

// =============================================================================================
// C ABI
// =============================================================================================
using namespace emp;

EMP_API int emp_version(void) { return 100; }

EMP_API int emp_profile_enable(int on)
{
    g_prof_on = on != 0;
    g_prof_used = 0;
    return EMP_OK;
}

EMP_API int emp_profile_read(double* ms_per_stage, int* launches_per_stage)
{
    for (int i = 0; i < ST_COUNT; ++i) { ms_per_stage[i] = 0.0; launches_per_stage[i] = 0; }
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
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.off = off; a.off_stride = (size_t)2 * H * W;
    a.out = ids_out; a.out_stride = (size_t)H * W;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    { Things none; memset(&none, 0, sizeof(none)); fill_assign_common(a, L, none, 0, 0); }
    a.H = H; a.W = W; a.step = step; a.chunksize = chunksize; a.k_cap = K; a.k_fixed = K;
    a.vec = (W % 4 == 0) && aligned16(off) && aligned16(ids_out);
    return launch_assign(1, SEM_NONE, ID_ARGMIN, ids_i32 ? OUT_IDS32 : OUT_IDS64, a, st);
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
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.off = off; a.off_stride = (size_t)2 * h * w;
    a.out = ids_out; a.out_stride = (size_t)h * w;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    { Things none; memset(&none, 0, sizeof(none)); fill_assign_common(a, L, none, 0, 0); }
    a.H = h; a.W = w; a.step = step; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
    a.vec = (w % 4 == 0) && aligned16(off) && aligned16(ids_out);
    return launch_assign(1, SEM_NONE, ID_ARGMIN, OUT_IDS32, a, st);
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
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.sem = sem; a.sem_stride = (size_t)H * W;
    a.off = off; a.off_stride = (size_t)2 * H * W;
    a.out = ins_out; a.out_stride = (size_t)H * W;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    fill_assign_common(a, L, th, 0, 0);
    a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
    a.vec = (W % 4 == 0) && aligned16(sem) && aligned16(off) && aligned16(ins_out);
    return launch_assign(1, SEM_I64, ID_ARGMIN, OUT_IDS64, a, st);
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
    fill_assign_common(a, L, th, label_divisor, void_label);
    a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = k_cap;
    a.max_id = max_id; a.k_dev = k_dev;
    a.vec = (W % 4 == 0) && aligned16(sem) && (id_mode != ID_DENSE || aligned16(ids_in));
    if ((rc = launch_assign(1, sem_mode, id_mode, L.code16 ? OUT_CODE16 : OUT_CODE32, a, st))) return rc;
    if ((rc = launch_build_lut(L, static_cast<char*>(ws), k_cap, k_cap, k_dev, th, label_divisor, void_label, st))) return rc;
    return launch_apply(1, L, static_cast<char*>(ws), L.total, label_divisor, stuff_area, void_label, pan_out,
                        (size_t)H * W, st);
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

    for (int b = 0; b < B; ++b)
        EMP_CUDA_CHECK(cudaMemsetAsync(wsb + (size_t)b * ws_bytes_per_tile, 0, L.zero_bytes, st));
    if ((rc = launch_centers(B, hm, H, W, threshold, nms_kernel, 1.0f, L, wsb, ws_bytes_per_tile, k_cap, ctr_out, cap, st))) return rc;

    const size_t sem_elt = sem_u8 ? 1 : 8;
    // classify -> argmin over thing tiles (+ label LUT) -> apply, tile by tile, so that a tile's code
    // map (2 B/px) is still in L2 when the next kernel reads it back.
    for (int b = 0; b < B; ++b) {
        char* wst = wsb + (size_t)b * ws_bytes_per_tile;
        if ((rc = launch_classify_argmin(static_cast<const char*>(sem) + (size_t)b * n_px * sem_elt, sem_u8,
                                         off + (size_t)b * 2 * n_px, H, W, L, wst, k_cap, th, label_divisor,
                                         void_label, st))) return rc;
        if ((rc = launch_apply(1, L, wst, ws_bytes_per_tile, label_divisor, stuff_area, void_label,
                               pan_out + (size_t)b * n_px, n_px, st))) return rc;
    }
    return EMP_OK;
}
