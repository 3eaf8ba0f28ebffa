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
// One lane per pixel column, 32 rows (+1 halo row above and below) per warp, all held in
// registers.  A row in which some lane is above threshold first checks the 3x3 neighbourhood with
// register / shuffle operands only (that alone rejects every slope pixel of a smooth heat-map);
// only 3x3-maxima walk the rest of the k x k window through L1.  Output: one ballot word per 32
// pixels and a per-row popcount.
// ---------------------------------------------------------------------------------------------
constexpr int kNmsRows = 32;

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
    const int wordcol = blockIdx.x * 4 + (warp & 3);
    const int y0 = blockIdx.y * (2 * kNmsRows) + (warp >> 2) * kNmsRows;
    const int x = wordcol * 32 + lane;
    const bool xin = x < W;
    if (wordcol >= wd || y0 >= H) return;       // warp-uniform
    const bool before = lo >= 1, after = hi >= 1;   // window reaches to -1 / +1 at all?

    float v[kNmsRows + 2];                      // v[i] = row y0 - 1 + i
#pragma unroll
    for (int i = 0; i < kNmsRows + 2; ++i) {
        const int y = y0 - 1 + i;
        v[i] = (xin && y >= 0 && y < H) ? __ldg(hm + (size_t)y * W + x) : -CUDART_INF_F;
    }
    unsigned myword = 0;
#pragma unroll
    for (int r = 0; r < kNmsRows; ++r) {
        const int y = y0 + r;
        const float c = v[r + 1];
        const bool cand = xin && c > thr && c > 0.0f;      // rows beyond H hold -inf
        bool peak = false;
        if (__any_sync(0xffffffffu, cand)) {
            const float up = v[r], dn = v[r + 2];
            float l = __shfl_up_sync(0xffffffffu, c, 1), rt = __shfl_down_sync(0xffffffffu, c, 1);
            float ul = __shfl_up_sync(0xffffffffu, up, 1), ur = __shfl_down_sync(0xffffffffu, up, 1);
            float dl = __shfl_up_sync(0xffffffffu, dn, 1), dr = __shfl_down_sync(0xffffffffu, dn, 1);
            if (cand && (lane == 0 || lane == 31)) {
                const int xx = lane == 0 ? x - 1 : x + 1;
                float e[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int yy = y - 1 + d;
                    e[d] = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? __ldg(hm + (size_t)yy * W + xx) : -CUDART_INF_F;
                }
                if (lane == 0) { ul = e[0]; l = e[1]; dl = e[2]; }
                else { ur = e[0]; rt = e[1]; dr = e[2]; }
            }
            float m = -CUDART_INF_F;
            if (before) { m = fmaxf(m, fmaxf(l, up)); m = fmaxf(m, ul); if (after) m = fmaxf(m, fmaxf(ur, dl)); }
            if (after) { m = fmaxf(m, fmaxf(rt, dn)); m = fmaxf(m, dr); }
            peak = cand && !(m > c);
            if (lo >= 2 && __any_sync(0xffffffffu, peak)) {
                if (peak) peak = window_is_peak(hm, H, W, y, x, c, lo, hi);
            }
        }
        const unsigned word = __ballot_sync(0xffffffffu, peak);
        if (lane == r) myword = word;
    }
    const int yw = y0 + lane;
    if (yw < H) {
        mask[(size_t)yw * wd + wordcol] = myword;
        if (myword) atomicAdd(rowcnt + yw, (uint32_t)__popc(myword));
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

struct AssignSmem {
    float cy[kCandCap], cx[kCandCap];
    int ck[kCandCap];
    float red[8][4];
    int redi[8];
    int wcnt[8];
    unsigned area[kAreaBins];
    unsigned vkey[kVoteSlots], vcnt[kVoteSlots];
    int run[EMP_MAX_THINGS];
    int last;
};

__device__ __forceinline__ void vote_insert(AssignSmem& sm, uint32_t* votes, unsigned key, int cnt)
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

__device__ __forceinline__ void area_insert(AssignSmem& sm, uint32_t* areas, unsigned cls, int cnt)
{
    if (cls < (unsigned)kAreaBins) atomicAdd(&sm.area[cls], (unsigned)cnt);
    else atomicAdd(areas + cls, (uint32_t)cnt);
}

// 16-bit info word of one pixel: thing -> kInfoThing | thing index, stuff -> class id,
// class outside [0, 4096) -> kInfoBad (reported through EMP_FLAG_CLASS_RANGE).
__device__ __forceinline__ unsigned classify(long long v, const AssignArgs& a)
{
    if ((unsigned long long)v < 64ull) {
        const unsigned c = (unsigned)v;
        if ((a.thing_bits >> c) & 1ull)
            return kInfoThing | (unsigned)__popcll(a.thing_bits & ((1ull << c) - 1ull));
        return c;
    }
    if (!a.things_small) {
        const int t = thing_index(v, a.things);
        if (t >= 0) return kInfoThing | (unsigned)t;
    }
    if (v < 0 || v >= kNumClasses) return kInfoBad;
    return (unsigned)v;
}

// merge_semantic_and_instance's bookkeeping (postprocess.py:263-281) by the last CTA of the grid:
//   id -> majority thing class (ties -> smallest class, torch.mode) * L + 1-based rank among voted
//   ids of that class in ascending id order.
__device__ void build_lut_tail(const AssignArgs& a, AssignSmem& sm, const int32_t* status,
                               const uint32_t* votes, long long* lut)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = a.things.n;
    const int T = nt > 0 ? nt : 1;
    long long K = a.k_fixed >= 0 ? (long long)a.k_fixed : (long long)min(__ldcg(status + EMP_ST_K), a.k_cap);
    if (a.k_dev) K = min(K, (long long)max(__ldcg(a.k_dev), 0));
    if (tid < EMP_MAX_THINGS) sm.run[tid] = 0;
    if (tid == 0) lut[0] = a.void_label;
    __syncthreads();
    for (long long base = 1; base <= K; base += kAssignThreads) {
        const long long id = base + tid;
        int t = -1;
        if (id <= K) {
            uint32_t best = 0;
            for (int c = 0; c < T; ++c) {
                const uint32_t v = __ldcg(votes + (size_t)id * T + c);
                if (v > best) { best = v; t = c; }
            }
            if (t < 0) lut[id] = a.void_label;
        }
        for (int c = 0; c < nt; ++c) {
            const bool f = (t == c);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) sm.wcnt[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const int x = sm.wcnt[w]; if (w < warp) woff += x; tot += x; }
            if (f) lut[id] = a.things.v[c] * a.label_divisor + (long long)(sm.run[c] + woff + __popc(bal & lanemask_lt()) + 1);
            __syncthreads();
            if (tid == 0) sm.run[c] += tot;
        }
        __syncthreads();
    }
}

template <int SEM, int IDM, int OUT>
__global__ void __launch_bounds__(kAssignThreads, 3)
assign_kernel(const AssignArgs a)
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

    // K is needed only after the first barrier: issue its load now so the latency overlaps
    int K = 0;
    if (IDM == ID_ARGMIN) K = a.k_fixed >= 0 ? a.k_fixed : min(__ldg(status + EMP_ST_K), a.k_cap);

    if (kCodes) {
        if (tid < kAreaBins) sm.area[tid] = 0;
        if (tid < kVoteSlots) { sm.vkey[tid] = kEmptyKey; sm.vcnt[tid] = 0; }
    }

    const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
    const int col0 = tx0 + (tid & 15) * 4;
    const int rowA = ty0 + (tid >> 4);              // second row group is rowA + 16
    const bool cin = col0 < W;

    // ---- phase 1: load sem / ids, classify --------------------------------------------------
    unsigned inb = 0;           // bit p: pixel p is inside the image
    unsigned thing = 0;         // bit p: pixel p takes an instance id
    unsigned info[kPx / 2];     // two 16-bit info words per register
    int idv[kPx];               // instance id (ID_DENSE / ID_COARSE) or argmin result
    int flags = 0;
#pragma unroll
    for (int q = 0; q < kPx / 2; ++q) info[q] = 0;
#pragma unroll
    for (int p = 0; p < kPx; ++p) idv[p] = 0;

#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = rowA + i * 16;
        const bool rin = row < H && cin;
        const size_t rbase = (size_t)row * W;
        unsigned w4[4] = {0, 0, 0, 0};
        if (SEM == SEM_NONE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) w4[j] = kInfoThing;
        } else if (SEM == SEM_I64) {
            const long long* sp = reinterpret_cast<const long long*>(a.sem) + (size_t)b * a.sem_stride + rbase + col0;
            long long sv[4] = {0, 0, 0, 0};
            if (rin && a.vec) {
                const longlong2 u0 = __ldcs(reinterpret_cast<const longlong2*>(sp));
                const longlong2 u1 = __ldcs(reinterpret_cast<const longlong2*>(sp) + 1);
                sv[0] = u0.x; sv[1] = u0.y; sv[2] = u1.x; sv[3] = u1.y;
            } else if (rin) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col0 + j < W) sv[j] = __ldcs(sp + j);
            }
            const unsigned long long any = (unsigned long long)(sv[0] | sv[1] | sv[2] | sv[3]);
            if (any == 0ull) {
                // all background: class 0 (a thing only if 0 is in thing_list)
                const unsigned w0 = (a.thing_bits & 1ull) ? kInfoThing : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) w4[j] = w0;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) w4[j] = classify(sv[j], a);
            }
        } else {
            const unsigned char* sp = reinterpret_cast<const unsigned char*>(a.sem) + (size_t)b * a.sem_stride + rbase + col0;
            unsigned u = 0;
            if (rin && a.vec) {
                u = __ldcs(reinterpret_cast<const unsigned*>(sp));
            } else if (rin) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col0 + j < W) u |= (unsigned)sp[j] << (8 * j);
            }
            if (u == 0u) {
                const unsigned w0 = (a.thing_bits & 1ull) ? kInfoThing : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) w4[j] = w0;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) w4[j] = classify((long long)((u >> (8 * j)) & 255u), a);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = i * 4 + j;
            const bool in = rin && (col0 + j < W);
            if (!in) w4[j] = 0;
            else inb |= 1u << p;
            if (w4[j] & kInfoThing) thing |= 1u << p;
            if (w4[j] & kInfoBad) flags |= EMP_FLAG_CLASS_RANGE;
        }
        info[i * 2] = w4[0] | (w4[1] << 16);
        info[i * 2 + 1] = w4[2] | (w4[3] << 16);

        if (IDM == ID_DENSE) {
            const long long* ip = reinterpret_cast<const long long*>(a.ids_in) + (size_t)b * a.ids_stride + rbase + col0;
            long long iv[4] = {0, 0, 0, 0};
            if (rin && a.vec) {
                const longlong2 u0 = __ldcs(reinterpret_cast<const longlong2*>(ip));
                const longlong2 u1 = __ldcs(reinterpret_cast<const longlong2*>(ip) + 1);
                iv[0] = u0.x; iv[1] = u0.y; iv[2] = u1.x; iv[3] = u1.y;
            } else if (rin) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col0 + j < W) iv[j] = __ldcs(ip + j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (iv[j] < 0 || iv[j] > a.max_id) { flags |= EMP_FLAG_ID_RANGE; iv[j] = 0; }
                idv[i * 4 + j] = (int)iv[j];
            }
        } else if (IDM == ID_COARSE) {
            const int* ip = reinterpret_cast<const int*>(a.ids_in) + (size_t)b * a.ids_stride;
            if (rin) {
                const size_t crow = (size_t)(row >> a.shift) * a.wc;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (col0 + j < W) {
                        int v = __ldg(ip + crow + ((col0 + j) >> a.shift));
                        if (v < 0 || v > a.max_id) { flags |= EMP_FLAG_ID_RANGE; v = 0; }
                        idv[i * 4 + j] = v;
                    }
                }
            }
        }
    }

    // ---- phase 2: nearest center over the culled candidate list ------------------------------
    if (IDM == ID_ARGMIN) {
        const int any = __syncthreads_or(thing != 0);       // also publishes the smem tables
        if (any && K > 0) {                                 // block-uniform
            float ly[kPx], lx[kPx];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = rowA + i * 16;
                float fy[4] = {0.f, 0.f, 0.f, 0.f}, fx[4] = {0.f, 0.f, 0.f, 0.f};
                if ((thing >> (4 * i)) & 15u) {
                    const float* oy = a.off + (size_t)b * a.off_stride + (size_t)row * W + col0;
                    const float* ox = oy + HW;
                    if (a.vec) {
                        const float4 u = __ldcs(reinterpret_cast<const float4*>(oy));
                        const float4 w = __ldcs(reinterpret_cast<const float4*>(ox));
                        fy[0] = u.x; fy[1] = u.y; fy[2] = u.z; fy[3] = u.w;
                        fx[0] = w.x; fx[1] = w.y; fx[2] = w.z; fx[3] = w.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (col0 + j < W) { fy[j] = __ldcs(oy + j); fx[j] = __ldcs(ox + j); }
                        }
                    }
                }
                const float ycoord = __fmul_rn((float)row, a.step);     // arange(0, H*step, step)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ly[i * 4 + j] = __fadd_rn(ycoord, fy[j]);
                    lx[i * 4 + j] = __fadd_rn(__fmul_rn((float)(col0 + j), a.step), fx[j]);
                }
            }
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
            if (lane == 0) {
                sm.red[warp][0] = by0; sm.red[warp][1] = by1; sm.red[warp][2] = bx0; sm.red[warp][3] = bx1;
                sm.redi[warp] = nonfinite;
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                by0 = fminf(by0, sm.red[w][0]); by1 = fmaxf(by1, sm.red[w][1]);
                bx0 = fminf(bx0, sm.red[w][2]); bx1 = fmaxf(bx1, sm.red[w][3]);
                nonfinite |= sm.redi[w];
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
            for (int w = 0; w < 8; ++w) u2 = fminf(u2, sm.red[w][0]);
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
                for (int w = 0; w < 8; ++w) { const int cc = sm.wcnt[w]; if (w < warp) woff += cc; tot += cc; }
                if (keep) {
                    const int pos = n_list + woff + __popc(bal & lanemask_lt());
                    sm.cy[pos] = c.x; sm.cx[pos] = c.y; sm.ck[pos] = k;
                }
                n_list += tot;
                __syncthreads();
                if (n_list > kCandCap - kAssignThreads || base + kAssignThreads >= K) {
                    if (thing != 0) {
                        for (int j = 0; j < n_list; ++j) {
                            const float ccy = sm.cy[j], ccx = sm.cx[j];
                            const int ck = sm.ck[j];
#pragma unroll
                            for (int p = 0; p < kPx; ++p) {
                                if (thing & (1u << p)) {
                                    const float dy = __fsub_rn(ccy, ly[p]);
                                    const float dx = __fsub_rn(ccx, lx[p]);
                                    const float s = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
                                    if (s < best_s[p]) {
                                        if (__fsqrt_rn(s) < __fsqrt_rn(best_s[p])) best_k[p] = ck;
                                        best_s[p] = s;
                                    }
                                }
                            }
                        }
                    }
                    n_list = 0;
                    __syncthreads();
                }
            }
            const bool chunked = K > a.chunksize;
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
    } else if (kCodes) {
        __syncthreads();                                    // publish the smem tables
    }

    // ---- phase 3: outputs, votes, stuff areas --------------------------------------------------
    unsigned vkey = kEmptyKey, akey = kEmptyKey;
    int vcnt = 0, acnt = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = rowA + i * 16;
        unsigned code[4] = {0, 0, 0, 0};
        long long idout[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = i * 4 + j;
            if (!(inb & (1u << p))) continue;
            const unsigned w = (info[p >> 1] >> (16 * (p & 1))) & 0xFFFFu;
            const bool th = (w & kInfoThing) != 0;
            const int id = idv[p];
            if (!kCodes) { idout[j] = th ? id : 0; continue; }
            if (th) {
                if (id != 0) {
                    code[j] = (unsigned)id;
                    const unsigned key = (unsigned)id * (unsigned)T + (w & 15u);
                    if (key == vkey) ++vcnt;
                    else { if (vcnt) vote_insert(sm, votes, vkey, vcnt); vkey = key; vcnt = 1; }
                }
            } else if (!(IDM == ID_DENSE && id > 0) && !(w & kInfoBad)) {
                code[j] = kClsBase + w;
                if (w == akey) ++acnt;
                else { if (acnt) area_insert(sm, areas, akey, acnt); akey = w; acnt = 1; }
            }
        }
        if (row < H && cin) {
            const size_t o = (size_t)b * a.out_stride + (size_t)row * W + col0;
            if (OUT == OUT_CODE16) {
                unsigned short* op = reinterpret_cast<unsigned short*>(a.out) + o;
                if (a.vec) {
                    *reinterpret_cast<uint2*>(op) = make_uint2(code[0] | (code[1] << 16), code[2] | (code[3] << 16));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (col0 + j < W) op[j] = (unsigned short)code[j];
                }
            } else if (OUT == OUT_CODE32) {
                unsigned* op = reinterpret_cast<unsigned*>(a.out) + o;
                if (a.vec) {
                    *reinterpret_cast<uint4*>(op) = make_uint4(code[0], code[1], code[2], code[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (col0 + j < W) op[j] = code[j];
                }
            } else if (OUT == OUT_IDS64) {
                long long* op = reinterpret_cast<long long*>(a.out) + o;
                if (a.vec) {
                    __stcs(reinterpret_cast<longlong2*>(op), make_longlong2(idout[0], idout[1]));
                    __stcs(reinterpret_cast<longlong2*>(op) + 1, make_longlong2(idout[2], idout[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (col0 + j < W) op[j] = idout[j];
                }
            } else {
                int* op = reinterpret_cast<int*>(a.out) + o;
                if (a.vec) {
                    *reinterpret_cast<int4*>(op) = make_int4((int)idout[0], (int)idout[1], (int)idout[2], (int)idout[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (col0 + j < W) op[j] = (int)idout[j];
                }
            }
        }
    }

    if (flags) atomicOr(status + EMP_ST_FLAGS, flags);
    if (kCodes) {
        // warp-aggregated flush of each thread's last run, then one global atomic per live bin
        {
            const unsigned peers = __match_any_sync(0xffffffffu, vkey);
            const int sum = __reduce_add_sync(peers, vcnt);
            if (vkey != kEmptyKey && sum > 0 && lane == __ffs(peers) - 1) vote_insert(sm, votes, vkey, sum);
        }
        {
            const unsigned peers = __match_any_sync(0xffffffffu, akey);
            const int sum = __reduce_add_sync(peers, acnt);
            if (akey != kEmptyKey && sum > 0 && lane == __ffs(peers) - 1) area_insert(sm, areas, akey, sum);
        }
        __syncthreads();
        if (tid < kAreaBins && sm.area[tid]) atomicAdd(areas + tid, sm.area[tid]);
        if (tid < kVoteSlots && sm.vkey[tid] != kEmptyKey && sm.vcnt[tid]) atomicAdd(votes + sm.vkey[tid], sm.vcnt[tid]);

        // last CTA of this tile builds the label LUT (threadFenceReduction pattern)
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int total = (int)(gridDim.x * gridDim.y);
            sm.last = (atomicAdd(status + EMP_ST_TICKET, 1) == total - 1) ? 1 : 0;
        }
        __syncthreads();
        if (sm.last) {
            __threadfence();
            build_lut_tail(a, sm, status, votes, reinterpret_cast<long long*>(ws + a.o_lut));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5  apply_lut — code map -> int64 panoptic labels (postprocess.py:281, :287-294).
//   code 0 -> void; 1..CLS_BASE-1 -> lut[id]; CLS_BASE + c -> c*L if area[c] >= stuff_area else void
// (a thing-class pixel never carries a class code, so no thing test is needed here).
// 16 codes per thread per iteration; a vector of identical codes (background) decodes once.
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
        return ((long long)__ldg(areas + c) >= a.stuff_area) ? (long long)c * a.label_divisor : a.void_label;
    }
    return __ldg(lut + code);
}

template <bool C16>
__device__ __forceinline__ void decode_store8(const unsigned (&w)[C16 ? 4 : 8], long long* out,
                                              const long long* __restrict__ lut,
                                              const uint32_t* __restrict__ areas, const ApplyArgs& a)
{
    longlong2* op = reinterpret_cast<longlong2*>(out);
    if (C16) {
        const unsigned w0 = w[0];
        if (w[1] == w0 && w[2] == w0 && w[3] == w0 && (w0 >> 16) == (w0 & 0xFFFFu)) {
            const long long v = decode<true>(w0 & 0xFFFFu, lut, areas, a);
#pragma unroll
            for (int q = 0; q < 4; ++q) __stcs(op + q, make_longlong2(v, v));
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                __stcs(op + q, make_longlong2(decode<true>(w[q] & 0xFFFFu, lut, areas, a),
                                              decode<true>(w[q] >> 16, lut, areas, a)));
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            __stcs(op + q, make_longlong2(decode<false>(w[2 * q], lut, areas, a),
                                          decode<false>(w[2 * q + 1], lut, areas, a)));
    }
}

template <bool C16>
__global__ void __launch_bounds__(256)
apply_lut_kernel(const ApplyArgs a)
{
    char* ws = a.ws + (size_t)blockIdx.z * a.ws_stride;
    const long long* lut = reinterpret_cast<const long long*>(ws + a.o_lut);
    const uint32_t* areas = reinterpret_cast<const uint32_t*>(ws + a.o_areas);
    long long* pan = a.pan + (size_t)blockIdx.z * a.n_px;
    const size_t n16 = a.vec ? a.n_px / 16 : 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4* cp = reinterpret_cast<const uint4*>(ws + a.o_codes);

    if (C16) {
        for (size_t i = t0; i < n16; i += stride) {
            const uint4 u0 = __ldcs(cp + 2 * i), u1 = __ldcs(cp + 2 * i + 1);
            const unsigned wa[4] = {u0.x, u0.y, u0.z, u0.w};
            const unsigned wb[4] = {u1.x, u1.y, u1.z, u1.w};
            decode_store8<true>(wa, pan + i * 16, lut, areas, a);
            decode_store8<true>(wb, pan + i * 16 + 8, lut, areas, a);
        }
        const unsigned short* cs = reinterpret_cast<const unsigned short*>(ws + a.o_codes);
        for (size_t i = n16 * 16 + t0; i < a.n_px; i += stride) pan[i] = decode<true>(cs[i], lut, areas, a);
    } else {
        for (size_t i = t0; i < n16; i += stride) {
            const uint4 u0 = __ldcs(cp + 4 * i), u1 = __ldcs(cp + 4 * i + 1);
            const uint4 u2 = __ldcs(cp + 4 * i + 2), u3 = __ldcs(cp + 4 * i + 3);
            const unsigned wa[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            const unsigned wb[8] = {u2.x, u2.y, u2.z, u2.w, u3.x, u3.y, u3.z, u3.w};
            decode_store8<false>(wa, pan + i * 16, lut, areas, a);
            decode_store8<false>(wb, pan + i * 16 + 8, lut, areas, a);
        }
        const unsigned* cs = reinterpret_cast<const unsigned*>(ws + a.o_codes);
        for (size_t i = n16 * 16 + t0; i < a.n_px; i += stride) pan[i] = decode<false>(cs[i], lut, areas, a);
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------------------------
static int g_sm_count = 0;

static int sm_count()
{
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count = n;
    }
    return g_sm_count;
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
    dim3 g1((L.wd + 3) / 4, (H + 2 * kNmsRows - 1) / (2 * kNmsRows), B);
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
    a.ws = ws; a.ws_stride = ws_stride;
    a.o_codes = L.codes; a.o_lut = L.lut; a.o_areas = L.areas;
    a.pan = reinterpret_cast<long long*>(pan_out); a.n_px = n_px;
    a.label_divisor = label_divisor; a.stuff_area = stuff_area; a.void_label = void_label;
    a.vec = aligned16(pan_out) && (n_px % 16 == 0);
    // one 16-pixel item per thread: short CTAs, so the last partial wave costs almost nothing
    size_t blocks = ((a.vec ? n_px / 16 : n_px) + 255) / 256;
    if (blocks > (1u << 30)) blocks = 1u << 30;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, 1, B);
    ProfScope ps(ST_APPLY, st);
    if (L.code16) apply_lut_kernel<true><<<grid, 256, 0, st>>>(a);
    else apply_lut_kernel<false><<<grid, 256, 0, st>>>(a);
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
    // assign (+ label LUT in its last CTA) -> apply tile by tile so that a tile's code map (2 B/px) is still in L2
    // when apply_lut reads it back.
    for (int b = 0; b < B; ++b) {
        char* wst = wsb + (size_t)b * ws_bytes_per_tile;
        AssignArgs a;
        memset(&a, 0, sizeof(a));
        a.sem = static_cast<const char*>(sem) + (size_t)b * n_px * sem_elt; a.sem_stride = n_px;
        a.off = off + (size_t)b * 2 * n_px; a.off_stride = 2 * n_px;
        a.out = wst + L.codes; a.out_stride = 0;
        a.ws = wst; a.ws_stride = ws_bytes_per_tile;
        fill_assign_common(a, L, th, label_divisor, void_label);
        a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
        a.vec = (W % 4 == 0) && aligned16(a.sem) && aligned16(a.off);
        if ((rc = launch_assign(1, sem_u8 ? SEM_U8 : SEM_I64, ID_ARGMIN, L.code16 ? OUT_CODE16 : OUT_CODE32, a, st))) return rc;
        if ((rc = launch_apply(1, L, wst, ws_bytes_per_tile, label_divisor, stuff_area, void_label,
                               pan_out + (size_t)b * n_px, n_px, st))) return rc;
    }
    return EMP_OK;
}
