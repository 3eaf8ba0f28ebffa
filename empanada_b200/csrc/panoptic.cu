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
// One lane per pixel column, 16 rows per warp; only above-threshold pixels (a few % of an EM
// heat-map) walk their window, nearest neighbours first so that slope pixels leave after a
// couple of L1 hits.  Output: one ballot word per 32 pixels and a per-row popcount.
// ---------------------------------------------------------------------------------------------
constexpr int kNmsRowsPerWarp = 16;

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
    const int y0 = blockIdx.y * (2 * kNmsRowsPerWarp) + (warp >> 2) * kNmsRowsPerWarp;
    const int x = wordcol * 32 + lane;
    const bool xin = x < W;
    if (wordcol >= wd || y0 >= H) return;       // warp-uniform

    float v[kNmsRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kNmsRowsPerWarp; ++r) {
        const int y = y0 + r;
        v[r] = (xin && y < H) ? __ldg(hm + (size_t)y * W + x) : -CUDART_INF_F;
    }
#pragma unroll
    for (int r = 0; r < kNmsRowsPerWarp; ++r) {
        const int y = y0 + r;
        if (y >= H) break;                      // warp-uniform
        const bool cand = xin && v[r] > thr && v[r] > 0.0f;
        bool peak = false;
        if (__any_sync(0xffffffffu, cand)) {
            if (cand) peak = window_is_peak(hm, H, W, y, x, v[r], lo, hi);
        }
        const unsigned word = __ballot_sync(0xffffffffu, peak);
        if (lane == 0) {
            mask[(size_t)y * wd + wordcol] = word;
            if (word) atomicAdd(rowcnt + y, (uint32_t)__popc(word));
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
                    size_t o_centers, size_t o_status, int H, int wd, int k_cap,
                    int64_t* __restrict__ ctr_out_base, size_t ctr_out_stride, int cap)
{
    char* ws = ws_base + (size_t)blockIdx.z * ws_stride;
    const uint32_t* mask = reinterpret_cast<const uint32_t*>(ws + o_mask);
    const uint32_t* rowcnt = reinterpret_cast<const uint32_t*>(ws + o_rowcnt);
    int2* centers = reinterpret_cast<int2*>(ws + o_centers);
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
                if (pos < k_cap) centers[pos] = make_int2(y, x);
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
__global__ void load_centers_kernel(const int64_t* __restrict__ ctr, int K, int2* __restrict__ centers)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) centers[i] = make_int2((int)ctr[2 * (size_t)i], (int)ctr[2 * (size_t)i + 1]);
}

// ---------------------------------------------------------------------------------------------
// K3  assign — group_pixels (postprocess.py:146-167, :97-116) fused with the thing mask of
// get_instance_segmentation (:207-221) and the vote / stuff-area pass of
// merge_semantic_and_instance (:253-294).
//
// One CTA per 64x32 pixel tile, 8 pixels per thread (2 row groups x 4 consecutive columns, so
// sem / offsets / codes move as 128-bit / 64-bit vectors and every warp touches whole lines).
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
    size_t o_status, o_centers, o_votes, o_areas;
    int H, W, wc, shift;
    float step;
    int chunksize, k_cap, k_fixed;             // k_fixed >= 0: K known on the host
    long long max_id;
    int vec;                                   // 1: W % 4 == 0 and all planes 16-byte aligned
    Things things;
};

constexpr int kTileW = 64, kTileH = 32, kAssignThreads = 256, kPx = 8;
constexpr int kCandCap = 1024;
constexpr int kAreaBins = 64, kVoteSlots = 64;
constexpr unsigned kEmptyKey = 0xFFFFFFFFu;

struct AssignSmem {
    float cy[kCandCap], cx[kCandCap];
    int ck[kCandCap];
    float red[8][4];
    int redi[8];
    int wcnt[8];
    unsigned area[kAreaBins];
    unsigned vkey[kVoteSlots], vcnt[kVoteSlots];
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

template <int SEM, int IDM, int OUT>
__global__ void __launch_bounds__(kAssignThreads)
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
    const int2* centers = reinterpret_cast<const int2*>(ws + a.o_centers);
    uint32_t* votes = reinterpret_cast<uint32_t*>(ws + a.o_votes);
    uint32_t* areas = reinterpret_cast<uint32_t*>(ws + a.o_areas);
    const int T = a.things.n > 0 ? a.things.n : 1;

    if (kCodes) {
        if (tid < kAreaBins) sm.area[tid] = 0;
        if (tid < kVoteSlots) { sm.vkey[tid] = kEmptyKey; sm.vcnt[tid] = 0; }
    }

    const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
    const int col0 = tx0 + (tid & 15) * 4;
    const int rowA = ty0 + (tid >> 4);              // second row group is rowA + 16

    // ---- phase 1: load sem / ids / offsets, classify -----------------------------------------
    unsigned inb = 0;           // bit p: pixel p is inside the image
    unsigned thing = 0;         // bit p: pixel p takes an instance id
    unsigned tcls = 0;          // 4 bits per pixel: index into things
    int cls[kPx];               // semantic class of non-thing pixels (for the code / area)
    int idv[kPx];               // instance id (ID_DENSE / ID_COARSE) or argmin result
    float ly[kPx], lx[kPx];
    int flags = 0;

#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = rowA + i * 16;
        const bool rin = row < H;
        const size_t rbase = (size_t)row * W;
        long long sv[4] = {0, 0, 0, 0};
        if (SEM == SEM_I64) {
            const long long* sp = reinterpret_cast<const long long*>(a.sem) + (size_t)b * a.sem_stride + rbase;
            if (rin && a.vec && col0 < W) {
                const longlong2 u0 = __ldcs(reinterpret_cast<const longlong2*>(sp + col0));
                const longlong2 u1 = __ldcs(reinterpret_cast<const longlong2*>(sp + col0) + 1);
                sv[0] = u0.x; sv[1] = u0.y; sv[2] = u1.x; sv[3] = u1.y;
            } else if (rin) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col0 + j < W) sv[j] = __ldcs(sp + col0 + j);
            }
        } else if (SEM == SEM_U8) {
            const unsigned char* sp = reinterpret_cast<const unsigned char*>(a.sem) + (size_t)b * a.sem_stride + rbase;
            if (rin && a.vec && col0 < W) {
                const unsigned u = __ldcs(reinterpret_cast<const unsigned*>(sp + col0));
                sv[0] = u & 255u; sv[1] = (u >> 8) & 255u; sv[2] = (u >> 16) & 255u; sv[3] = u >> 24;
            } else if (rin) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col0 + j < W) sv[j] = sp[col0 + j];
            }
        }
        unsigned grp_thing = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = i * 4 + j;
            const bool in = rin && (col0 + j < W);
            cls[p] = 0; idv[p] = 0; ly[p] = 0.f; lx[p] = 0.f;
            if (!in) continue;
            inb |= 1u << p;
            if (SEM == SEM_NONE) {
                thing |= 1u << p; grp_thing |= 1u << j;
            } else {
                const int t = thing_index(sv[j], a.things);
                if (t >= 0) { thing |= 1u << p; tcls |= (unsigned)t << (4 * p); grp_thing |= 1u << j; }
                else if (sv[j] < 0 || sv[j] >= kNumClasses) { flags |= EMP_FLAG_CLASS_RANGE; cls[p] = -1; }
                else cls[p] = (int)sv[j];
            }
        }
        if (IDM == ID_DENSE) {
            const long long* ip = reinterpret_cast<const long long*>(a.ids_in) + (size_t)b * a.ids_stride + rbase;
            long long iv[4] = {0, 0, 0, 0};
            if (rin && a.vec && col0 < W) {
                const longlong2 u0 = __ldcs(reinterpret_cast<const longlong2*>(ip + col0));
                const longlong2 u1 = __ldcs(reinterpret_cast<const longlong2*>(ip + col0) + 1);
                iv[0] = u0.x; iv[1] = u0.y; iv[2] = u1.x; iv[3] = u1.y;
            } else if (rin) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col0 + j < W) iv[j] = __ldcs(ip + col0 + j);
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
        } else {    // ID_ARGMIN: shifted locations of thing pixels
            if (grp_thing) {
                const float* oy = a.off + (size_t)b * a.off_stride + rbase;
                const float* ox = oy + HW;
                float fy[4], fx[4];
                if (a.vec) {
                    const float4 u = __ldcs(reinterpret_cast<const float4*>(oy + col0));
                    const float4 w = __ldcs(reinterpret_cast<const float4*>(ox + col0));
                    fy[0] = u.x; fy[1] = u.y; fy[2] = u.z; fy[3] = u.w;
                    fx[0] = w.x; fx[1] = w.y; fx[2] = w.z; fx[3] = w.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool in = col0 + j < W;
                        fy[j] = in ? __ldcs(oy + col0 + j) : 0.f;
                        fx[j] = in ? __ldcs(ox + col0 + j) : 0.f;
                    }
                }
                const float ycoord = __fmul_rn((float)row, a.step);     // arange(0, H*step, step)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ly[i * 4 + j] = __fadd_rn(ycoord, fy[j]);
                    lx[i * 4 + j] = __fadd_rn(__fmul_rn((float)(col0 + j), a.step), fx[j]);
                }
            }
        }
    }

    // ---- phase 2: nearest center over the culled candidate list ------------------------------
    if (IDM == ID_ARGMIN) {
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
        int any = thing != 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, d));
            by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, d));
            bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, d));
            bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, d));
        }
        any = __any_sync(0xffffffffu, any) ? 1 : 0;
        nonfinite = __any_sync(0xffffffffu, nonfinite) ? 1 : 0;
        if (lane == 0) {
            sm.red[warp][0] = by0; sm.red[warp][1] = by1; sm.red[warp][2] = bx0; sm.red[warp][3] = bx1;
            sm.redi[warp] = any | (nonfinite << 1);
        }
        __syncthreads();
        int fl = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            by0 = fminf(by0, sm.red[w][0]); by1 = fmaxf(by1, sm.red[w][1]);
            bx0 = fminf(bx0, sm.red[w][2]); bx1 = fmaxf(bx1, sm.red[w][3]);
            fl |= sm.redi[w];
        }
        any = fl & 1;
        nonfinite = (fl >> 1) & 1;

        int K = a.k_fixed >= 0 ? a.k_fixed : min(status[EMP_ST_K], a.k_cap);
        float best_s[kPx];
        int best_k[kPx];
#pragma unroll
        for (int p = 0; p < kPx; ++p) { best_s[p] = CUDART_INF_F; best_k[p] = -1; }

        if (any && K > 0) {         // block-uniform
            // sweep 1: U2 = min_k maxdist^2(box, c_k)
            float u2 = CUDART_INF_F;
            for (int k = tid; k < K; k += kAssignThreads) {
                const int2 c = __ldg(centers + k);
                const float cy = __fmul_rn(a.step, (float)c.x), cx = __fmul_rn(a.step, (float)c.y);
                const float my = fmaxf(fabsf(cy - by0), fabsf(cy - by1));
                const float mx = fmaxf(fabsf(cx - bx0), fabsf(cx - bx1));
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

            // sweep 2: ordered compaction of survivors, evaluated in batches of <= kCandCap
            int n_list = 0;
            for (int base = 0; base < K; base += kAssignThreads) {
                const int k = base + tid;
                bool keep = false;
                float cy = 0.f, cx = 0.f;
                if (k < K) {
                    const int2 c = __ldg(centers + k);
                    cy = __fmul_rn(a.step, (float)c.x);      // ctr = step * ctr (postprocess.py:152)
                    cx = __fmul_rn(a.step, (float)c.y);
                    const float dy = fmaxf(fmaxf(by0 - cy, cy - by1), 0.f);
                    const float dx = fmaxf(fmaxf(bx0 - cx, cx - bx1), 0.f);
                    keep = nonfinite || !(dy * dy + dx * dx > thr2);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) sm.wcnt[warp] = __popc(bal);
                __syncthreads();
                int woff = 0, tot = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { const int c = sm.wcnt[w]; if (w < warp) woff += c; tot += c; }
                if (keep) {
                    const int pos = n_list + woff + __popc(bal & lanemask_lt());
                    sm.cy[pos] = cy; sm.cx[pos] = cx; sm.ck[pos] = k;
                }
                n_list += tot;
                __syncthreads();
                if (n_list > kCandCap - kAssignThreads || base + kAssignThreads >= K) {
                    if (__any_sync(0xffffffffu, thing != 0)) {
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
        }
        const bool chunked = K > a.chunksize;
#pragma unroll
        for (int p = 0; p < kPx; ++p) {
            int id = 0;
            if ((thing & (1u << p)) && K > 0) {
                if (chunked) id = (best_k[p] >= 0 && __fsqrt_rn(best_s[p]) < 1e5f) ? best_k[p] + 1 : 0;
                else id = best_k[p] >= 0 ? best_k[p] + 1 : 1;
            }
            idv[p] = id;
        }
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
            const bool th = (thing >> p) & 1u;
            const int id = idv[p];
            if (!kCodes) { idout[j] = th ? id : 0; continue; }
            if (th) {
                if (id != 0) {
                    code[j] = (unsigned)id;
                    const unsigned key = (unsigned)id * (unsigned)T + ((tcls >> (4 * p)) & 15u);
                    if (key == vkey) ++vcnt;
                    else { if (vcnt) vote_insert(sm, votes, vkey, vcnt); vkey = key; vcnt = 1; }
                }
            } else if (!(IDM == ID_DENSE && id > 0) && cls[p] >= 0) {
                code[j] = kClsBase + (unsigned)cls[p];
                const unsigned key = (unsigned)cls[p];
                if (key == akey) ++acnt;
                else { if (acnt) area_insert(sm, areas, akey, acnt); akey = key; acnt = 1; }
            }
        }
        if (row < H && col0 < W) {
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
    }
    if (flags) atomicOr(status + EMP_ST_FLAGS, flags);
}

// ---------------------------------------------------------------------------------------------
// K4  build_lut — merge_semantic_and_instance's bookkeeping (postprocess.py:263-294):
//   id -> majority thing class (ties -> smallest class, torch.mode) * L + 1-based rank among
//   voted ids of that class in ascending id order; class c -> c*L if area >= stuff_area.
// One CTA per tile; per class a ballot/prefix scan over ids in chunks of 1024.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
build_lut_kernel(char* __restrict__ ws_base, size_t ws_stride, size_t o_status, size_t o_votes,
                 size_t o_areas, size_t o_lut, size_t o_clut, int k_cap, long long k_fixed,
                 const int32_t* __restrict__ k_dev, Things things, long long L, long long stuff_area,
                 long long void_label)
{
    char* ws = ws_base + (size_t)blockIdx.x * ws_stride;
    const int32_t* status = reinterpret_cast<const int32_t*>(ws + o_status);
    const uint32_t* votes = reinterpret_cast<const uint32_t*>(ws + o_votes);
    const uint32_t* areas = reinterpret_cast<const uint32_t*>(ws + o_areas);
    long long* lut = reinterpret_cast<long long*>(ws + o_lut);
    long long* clut = reinterpret_cast<long long*>(ws + o_clut);

    __shared__ int s_run[EMP_MAX_THINGS];
    __shared__ int s_w[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = things.n > 0 ? things.n : 1;
    long long K = k_fixed >= 0 ? k_fixed : (long long)min(status[EMP_ST_K], k_cap);
    if (k_dev) K = min(K, (long long)max(*k_dev, 0));

    if (tid < EMP_MAX_THINGS) s_run[tid] = 0;
    if (tid == 0) lut[0] = void_label;
    for (int c = tid; c < kNumClasses; c += 1024) {
        const bool is_thing = thing_index(c, things) >= 0;
        clut[c] = (!is_thing && (long long)areas[c] >= stuff_area) ? (long long)c * L : void_label;
    }
    __syncthreads();

    for (long long base = 1; base <= K; base += 1024) {
        const long long id = base + tid;
        int t = -1;
        if (id <= K) {
            uint32_t best = 0;
            for (int c = 0; c < T; ++c) {
                const uint32_t v = votes[(size_t)id * T + c];
                if (v > best) { best = v; t = c; }
            }
            if (t < 0) lut[id] = void_label;
        }
        for (int c = 0; c < things.n; ++c) {
            const bool f = (t == c);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_w[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < 32; ++w) { const int x = s_w[w]; if (w < warp) woff += x; tot += x; }
            if (f) lut[id] = things.v[c] * L + (long long)(s_run[c] + woff + __popc(bal & lanemask_lt()) + 1);
            __syncthreads();
            if (tid == 0) s_run[c] += tot;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K5  apply_lut — code map -> int64 panoptic labels (postprocess.py:281, :294).
// ---------------------------------------------------------------------------------------------
template <bool C16>
__device__ __forceinline__ long long decode(unsigned code, const long long* __restrict__ lut,
                                            const long long* __restrict__ clut)
{
    constexpr uint32_t base = C16 ? kClsBase16 : kClsBase32;
    return code >= base ? __ldg(clut + (code - base)) : __ldg(lut + code);
}

template <bool C16>
__global__ void __launch_bounds__(256)
apply_lut_kernel(char* __restrict__ ws_base, size_t ws_stride, size_t o_codes, size_t o_lut,
                 size_t o_clut, long long* __restrict__ pan_base, size_t n_px, int vec)
{
    char* ws = ws_base + (size_t)blockIdx.z * ws_stride;
    const long long* lut = reinterpret_cast<const long long*>(ws + o_lut);
    const long long* clut = reinterpret_cast<const long long*>(ws + o_clut);
    long long* pan = pan_base + (size_t)blockIdx.z * n_px;
    const size_t n8 = vec ? n_px / 8 : 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;

    if (C16) {
        const uint4* cp = reinterpret_cast<const uint4*>(ws + o_codes);
        for (size_t i = t0; i < n8; i += stride) {
            const uint4 u = __ldcs(cp + i);
            const unsigned w[4] = {u.x, u.y, u.z, u.w};
            longlong2* op = reinterpret_cast<longlong2*>(pan + i * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                __stcs(op + q, make_longlong2(decode<true>(w[q] & 0xFFFFu, lut, clut),
                                              decode<true>(w[q] >> 16, lut, clut)));
        }
        const unsigned short* cs = reinterpret_cast<const unsigned short*>(ws + o_codes);
        for (size_t i = n8 * 8 + t0; i < n_px; i += stride) pan[i] = decode<true>(cs[i], lut, clut);
    } else {
        const uint4* cp = reinterpret_cast<const uint4*>(ws + o_codes);
        for (size_t i = t0; i < n8; i += stride) {
            const uint4 u0 = __ldcs(cp + 2 * i), u1 = __ldcs(cp + 2 * i + 1);
            const unsigned w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            longlong2* op = reinterpret_cast<longlong2*>(pan + i * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                __stcs(op + q, make_longlong2(decode<false>(w[2 * q], lut, clut),
                                              decode<false>(w[2 * q + 1], lut, clut)));
        }
        const unsigned* cs = reinterpret_cast<const unsigned*>(ws + o_codes);
        for (size_t i = n8 * 8 + t0; i < n_px; i += stride) pan[i] = decode<false>(cs[i], lut, clut);
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

int launch_centers(int B, const float* hm, int H, int W, float thr, int k, const WsLayout& L,
                   char* ws, size_t ws_stride, int k_cap, int64_t* ctr_out, int cap, cudaStream_t st)
{
    const int lo = k / 2, hi = k - 1 - lo;
    dim3 g1((L.wd + 3) / 4, (H + 2 * kNmsRowsPerWarp - 1) / (2 * kNmsRowsPerWarp), B);
    {
        ProfScope ps(ST_NMS, st);
        nms_peaks_kernel<<<g1, 256, 0, st>>>(hm, (size_t)H * W, H, W, thr, lo, hi, ws, ws_stride, L.mask, L.rowcnt, L.wd);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    dim3 g2((H + 31) / 32, 1, B);
    {
        ProfScope ps(ST_EMIT, st);
        emit_centers_kernel<<<g2, 256, 0, st>>>(ws, ws_stride, L.mask, L.rowcnt, L.centers, L.status, H, L.wd, k_cap,
                                                ctr_out, (size_t)cap * 2, cap);
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

int launch_lut_and_apply(int B, const WsLayout& L, char* ws, size_t ws_stride, int k_cap, long long k_fixed,
                         const int32_t* k_dev, const Things& things, long long label_divisor, long long stuff_area,
                         long long void_label, int64_t* pan_out, size_t n_px, cudaStream_t st)
{
    {
        ProfScope ps(ST_LUT, st);
        build_lut_kernel<<<B, 1024, 0, st>>>(ws, ws_stride, L.status, L.votes, L.areas, L.lut, L.clut, k_cap, k_fixed,
                                             k_dev, things, label_divisor, stuff_area, void_label);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    const int vec = aligned16(pan_out) && (n_px % 8 == 0);
    size_t blocks = (n_px / 8 + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, 1, B);
    ProfScope ps(ST_APPLY, st);
    if (L.code16)
        apply_lut_kernel<true><<<grid, 256, 0, st>>>(ws, ws_stride, L.codes, L.lut, L.clut,
                                                      reinterpret_cast<long long*>(pan_out), n_px, vec);
    else
        apply_lut_kernel<false><<<grid, 256, 0, st>>>(ws, ws_stride, L.codes, L.lut, L.clut,
                                                       reinterpret_cast<long long*>(pan_out), n_px, vec);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

int load_centers(const int64_t* ctr, int K, const WsLayout& L, char* ws, cudaStream_t st)
{
    if (K > 0) {
        load_centers_kernel<<<(K + 255) / 256, 256, 0, st>>>(ctr, K, reinterpret_cast<int2*>(ws + L.centers));
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
    return launch_centers(1, hm, H, W, threshold, nms_kernel, L, static_cast<char*>(ws), L.total, cap, ctr_out, cap, st);
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
    if ((rc = load_centers(ctr, K, L, static_cast<char*>(ws), st))) return rc;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.off = off; a.off_stride = (size_t)2 * H * W;
    a.out = ids_out; a.out_stride = (size_t)H * W;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    a.o_status = L.status; a.o_centers = L.centers; a.o_votes = L.votes; a.o_areas = L.areas;
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
    if ((rc = launch_centers(1, hm, h, w, threshold, nms_kernel, L, static_cast<char*>(ws), L.total, k_cap, nullptr, 0, st))) return rc;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.off = off; a.off_stride = (size_t)2 * h * w;
    a.out = ids_out; a.out_stride = (size_t)h * w;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    a.o_status = L.status; a.o_centers = L.centers; a.o_votes = L.votes; a.o_areas = L.areas;
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
    if ((rc = launch_centers(1, hm, H, W, threshold, nms_kernel, L, static_cast<char*>(ws), L.total, k_cap, ctr_out, cap, st))) return rc;
    AssignArgs a;
    memset(&a, 0, sizeof(a));
    a.sem = sem; a.sem_stride = (size_t)H * W;
    a.off = off; a.off_stride = (size_t)2 * H * W;
    a.out = ins_out; a.out_stride = (size_t)H * W;
    a.ws = static_cast<char*>(ws); a.ws_stride = L.total;
    a.o_status = L.status; a.o_centers = L.centers; a.o_votes = L.votes; a.o_areas = L.areas;
    a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
    a.things = th;
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
    a.o_status = L.status; a.o_centers = L.centers; a.o_votes = L.votes; a.o_areas = L.areas;
    a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = k_cap;
    a.max_id = max_id; a.things = th;
    a.vec = (W % 4 == 0) && aligned16(sem) && (id_mode != ID_DENSE || aligned16(ids_in));
    if ((rc = launch_assign(1, sem_mode, id_mode, L.code16 ? OUT_CODE16 : OUT_CODE32, a, st))) return rc;
    return launch_lut_and_apply(1, L, static_cast<char*>(ws), L.total, k_cap, max_id, k_dev, th, label_divisor,
                                stuff_area, void_label, pan_out, (size_t)H * W, st);
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
    if ((rc = launch_centers(B, hm, H, W, threshold, nms_kernel, L, wsb, ws_bytes_per_tile, k_cap, ctr_out, cap, st))) return rc;

    const size_t sem_elt = sem_u8 ? 1 : 8;
    // assign -> lut -> apply tile by tile so that a tile's code map (2 B/px) is still in L2
    // when apply_lut reads it back.
    for (int b = 0; b < B; ++b) {
        char* wst = wsb + (size_t)b * ws_bytes_per_tile;
        AssignArgs a;
        memset(&a, 0, sizeof(a));
        a.sem = static_cast<const char*>(sem) + (size_t)b * n_px * sem_elt; a.sem_stride = n_px;
        a.off = off + (size_t)b * 2 * n_px; a.off_stride = 2 * n_px;
        a.out = wst + L.codes; a.out_stride = 0;
        a.ws = wst; a.ws_stride = ws_bytes_per_tile;
        a.o_status = L.status; a.o_centers = L.centers; a.o_votes = L.votes; a.o_areas = L.areas;
        a.H = H; a.W = W; a.step = 1.0f; a.chunksize = 20; a.k_cap = k_cap; a.k_fixed = -1;
        a.things = th;
        a.vec = (W % 4 == 0) && aligned16(a.sem) && aligned16(a.off);
        if ((rc = launch_assign(1, sem_u8 ? SEM_U8 : SEM_I64, ID_ARGMIN, L.code16 ? OUT_CODE16 : OUT_CODE32, a, st))) return rc;
        if ((rc = launch_lut_and_apply(1, L, wst, ws_bytes_per_tile, k_cap, -1, nullptr, th, label_divisor, stuff_area,
                                       void_label, pan_out + (size_t)b * n_px, n_px, st))) return rc;
    }
    return EMP_OK;
}
