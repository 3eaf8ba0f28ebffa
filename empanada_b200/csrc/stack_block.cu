// stack_block.cu — the 3D stack path, a z-block at a time (one launch per kernel for B slices).
//
//   emp_median_chain         _MedianQueue + _harden_seg over a whole z-block (engines.py:47-90, :114-121): every thread
//                            owns 4 pixels and walks the block in z with the recursion's state in registers — each raw
//                            probability is read ONCE (4 B/voxel) and only the hardened class byte is written (1 B/voxel)
//   emp_median_chain_repair  the same block re-run from a corrected carry, per pixel only as far as the two chains differ
//                            (the z-sharded stack starts every rank's chain from a guessed carry: inference/stack.py)
//   emp_stack_block          hardened classes + coarse heat-maps / offsets of B slices -> RLE tables, packed for one D2H:
//                              coarse centers + nearest-center ids          (panoptic.cu, batched; engines.py:257-272)
//                              upsample-fused merge -> 16-bit code map, strip flags, label LUT
//                                                                           (panoptic.cu, batched; postprocess.py:253-294)
//                              rle_block_keys   label LUT -> 32-bit run keys (class index, label - class base)
//                              rle_block_mark   code map + strip flags -> run start / end bit masks (rle.py:57-71)
//                              rle_block_emit   masks -> row-runs in raster order
//                              rle_block_runs   one CTA per slice: 8-connected union-find over row-runs, instance slots
//                                               in the reference's dict order, per-instance run lists merged across row
//                                               ends (array_utils.rle_encode :209-235), boxes, areas
//                              rle_block_pack   instance tables of all slices packed behind the run lists
//                            The int64 label map the reference materialises between merge and RLE (8 B/px written, 8 B/px
//                            read back) never exists: runs are cut from the 2 B/px code map, and only for strips that
//                            hold something other than class-0 background.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "rle_common.cuh"

namespace emp {

// =====================================================================================================
// 1. recursive median chain
// =====================================================================================================
constexpr unsigned kNoVote = 0xFFFFFFFFu;
#ifndef EMP_CHAIN_PF
#define EMP_CHAIN_PF 3
#endif
constexpr int kChainPf = EMP_CHAIN_PF;   // raw planes loaded ahead of the window (loads in flight per thread)

struct ChainArgs {
    const float* const* planes;     // device array: planes[j] = raw (C,H,W) probabilities of slice z0 + j, j < n_planes
    const float* const* carry_in;   // device array of mid pointers: filtered planes z0-mid .. z0-1 (null: none)
    const float* const* carry_new;  // repair: the corrected carry
    float* const* carry_out;        // device array of mid pointers (null: not wanted): filtered planes z0+n-mid .. z0+n-1
    int n, n_planes, z0, depth, c;
    size_t hw;                      // elements per channel plane
    float thr;
    unsigned char* sem8; size_t sem8_stride;
    float* best;                    // C > 1: running maximum over the channels seen so far, [n][hw]
    int multi;                      // C > 1
    int32_t* changed;               // repair: set to 1 if the block's outgoing carry changed
    unsigned char* need; size_t need_stride;        // optional (n, need_stride) coarse maps: 1 where a cell holds a thing pixel
    int W, shift, wc;               // plane width, log2 of the cell size, cells per coarse row
    unsigned long long thing_bits;
};

// min / max that return NaN when either operand is NaN (fminf / fmaxf drop it): with these a sorting network carries
// a NaN to every output it can reach — for an odd-even transposition network of KS passes that includes the middle
// one — which is torch.median's "a window holding a NaN gives NaN"
__device__ __forceinline__ float min_nan(float a, float b)
{
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

__device__ __forceinline__ float max_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// middle order statistic of KS values; NaN if any of them is NaN
template <int KS>
__device__ __forceinline__ float median_nan(const float (&w)[KS])
{
    if constexpr (KS == 3) {
        return max_nan(min_nan(w[0], w[1]), min_nan(max_nan(w[0], w[1]), w[2]));
    } else {
        float v[KS];
#pragma unroll
        for (int i = 0; i < KS; ++i) v[i] = w[i];
#pragma unroll
        for (int pass = 0; pass < KS; ++pass) {
#pragma unroll
            for (int i = (pass & 1); i + 1 < KS; i += 2) {
                const float lo = min_nan(v[i], v[i + 1]);
                const float hi = max_nan(v[i], v[i + 1]);
                v[i] = lo; v[i + 1] = hi;
            }
        }
        return v[(KS - 1) / 2];
    }
}

template <int N>
__device__ __forceinline__ void load_px(float (&dst)[N], const float* p)
{
    if (N == 4) {
        const float4 u = __ldcs(reinterpret_cast<const float4*>(p));
        dst[0] = u.x; dst[1 % N] = u.y; dst[2 % N] = u.z; dst[3 % N] = u.w;
    } else if (N == 2) {
        const float2 u = __ldcs(reinterpret_cast<const float2*>(p));
        dst[0] = u.x; dst[1 % N] = u.y;
    } else {
        dst[0] = __ldcs(p);
    }
}

template <int N>
__device__ __forceinline__ void store_px(float* p, const float (&src)[N])
{
    if (N == 4) *reinterpret_cast<float4*>(p) = make_float4(src[0], src[1 % N], src[2 % N], src[3 % N]);
    else if (N == 2) *reinterpret_cast<float2*>(p) = make_float2(src[0], src[1 % N]);
    else p[0] = src[0];
}

// One thread per N consecutive pixels of one channel; the z loop keeps the last MID filtered values and the raw
// window (+ kChainPf planes of look-ahead) in registers — the window as a ring whose slots are named at compile time
// (the loop is unrolled by the ring's length), so advancing costs no moves.  REPAIR runs the chain from two carries at
// once and stops as soon as their states agree bit for bit (from there on the outputs are identical by construction).
// Small CTAs: every thread lives for the whole z loop, so the grid drains in waves, and 64-thread CTAs make the waves
// finer (measured on 512 x 2048^2: 2.41 ms with 256 threads, 2.31 with 64 or 32; forcing 48 registers, a pointer
// look-ahead or a deeper prefetch change nothing: the 512-plane access pattern holds DRAM at ~4.6 TB/s).
#ifndef EMP_CHAIN_THREADS
#define EMP_CHAIN_THREADS 64
#endif
#ifndef EMP_CHAIN_MIN_CTAS
#define EMP_CHAIN_MIN_CTAS 1
#endif
constexpr int kChainThreads = EMP_CHAIN_THREADS;
template <int KS, int N, bool REPAIR>
__global__ void __launch_bounds__(kChainThreads, EMP_CHAIN_MIN_CTAS)
median_chain_kernel(const ChainArgs a)
{
    constexpr int MID = KS / 2, MS = MID > 0 ? MID : 1, WIN = MID + 1 + kChainPf;
    const size_t items = a.hw / N;
    const size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= items) return;
    const size_t px = it * N;
    const size_t e = (size_t)a.c * a.hw + px;
    int cell[N];                                                    // coarse cell of each pixel (need map)
    if (a.need) {
        const int y = (int)(px / (size_t)a.W), x = (int)(px - (size_t)y * a.W);
#pragma unroll
        for (int q = 0; q < N; ++q) {                                // N > 1 only with hw % N == 0; pixels may wrap into the next row
            const int xx = x + q;
            cell[q] = ((y + xx / a.W) >> a.shift) * a.wc + ((xx % a.W) >> a.shift);
        }
    }
    float f[MS][N], g[MS][N], r[WIN][N];
#pragma unroll
    for (int j = 0; j < MS; ++j)
#pragma unroll
        for (int q = 0; q < N; ++q) { f[j][q] = 0.f; g[j][q] = 0.f; }
    if (MID > 0 && a.carry_in) {
#pragma unroll
        for (int j = 0; j < MID; ++j) load_px<N>(f[j], a.carry_in[j] + e);
    }
    if (REPAIR) {
#pragma unroll
        for (int j = 0; j < MID; ++j) load_px<N>(g[j], a.carry_new[j] + e);
    }
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
#pragma unroll
        for (int q = 0; q < N; ++q) r[j][q] = 0.f;
        if (j < a.n_planes) load_px<N>(r[j], a.planes[j] + e);
    }
    bool differs = REPAIR;
    bool done = false;
#pragma unroll 1
    for (int i0 = 0; i0 < a.n && !done; i0 += WIN) {
#pragma unroll
        for (int u = 0; u < WIN; ++u) {                             // slice i0 + u: its window is ring slots u, u+1, ... (mod WIN)
            const int i = i0 + u;
            if (i >= a.n) { done = true; break; }
            if (REPAIR) {
                differs = false;
#pragma unroll
                for (int j = 0; j < MID; ++j)
#pragma unroll
                    for (int q = 0; q < N; ++q) differs |= __float_as_uint(f[j][q]) != __float_as_uint(g[j][q]);
                if (!differs) { done = true; break; }
            }
            const int z = a.z0 + i;
            const bool raw = (z < MID) || (z >= a.depth - MID);    // the queue passes the stack's ends through unfiltered
            float o[N], og[N];
#pragma unroll
            for (int q = 0; q < N; ++q) {
                o[q] = r[u][q]; og[q] = r[u][q];
                if (KS > 1 && !raw) {
                    float w[KS];
#pragma unroll
                    for (int j = 0; j < MID; ++j) w[j] = f[j][q];
#pragma unroll
                    for (int j = 0; j <= MID; ++j) w[MID + j] = r[(u + j) % WIN][q];
                    o[q] = median_nan<KS>(w);
                    if (REPAIR) {
#pragma unroll
                        for (int j = 0; j < MID; ++j) w[j] = g[j][q];
                        og[q] = median_nan<KS>(w);
                    }
                }
            }
            const float (&res)[N] = REPAIR ? og : o;                // the values that count
            unsigned char* sp = a.sem8 + (size_t)i * a.sem8_stride + px;
            unsigned set = 0;                                       // bit q: pixel q's class was (re)written by this launch
            unsigned bits = 0;                                      // class bytes
            if (!a.multi) {
#pragma unroll
                for (int q = 0; q < N; ++q) bits |= (res[q] >= a.thr ? 1u : 0u) << (8 * q);
                if (N == 4) *reinterpret_cast<unsigned*>(sp) = bits;
                else if (N == 2) *reinterpret_cast<unsigned short*>(sp) = (unsigned short)bits;
                else sp[0] = (unsigned char)bits;
                set = (1u << N) - 1u;
            } else {                                                // first arg-max over channels, NaN counts as the maximum
                float* bp = a.best + (size_t)i * a.hw + px;
                if (a.c == 0) {
                    store_px<N>(bp, res);
#pragma unroll
                    for (int q = 0; q < N; ++q) sp[q] = 0;
                    set = (1u << N) - 1u;
                } else {
                    float b[N];
                    load_px<N>(b, bp);
#pragma unroll
                    for (int q = 0; q < N; ++q)
                        if (res[q] > b[q] || (res[q] != res[q] && b[q] == b[q])) {
                            bp[q] = res[q]; sp[q] = (unsigned char)a.c;
                            set |= 1u << q; bits |= (unsigned)a.c << (8 * q);
                        }
                }
            }
            if (a.need && (bits != 0u || (a.thing_bits & 1ull))) {  // a superset is fine: a stale 1 only costs an unused id
                unsigned char* np = a.need + (size_t)i * a.need_stride;
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    const unsigned cls = (bits >> (8 * q)) & 0xFFu;
                    if (((set >> q) & 1u) && cls < 64u && ((a.thing_bits >> cls) & 1ull)) np[cell[q]] = 1;
                }
            }
#pragma unroll
            for (int q = 0; q < N; ++q) {
#pragma unroll
                for (int j = 0; j + 1 < MID; ++j) { f[j][q] = f[j + 1][q]; g[j][q] = g[j + 1][q]; }
                if constexpr (MID > 0) { f[MID - 1][q] = o[q]; g[MID - 1][q] = og[q]; }
            }
            if (i + WIN < a.n_planes) load_px<N>(r[u], a.planes[i + WIN] + e);     // slot u is free: it takes slice i + WIN
        }
    }
    if (REPAIR && differs) {                                        // ran to the block's end: are the final states still apart?
        differs = false;
#pragma unroll
        for (int j = 0; j < MID; ++j)
#pragma unroll
            for (int q = 0; q < N; ++q) differs |= __float_as_uint(f[j][q]) != __float_as_uint(g[j][q]);
    }
    if (MID > 0 && a.carry_out) {
        if (!REPAIR) {
#pragma unroll
            for (int j = 0; j < MID; ++j) store_px<N>(a.carry_out[j] + e, f[j]);
        } else if (differs) {                                       // still apart at the block's end: the carry moves on
#pragma unroll
            for (int j = 0; j < MID; ++j) store_px<N>(a.carry_out[j] + e, g[j]);
            *a.changed = 1;
        }
    }
}

template <int KS, bool REPAIR>
static int launch_chain_ks(const ChainArgs& a, bool vec, cudaStream_t st)
{
    static int width = -1;                  // pixels per thread of the vector path (tuning knob: EMP_CHAIN_N = 2 | 4)
    if (width < 0) {
        const char* e = getenv("EMP_CHAIN_N");
        width = (e && atoi(e) == 2) ? 2 : 4;
    }
    const int N = vec ? width : 1;
    const size_t items = a.hw / N;
    const unsigned grid = (unsigned)((items + kChainThreads - 1) / kChainThreads);
    ProfScope ps(ST_CHAIN, st);
    if (N == 4) median_chain_kernel<KS, 4, REPAIR><<<grid, kChainThreads, 0, st>>>(a);
    else if (N == 2) median_chain_kernel<KS, 2, REPAIR><<<grid, kChainThreads, 0, st>>>(a);
    else median_chain_kernel<KS, 1, REPAIR><<<grid, kChainThreads, 0, st>>>(a);
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

template <bool REPAIR>
static int launch_chain(int ks, const ChainArgs& a, bool vec, cudaStream_t st)
{
    switch (ks) {
        case 1:  return REPAIR ? EMP_OK : launch_chain_ks<1, false>(a, vec, st);
        case 3:  return launch_chain_ks<3, REPAIR>(a, vec, st);
        case 5:  return launch_chain_ks<5, REPAIR>(a, vec, st);
        case 7:  return launch_chain_ks<7, REPAIR>(a, vec, st);
        case 9:  return launch_chain_ks<9, REPAIR>(a, vec, st);
        case 11: return launch_chain_ks<11, REPAIR>(a, vec, st);
        case 13: return launch_chain_ks<13, REPAIR>(a, vec, st);
        default: return launch_chain_ks<15, REPAIR>(a, vec, st);
    }
}

// =====================================================================================================
// 2. upsample-fused merge for uint8 class maps + coarse id maps (the stack path's common case)
// =====================================================================================================
// get_panoptic_seg (engines.py:277-292) up to the label LUT, as the general assign kernel of panoptic.cu does it for
// (SEM_U8, ID_COARSE), but without anything the nearest-center search needs: ~60 registers, 8 CTAs per SM, and a
// background strip costs four 16-byte loads and one flag byte.  A warp takes a strip row (4 rows) x 512 columns = 8
// strips of 4 x 64; four lanes share a strip, each owning 16 consecutive pixels of its 4 rows.  Same outputs as the
// general kernel: strip flags, 16-bit codes (only for strips that are not all class-0 background), votes per
// (id, thing class), stuff areas by complement.  Needs W % 16 == 0 and thing classes < 64 (else the general kernel).
struct LeanArgs {
    const unsigned char* sem8; size_t sem_stride;
    const int32_t* ids; size_t ids_stride;
    char* ws; size_t ws_stride;
    size_t o_votes, o_areas, o_sflags, o_codes;
    int B, H, W, wc, shift, blocks_x, blk_shift, T;
    unsigned long long thing_bits;
    int simple;                 // one thing class (thing_class, != 0) and cells of >= 4 pixels: the byte-mask path applies
    unsigned thing_class;
    // mark != 0: also write the RLE encoder's run start / end masks and row counts (possible when the void label is not
    // a selected label: then distinct codes other than void / class-0 background never share a label, so run boundaries
    // are code boundaries; runs of codes whose label turns out not to be selected are dropped at run level)
    int mark;
    char* rs; size_t rs_stride; size_t o_smask, o_emask, o_rowcnt;
    int wd, crop_h, crop_w;
};

// the code merge_lean gives pixel (y, x): instance id (0: none -> void) on thing pixels, class code elsewhere
__device__ __forceinline__ unsigned lean_code_at(const LeanArgs& a, const unsigned char* plane, const int32_t* ids, int y, int x)
{
    const unsigned c = plane[(size_t)y * a.W + x];
    if (c < 64u && ((a.thing_bits >> c) & 1ull)) return (unsigned)__ldg(ids + (y >> a.shift) * a.wc + (x >> a.shift));
    return kClsBase16 + c;
}

#ifndef EMP_LEAN_WARPS
#define EMP_LEAN_WARPS 2
#endif
constexpr int kLeanWarps = EMP_LEAN_WARPS;              // warps per CTA: each takes one item, a CTA lives as long as its slowest warp
__global__ void __launch_bounds__(kLeanWarps * 32, 32 / kLeanWarps)
merge_lean_kernel(const LeanArgs a)
{
    const int lane = threadIdx.x & 31, grp = lane >> 2, q = lane & 3;
    const unsigned groups = (a.W + 511) / 512, srows = (a.H + 3) / 4;
    const unsigned item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one item per warp, one slice per grid plane
    if (item >= srows * groups) return;                                 // warp-uniform
    const bool class0_stuff = !(a.thing_bits & 1ull);
    const bool multi = a.T > 1;
    unsigned deficit = 0;                                               // in-image pixels of the item that are NOT class-0 stuff
    const int b = blockIdx.z;
    char* ws = a.ws + (size_t)b * a.ws_stride;
    const int sy = (int)(item / groups), g = (int)(item - (unsigned)sy * groups);
    {
        // ---- streaming part: 4 rows x 512 columns, 16 bytes per lane and row; which of the 8 strips are pure background?
        const int y0 = sy * 4, xs = g * 512 + grp * 64;                 // the strip of this lane's group
        const int x0 = xs + q * 16;
        const unsigned char* plane = a.sem8 + (size_t)b * a.sem_stride;
        unsigned orall = 0;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            if (x0 < a.W && y0 + rr < a.H) {                            // W % 16 == 0: a lane is inside or outside as a whole
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(plane + (size_t)(y0 + rr) * a.W + x0));
                orall |= u.x | u.y | u.z | u.w;
            }
        }
        const unsigned zero4 = __ballot_sync(0xffffffffu, orall == 0u);
        const bool full = xs + 64 <= a.W && y0 + 4 <= a.H;
        const unsigned gmask = 0xFu << (lane & ~3);
        const bool bg = full && class0_stuff && ((zero4 & gmask) == gmask);
        if (q == 0 && xs < a.W)
            reinterpret_cast<unsigned char*>(ws + a.o_sflags)[((size_t)(y0 >> (a.blk_shift + 2)) * a.blocks_x + (xs >> 6)) * 16 +
                                                              ((y0 >> 2) & ((1 << a.blk_shift) - 1))] = bg ? 1 : 0;
        unsigned todo = __ballot_sync(0xffffffffu, q == 0 && xs < a.W && !bg);
        // ---- strips that hold something: the whole warp takes one strip at a time, lane l 8 pixels of row l / 8
        uint32_t* votes = reinterpret_cast<uint32_t*>(ws + a.o_votes);
        uint32_t* areas = reinterpret_cast<uint32_t*>(ws + a.o_areas);
        const int32_t* ids = a.ids + (size_t)b * a.ids_stride;
        unsigned short* codes = reinterpret_cast<unsigned short*>(ws + a.o_codes);
        for (; todo; todo &= todo - 1) {                                // warp-uniform trip count, no lane leaves early
            const int sx = g * 512 + ((__ffs(todo) - 1) >> 2) * 64;     // first column of the strip
            const int y = y0 + (lane >> 3), x = sx + (lane & 7) * 8;
            const bool valid = y < a.H && x < a.W;                      // lanes past the plane's edge (partial strips only)
            uint2 u = make_uint2(0u, 0u);
            if (valid) u = __ldg(reinterpret_cast<const uint2*>(plane + (size_t)y * a.W + x));     // just read: an L1 hit
            const int crow = (y >> a.shift) * a.wc;
            unsigned out[4];
            unsigned k1 = kNoVote, c1 = 0, k2 = kNoVote, c2 = 0;        // up to two (vote key, count) pairs leave the lane aggregated
            const unsigned tcx4 = a.thing_class * 0x01010101u;
            const unsigned t0 = __vcmpeq4(u.x, tcx4), t1 = __vcmpeq4(u.y, tcx4);          // 0xFF per byte of the (one) thing class
            const unsigned z0 = __vcmpeq4(u.x, 0u), z1 = __vcmpeq4(u.y, 0u);              // 0xFF per class-0 byte
            if (a.simple && ((t0 | z0) & (t1 | z1)) == 0xFFFFFFFFu) {
                // ---- one thing class, cells of >= 4 aligned pixels, nothing but that class and class 0 here: a word of 4
                // pixels meets one cell; codes and votes come from byte masks
                const unsigned tm[2] = {t0, t1};
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    unsigned id = 0;
                    if (tm[k]) id = (unsigned)__ldg(ids + crow + ((x + 4 * k) >> a.shift));
                    const unsigned n_thing = (unsigned)__popc(tm[k]) >> 3;
                    deficit += n_thing;
                    const unsigned idid = id | (id << 16);
                    const unsigned bg2 = kClsBase16 | (kClsBase16 << 16);
                    const unsigned m_lo = __byte_perm(tm[k], 0u, 0x1100), m_hi = __byte_perm(tm[k], 0u, 0x3322);
                    out[2 * k] = (m_lo & idid) | (~m_lo & bg2);          // a thing pixel without an instance (id 0) stays void
                    out[2 * k + 1] = (m_hi & idid) | (~m_hi & bg2);
                    if (id > 0u && n_thing) {
                        if (k == 0 || k1 == kNoVote) { k1 = id; c1 = n_thing; }              // T == 1: key = id
                        else if (k1 == id) c1 += n_thing;
                        else { k2 = id; c2 = n_thing; }
                    }
                }
            } else if (valid) {
                // ---- anything else: pixel by pixel
                const unsigned w2[2] = {u.x, u.y};
                unsigned vkey = kNoVote, vcnt = 0;                      // pending vote: neighbours of equal (id, class)
                unsigned akey = kNoVote, acnt = 0;                      // pending stuff-area count of one non-zero class
                int last_cell = -1, last_id = 0;
                auto push_vote = [&]() {
                    if (!vcnt) return;
                    if (k1 == kNoVote) { k1 = vkey; c1 = vcnt; }
                    else if (k1 == vkey) c1 += vcnt;
                    else if (k2 == kNoVote) { k2 = vkey; c2 = vcnt; }
                    else if (k2 == vkey) c2 += vcnt;
                    else atomicAdd(votes + vkey, vcnt);
                };
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const unsigned c = (w2[p >> 2] >> (8 * (p & 3))) & 0xFFu;
                    const bool thing = c < 64u && ((a.thing_bits >> c) & 1ull);
                    unsigned code;
                    if (thing) {
                        const int cell = crow + ((x + p) >> a.shift);
                        if (cell != last_cell) { last_cell = cell; last_id = __ldg(ids + cell); }
                        code = (unsigned)last_id;                       // 0: a thing pixel without an instance stays void
                        if (last_id > 0) {
                            const unsigned t = multi ? (unsigned)__popcll(a.thing_bits & ((1ull << c) - 1ull)) : 0u;
                            const unsigned key = (unsigned)last_id * (unsigned)a.T + t;
                            if (key != vkey) { push_vote(); vkey = key; vcnt = 0; }
                            ++vcnt;
                        }
                        ++deficit;
                    } else {
                        code = kClsBase16 + c;
                        if (c != 0u) {
                            ++deficit;
                            if (c != akey) {
                                if (acnt) atomicAdd(areas + akey, acnt);
                                akey = c; acnt = 0;
                            }
                            ++acnt;
                        }
                    }
                    if (p & 1) out[p >> 1] |= code << 16; else out[p >> 1] = code;
                }
                push_vote();
                if (acnt) atomicAdd(areas + akey, acnt);
            }
            if (valid) *reinterpret_cast<uint4*>(codes + (size_t)y * a.W + x) = make_uint4(out[0], out[1], out[2], out[3]);
            if (a.mark) {
                // ---- run boundaries of this strip row: a pixel starts (ends) a run if its code is neither void nor class-0
                // background and differs from its left (right) neighbour's; pixels outside the crop count as background
                const bool row_in = valid && y < a.crop_h;
                unsigned c8[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const unsigned v = (out[p >> 1] >> (16 * (p & 1))) & 0xFFFFu;
                    c8[p] = (row_in && x + p < a.crop_w) ? v : kClsBase16;
                }
                unsigned left = __shfl_up_sync(0xffffffffu, c8[7], 1);
                unsigned right = __shfl_down_sync(0xffffffffu, c8[0], 1);
                if ((lane & 7) == 0) left = (row_in && x > 0) ? lean_code_at(a, plane, ids, y, x - 1) : kClsBase16;
                if ((lane & 7) == 7) right = (row_in && x + 8 < a.crop_w) ? lean_code_at(a, plane, ids, y, x + 8) : kClsBase16;
                unsigned sb = 0, eb = 0;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const unsigned l = p ? c8[p - 1] : left, rn = p < 7 ? c8[p + 1] : right;
                    const bool sel = c8[p] != 0u && c8[p] != kClsBase16;
                    sb |= (sel && l != c8[p] ? 1u : 0u) << p;
                    eb |= (sel && rn != c8[p] ? 1u : 0u) << p;
                }
                const unsigned mine = sb | (eb << 8);
                const unsigned m1 = __shfl_down_sync(0xffffffffu, mine, 1), m2 = __shfl_down_sync(0xffffffffu, mine, 2),
                               m3 = __shfl_down_sync(0xffffffffu, mine, 3);
                unsigned cnt = 0;
                if ((lane & 3) == 0 && row_in) {
                    const unsigned sw = (mine & 0xFFu) | ((m1 & 0xFFu) << 8) | ((m2 & 0xFFu) << 16) | ((m3 & 0xFFu) << 24);
                    const unsigned ew = ((mine >> 8) & 0xFFu) | (((m1 >> 8) & 0xFFu) << 8) | (((m2 >> 8) & 0xFFu) << 16) | (((m3 >> 8) & 0xFFu) << 24);
                    const int wi = (x >> 5);
                    if (wi < a.wd && (sw | ew)) {
                        char* rsb = a.rs + (size_t)b * a.rs_stride;
                        reinterpret_cast<uint32_t*>(rsb + a.o_smask)[(size_t)y * a.wd + wi] = sw;
                        reinterpret_cast<uint32_t*>(rsb + a.o_emask)[(size_t)y * a.wd + wi] = ew;
                    }
                    cnt = (unsigned)__popc(sw);
                }
                cnt += __shfl_down_sync(0xffffffffu, cnt, 4);           // the row's two words
                if ((lane & 7) == 0 && cnt)
                    atomicAdd(reinterpret_cast<uint32_t*>(a.rs + (size_t)b * a.rs_stride + a.o_rowcnt) + y, cnt);
            }
            // votes: lanes holding the same key (the rows of a strip share their cells) add up first
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                const unsigned key = sl ? k2 : k1, cnt = sl ? c2 : c1;
                if (!__any_sync(0xffffffffu, key != kNoVote)) continue; // warp-uniform
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const unsigned sum = __reduce_add_sync(peers, cnt);
                if (key != kNoVote && lane == __ffs(peers) - 1) atomicAdd(votes + key, sum);
            }
        }
    }
    const unsigned d = __reduce_add_sync(0xffffffffu, deficit);
    if (d && lane == 0) atomicAdd(reinterpret_cast<uint32_t*>(ws + a.o_areas) + kNumClasses, d);
}

// =====================================================================================================
// 3. RLE tables from code maps
// =====================================================================================================
// per-slice scratch of the encoder (B of them, rs_stride apart)
struct BlkLayout {
    size_t status, rowcnt, flags, cnt, smask, emask, zero_bytes;    // [0, zero_bytes) is cleared per call
    size_t keylut, rowoff, r_y, r_xs, r_xe, r_key, parent, slot_of, ymin, ymax, inst, total;
    size_t flags_len;
    int wd;
};

static BlkLayout blk_layout(int crop_h, int crop_w, int run_cap, int inst_cap, int n_labels, long long L, int k_cap)
{
    BlkLayout R;
    R.wd = (crop_w + 31) / 32;
    if (n_labels < 1) n_labels = 1;
    // key space per class: row-runs for CCL classes, label - base (<= number of centers) otherwise
    const size_t key_space = (size_t)std::max<long long>((long long)run_cap, std::min<long long>(L, (long long)k_cap + 1));
    R.flags_len = key_space * (size_t)n_labels + 1;
    size_t o = 0;
    R.status = o;  o = align_up(o + sizeof(int32_t) * EMP_ST_WORDS, 256);
    R.rowcnt = o;  o = align_up(o + sizeof(uint32_t) * (size_t)crop_h, 256);
    R.flags = o;   o = align_up(o + sizeof(int) * R.flags_len, 256);
    R.cnt = o;     o = align_up(o + sizeof(int) * ((size_t)inst_cap + 1), 256);
    R.smask = o;   o = align_up(o + sizeof(uint32_t) * (size_t)crop_h * R.wd, 256);     // mark only writes words that hold a run boundary
    R.emask = o;   o = align_up(o + sizeof(uint32_t) * (size_t)crop_h * R.wd, 256);
    R.zero_bytes = o;
    R.keylut = o;  o = align_up(o + sizeof(uint32_t) * ((size_t)k_cap + 1 + kNumClasses), 256);
    R.rowoff = o;  o = align_up(o + sizeof(int) * ((size_t)crop_h + 1), 256);
    R.r_y = o;     o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.r_xs = o;    o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.r_xe = o;    o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.r_key = o;   o = align_up(o + sizeof(uint32_t) * (size_t)run_cap, 256);
    R.parent = o;  o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.slot_of = o; o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.ymin = o;    o = align_up(o + sizeof(int) * (size_t)inst_cap, 256);
    R.ymax = o;    o = align_up(o + sizeof(int) * (size_t)inst_cap, 256);
    R.inst = o;    o = align_up(o + sizeof(long long) * EMP_BLK_INST_WORDS * (size_t)inst_cap, 256);
    R.total = o;
    return R;
}

struct BlkArgs {
    char* ws; size_t ws_stride;             // merge workspaces (ws_layout): codes, strip flags, label LUT, status
    size_t o_codes, o_sflags, o_lut, o_status;
    char* cs; size_t cs_stride; size_t o_cstatus;   // coarse workspaces: status (center count K)
    char* rs; size_t rs_stride;             // encoder scratch (BlkLayout)
    BlkLayout R;
    RleClasses rc;
    int B, W, crop_h, crop_w;               // W: row pitch of the code map (the padded plane)
    int blocks_x, blk_shift;                // strip-flag geometry of the assign kernel (1 << blk_shift strips per block)
    unsigned cls_off;
    int k_cap, run_cap, inst_cap;
    int vec;                                // 16-byte code loads allowed
    long long* packed;                      // EMP_BLK_* layout
    long long* runs3; size_t runs3_stride;  // optional (start, length, slot) row-runs per slice, int64 triples
    long long* maxlab_all;                  // optional: per-class maxima accumulated over the blocks of a z-block
};

// label LUT -> run keys: 0 if the label belongs to no selected class, else (class index + 1) << 22 | (label - base)
__global__ void __launch_bounds__(256)
rle_block_keys_kernel(const BlkArgs a)
{
    const int b = blockIdx.z;
    const char* ws = a.ws + (size_t)b * a.ws_stride;
    const long long* lut = reinterpret_cast<const long long*>(ws + a.o_lut);
    uint32_t* keylut = reinterpret_cast<uint32_t*>(a.rs + (size_t)b * a.rs_stride + a.R.keylut);
    const int K = min(max(__ldcg(reinterpret_cast<const int32_t*>(a.cs + (size_t)b * a.cs_stride + a.o_cstatus) + EMP_ST_K), 0), a.k_cap);
    const int n = K + 1 + kNumClasses;              // ids 0..K, then the class entries at cls_off
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int idx = i <= K ? i : (int)a.cls_off + (i - K - 1);
        keylut[idx] = run_key(lut[idx], a.rc);
    }
}

// run key of pixel (y, x) of slice-local planes: strip flag -> code -> key
struct CodeView {
    const unsigned short* codes;
    const unsigned char* sflags;
    const uint32_t* keylut;
    int W, blocks_x, blk_shift;      // strips per block of the assign geometry: 1 << blk_shift
    unsigned cls_off;
};

__device__ __forceinline__ unsigned code_index(unsigned code, unsigned cls_off)
{
    return code >= kClsBase16 ? code - kClsBase16 + cls_off : code;
}

__device__ __forceinline__ size_t flag_index(const CodeView& v, int y, int x)
{
    return ((size_t)(y >> (v.blk_shift + 2)) * v.blocks_x + (x >> 6)) * 16 + ((y >> 2) & ((1 << v.blk_shift) - 1));
}

__device__ __forceinline__ unsigned key_at(const CodeView& v, int y, int x)
{
    const unsigned code = v.sflags[flag_index(v, y, x)] ? kClsBase16 : v.codes[(size_t)y * v.W + x];
    return __ldg(v.keylut + code_index(code, v.cls_off));
}

__device__ __forceinline__ CodeView code_view(const BlkArgs& a, int b)
{
    CodeView v;
    const char* ws = a.ws + (size_t)b * a.ws_stride;
    v.codes = reinterpret_cast<const unsigned short*>(ws + a.o_codes);
    v.sflags = reinterpret_cast<const unsigned char*>(ws + a.o_sflags);
    v.keylut = reinterpret_cast<const uint32_t*>(a.rs + (size_t)b * a.rs_stride + a.R.keylut);
    v.W = a.W; v.blocks_x = a.blocks_x; v.blk_shift = a.blk_shift; v.cls_off = a.cls_off;
    return v;
}

// The one pass over the (cropped) code maps.  A warp takes a strip row (4 rows) x 256 columns; lane l owns pixels
// 8l .. 8l+7 of each row (one 16-byte load of codes per row — or none at all when its 4 x 64 strip is flagged "all
// class-0 background", the bulk of an EM slice: such an item costs one flag byte).  Start / end bits are formed per
// lane and gathered into the row's mask words with three shuffles; the masks are pre-zeroed, so only words of items
// that hold a selected pixel are written.
__global__ void __launch_bounds__(256)
rle_block_mark_kernel(const BlkArgs a)
{
    const int lane = threadIdx.x & 31;
    const unsigned groups = (a.crop_w + 255) / 256;
    const unsigned srows = (a.crop_h + 3) / 4;
    const unsigned item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one item per warp, one slice per grid plane
    if (item >= srows * groups) return;                                 // warp-uniform
    const int wd = a.R.wd;
    const int b = blockIdx.z;
    const CodeView v = code_view(a, b);
    char* rs = a.rs + (size_t)b * a.rs_stride;
    const unsigned bgkey = __ldg(v.keylut + a.cls_off);                 // class-0 background
    const unsigned key0 = __ldg(v.keylut);                              // void
    const int sy = (int)(item / groups), g = (int)(item - (unsigned)sy * groups);
    {
        const int y0 = sy * 4;
        const int xb = g * 256, x0 = xb + lane * 8;
        const bool inside = x0 < a.crop_w;
        // the pixel just outside the item on this lane's side (lanes 0 / 31): its key decides whether a run starts / ends here
        const int ex = lane == 0 ? xb - 1 : (lane == 31 ? xb + 256 : -1);
        const bool edge = ex >= 0 && ex < a.crop_w;
        // ---- round trip 1: strip flags
        const bool flagged = inside && v.sflags[flag_index(v, y0, x0)] != 0;
        const bool eflagged = edge && v.sflags[flag_index(v, y0, ex)] != 0;
        if (bgkey == 0u && __all_sync(0xffffffffu, flagged || !inside)) return;        // nothing selected in this item
        // ---- round trip 2: codes of the 4 rows (and of the edge pixels)
        uint4 raw[4];
        unsigned ecode[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            raw[rr] = make_uint4(0u, 0u, 0u, 0u);
            ecode[rr] = kClsBase16;
            const bool row_in = y0 + rr < a.crop_h;
            if (inside && !flagged && row_in && a.vec)
                raw[rr] = __ldcs(reinterpret_cast<const uint4*>(v.codes + (size_t)(y0 + rr) * a.W + x0));
            if (edge && !eflagged && row_in) ecode[rr] = v.codes[(size_t)(y0 + rr) * a.W + ex];
        }
        if (key0 == 0u) {
            // ---- common case (void is not a selected label): distinct codes never share a non-zero key, so run boundaries
            // are code boundaries.  Two codes per 32-bit word are compared with their neighbours by the SIMD halfword
            // instructions; keys are looked up only at boundary pixels, to drop runs of unselected codes.
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int y = y0 + rr;
                if (y >= a.crop_h) break;                               // warp-uniform
                unsigned w[4];
                if (flagged || !inside) {
                    w[0] = w[1] = w[2] = w[3] = kClsBase16 | (kClsBase16 << 16);       // outside: treated as background
                } else if (a.vec) {
                    w[0] = raw[rr].x; w[1] = raw[rr].y; w[2] = raw[rr].z; w[3] = raw[rr].w;
                } else {
                    const unsigned short* cp = v.codes + (size_t)y * a.W + x0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned c0 = (x0 + 2 * k < a.crop_w) ? cp[2 * k] : kClsBase16;
                        const unsigned c1 = (x0 + 2 * k + 1 < a.crop_w) ? cp[2 * k + 1] : kClsBase16;
                        w[k] = c0 | (c1 << 16);
                    }
                }
                if (inside && x0 + 8 > a.crop_w && a.vec) {             // the crop ends inside this lane's 8 pixels
#pragma unroll
                    for (int p = 0; p < 8; ++p)
                        if (x0 + p >= a.crop_w) w[p >> 1] = (p & 1) ? (w[p >> 1] & 0xFFFFu) | (kClsBase16 << 16) : (w[p >> 1] & 0xFFFF0000u) | kClsBase16;
                }
                unsigned left = __shfl_up_sync(0xffffffffu, w[3] >> 16, 1);
                unsigned right = __shfl_down_sync(0xffffffffu, w[0] & 0xFFFFu, 1);
                if (lane == 0) left = edge ? ecode[rr] : kClsBase16;    // outside the image: background, never selected
                if (lane == 31) right = edge ? ecode[rr] : kClsBase16;
                // a row segment of one code (the inside of an instance, background between instances) has no boundary
                const bool flat = w[0] == w[1] && w[1] == w[2] && w[2] == w[3] && w[0] == __funnelshift_l(w[0], w[0], 16) &&
                                  left == (w[0] & 0xFFFFu) && right == (w[0] & 0xFFFFu);
                if (__all_sync(0xffffffffu, flat)) continue;            // warp-uniform
                // bit p of sb / eb: pixel p differs from its left / right neighbour
                unsigned sb = 0, eb = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned prev = k ? __funnelshift_l(w[k - 1], w[k], 16) : ((w[0] << 16) | left);
                    const unsigned next = k < 3 ? __funnelshift_r(w[k], w[k + 1], 16) : ((w[3] >> 16) | (right << 16));
                    const unsigned dp = __vcmpne2(w[k], prev), dn = __vcmpne2(w[k], next);
                    sb |= ((dp & 1u) | ((dp >> 15) & 2u)) << (2 * k);
                    eb |= ((dn & 1u) | ((dn >> 15) & 2u)) << (2 * k);
                }
                for (unsigned m = sb | eb; m; m &= m - 1) {             // boundary pixels: keep those of selected codes
                    const int p = __ffs(m) - 1;
                    const unsigned wp = p < 4 ? (p < 2 ? w[0] : w[1]) : (p < 6 ? w[2] : w[3]);   // no dynamic register index
                    const unsigned code = (wp >> (16 * (p & 1))) & 0xFFFFu;
                    const unsigned key = code == 0u ? 0u : code == kClsBase16 ? bgkey : __ldg(v.keylut + code_index(code, a.cls_off));
                    if (key == 0u) { sb &= ~(1u << p); eb &= ~(1u << p); }
                }
                if (!__any_sync(0xffffffffu, (sb | eb) != 0u)) continue;               // warp-uniform
                const unsigned mine = sb | (eb << 8);
                const unsigned m1 = __shfl_down_sync(0xffffffffu, mine, 1), m2 = __shfl_down_sync(0xffffffffu, mine, 2),
                               m3 = __shfl_down_sync(0xffffffffu, mine, 3);
                unsigned cnt = 0;
                if ((lane & 3) == 0) {
                    const unsigned sw = (mine & 0xFFu) | ((m1 & 0xFFu) << 8) | ((m2 & 0xFFu) << 16) | ((m3 & 0xFFu) << 24);
                    const unsigned ew = ((mine >> 8) & 0xFFu) | (((m1 >> 8) & 0xFFu) << 8) | (((m2 >> 8) & 0xFFu) << 16) | (((m3 >> 8) & 0xFFu) << 24);
                    const int wi = g * 8 + (lane >> 2);
                    if (wi < wd && (sw | ew)) {
                        reinterpret_cast<uint32_t*>(rs + a.R.smask)[(size_t)y * wd + wi] = sw;
                        reinterpret_cast<uint32_t*>(rs + a.R.emask)[(size_t)y * wd + wi] = ew;
                    }
                    cnt = (unsigned)__popc(sw);
                }
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                if (lane == 0 && cnt) atomicAdd(reinterpret_cast<uint32_t*>(rs + a.R.rowcnt) + y, cnt);
            }
            return;
        }
        // ---- void is a selected label (it may share its key with a stuff class or an instance): compare keys, pixel by pixel
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int y = y0 + rr;
            if (y >= a.crop_h) break;                                   // warp-uniform
            unsigned key[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) key[p] = 0u;
            if (flagged) {
#pragma unroll
                for (int p = 0; p < 8; ++p) key[p] = (x0 + p < a.crop_w) ? bgkey : 0u;
            } else if (inside) {
                unsigned code[8];
                if (a.vec) {                                            // W % 8 == 0: the whole group lies inside the plane
                    const uint4 u = raw[rr];
                    code[0] = u.x & 0xFFFFu; code[1] = u.x >> 16; code[2] = u.y & 0xFFFFu; code[3] = u.y >> 16;
                    code[4] = u.z & 0xFFFFu; code[5] = u.z >> 16; code[6] = u.w & 0xFFFFu; code[7] = u.w >> 16;
                } else {
                    const unsigned short* cp = v.codes + (size_t)y * a.W + x0;
#pragma unroll
                    for (int p = 0; p < 8; ++p) code[p] = (x0 + p < a.crop_w) ? cp[p] : 0u;
                }
                unsigned prev_code = 0u, prev_key = key0;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    if (code[p] != prev_code) {
                        prev_code = code[p];
                        prev_key = code[p] == 0u ? key0 : code[p] == kClsBase16 ? bgkey : __ldg(v.keylut + code_index(code[p], a.cls_off));
                    }
                    key[p] = (x0 + p < a.crop_w) ? prev_key : 0u;
                }
            }
            unsigned ekey = 0u;
            if (edge) {
                const unsigned c = ecode[rr];
                ekey = c == 0u ? key0 : c == kClsBase16 ? bgkey : __ldg(v.keylut + code_index(c, a.cls_off));
            }
            unsigned any = 0;
#pragma unroll
            for (int p = 0; p < 8; ++p) any |= key[p];
            if (!__any_sync(0xffffffffu, any != 0u)) continue;          // warp-uniform: nothing selected in this row
            unsigned left = __shfl_up_sync(0xffffffffu, key[7], 1);
            unsigned right = __shfl_down_sync(0xffffffffu, key[0], 1);
            if (lane == 0) left = ekey;
            if (lane == 31) right = ekey;
            unsigned sb = 0, eb = 0;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const unsigned l = p ? key[p - 1] : left, rn = p < 7 ? key[p + 1] : right;
                sb |= (key[p] != 0u && l != key[p] ? 1u : 0u) << p;
                eb |= (key[p] != 0u && rn != key[p] ? 1u : 0u) << p;
            }
            const unsigned mine = sb | (eb << 8);
            const unsigned m1 = __shfl_down_sync(0xffffffffu, mine, 1), m2 = __shfl_down_sync(0xffffffffu, mine, 2),
                           m3 = __shfl_down_sync(0xffffffffu, mine, 3);
            unsigned cnt = 0;
            if ((lane & 3) == 0) {
                const unsigned sw = (mine & 0xFFu) | ((m1 & 0xFFu) << 8) | ((m2 & 0xFFu) << 16) | ((m3 & 0xFFu) << 24);
                const unsigned ew = ((mine >> 8) & 0xFFu) | (((m1 >> 8) & 0xFFu) << 8) | (((m2 >> 8) & 0xFFu) << 16) | (((m3 >> 8) & 0xFFu) << 24);
                const int wi = g * 8 + (lane >> 2);
                if (wi < wd && (sw | ew)) {
                    reinterpret_cast<uint32_t*>(rs + a.R.smask)[(size_t)y * wd + wi] = sw;
                    reinterpret_cast<uint32_t*>(rs + a.R.emask)[(size_t)y * wd + wi] = ew;
                }
                cnt = (unsigned)__popc(sw);
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) atomicAdd(reinterpret_cast<uint32_t*>(rs + a.R.rowcnt) + y, cnt);
        }
    }
}

// masks -> row-runs in raster order: a CTA owns 32 rows of one slice (prefix = the row counts above them)
__global__ void __launch_bounds__(256)
rle_block_emit_kernel(const BlkArgs a)
{
    __shared__ int s_part[8];
    __shared__ int s_off[33];
    const int b = blockIdx.z;
    const CodeView v = code_view(a, b);
    char* rs = a.rs + (size_t)b * a.rs_stride;
    const uint32_t* rowcnt = reinterpret_cast<const uint32_t*>(rs + a.R.rowcnt);
    const uint32_t* smask = reinterpret_cast<const uint32_t*>(rs + a.R.smask);
    const uint32_t* emask = reinterpret_cast<const uint32_t*>(rs + a.R.emask);
    int* rowoff = reinterpret_cast<int*>(rs + a.R.rowoff);
    int* r_y = reinterpret_cast<int*>(rs + a.R.r_y);
    int* r_xs = reinterpret_cast<int*>(rs + a.R.r_xs);
    int* r_xe = reinterpret_cast<int*>(rs + a.R.r_xe);
    uint32_t* r_key = reinterpret_cast<uint32_t*>(rs + a.R.r_key);
    int* parent = reinterpret_cast<int*>(rs + a.R.parent);
    const int H = a.crop_h, wd = a.R.wd, run_cap = a.run_cap;
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
        if (r0 + lane < H) rowoff[r0 + lane] = prefix + ex;
        if (r0 + 32 >= H && lane == 31) rowoff[H] = prefix + tot;
    }
    __syncthreads();

    for (int j = 0; j < 4; ++j) {
        const int rr = warp * 4 + j;
        const int y = r0 + rr;
        if (y >= H) break;
        const int base = s_off[rr];
        const int cnt = s_off[rr + 1] - base;
        if (cnt == 0) continue;
        int run_s = 0, run_e = 0;
        for (int wb = 0; wb < wd; wb += 32) {
            const int wi = wb + lane;
            unsigned sw = wi < wd ? smask[(size_t)y * wd + wi] : 0u;
            unsigned ew = wi < wd ? emask[(size_t)y * wd + wi] : 0u;
            int tot_s, tot_e;
            int ps = base + run_s + warp_excl_scan(__popc(sw), lane, &tot_s);
            int pe = base + run_e + warp_excl_scan(__popc(ew), lane, &tot_e);
            while (sw) {
                const int bit = __ffs(sw) - 1;
                sw &= sw - 1;
                const int x = wi * 32 + bit;
                if (ps < run_cap) { r_y[ps] = y; r_xs[ps] = x; r_key[ps] = key_at(v, y, x); parent[ps] = ps; }
                ++ps;
            }
            while (ew) {
                const int bit = __ffs(ew) - 1;
                ew &= ew - 1;
                if (pe < run_cap) r_xe[pe] = wi * 32 + bit + 1;         // exclusive end
                ++pe;
            }
            run_s += tot_s;
            run_e += tot_e;
            if (run_s >= cnt && run_e >= cnt) break;
        }
    }
}

constexpr int kStBlkOff = 7, kStBlkTotal = 8;       // internal status words: this slice's first run row / run rows of the block

// Everything behind the row-runs (a 2048^2 EM slice has a few thousand of them), one thread per row-run and all slices
// of the block per launch:
//   union    8-connected union of touching equal-key runs of CCL classes (root = lowest run index = raster-first pixel)
//   flags    key flags (CCL: root runs; else label - base)
//   slots    per slice: exclusive scan of the flags -> instance slots in the reference's dict order (class order of
//            `labels`, ascending label: rle.py:57-84), class / label of every table row
//   assign   slot per row-run, run count and row band per slot, (start, length, slot) triples for the matcher
//   offsets  per slice: scan of the counts -> where each slot's run list starts; the slice's place in the packed output
//   lists    a warp per slot walks the row-runs of its row band, keeps its own, merges a run that starts where the
//            slot's previous run ended (array_utils.rle_encode only breaks where idx[i] != idx[i-1]+1: a run reaching
//            the last column continues in column 0 of the next row) and writes (start, length) in ascending order; box
//            and area on the way
struct SliceRuns {
    int32_t* status;
    const int *rowoff, *r_y, *r_xs, *r_xe;
    const uint32_t* r_key;
    int *parent, *slot_of, *flags, *cnt, *ymin, *ymax;
    long long* inst;
    int n_all, n;
};

__device__ __forceinline__ SliceRuns slice_runs(const BlkArgs& a, int b)
{
    SliceRuns s;
    char* rs = a.rs + (size_t)b * a.rs_stride;
    s.status = reinterpret_cast<int32_t*>(rs + a.R.status);
    s.rowoff = reinterpret_cast<const int*>(rs + a.R.rowoff);
    s.r_y = reinterpret_cast<const int*>(rs + a.R.r_y);
    s.r_xs = reinterpret_cast<const int*>(rs + a.R.r_xs);
    s.r_xe = reinterpret_cast<const int*>(rs + a.R.r_xe);
    s.r_key = reinterpret_cast<const uint32_t*>(rs + a.R.r_key);
    s.parent = reinterpret_cast<int*>(rs + a.R.parent);
    s.slot_of = reinterpret_cast<int*>(rs + a.R.slot_of);
    s.flags = reinterpret_cast<int*>(rs + a.R.flags);
    s.cnt = reinterpret_cast<int*>(rs + a.R.cnt);
    s.ymin = reinterpret_cast<int*>(rs + a.R.ymin);
    s.ymax = reinterpret_cast<int*>(rs + a.R.ymax);
    s.inst = reinterpret_cast<long long*>(rs + a.R.inst);
    s.n_all = s.rowoff[a.crop_h];
    s.n = min(s.n_all, a.run_cap);
    return s;
}

// key-space offset of class ci when the slice has n row-runs: CCL classes take n keys, the others min(L, k_cap + 1)
__device__ __forceinline__ long long blk_key_off(const BlkArgs& a, int ci, int n)
{
    const long long plain = min(a.rc.L, (long long)a.k_cap + 1);
    long long o = 0;
    for (int i = 0; i < ci; ++i) o += a.rc.ccl[i] ? (long long)n : plain;
    return o;
}

__global__ void __launch_bounds__(256)
rle_block_union_kernel(const BlkArgs a)
{
    const SliceRuns s = slice_runs(a, blockIdx.z);
    const int n = s.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = s.r_key[i];
        const int c = (int)(key >> 22) - 1;
        const int y = s.r_y[i];
        if (c < 0 || !a.rc.ccl[c] || y == 0) continue;
        const int lo = min(s.rowoff[y - 1], n), hi = min(s.rowoff[y], n);
        const int xs = s.r_xs[i], xe = s.r_xe[i];
        int p = lo, q = hi;                                     // first run of the row above with r_xe >= xs
        while (p < q) {
            const int m = (p + q) >> 1;
            if (s.r_xe[m] >= xs) q = m; else p = m + 1;
        }
        for (int j = p; j < hi && s.r_xs[j] <= xe; ++j)
            if (s.r_key[j] == key) uf_union(s.parent, i, j);
    }
}

__global__ void __launch_bounds__(256)
rle_block_flags_kernel(const BlkArgs a)
{
    const SliceRuns s = slice_runs(a, blockIdx.z);
    const int n = s.n;
    const long long plain = min(a.rc.L, (long long)a.k_cap + 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        s.status[EMP_ST_NROWRUNS] = s.n_all;
        if (s.n_all > a.run_cap) atomicOr(s.status + EMP_ST_FLAGS, EMP_FLAG_RLE_OVERFLOW);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = s.r_key[i];
        const int c = (int)(key >> 22) - 1;
        if (c < 0) continue;
        const long long off = blk_key_off(a, c, n);
        if (a.rc.ccl[c]) {
            // no path compression: parent[] is read by other threads' finds and a root's self-link must survive
            if (uf_find(s.parent, i) == i) s.flags[off + i] = 1;
        } else {
            const long long kv = (long long)(key & 0x3FFFFFu);
            if (kv < plain) s.flags[off + kv] = 1;
            else atomicOr(s.status + EMP_ST_FLAGS, EMP_FLAG_RLE_OVERFLOW);      // cannot happen: labels are 1 + rank <= K
        }
    }
}

__global__ void __launch_bounds__(1024)
rle_block_slots_kernel(const BlkArgs a)
{
    __shared__ int s_w[32];
    __shared__ int s_base[EMP_MAX_LABELS + 1];
    const SliceRuns s = slice_runs(a, blockIdx.x);
    const RleClasses& rc = a.rc;
    const int tid = threadIdx.x, n = s.n, inst_cap = a.inst_cap;
    const long long plain = min(rc.L, (long long)a.k_cap + 1);
    const long long F = blk_key_off(a, rc.n, n);
    const int n_inst_all = cta_scan_inplace(s.flags, F, s_w);
    if (tid == 0) {
        s.flags[F] = n_inst_all;
        s.status[EMP_ST_NINST] = n_inst_all;
        if (n_inst_all > inst_cap) atomicOr(s.status + EMP_ST_FLAGS, EMP_FLAG_RLE_OVERFLOW);
    }
    __syncthreads();
    if (tid <= rc.n) s_base[tid] = s.flags[blk_key_off(a, tid, n)];
    __syncthreads();
    for (long long p = tid; p < F; p += 1024) {
        const int slot = s.flags[p];
        if (s.flags[p + 1] - slot != 1 || slot >= inst_cap) continue;
        int ci = 0;
        long long off = 0;
        for (; ci < rc.n; ++ci) {
            const long long sz = rc.ccl[ci] ? (long long)n : plain;
            if (p < off + sz) break;
            off += sz;
        }
        long long* row = s.inst + (size_t)slot * EMP_BLK_INST_WORDS;
        row[0] = rc.label[ci];
        row[1] = rc.ccl[ci] ? rc.lo[ci] + (long long)(slot - s_base[ci]) + 1 : rc.lo[ci] + (p - off);
        s.ymin[slot] = INT_MAX; s.ymax[slot] = -1;
    }
}

__global__ void __launch_bounds__(256)
rle_block_assign_kernel(const BlkArgs a)
{
    const int b = blockIdx.z;
    const SliceRuns s = slice_runs(a, b);
    const int n = s.n, Wc = a.crop_w, inst_cap = a.inst_cap;
    const long long plain = min(a.rc.L, (long long)a.k_cap + 1);
    long long* r3 = a.runs3 ? a.runs3 + (size_t)b * a.runs3_stride : nullptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = s.r_key[i];
        const int c = (int)(key >> 22) - 1;
        int slot = -1;
        if (c >= 0) {
            const long long off = blk_key_off(a, c, n);
            const long long kv = a.rc.ccl[c] ? (long long)uf_find(s.parent, i) : (long long)(key & 0x3FFFFFu);
            slot = (a.rc.ccl[c] || kv < plain) ? s.flags[off + kv] : -1;
        }
        s.slot_of[i] = slot;
        const int y = s.r_y[i];
        if (slot >= 0 && slot < inst_cap) {
            atomicAdd(s.cnt + slot, 1);
            atomicMin(s.ymin + slot, y);
            atomicMax(s.ymax + slot, y);
        }
        if (r3) {
            r3[(size_t)i * 3] = (long long)y * Wc + s.r_xs[i];
            r3[(size_t)i * 3 + 1] = s.r_xe[i] - s.r_xs[i];
            r3[(size_t)i * 3 + 2] = slot;
        }
    }
}

__global__ void __launch_bounds__(1024)
rle_block_offsets_kernel(const BlkArgs a)
{
    __shared__ int s_w[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SliceRuns s = slice_runs(a, b);
    const int n_inst = min(s.status[EMP_ST_NINST], a.inst_cap);
    cta_scan_inplace(s.cnt, n_inst, s_w);                       // cnt[slot] = first row of the slot's run list (slice-local)
    // this slice's place in the packed run region: the row-runs of the slices before it
    int before = 0, all = 0;
    for (int j = tid; j < a.B; j += 1024) {
        const int* ro = reinterpret_cast<const int*>(a.rs + (size_t)j * a.rs_stride + a.R.rowoff);
        const int nj = min(ro[a.crop_h], a.run_cap);
        all += nj;
        if (j < b) before += nj;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, d); all += __shfl_xor_sync(0xffffffffu, all, d); }
    __shared__ int s_b[32], s_a[32];
    if (lane == 0) { s_b[warp] = before; s_a[warp] = all; }
    __syncthreads();
    if (tid == 0) {
        int tb = 0, ta = 0;
        for (int w = 0; w < 32; ++w) { tb += s_b[w]; ta += s_a[w]; }
        s.status[kStBlkOff] = tb;
        s.status[kStBlkTotal] = ta;
    }
}

__global__ void __launch_bounds__(256)
rle_block_lists_kernel(const BlkArgs a)
{
    const int b = blockIdx.z;
    const SliceRuns s = slice_runs(a, b);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n = s.n, Wc = a.crop_w;
    const int n_inst = min(s.status[EMP_ST_NINST], a.inst_cap);
    const int my_off = s.status[kStBlkOff], total_rr = s.status[kStBlkTotal];
    // run lists as int32 pairs of planes (flat indices are < 2^31): half the bytes of the one copy that crosses PCIe
    const int rr_even = (total_rr + 1) & ~1;
    int* starts_out = reinterpret_cast<int*>(a.packed + EMP_BLK_HDR_WORDS + (size_t)EMP_BLK_SLICE_WORDS * a.B) + my_off;
    int* lens_out = starts_out + rr_even;
    for (int slot = warp; slot < n_inst; slot += n_warps) {
        const int ya = s.ymin[slot], yb = s.ymax[slot];
        const int base = s.cnt[slot];
        long long* row = s.inst + (size_t)slot * EMP_BLK_INST_WORDS;
        if (yb < 0) {                                           // cannot happen (a slot exists because a run carries it)
            if (lane == 0) { row[2] = 0; row[3] = 0; row[4] = 0; row[5] = 0; row[6] = 0; row[7] = my_off + base; row[8] = 0; }
            continue;
        }
        const int ja = min(s.rowoff[ya], n), jb = min(s.rowoff[yb + 1], n);
        int n_final = 0, x0 = INT_MAX, x1 = -1;
        long long carry_start = -1, carry_end = -1, area = 0;   // the slot's last run so far (flat indices)
        for (int j0 = ja; j0 < jb; j0 += 32) {
            const int j = j0 + lane;
            const bool mine = j < jb && s.slot_of[j] == slot;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (!m) continue;                                   // warp-uniform
            long long start = 0, end = 0;
            if (mine) {
                const int y = s.r_y[j], xs = s.r_xs[j], xe = s.r_xe[j];
                start = (long long)y * Wc + xs;
                end = (long long)y * Wc + xe;
                x0 = min(x0, xs); x1 = max(x1, xe);
                area += xe - xs;
            }
            const unsigned below = m & lanemask_lt();
            const int prev_lane = below ? 31 - __clz(below) : 0;
            long long prev_end = __shfl_sync(0xffffffffu, end, prev_lane);
            if (!below) prev_end = carry_end;
            const bool head = mine && prev_end != start;
            const unsigned hb = __ballot_sync(0xffffffffu, head);
            // start of the final run this row-run belongs to: the nearest head at or below this lane, else the carried one
            const unsigned hle = hb & (lanemask_lt() | (1u << lane));
            const int head_lane = hle ? 31 - __clz(hle) : 0;
            long long fstart = __shfl_sync(0xffffffffu, start, head_lane);
            if (!hle) fstart = carry_start;
            const int fidx = n_final + __popc(hle) - 1;         // index of that final run within the slot
            // the last row-run of a final run inside this chunk writes its length (a continuation in a later chunk overwrites it)
            const unsigned above = m & ~(lanemask_lt() | (1u << lane));
            const int next_lane = above ? __ffs(above) - 1 : 32;
            const bool last_of_final = mine && (next_lane == 32 || ((hb >> next_lane) & 1u));
            if (head) starts_out[base + fidx] = (int)start;
            if (last_of_final) lens_out[base + fidx] = (int)(end - fstart);
            const int last_lane = 31 - __clz(m);                // carry: the slot's last row-run of this chunk
            carry_end = __shfl_sync(0xffffffffu, end, last_lane);
            carry_start = __shfl_sync(0xffffffffu, fstart, last_lane);
            n_final += __popc(hb);
            __syncwarp();
        }
        x0 = __reduce_min_sync(0xffffffffu, x0);
        x1 = __reduce_max_sync(0xffffffffu, x1);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) area += __shfl_xor_sync(0xffffffffu, area, d);
        if (lane == 0) {
            row[2] = ya; row[3] = x0; row[4] = yb + 1; row[5] = x1;
            row[6] = n_final; row[7] = my_off + base; row[8] = area;
        }
    }
}

// instance tables of all slices packed behind the run lists + the header the host parses
__global__ void __launch_bounds__(256)
rle_block_pack_kernel(const BlkArgs a)
{
    __shared__ int s_w[32];
    __shared__ long long s_max[EMP_MAX_LABELS];
    const int b = blockIdx.x, tid = threadIdx.x;
    int before = 0, all = 0, rr_before = 0, rr_all = 0;
    for (int j = tid; j < a.B; j += blockDim.x) {
        const char* rs = a.rs + (size_t)j * a.rs_stride;
        const int ni = min(__ldcg(reinterpret_cast<const int32_t*>(rs + a.R.status) + EMP_ST_NINST), a.inst_cap);
        const int nr = min(__ldcg(reinterpret_cast<const int*>(rs + a.R.rowoff) + a.crop_h), a.run_cap);
        all += ni; rr_all += nr;
        if (j < b) { before += ni; rr_before += nr; }
    }
    // block_sum needs 32 warps' worth of slots; with 256 threads the upper ones are zero
    if (tid < 32) s_w[tid] = 0;
    __syncthreads();
    auto sum256 = [&](int v) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        __syncthreads();
        if ((tid & 31) == 0) s_w[tid >> 5] = v;
        __syncthreads();
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_w[w];
        return t;
    };
    const int inst_off = sum256(before), total_inst = sum256(all), rr_off = sum256(rr_before), total_rr = sum256(rr_all);
    const char* rs = a.rs + (size_t)b * a.rs_stride;
    const int32_t* status = reinterpret_cast<const int32_t*>(rs + a.R.status);
    const int n_inst = min(__ldcg(status + EMP_ST_NINST), a.inst_cap);
    const long long* inst = reinterpret_cast<const long long*>(rs + a.R.inst);
    long long* hdr = a.packed;
    long long* out = a.packed + EMP_BLK_HDR_WORDS + (size_t)EMP_BLK_SLICE_WORDS * a.B + (size_t)((total_rr + 1) & ~1) +
                     (size_t)inst_off * EMP_BLK_INST_WORDS;
    for (int i = tid; i < n_inst * EMP_BLK_INST_WORDS; i += blockDim.x) out[i] = inst[i];
    // largest label - class base per class over the block (the z-sharded stack's label offsets)
    if (tid < EMP_MAX_LABELS) s_max[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n_inst; i += blockDim.x) {
        const long long cls = inst[(size_t)i * EMP_BLK_INST_WORDS], lab = inst[(size_t)i * EMP_BLK_INST_WORDS + 1];
        for (int ci = 0; ci < a.rc.n; ++ci)
            if (a.rc.label[ci] == cls) atomicMax(reinterpret_cast<unsigned long long*>(s_max + ci), (unsigned long long)(lab - a.rc.lo[ci]));
    }
    __syncthreads();
    if (tid < a.rc.n && s_max[tid] > 0) {
        atomicMax(reinterpret_cast<unsigned long long*>(hdr + EMP_BLK_HDR_MAXLAB + tid), (unsigned long long)s_max[tid]);
        if (a.maxlab_all) atomicMax(reinterpret_cast<unsigned long long*>(a.maxlab_all + tid), (unsigned long long)s_max[tid]);
    }
    if (tid == 0) {
        const int32_t* cstat = reinterpret_cast<const int32_t*>(a.cs + (size_t)b * a.cs_stride + a.o_cstatus);
        const int32_t* mstat = reinterpret_cast<const int32_t*>(a.ws + (size_t)b * a.ws_stride + a.o_status);
        long long* s = hdr + EMP_BLK_HDR_WORDS + (size_t)EMP_BLK_SLICE_WORDS * b;
        s[0] = n_inst;
        s[1] = inst_off;
        s[2] = rr_off;
        s[3] = min(__ldcg(status + EMP_ST_NROWRUNS), a.run_cap);
        const int fl = __ldcg(cstat + EMP_ST_FLAGS) | __ldcg(mstat + EMP_ST_FLAGS) | __ldcg(status + EMP_ST_FLAGS);
        s[4] = (long long)fl;
        s[5] = __ldcg(cstat + EMP_ST_K);
        // a slice that overflowed a table is not in the maxima: poison them, so that every rank of a sharded stack learns
        // of it from the gathered maxima (EMP_BLK_MAXLAB_OVERFLOW) and all of them fail together
        if ((fl & (EMP_FLAG_K_OVERFLOW | EMP_FLAG_RLE_OVERFLOW)) && a.maxlab_all)
            atomicMax(reinterpret_cast<unsigned long long*>(a.maxlab_all), (unsigned long long)EMP_BLK_MAXLAB_OVERFLOW);
        if (b == 0) { hdr[0] = a.B; hdr[1] = total_rr; hdr[2] = total_inst; hdr[3] = EMP_BLK_INST_WORDS; }
    }
}

}  // namespace emp

using namespace emp;

// =====================================================================================================
// C ABI
// =====================================================================================================
static int set_need(ChainArgs& a, const emp_need_map* need, size_t hw)
{
    if (!need || !need->map) return EMP_OK;
    EMP_REQUIRE(need->W > 0 && hw % (size_t)need->W == 0 && need->shift >= 0 && need->shift < 16 && need->wc > 0 &&
                ((need->W - 1) >> need->shift) < need->wc, EMP_ERR_INVALID, "bad need-map geometry");
    EMP_REQUIRE(need->stride >= (size_t)need->wc * (((hw / need->W - 1) >> need->shift) + 1), EMP_ERR_INVALID, "need-map stride too small");
    a.need = need->map; a.need_stride = need->stride; a.W = need->W; a.shift = need->shift; a.wc = need->wc;
    a.thing_bits = need->thing_bits;
    return EMP_OK;
}

EMP_API int emp_median_chain(const float* const* planes_dev, int n, int n_planes, int z0, int depth, int ks, int C, size_t hw,
                             const float* const* carry_in_dev, float confidence_thr, uint8_t* sem8_out, size_t sem8_stride,
                             float* best_scratch, float* const* carry_out_dev, const emp_need_map* need, void* stream)
{
    EMP_REQUIRE(planes_dev && sem8_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(ks >= 1 && ks <= 15 && (ks & 1), EMP_ERR_INVALID, "median kernel size must be odd and <= 15 (got %d)", ks);
    EMP_REQUIRE(n >= 1 && n_planes >= n && z0 >= 0 && z0 + n <= depth && C >= 1 && C <= 256 && hw > 0, EMP_ERR_INVALID,
                "bad block (n=%d, n_planes=%d, z0=%d, depth=%d, C=%d)", n, n_planes, z0, depth, C);
    const int mid = ks / 2;
    EMP_REQUIRE(n_planes >= std::min(n + mid, depth - z0), EMP_ERR_INVALID, "the block needs %d raw planes (got %d)",
                std::min(n + mid, depth - z0), n_planes);
    EMP_REQUIRE(mid == 0 || z0 == 0 || (z0 >= mid && carry_in_dev), EMP_ERR_INVALID,
                "a block that starts at z0=%d needs z0 >= %d and the carry planes", z0, mid);
    EMP_REQUIRE(C == 1 || best_scratch, EMP_ERR_INVALID, "multi-channel chains need the running-maximum scratch");
    EMP_REQUIRE(sem8_stride >= hw, EMP_ERR_INVALID, "sem8 stride smaller than a plane");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the planes themselves are only known on the device; torch planes are 256-byte aligned, channel planes hw apart
    const bool vec = (hw % 4 == 0) && (sem8_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(sem8_out) & 3u) == 0) &&
                     (!best_scratch || (reinterpret_cast<uintptr_t>(best_scratch) & 15u) == 0);
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.planes = planes_dev; a.carry_in = (mid > 0 && z0 >= mid) ? carry_in_dev : nullptr; a.carry_out = carry_out_dev;
    a.n = n; a.n_planes = n_planes; a.z0 = z0; a.depth = depth; a.hw = hw; a.thr = confidence_thr;
    a.sem8 = sem8_out; a.sem8_stride = sem8_stride; a.best = best_scratch; a.multi = C > 1;
    int rcn;
    if ((rcn = set_need(a, need, hw))) return rcn;
    for (int c = 0; c < C; ++c) {
        a.c = c;
        const int rc = launch_chain<false>(ks, a, vec, st);
        if (rc) return rc;
    }
    return EMP_OK;
}

EMP_API int emp_median_chain_repair(const float* const* planes_dev, int n, int n_planes, int z0, int depth, int ks, size_t hw,
                                    const float* const* carry_old_dev, const float* const* carry_new_dev,
                                    float confidence_thr, uint8_t* sem8, size_t sem8_stride, float* const* carry_out_dev,
                                    int32_t* changed, const emp_need_map* need, void* stream)
{
    EMP_REQUIRE(planes_dev && sem8 && carry_old_dev && carry_new_dev && carry_out_dev && changed, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(ks >= 3 && ks <= 15 && (ks & 1), EMP_ERR_INVALID, "median kernel size must be odd, 3 .. 15 (got %d)", ks);
    const int mid = ks / 2;
    EMP_REQUIRE(n >= 1 && z0 >= mid && z0 + n <= depth && hw > 0 && n_planes >= std::min(n + mid, depth - z0), EMP_ERR_INVALID,
                "bad block (n=%d, n_planes=%d, z0=%d, depth=%d)", n, n_planes, z0, depth);
    EMP_REQUIRE(sem8_stride >= hw, EMP_ERR_INVALID, "sem8 stride smaller than a plane");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (hw % 4 == 0) && (sem8_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(sem8) & 3u) == 0);
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.planes = planes_dev; a.carry_in = carry_old_dev; a.carry_new = carry_new_dev; a.carry_out = carry_out_dev;
    a.n = n; a.n_planes = n_planes; a.z0 = z0; a.depth = depth; a.hw = hw; a.thr = confidence_thr;
    a.sem8 = sem8; a.sem8_stride = sem8_stride; a.changed = changed;
    int rcn;
    if ((rcn = set_need(a, need, hw))) return rcn;
    return launch_chain<true>(ks, a, vec, st);
}

namespace {

struct BlockPlan {
    WsLayout Lc, Lm;            // coarse / merge workspaces
    BlkLayout R;
    size_t o_ids, o_cs, o_ws, o_rs, total;
    size_t ids_stride;          // int32 elements
    Things th;
    RleClasses rc;
};

int plan_block(const emp_stack_cfg* c, int B, BlockPlan* P)
{
    EMP_REQUIRE(c != nullptr, EMP_ERR_INVALID, "cfg is null");
    EMP_REQUIRE(B >= 1 && B <= 4096, EMP_ERR_INVALID, "bad block size %d", B);
    EMP_REQUIRE(c->H > 0 && c->W > 0 && c->h > 0 && c->w > 0 && (long long)c->H * c->W < (1ll << 31), EMP_ERR_INVALID, "bad shape");
    EMP_REQUIRE(c->shift >= 0 && c->shift < 16 && ((c->H - 1) >> c->shift) < c->h && ((c->W - 1) >> c->shift) < c->w, EMP_ERR_INVALID,
                "coarse map %d x %d << %d does not cover %d x %d", c->h, c->w, c->shift, c->H, c->W);
    EMP_REQUIRE(c->crop_h >= 1 && c->crop_h <= c->H && c->crop_w >= 1 && c->crop_w <= c->W, EMP_ERR_INVALID, "bad crop %d x %d",
                c->crop_h, c->crop_w);
    EMP_REQUIRE(c->nms_kernel >= 1 && c->k_cap >= 1 && c->run_cap >= 1 && c->inst_cap >= 1, EMP_ERR_INVALID, "bad nms_kernel / capacities");
    int rc;
    if ((rc = make_things(c->thing_list, c->n_things, &P->th))) return rc;
    if ((rc = make_rle_classes(c->labels, c->n_labels, c->label_divisor, c->thing_list, c->n_things, c->force_connected, &P->rc))) return rc;
    P->Lc = ws_layout(c->h, c->w, c->k_cap, 1);
    P->Lm = ws_layout(c->H, c->W, c->k_cap, P->th.n);
    EMP_REQUIRE(P->Lm.code16, EMP_ERR_INVALID, "k_cap must stay below %u", kClsBase16);
    P->R = blk_layout(c->crop_h, c->crop_w, c->run_cap, c->inst_cap, c->n_labels, c->label_divisor, c->k_cap);
    P->ids_stride = align_up((size_t)c->h * c->w, 64);
    size_t o = 0;
    P->o_ids = o; o = align_up(o + sizeof(int32_t) * P->ids_stride * B, 256);
    P->o_cs = o;  o = align_up(o + P->Lc.total * B, 256);
    P->o_ws = o;  o = align_up(o + P->Lm.total * B, 256);
    P->o_rs = o;  o = align_up(o + P->R.total * B, 256);
    P->total = o;
    return EMP_OK;
}

}  // namespace

EMP_API size_t emp_stack_block_scratch_bytes(const emp_stack_cfg* cfg, int B)
{
    BlockPlan P;
    if (plan_block(cfg, B, &P)) return 0;
    return P.total;
}

EMP_API size_t emp_stack_block_packed_words(const emp_stack_cfg* cfg, int B)
{
    if (!cfg || B < 1 || cfg->run_cap < 1 || cfg->inst_cap < 1) return 0;
    return (size_t)EMP_BLK_HDR_WORDS + (size_t)EMP_BLK_SLICE_WORDS * B +
           (size_t)B * ((size_t)cfg->run_cap + 1 + (size_t)EMP_BLK_INST_WORDS * cfg->inst_cap);
}

static int stack_block_impl(const emp_stack_cfg* cfg, int B, const uint8_t* sem8, size_t sem8_stride, const float* hm,
                            size_t hm_stride, const float* off, size_t off_stride, const uint8_t* need, size_t need_stride,
                            void* scratch, size_t scratch_bytes, int64_t* packed_out, size_t packed_words, int64_t* runs3_out,
                            int64_t* maxlab_all, void* stream)
{
    BlockPlan P;
    int rc = plan_block(cfg, B, &P);
    if (rc) return rc;
    EMP_REQUIRE(sem8 && hm && off && scratch && packed_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255u) == 0 && scratch_bytes >= P.total, EMP_ERR_WORKSPACE,
                "scratch too small or misaligned: %zu < %zu", scratch_bytes, P.total);
    EMP_REQUIRE(packed_words >= emp_stack_block_packed_words(cfg, B), EMP_ERR_WORKSPACE, "packed output too small");
    EMP_REQUIRE(sem8_stride >= (size_t)cfg->H * cfg->W && hm_stride >= (size_t)cfg->h * cfg->w &&
                off_stride >= 2 * (size_t)cfg->h * cfg->w && (!need || need_stride >= (size_t)cfg->h * cfg->w), EMP_ERR_INVALID,
                "plane strides smaller than the planes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(scratch);
    int32_t* ids = reinterpret_cast<int32_t*>(base + P.o_ids);
    char* cs = base + P.o_cs;
    char* ws = base + P.o_ws;
    char* rs = base + P.o_rs;

    if ((rc = coarse_ids_batched(B, hm, hm_stride, off, off_stride, cfg->h, cfg->w, cfg->nms_threshold, cfg->nms_kernel, cfg->step,
                                 ids, P.ids_stride, cfg->k_cap, cs, P.Lc.total, st, need, need_stride)))
        return rc;
    const int32_t* k_dev = reinterpret_cast<const int32_t*>(cs + P.Lc.status) + EMP_ST_K;
    {
        ProfScope ps(ST_MEMSET, st);
        EMP_CUDA_CHECK(cudaMemset2DAsync(rs, P.R.total, 0, P.R.zero_bytes, (size_t)B, st));
        EMP_CUDA_CHECK(cudaMemsetAsync(packed_out, 0, sizeof(int64_t) * EMP_BLK_HDR_WORDS, st));
    }
    // is the void label itself a selected label?  (then it may share its label with a stuff class or an instance, and run
    // boundaries have to be decided on labels, by rle_block_mark, not on codes inside the merge kernel)
    bool void_selected = false;
    for (int i = 0; i < P.rc.n; ++i)
        void_selected |= cfg->void_label != 0 && cfg->void_label >= P.rc.lo[i] && cfg->void_label < P.rc.lo[i] + P.rc.L;
    bool marked = false;
    unsigned long long thing_bits = 0ull;
    bool things_small = true;
    for (int i = 0; i < P.th.n; ++i) {
        if (P.th.v[i] >= 0 && P.th.v[i] < 64) thing_bits |= 1ull << P.th.v[i];
        else things_small = false;
    }
    if (things_small && cfg->W % 16 == 0 && (reinterpret_cast<uintptr_t>(sem8) & 15u) == 0 && sem8_stride % 16 == 0) {
        {
            ProfScope ps(ST_MEMSET, st);
            EMP_CUDA_CHECK(cudaMemset2DAsync(ws, P.Lm.total, 0, P.Lm.zero_bytes, (size_t)B, st));
        }
        LeanArgs m;
        memset(&m, 0, sizeof(m));
        m.sem8 = sem8; m.sem_stride = sem8_stride; m.ids = ids; m.ids_stride = P.ids_stride;
        m.ws = ws; m.ws_stride = P.Lm.total;
        m.o_votes = P.Lm.votes; m.o_areas = P.Lm.areas; m.o_sflags = P.Lm.sflags; m.o_codes = P.Lm.codes;
        m.B = B; m.H = cfg->H; m.W = cfg->W; m.wc = cfg->w; m.shift = cfg->shift;
        m.blocks_x = (cfg->W + 63) / 64;
        m.blk_shift = 0;
        while ((1 << m.blk_shift) < assign_block_items(cfg->H, cfg->W)) ++m.blk_shift;     // strips per block: a power of two
        m.T = P.th.n > 0 ? P.th.n : 1; m.thing_bits = thing_bits;
        m.simple = (P.th.n == 1 && P.th.v[0] > 0 && cfg->shift >= 2) ? 1 : 0;
        m.thing_class = P.th.n == 1 ? (unsigned)P.th.v[0] : 0u;
        m.mark = void_selected ? 0 : 1;
        m.rs = rs; m.rs_stride = P.R.total; m.o_smask = P.R.smask; m.o_emask = P.R.emask; m.o_rowcnt = P.R.rowcnt;
        m.wd = P.R.wd; m.crop_h = cfg->crop_h; m.crop_w = cfg->crop_w;
        marked = m.mark != 0;
        const unsigned items = (unsigned)(((cfg->H + 3) / 4) * ((cfg->W + 511) / 512));
        {
            ProfScope ps(ST_ASSIGN, st);
            merge_lean_kernel<<<dim3((items + kLeanWarps - 1) / kLeanWarps, 1, B), kLeanWarps * 32, 0, st>>>(m);
        }
        EMP_CUDA_CHECK(cudaGetLastError());
        if ((rc = build_luts_batched(B, cfg->H, cfg->W, P.th, cfg->label_divisor, cfg->stuff_area, cfg->void_label, cfg->k_cap, k_dev,
                                     P.Lc.total / 4, ws, P.Lm.total, st)))
            return rc;
    } else if ((rc = merge_codes_batched(B, sem8, sem8_stride, ids, P.ids_stride, cfg->h, cfg->w, cfg->shift, cfg->H, cfg->W, P.th,
                                         cfg->label_divisor, cfg->stuff_area, cfg->void_label, cfg->k_cap, k_dev, P.Lc.total / 4,
                                         ws, P.Lm.total, st))) {
        return rc;
    }

    BlkArgs a;
    memset(&a, 0, sizeof(a));
    a.ws = ws; a.ws_stride = P.Lm.total; a.o_codes = P.Lm.codes; a.o_sflags = P.Lm.sflags; a.o_lut = P.Lm.lut; a.o_status = P.Lm.status;
    a.cs = cs; a.cs_stride = P.Lc.total; a.o_cstatus = P.Lc.status;
    a.rs = rs; a.rs_stride = P.R.total; a.R = P.R; a.rc = P.rc;
    a.B = B; a.W = cfg->W; a.crop_h = cfg->crop_h; a.crop_w = cfg->crop_w;
    a.blk_shift = 0;
    while ((1 << a.blk_shift) < assign_block_items(cfg->H, cfg->W)) ++a.blk_shift;
    a.blocks_x = (cfg->W + 63) / 64;
    a.cls_off = (unsigned)P.Lm.cls_off;
    a.k_cap = cfg->k_cap; a.run_cap = cfg->run_cap; a.inst_cap = cfg->inst_cap;
    a.vec = (cfg->W % 8 == 0);
    a.packed = reinterpret_cast<long long*>(packed_out);
    a.runs3 = reinterpret_cast<long long*>(runs3_out); a.runs3_stride = 3 * (size_t)cfg->run_cap;
    a.maxlab_all = reinterpret_cast<long long*>(maxlab_all);

    {
        ProfScope ps(ST_BLK_KEYS, st);
        rle_block_keys_kernel<<<dim3(8, 1, B), 256, 0, st>>>(a);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    if (!marked) {                                              // else merge_lean has written the masks already
        const unsigned items = (unsigned)(((cfg->crop_h + 3) / 4) * ((cfg->crop_w + 255) / 256));
        ProfScope ps(ST_BLK_MARK, st);
        rle_block_mark_kernel<<<dim3((items + 7) / 8, 1, B), 256, 0, st>>>(a);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    {
        ProfScope ps(ST_BLK_EMIT, st);
        rle_block_emit_kernel<<<dim3((cfg->crop_h + 31) / 32, 1, B), 256, 0, st>>>(a);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    {
        ProfScope ps(ST_BLK_RUNS, st);                          // union .. lists as one interval
        const unsigned gx = (unsigned)std::min(((size_t)cfg->run_cap + 255) / 256, (size_t)64);
        bool any_ccl = false;
        for (int i = 0; i < P.rc.n; ++i) any_ccl |= P.rc.ccl[i] != 0;
        if (any_ccl) rle_block_union_kernel<<<dim3(gx, 1, B), 256, 0, st>>>(a);
        rle_block_flags_kernel<<<dim3(gx, 1, B), 256, 0, st>>>(a);
        rle_block_slots_kernel<<<B, 1024, 0, st>>>(a);
        rle_block_assign_kernel<<<dim3(gx, 1, B), 256, 0, st>>>(a);
        rle_block_offsets_kernel<<<B, 1024, 0, st>>>(a);
        const unsigned gl = (unsigned)std::min(((size_t)cfg->inst_cap + 7) / 8, (size_t)32);
        rle_block_lists_kernel<<<dim3(gl, 1, B), 256, 0, st>>>(a);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    {
        ProfScope ps(ST_BLK_PACK, st);
        rle_block_pack_kernel<<<B, 256, 0, st>>>(a);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}

EMP_API int emp_stack_block(const emp_stack_cfg* cfg, int B, const uint8_t* sem8, size_t sem8_stride, const float* hm,
                            size_t hm_stride, const float* off, size_t off_stride, const uint8_t* need, size_t need_stride,
                            void* scratch, size_t scratch_bytes, int64_t* packed_out, size_t packed_words, int64_t* runs3_out,
                            void* stream)
{
    return stack_block_impl(cfg, B, sem8, sem8_stride, hm, hm_stride, off, off_stride, need, need_stride, scratch, scratch_bytes,
                            packed_out, packed_words, runs3_out, nullptr, stream);
}

// n slices as ceil(n / SB) blocks, each followed by the device -> host copy of the first host_words words of its packed
// tables on copy_stream and by an 8-byte copy of its header word 0 (= B, never 0) into host_flags[block]: copies on
// one stream land in order, so a host thread that sees host_flags[block] != 0 may parse host_out[block].
EMP_API int emp_stack_blocks(const emp_stack_cfg* cfg, int n, int SB, const uint8_t* sem8, size_t sem8_stride, const float* hm,
                             size_t hm_stride, const float* off, size_t off_stride, const uint8_t* need, size_t need_stride,
                             void* scratch, size_t scratch_bytes, int64_t* packed_all, size_t packed_words, int64_t* runs3_all,
                             int64_t* maxlab_all, int64_t* host_out, size_t host_stride, size_t host_words,
                             int64_t* host_flags, void* stream, void* copy_stream)
{
    EMP_REQUIRE(n >= 1 && SB >= 1, EMP_ERR_INVALID, "bad block sizes (n=%d, SB=%d)", n, SB);
    EMP_REQUIRE(packed_all && host_out && host_flags && copy_stream, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(host_words >= (size_t)EMP_BLK_HDR_WORDS && host_words <= host_stride && host_words <= packed_words, EMP_ERR_INVALID,
                "bad host copy size");
    static thread_local std::vector<cudaEvent_t> events;        // reused from call to call; never destroyed
    cudaStream_t st = static_cast<cudaStream_t>(stream), cs = static_cast<cudaStream_t>(copy_stream);
    const int n_sub = (n + SB - 1) / SB;
    while ((int)events.size() < n_sub) {
        cudaEvent_t e;
        EMP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        events.push_back(e);
    }
    for (int bi = 0; bi < n_sub; ++bi) {
        const int i0 = bi * SB, B = std::min(SB, n - i0);
        int64_t* packed = packed_all + (size_t)bi * packed_words;
        const int rc = stack_block_impl(cfg, B, sem8 + (size_t)i0 * sem8_stride, sem8_stride, hm + (size_t)i0 * hm_stride, hm_stride,
                                        off + (size_t)i0 * off_stride, off_stride, need ? need + (size_t)i0 * need_stride : nullptr,
                                        need_stride, scratch, scratch_bytes, packed, packed_words,
                                        runs3_all ? runs3_all + (size_t)i0 * 3 * cfg->run_cap : nullptr, maxlab_all, stream);
        if (rc) return rc;
        EMP_CUDA_CHECK(cudaEventRecord(events[bi], st));
        EMP_CUDA_CHECK(cudaStreamWaitEvent(cs, events[bi], 0));
        const size_t full = emp_stack_block_packed_words(cfg, B);
        EMP_CUDA_CHECK(cudaMemcpyAsync(host_out + (size_t)bi * host_stride, packed, sizeof(int64_t) * std::min(host_words, full),
                                       cudaMemcpyDeviceToHost, cs));
        EMP_CUDA_CHECK(cudaMemcpyAsync(host_flags + bi, packed, sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
    }
    return EMP_OK;
}
