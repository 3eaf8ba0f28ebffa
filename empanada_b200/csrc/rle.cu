// rle.cu — pan_seg_to_rle_seg (reference empanada/inference/rle.py:26-86, connected_components
// :18-24, array_utils.rle_encode array_utils.py:209-235) on the GPU.
//
// Only ONE kernel touches pixels (rle_mark: 8 B/px read, 2 bits/px written); everything after
// works on row-runs — maximal horizontal segments of one selected value — of which a 4096^2 EM
// tile has tens of thousands, not millions:
//   rle_mark     pan -> start / end ballot words + per-row run counts           (warp ballot)
//   rle_emit     ordered expansion into run records, per-row offsets            (per-row prefix)
//   rle_union    8-connected union-find over runs of CCL classes (root = lowest run index,
//                i.e. the run holding the component's raster-first pixel)
//   rle_flags    per run: find root; mark its instance key (CCL: root run, else value - lo)
//   rle_slots    exclusive scan of the key flags -> instance slots in the reference's dict order
//                (class order of `labels`, ascending label), instance table rows
//   rle_assign   per run: slot, box atomics, "head" flag (not mergeable with the previous run:
//                array_utils.rle_encode only breaks where idx[i] != idx[i-1]+1, so a run that
//                ends at column W-1 continues into column 0 of the next row)
//   rle_finish   scan of head flags -> final (start, length, slot) runs in ascending start order
#include <string.h>
#include <algorithm>
#include "rle_common.cuh"

namespace emp {

struct RleLayout {
    size_t status, rowcnt, zero_bytes;
    size_t smask, emask, rowoff, r_y, r_xs, r_xe, r_cls, r_val, parent, slot, head, flags, total;
    size_t flags_len;
    int wd;
};

static RleLayout rle_layout(int H, int W, int run_cap, int n_labels, long long L)
{
    RleLayout R;
    R.wd = (W + 31) / 32;
    if (n_labels < 1) n_labels = 1;
    const size_t key_space = (size_t)std::max<long long>((long long)run_cap, L);
    R.flags_len = key_space * (size_t)n_labels + 1;
    size_t o = 0;
    R.status = o; o = align_up(o + sizeof(int32_t) * EMP_ST_WORDS, 256);
    R.rowcnt = o; o = align_up(o + sizeof(uint32_t) * (size_t)H, 256);
    R.zero_bytes = o;
    R.smask = o;  o = align_up(o + sizeof(uint32_t) * (size_t)H * R.wd, 256);
    R.emask = o;  o = align_up(o + sizeof(uint32_t) * (size_t)H * R.wd, 256);
    R.rowoff = o; o = align_up(o + sizeof(int) * ((size_t)H + 1), 256);
    R.r_y = o;    o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.r_xs = o;   o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.r_xe = o;   o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.r_cls = o;  o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.r_val = o;  o = align_up(o + sizeof(long long) * (size_t)run_cap, 256);
    R.parent = o; o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.slot = o;   o = align_up(o + sizeof(int) * (size_t)run_cap, 256);
    R.head = o;   o = align_up(o + sizeof(int) * ((size_t)run_cap + 1), 256);
    R.flags = o;  o = align_up(o + sizeof(int) * R.flags_len, 256);
    R.total = o;
    return R;
}

// ---------------------------------------------------------------------------------------------
// rle_mark — the one pass over the pixels.  A warp takes 256 consecutive pixels of a row (8 mask
// words); lane l loads pixels 32j + l (j = 0..7: eight independent 8-byte loads in flight per lane,
// 256 B per warp-wide access) and reduces each to a 32-bit run key — 0 if the value belongs to no
// selected class, else (class index + 1) << 22 | (value - class base), injective because
// label_divisor <= 2^22 — so that neighbours compare with one shuffle.  An item with no selected pixel
// (the bulk of an EM slice) stores zeros and moves on.
constexpr int kMarkWords = 8;

__global__ void __launch_bounds__(256)
rle_mark_kernel(const long long* __restrict__ pan, int H, int W, int wd, RleClasses rc,
                uint32_t* __restrict__ smask, uint32_t* __restrict__ emask, uint32_t* __restrict__ rowcnt)
{
    const int lane = threadIdx.x & 31;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int groups = (wd + kMarkWords - 1) / kMarkWords;             // items per row
    const size_t items = (size_t)H * groups;
    for (size_t it = warp_global; it < items; it += n_warps) {
        const int y = (int)(it / groups), wg = (int)(it % groups);
        const int w0 = wg * kMarkWords;
        const int xb = w0 * 32;                                         // first pixel of the item
        const long long* row = pan + (size_t)y * W;
        long long v[kMarkWords];
#pragma unroll
        for (int j = 0; j < kMarkWords; ++j) {
            const int x = xb + 32 * j + lane;
            v[j] = x < W ? __ldcs(row + x) : 0;
        }
        unsigned key[kMarkWords];
        unsigned any = 0;
#pragma unroll
        for (int j = 0; j < kMarkWords; ++j) {
            key[j] = (v[j] != 0) ? run_key(v[j], rc) : 0u;
            any |= key[j];
        }
        if (!__any_sync(0xffffffffu, any != 0u)) {                      // warp-uniform: nothing selected here
            if (lane < kMarkWords && w0 + lane < wd) {
                smask[(size_t)y * wd + w0 + lane] = 0u;
                emask[(size_t)y * wd + w0 + lane] = 0u;
            }
            continue;
        }
        // keys of the pixels just outside the item (0 at the image border: the row starts / ends a run)
        unsigned edge = 0;
        if (lane == 0 && xb > 0) edge = run_key(__ldg(row + xb - 1), rc);
        if (lane == 31 && xb + 32 * kMarkWords < W) edge = run_key(__ldg(row + xb + 32 * kMarkWords), rc);
        unsigned mys = 0, mye = 0, cnt = 0;
#pragma unroll
        for (int j = 0; j < kMarkWords; ++j) {
            unsigned left = __shfl_up_sync(0xffffffffu, key[j], 1);
            unsigned right = __shfl_down_sync(0xffffffffu, key[j], 1);
            const unsigned prev_last = __shfl_sync(0xffffffffu, j > 0 ? key[j > 0 ? j - 1 : 0] : edge, j > 0 ? 31 : 0);
            const unsigned next_first = __shfl_sync(0xffffffffu, j + 1 < kMarkWords ? key[j + 1 < kMarkWords ? j + 1 : j] : edge,
                                                    j + 1 < kMarkWords ? 0 : 31);
            if (lane == 0) left = prev_last;
            if (lane == 31) right = next_first;
            const int x = xb + 32 * j + lane;
            if (x == W - 1) right = 0;                                  // the last pixel of the row always ends its run
            const bool sel = key[j] != 0u;
            const unsigned sw = __ballot_sync(0xffffffffu, sel && left != key[j]);
            const unsigned ew = __ballot_sync(0xffffffffu, sel && right != key[j]);
            if (lane == j) { mys = sw; mye = ew; }
            cnt += (unsigned)__popc(sw);
        }
        if (lane < kMarkWords && w0 + lane < wd) {
            smask[(size_t)y * wd + w0 + lane] = mys;
            emask[(size_t)y * wd + w0 + lane] = mye;
        }
        if (lane == 0 && cnt) atomicAdd(rowcnt + y, cnt);
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rle_emit_kernel(const long long* __restrict__ pan, int H, int W, int wd, RleClasses rc,
                const uint32_t* __restrict__ smask, const uint32_t* __restrict__ emask,
                const uint32_t* __restrict__ rowcnt, int* __restrict__ rowoff, int* __restrict__ r_y,
                int* __restrict__ r_xs, int* __restrict__ r_xe, int* __restrict__ r_cls,
                long long* __restrict__ r_val, int* __restrict__ parent, int32_t* __restrict__ status, int run_cap)
{
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
        if (r0 + lane < H) rowoff[r0 + lane] = prefix + ex;
        if (r0 + 32 >= H && lane == 31) {
            rowoff[H] = prefix + tot;
            status[EMP_ST_NROWRUNS] = prefix + tot;
            if (prefix + tot > run_cap) atomicOr(status + EMP_ST_FLAGS, EMP_FLAG_RLE_OVERFLOW);
        }
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
                const int b = __ffs(sw) - 1;
                sw &= sw - 1;
                const int x = wi * 32 + b;
                if (ps < run_cap) {
                    const long long v = __ldg(pan + (size_t)y * W + x);
                    r_y[ps] = y; r_xs[ps] = x; r_val[ps] = v; r_cls[ps] = class_of(v, rc); parent[ps] = ps;
                }
                ++ps;
            }
            while (ew) {
                const int b = __ffs(ew) - 1;
                ew &= ew - 1;
                if (pe < run_cap) r_xe[pe] = wi * 32 + b + 1;       // exclusive end
                ++pe;
            }
            run_s += tot_s;
            run_e += tot_e;
            if (run_s >= cnt && run_e >= cnt) break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rle_union_kernel(const int32_t* __restrict__ status, int run_cap, RleClasses rc, const int* __restrict__ rowoff,
                 const int* __restrict__ r_y, const int* __restrict__ r_xs, const int* __restrict__ r_xe,
                 const int* __restrict__ r_cls, const long long* __restrict__ r_val, int* parent)
{
    const int n = min(status[EMP_ST_NROWRUNS], run_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = r_cls[i];
        if (c < 0 || !rc.ccl[c]) continue;
        const int y = r_y[i];
        if (y == 0) continue;
        int lo = min(rowoff[y - 1], n), hi = min(rowoff[y], n);
        const int xs = r_xs[i], xe = r_xe[i];
        const long long v = r_val[i];
        // first run j of the row above with r_xe[j] >= xs  (touches or overlaps column xs-1)
        int a = lo, b = hi;
        while (a < b) {
            const int m = (a + b) >> 1;
            if (r_xe[m] >= xs) b = m; else a = m + 1;
        }
        for (int j = a; j < hi && r_xs[j] <= xe; ++j)
            if (r_val[j] == v) uf_union(parent, i, j);
    }
}

__global__ void __launch_bounds__(256)
rle_flags_kernel(const int32_t* __restrict__ status, int run_cap, RleClasses rc, const int* __restrict__ r_cls,
                 const long long* __restrict__ r_val, int* parent, int* __restrict__ flags)
{
    const int n = min(status[EMP_ST_NROWRUNS], run_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = r_cls[i];
        if (c < 0) continue;
        const long long off = key_offset(c, n, rc);
        if (rc.ccl[c]) {
            const int root = uf_find(parent, i);
            // no path compression here: parent[] is still being read by other threads' finds and a
            // root's self-link must survive; rle_assign re-finds (chains are short after hooking)
            if (root == i) flags[off + i] = 1;
        } else {
            flags[off + (r_val[i] - rc.lo[c])] = 1;
        }
    }
}

__global__ void __launch_bounds__(1024)
rle_slots_kernel(int32_t* __restrict__ status, int run_cap, int inst_cap, RleClasses rc, int* __restrict__ flags,
                 long long* __restrict__ inst, int H, int W)
{
    __shared__ int s_w[32];
    __shared__ int s_base[EMP_MAX_LABELS + 1];
    const int n = min(status[EMP_ST_NROWRUNS], run_cap);
    const long long F = key_offset(rc.n, n, rc);
    const int total = cta_scan_inplace(flags, F, s_w);
    if (threadIdx.x == 0) {
        flags[F] = total;
        status[EMP_ST_NINST] = total;
        if (total > inst_cap) atomicOr(status + EMP_ST_FLAGS, EMP_FLAG_RLE_OVERFLOW);
    }
    __syncthreads();
    if (threadIdx.x <= rc.n) s_base[threadIdx.x] = flags[key_offset(threadIdx.x, n, rc)];
    __syncthreads();
    for (long long p = threadIdx.x; p < F; p += blockDim.x) {
        const int slot = flags[p];
        if (flags[p + 1] - slot != 1 || slot >= inst_cap) continue;
        int ci = 0;
        long long off = 0;
        for (; ci < rc.n; ++ci) {
            const long long sz = rc.ccl[ci] ? (long long)n : rc.L;
            if (p < off + sz) break;
            off += sz;
        }
        long long* row = inst + (size_t)slot * 8;
        row[0] = rc.label[ci];
        row[1] = rc.ccl[ci] ? rc.lo[ci] + (long long)(slot - s_base[ci]) + 1 : rc.lo[ci] + (p - off);
        row[2] = H; row[3] = W; row[4] = 0; row[5] = 0; row[6] = 0; row[7] = 0;
    }
}

__device__ __forceinline__ int run_slot(int i, int n, const RleClasses& rc, const int* r_cls,
                                        const long long* r_val, const int* parent, const int* flags)
{
    const int c = r_cls[i];
    if (c < 0) return -1;
    const long long off = key_offset(c, n, rc);
    const long long key = rc.ccl[c] ? (long long)uf_find(parent, i) : (r_val[i] - rc.lo[c]);
    return flags[off + key];
}

__global__ void __launch_bounds__(256)
rle_assign_kernel(const int32_t* __restrict__ status, int run_cap, int inst_cap, RleClasses rc, int W,
                  const int* __restrict__ r_y, const int* __restrict__ r_xs, const int* __restrict__ r_xe,
                  const int* __restrict__ r_cls, const long long* __restrict__ r_val, const int* __restrict__ parent,
                  const int* __restrict__ flags, int* __restrict__ slot_out, int* __restrict__ head,
                  long long* __restrict__ inst)
{
    const int n = min(status[EMP_ST_NROWRUNS], run_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = run_slot(i, n, rc, r_cls, r_val, parent, flags);
        slot_out[i] = slot;
        const int y = r_y[i], xs = r_xs[i], xe = r_xe[i];
        if (slot >= 0 && slot < inst_cap) {
            long long* row = inst + (size_t)slot * 8;
            atomicMin(row + 2, (long long)y);
            atomicMin(row + 3, (long long)xs);
            atomicMax(row + 4, (long long)y + 1);
            atomicMax(row + 5, (long long)xe);
        }
        int h = 1;
        if (i > 0) {
            const long long prev_end = (long long)r_y[i - 1] * W + r_xe[i - 1];
            const long long start = (long long)y * W + xs;
            if (prev_end == start && run_slot(i - 1, n, rc, r_cls, r_val, parent, flags) == slot) h = 0;
        }
        head[i] = h;
    }
}

__global__ void __launch_bounds__(1024)
rle_finish_kernel(int32_t* __restrict__ status, int run_cap, int out_cap, int inst_cap, int W,
                  const int* __restrict__ r_y, const int* __restrict__ r_xs, const int* __restrict__ r_xe,
                  const int* __restrict__ slot, int* __restrict__ head, long long* __restrict__ runs_out,
                  long long* __restrict__ inst)
{
    __shared__ int s_w[32];
    const int n = min(status[EMP_ST_NROWRUNS], run_cap);
    const int total = cta_scan_inplace(head, n, s_w);       // head[i] = index of the final run i opens / belongs to
    if (threadIdx.x == 0) {
        head[n] = total;
        status[EMP_ST_NRUNS] = total;
        if (total > out_cap) atomicOr(status + EMP_ST_FLAGS, EMP_FLAG_RLE_OVERFLOW);
    }
    __syncthreads();
    // after the exclusive scan, run i is a head iff head[i+1] == head[i] + 1; its final index is head[i];
    // a non-head run belongs to final run head[i] - 1 ... but only tails need that: the last row-run of
    // final run f is the i with head[i+1] - 1 == f and (i+1 == n or i+1 is a head).
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int f = head[i];
        if (head[i + 1] - f == 1 && f < out_cap) {
            runs_out[(size_t)f * 3] = (long long)r_y[i] * W + r_xs[i];
            runs_out[(size_t)f * 3 + 2] = slot[i];
            const int s = slot[i];
            if (s >= 0 && s < inst_cap) atomicAdd(reinterpret_cast<unsigned long long*>(inst + (size_t)s * 8 + 6), 1ull);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int next = head[i + 1];                       // number of heads among runs 0..i
        const bool next_is_head = (i + 1 == n) || (head[i + 2 <= n ? i + 2 : n] - next == 1);
        if (next_is_head) {
            const int f = next - 1;
            if (f >= 0 && f < out_cap) {
                const long long end = (long long)r_y[i] * W + r_xe[i];
                runs_out[(size_t)f * 3 + 1] = end - runs_out[(size_t)f * 3];
            }
        }
    }
}

}  // namespace emp

using namespace emp;

EMP_API size_t emp_rle_workspace_bytes(int H, int W, int run_cap, int n_labels, int64_t label_divisor)
{
    if (H <= 0 || W <= 0 || run_cap < 1 || label_divisor <= 0) return 0;
    return rle_layout(H, W, run_cap, n_labels, label_divisor).total;
}

EMP_API int emp_rle(const int64_t* pan, int H, int W, const int64_t* labels, int n_labels, int64_t label_divisor,
                    const int64_t* thing_list, int n_things, int force_connected, int64_t* runs_out, int run_cap,
                    int64_t* inst_out, int inst_cap, void* ws, size_t ws_bytes, void* stream)
{
    EMP_REQUIRE(pan && runs_out && inst_out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 31), EMP_ERR_INVALID, "bad image size %d x %d", H, W);
    EMP_REQUIRE(run_cap >= 1 && inst_cap >= 1, EMP_ERR_INVALID, "bad capacities");
    EMP_REQUIRE(n_things == 0 || thing_list, EMP_ERR_INVALID, "thing_list is null");
    RleClasses rc;
    int rcode = make_rle_classes(labels, n_labels, label_divisor, thing_list, n_things, force_connected, &rc);
    if (rcode) return rcode;
    const RleLayout R = rle_layout(H, W, run_cap, n_labels, label_divisor);
    EMP_REQUIRE(ws && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, EMP_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    EMP_REQUIRE(ws_bytes >= R.total, EMP_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, R.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* w = static_cast<char*>(ws);
    int32_t* status = reinterpret_cast<int32_t*>(w + R.status);
    uint32_t* rowcnt = reinterpret_cast<uint32_t*>(w + R.rowcnt);
    uint32_t* smask = reinterpret_cast<uint32_t*>(w + R.smask);
    uint32_t* emask = reinterpret_cast<uint32_t*>(w + R.emask);
    int* rowoff = reinterpret_cast<int*>(w + R.rowoff);
    int* r_y = reinterpret_cast<int*>(w + R.r_y);
    int* r_xs = reinterpret_cast<int*>(w + R.r_xs);
    int* r_xe = reinterpret_cast<int*>(w + R.r_xe);
    int* r_cls = reinterpret_cast<int*>(w + R.r_cls);
    long long* r_val = reinterpret_cast<long long*>(w + R.r_val);
    int* parent = reinterpret_cast<int*>(w + R.parent);
    int* slot = reinterpret_cast<int*>(w + R.slot);
    int* head = reinterpret_cast<int*>(w + R.head);
    int* flags = reinterpret_cast<int*>(w + R.flags);
    const long long* panp = reinterpret_cast<const long long*>(pan);

    EMP_CUDA_CHECK(cudaMemsetAsync(w, 0, R.zero_bytes, st));
    EMP_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int) * R.flags_len, st));

    const size_t chunks = (size_t)H * ((R.wd + kMarkWords - 1) / kMarkWords);      // 256-pixel items, one per warp
    unsigned g_mark = (unsigned)std::min<size_t>((chunks + 7) / 8, (size_t)device_sm_count() * 16);
    if (g_mark < 1) g_mark = 1;
    {
        ProfScope ps(ST_RLE_MARK, st);
        rle_mark_kernel<<<g_mark, 256, 0, st>>>(panp, H, W, R.wd, rc, smask, emask, rowcnt);
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    ProfScope ps_runs(ST_RLE_RUNS, st);          // everything after the pixel pass, as one interval
    rle_emit_kernel<<<(H + 31) / 32, 256, 0, st>>>(panp, H, W, R.wd, rc, smask, emask, rowcnt, rowoff, r_y, r_xs, r_xe,
                                                   r_cls, r_val, parent, status, run_cap);
    EMP_CUDA_CHECK(cudaGetLastError());
    const unsigned g_runs = (unsigned)std::min<size_t>(((size_t)run_cap + 255) / 256, (size_t)device_sm_count() * 8);
    bool any_ccl = false;
    for (int i = 0; i < rc.n; ++i) any_ccl |= rc.ccl[i] != 0;
    if (any_ccl) {
        rle_union_kernel<<<g_runs, 256, 0, st>>>(status, run_cap, rc, rowoff, r_y, r_xs, r_xe, r_cls, r_val, parent);
        EMP_CUDA_CHECK(cudaGetLastError());
    }
    rle_flags_kernel<<<g_runs, 256, 0, st>>>(status, run_cap, rc, r_cls, r_val, parent, flags);
    EMP_CUDA_CHECK(cudaGetLastError());
    rle_slots_kernel<<<1, 1024, 0, st>>>(status, run_cap, inst_cap, rc, flags, reinterpret_cast<long long*>(inst_out), H, W);
    EMP_CUDA_CHECK(cudaGetLastError());
    rle_assign_kernel<<<g_runs, 256, 0, st>>>(status, run_cap, inst_cap, rc, W, r_y, r_xs, r_xe, r_cls, r_val, parent,
                                              flags, slot, head, reinterpret_cast<long long*>(inst_out));
    EMP_CUDA_CHECK(cudaGetLastError());
    rle_finish_kernel<<<1, 1024, 0, st>>>(status, run_cap, run_cap, inst_cap, W, r_y, r_xs, r_xe, slot, head,
                                          reinterpret_cast<long long*>(runs_out), reinterpret_cast<long long*>(inst_out));
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}
