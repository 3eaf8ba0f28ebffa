/*
 * oracle.c — CPU restatement of empanada's panoptic post-processing + RLE path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker.  The product path (empanada_b200/) never imports it.
 *
 * Parity: pinned.  Every function below is checked bit-for-bit against outputs of the
 * reference itself (imported from /root/reference in the build container) by
 * tests/golden/make_golden.py -> tests/golden/<case>.npz -> tests/test_oracle_golden.py.
 *
 * Each function cites the reference file:line whose behaviour it restates.  The code is a
 * restatement of the *closed forms* (SURVEY.md Appendix A), not a transliteration of the
 * torch op sequence: e.g. group_pixels is one brute-force argmin per pixel instead of
 * chunks of 20 centers with (20,H*W,2) temporaries.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp oracle.c -o _build/liboracle.so -lm
 *        (-ffp-contract=off matters: the distance is sqrt(fma(dx,dx, rn(dy*dy))) with the
 *         y-term rounded first; letting gcc contract dy*dy+... would break bit parity.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * find_instance_center — empanada/inference/postprocess.py:38-76
 *   F.threshold(x, t, -1) -> max_pool2d(k, stride 1, pad k//2) [even k: drop last row/col]
 *   -> keep x == pooled -> nonzero(x > 0), row-major.
 * Closed form: (y,x) is a center iff v > t32 and v > 0 and v >= every *thresholded* value in
 * the window rows y-lo..y+hi, cols x-lo..x+hi (clipped), lo = k//2, hi = k-1-lo.
 * Returns K; writes min(K, cap) rows of (y, x).
 * ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_find_centers(const float* hm, int H, int W, float thr, int k,
                                 int64_t* out, int64_t cap)
{
    const int lo = k / 2, hi = k - 1 - lo;
    int64_t n = 0;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const float v = hm[(int64_t)y * W + x];
            if (!(v > thr) || !(v > 0.0f)) continue;
            int peak = 1;
            const int y0 = y - lo < 0 ? 0 : y - lo, y1 = y + hi >= H ? H - 1 : y + hi;
            const int x0 = x - lo < 0 ? 0 : x - lo, x1 = x + hi >= W ? W - 1 : x + hi;
            for (int yy = y0; yy <= y1 && peak; ++yy)
                for (int xx = x0; xx <= x1; ++xx) {
                    float u = hm[(int64_t)yy * W + xx];
                    u = (u > thr) ? u : -1.0f;          /* F.threshold, postprocess.py:55 */
                    if (u > v) { peak = 0; break; }
                }
            if (peak) {
                if (n < cap) { out[2 * n] = y; out[2 * n + 1] = x; }
                ++n;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * group_pixels / chunked_pixel_grouping — postprocess.py:118-169, :78-116
 *   ly = fl(y*step + off_y), lx = fl(x*step + off_x); cy = step*ctr_y; dy = fl(cy - ly) ...
 *   d_k = sqrt_rn( fma(dx, dx, rn(dy*dy)) )   (torch CPU norm over the last dim)
 *   K <= chunksize : id = 1 + argmin_k d_k (first minimum)
 *   K  > chunksize : running strict-< minimum from 1e5 -> id 0 if every d_k >= 1e5
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_group_pixels_masked(const int64_t* ctr, int64_t K, const float* off, int H, int W,
                                     float step, int chunksize, const uint8_t* mask, int64_t* ids);

ORC_API void orc_group_pixels(const int64_t* ctr, int64_t K, const float* off, int H, int W,
                              float step, int chunksize, int64_t* ids)
{
    orc_group_pixels_masked(ctr, K, off, H, W, step, chunksize, NULL, ids);
}

/* mask != NULL: only pixels with mask[p] != 0 are searched, the rest get id 0 (what
 * get_instance_segmentation's `instance_seg * instance_id` keeps, postprocess.py:221). */
ORC_API void orc_group_pixels_masked(const int64_t* ctr, int64_t K, const float* off, int H, int W,
                                     float step, int chunksize, const uint8_t* mask, int64_t* ids)
{
    const int64_t HW = (int64_t)H * W;
    float* cy = (float*)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
    float* cx = (float*)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
    for (int64_t i = 0; i < K; ++i) {
        cy[i] = step * (float)ctr[2 * i];
        cx[i] = step * (float)ctr[2 * i + 1];
    }
    const int chunked = K > chunksize;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        const float fy = (float)y * step;
        for (int x = 0; x < W; ++x) {
            const int64_t p = (int64_t)y * W + x;
            if (mask && !mask[p]) { ids[p] = 0; continue; }
            const float ly = fy + off[p];
            const float lx = (float)x * step + off[HW + p];
            float best = chunked ? 1e5f : 0.0f;
            int64_t bid = 0;
            for (int64_t i = 0; i < K; ++i) {
                const float dy = cy[i] - ly;
                const float dx = cx[i] - lx;
                const float t = dy * dy;
                const float d = sqrtf(fmaf(dx, dx, t));
                if (!chunked && i == 0) { best = d; bid = 1; continue; }
                /* strict <: first minimum wins.  A non-finite location makes every d_k the
                 * same inf/NaN, so K<=chunksize gives id 1 and the chunked path id 0, which is
                 * also what torch.argmin / the 1e5 running minimum produce. */
                if (d < best) { best = d; bid = i + 1; }
            }
            ids[p] = bid;
        }
    }
    free(cy); free(cx);
}

/* ------------------------------------------------------------------------------------------
 * merge_semantic_and_instance — postprocess.py:223-296
 * ------------------------------------------------------------------------------------------ */
static int cmp_i64(const void* a, const void* b)
{
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

static int in_list(int64_t v, const int64_t* list, int n)
{
    for (int i = 0; i < n; ++i) if (list[i] == v) return 1;
    return 0;
}

typedef struct { int64_t id; int64_t cls; } vote_key;

static int cmp_vote(const void* a, const void* b)
{
    const vote_key* x = (const vote_key*)a; const vote_key* y = (const vote_key*)b;
    if (x->id != y->id) return (x->id > y->id) - (x->id < y->id);
    return (x->cls > y->cls) - (x->cls < y->cls);
}

typedef struct { int64_t id; int64_t newlabel; } id_map;

static int64_t lookup_id(const id_map* m, int64_t n, int64_t id)
{
    int64_t lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int64_t mid = (lo + hi) / 2;
        if (m[mid].id == id) return m[mid].newlabel;
        if (m[mid].id < id) lo = mid + 1; else hi = mid - 1;
    }
    return INT64_MIN;
}

ORC_API int orc_merge(const int64_t* sem, const int64_t* ins, int64_t n, int64_t L,
                      const int64_t* things, int nt, int64_t stuff_area, int64_t void_label,
                      int64_t* pan)
{
    /* votes: (id, class) pairs over tm = ins != 0-id pixels whose sem is a thing class.
     * The reference loops over torch.unique(ins) skipping only id 0 (:263-266), so negative
     * ids vote too. */
    int64_t nv = 0;
    for (int64_t p = 0; p < n; ++p) if (ins[p] != 0 && in_list(sem[p], things, nt)) ++nv;
    vote_key* v = (vote_key*)malloc(sizeof(vote_key) * (size_t)(nv > 0 ? nv : 1));
    int64_t j = 0;
    for (int64_t p = 0; p < n; ++p)
        if (ins[p] != 0 && in_list(sem[p], things, nt)) { v[j].id = ins[p]; v[j].cls = sem[p]; ++j; }
    qsort(v, (size_t)nv, sizeof(vote_key), cmp_vote);

    /* per id (ascending): majority class, ties -> smallest class (torch.mode, :273);
     * new id = 1-based rank among ids with the same majority class (:274-280). */
    id_map* map = (id_map*)malloc(sizeof(id_map) * (size_t)(nv > 0 ? nv : 1));
    int64_t nmap = 0;
    /* class tracker: small open list */
    int64_t trk_cls[4096]; int64_t trk_cnt[4096]; int ntrk = 0;
    int64_t i = 0;
    while (i < nv) {
        const int64_t id = v[i].id;
        int64_t best_cls = 0, best_cnt = -1;
        while (i < nv && v[i].id == id) {
            const int64_t c = v[i].cls; int64_t cnt = 0;
            while (i < nv && v[i].id == id && v[i].cls == c) { ++cnt; ++i; }
            if (cnt > best_cnt) { best_cnt = cnt; best_cls = c; }   /* ascending c: first max */
        }
        int t = 0;
        for (; t < ntrk; ++t) if (trk_cls[t] == best_cls) break;
        if (t == ntrk) { if (ntrk >= 4096) { free(v); free(map); return -1; } trk_cls[ntrk] = best_cls; trk_cnt[ntrk] = 0; ++ntrk; }
        trk_cnt[t] += 1;
        map[nmap].id = id; map[nmap].newlabel = best_cls * L + trk_cnt[t]; ++nmap;
    }

    /* stuff areas: sem == c and not (ins > 0), for c not in things (:284-294) */
    int64_t ns = 0;
    for (int64_t p = 0; p < n; ++p)
        if (!(ins[p] > 0) && !in_list(sem[p], things, nt)) ++ns;
    int64_t* cls_list = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ns > 0 ? ns : 1));
    int64_t nn = 0;
    for (int64_t p = 0; p < n; ++p)
        if (!(ins[p] > 0) && !in_list(sem[p], things, nt)) cls_list[nn++] = sem[p];
    qsort(cls_list, (size_t)nn, sizeof(int64_t), cmp_i64);
    int64_t* ucls = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nn > 0 ? nn : 1));
    int64_t* uarea = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nn > 0 ? nn : 1));
    int64_t nu = 0;
    for (int64_t a = 0; a < nn;) {
        int64_t b = a; while (b < nn && cls_list[b] == cls_list[a]) ++b;
        ucls[nu] = cls_list[a]; uarea[nu] = b - a; ++nu; a = b;
    }

    for (int64_t p = 0; p < n; ++p) {
        int64_t out = void_label;
        const int thing = in_list(sem[p], things, nt);
        if (thing) {
            if (ins[p] != 0) out = lookup_id(map, nmap, ins[p]);
        } else if (!(ins[p] > 0)) {
            /* binary search class */
            int64_t lo = 0, hi = nu - 1;
            while (lo <= hi) {
                const int64_t mid = (lo + hi) / 2;
                if (ucls[mid] == sem[p]) { if (uarea[mid] >= stuff_area) out = sem[p] * L; break; }
                if (ucls[mid] < sem[p]) lo = mid + 1; else hi = mid - 1;
            }
        }
        pan[p] = out;
    }
    free(v); free(map); free(cls_list); free(ucls); free(uarea);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * _MedianQueue.get_median — engines.py:59-66: middle order statistic of ks (odd) planes.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_median(const float* const* planes, int ks, int64_t n, float* out)
{
    float buf[64] = {0};
    if (ks > 64) ks = 64;
    for (int64_t p = 0; p < n; ++p) {
        for (int i = 0; i < ks; ++i) buf[i] = planes[i][p];
        for (int i = 1; i < ks; ++i) {          /* insertion sort; NaN sorts last like torch */
            const float key = buf[i]; int j2 = i - 1;
            while (j2 >= 0 && (buf[j2] > key || (buf[j2] != buf[j2] && key == key))) { buf[j2 + 1] = buf[j2]; --j2; }
            buf[j2 + 1] = key;
        }
        out[p] = buf[(ks - 1) / 2];
    }
}

/* _harden_seg — engines.py:114-121: C>1 argmax over channels (first max), C==1 >= thr. */
ORC_API void orc_harden(const float* prob, int C, int64_t hw, float thr, int64_t* out)
{
    for (int64_t p = 0; p < hw; ++p) {
        if (C == 1) { out[p] = prob[p] >= thr ? 1 : 0; continue; }
        int best = 0; float bv = prob[p];
        for (int c = 1; c < C; ++c) {
            const float v = prob[(int64_t)c * hw + p];
            if (v > bv || (v != v && bv == bv)) { bv = v; best = c; }   /* NaN counts as max */
        }
        out[p] = best;
    }
}

/* ------------------------------------------------------------------------------------------
 * connected_components — empanada/inference/rle.py:18-24 (cc3d connectivity=8 /
 * skimage.measure.label default): components of equal non-zero value, 8-connected,
 * numbered 1..n by raster order of each component's first pixel.  Returns n.
 * ------------------------------------------------------------------------------------------ */
static int64_t uf_find(int64_t* parent, int64_t a)
{
    while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
    return a;
}

static void uf_union(int64_t* parent, int64_t a, int64_t b)
{
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
}

ORC_API int64_t orc_ccl8(const int64_t* seg, int H, int W, int64_t* out)
{
    const int64_t n = (int64_t)H * W;
    int64_t* parent = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t p = 0; p < n; ++p) parent[p] = p;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int64_t p = (int64_t)y * W + x; const int64_t v = seg[p];
            if (v == 0) continue;
            if (x > 0 && seg[p - 1] == v) uf_union(parent, p, p - 1);
            if (y > 0) {
                if (seg[p - W] == v) uf_union(parent, p, p - W);
                if (x > 0 && seg[p - W - 1] == v) uf_union(parent, p, p - W - 1);
                if (x + 1 < W && seg[p - W + 1] == v) uf_union(parent, p, p - W + 1);
            }
        }
    int64_t next = 0;
    for (int64_t p = 0; p < n; ++p) {
        if (seg[p] == 0) { out[p] = 0; continue; }
        const int64_t r = uf_find(parent, p);
        if (r == p) out[p] = ++next;        /* root = min flat index = raster-first pixel */
        else out[p] = out[r];
    }
    free(parent);
    return next;
}

/* ------------------------------------------------------------------------------------------
 * pan_seg_to_rle_seg — rle.py:26-86 (+ array_utils.rle_encode, array_utils.py:209-235),
 * flattened to tables.  For each class label in `labels` (in the given order):
 *   seg = pan with values outside [label*L, (label+1)*L) zeroed; thing + force_connected ->
 *   CCL8 then += label*L; for each distinct non-zero value ascending: bbox and runs over the
 *   FLAT index (a run continues from column W-1 into column 0 of the next row).
 * Outputs (caller allocates with generous capacity):
 *   inst[ni*7 + {0..6}] = class_label, instance_label, y0, x0, y1, x1, n_runs   (ni rows)
 *   runs[2*r + {0,1}]   = start, length, grouped per instance in inst order.
 * Returns 0, or -1 if a capacity was exceeded.  *n_inst, *n_runs receive the totals.
 * ------------------------------------------------------------------------------------------ */
typedef struct { int64_t label; int64_t idx; } lab_idx;

static int cmp_lab_idx(const void* a, const void* b)
{
    const lab_idx* x = (const lab_idx*)a; const lab_idx* y = (const lab_idx*)b;
    if (x->label != y->label) return (x->label > y->label) - (x->label < y->label);
    return (x->idx > y->idx) - (x->idx < y->idx);
}

ORC_API int orc_pan_to_rle(const int64_t* pan, int H, int W, const int64_t* labels, int nl,
                           int64_t L, const int64_t* things, int nt, int force_connected,
                           int64_t* inst, int64_t inst_cap, int64_t* runs, int64_t runs_cap,
                           int64_t* n_inst, int64_t* n_runs)
{
    const int64_t n = (int64_t)H * W;
    int64_t* seg = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t* cc = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    lab_idx* li = (lab_idx*)malloc(sizeof(lab_idx) * (size_t)(n > 0 ? n : 1));
    int64_t ni = 0, nr = 0; int rc = 0;
    for (int l = 0; l < nl && rc == 0; ++l) {
        const int64_t lo = labels[l] * L, hi = lo + L;
        for (int64_t p = 0; p < n; ++p) seg[p] = (pan[p] < lo || pan[p] >= hi) ? 0 : pan[p];
        if (force_connected && in_list(labels[l], things, nt)) {
            orc_ccl8(seg, H, W, cc);
            for (int64_t p = 0; p < n; ++p) seg[p] = cc[p] > 0 ? cc[p] + lo : 0;
        }
        int64_t m = 0;
        for (int64_t p = 0; p < n; ++p) if (seg[p] != 0) { li[m].label = seg[p]; li[m].idx = p; ++m; }
        qsort(li, (size_t)m, sizeof(lab_idx), cmp_lab_idx);
        for (int64_t a = 0; a < m && rc == 0;) {
            int64_t b = a; while (b < m && li[b].label == li[a].label) ++b;
            if (ni >= inst_cap) { rc = -1; break; }
            int64_t y0 = H, x0 = W, y1 = -1, x1 = -1, cnt = 0;
            for (int64_t q = a; q < b; ++q) {
                const int64_t yy = li[q].idx / W, xx = li[q].idx % W;
                if (yy < y0) y0 = yy;
                if (yy > y1) y1 = yy;
                if (xx < x0) x0 = xx;
                if (xx > x1) x1 = xx;
                if (q == a || li[q].idx != li[q - 1].idx + 1) {     /* array_utils.py:221 */
                    if (nr >= runs_cap) { rc = -1; break; }
                    runs[2 * nr] = li[q].idx; runs[2 * nr + 1] = 1; ++nr; ++cnt;
                } else {
                    runs[2 * (nr - 1) + 1] += 1;
                }
            }
            int64_t* row = inst + 7 * ni;
            row[0] = labels[l]; row[1] = li[a].label; row[2] = y0; row[3] = x0;
            row[4] = y1 + 1; row[5] = x1 + 1; row[6] = cnt;
            ++ni; a = b;
        }
    }
    free(seg); free(cc); free(li);
    *n_inst = ni; *n_runs = nr;
    return rc;
}
