/*
 * empanada_b200.h — C ABI of libempanada_b200.so: B200 (sm_100a) replacement for empanada's
 * panoptic post-processing, 3D median/harden and pan_seg -> RLE path.
 *
 * The reference has no FFI layer (it is pure Python on torch ops); its boundary for this path is
 * the call surface of empanada/inference/{postprocess,engines,rle}.py.  Each entry point below
 * names the reference function (file:line under the reference tree) it replaces.  The Python
 * package empanada_b200.inference mirrors those functions name for name and reaches this library
 * through ctypes with tensor.data_ptr() / torch.cuda.current_stream().cuda_stream — no torch
 * types cross this interface.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless it says "host";
 *   - all work is enqueued on `stream` (a cudaStream_t); no entry point synchronises or allocates
 *     device memory — the caller passes a workspace of emp_workspace_bytes() bytes per tile, 256-B
 *     aligned (the Python shim allocates it with torch.empty so the caching allocator owns it);
 *   - dynamic results (number of centers, runs, instances) are written to a device status block
 *     at the start of the workspace (see EMP_ST_*); the caller reads it back when it needs them;
 *   - images are row-major contiguous planes; batched entry points take B tiles laid out
 *     back to back (tile stride = plane size) and B workspaces back to back;
 *   - return value: EMP_OK or an EMP_ERR_* code; emp_last_error() gives a thread-local message.
 *   - data-dependent problems (class id out of range, more centers than k_cap) cannot be known at
 *     enqueue time; they are reported through EMP_ST_FLAGS in the status block.
 */
#ifndef EMPANADA_B200_H
#define EMPANADA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMP_OK 0
#define EMP_ERR_INVALID 1   /* bad argument (null pointer, size, too many thing classes ...) */
#define EMP_ERR_CUDA 2      /* a CUDA runtime call failed; see emp_last_error() */
#define EMP_ERR_WORKSPACE 3 /* workspace too small or misaligned */

#define EMP_MAX_THINGS 16    /* thing classes per call */
#define EMP_MAX_CLASSES 4096 /* semantic class ids must lie in [0, EMP_MAX_CLASSES) */
#define EMP_MAX_LABELS 64    /* classes per emp_rle call */

/* status block: int32 words at the start of each tile's workspace */
#define EMP_ST_K 0          /* number of centers found (may exceed k_cap; then only k_cap are used) */
#define EMP_ST_FLAGS 1      /* EMP_FLAG_* bits */
#define EMP_ST_NROWRUNS 2   /* rle: row-runs found */
#define EMP_ST_NRUNS 3      /* rle: final runs */
#define EMP_ST_NINST 4      /* rle: instances (distinct output labels) */
#define EMP_ST_GSHIFT 5     /* internal: log2 of the center-index cell size in pixels */
#define EMP_ST_TICKET 6     /* internal: block hand-out counter of the assign kernel */
#define EMP_ST_WORDS 16
#define EMP_PROFILE_STAGES 16

#define EMP_FLAG_K_OVERFLOW 1     /* more centers than k_cap: result invalid, retry with larger cap */
#define EMP_FLAG_CLASS_RANGE 2    /* a semantic class id outside [0, EMP_MAX_CLASSES) was seen */
#define EMP_FLAG_ID_RANGE 4       /* an instance id outside [0, max_id] was seen (emp_merge) */
#define EMP_FLAG_RLE_OVERFLOW 8   /* rle run / instance capacity exceeded */

int emp_version(void);
const char* emp_last_error(void);

/* Optional per-stage device timing for benchmarks: while enabled, every kernel launch is
 * bracketed by a CUDA event pair on its stream.  emp_profile_read() waits for the recorded events
 * and returns, per stage (0 nms_peaks, 1 emit_centers, 2 assign, 3 build_lut, 4 apply_lut,
 * 5 median_harden, 6 rle_mark, 7 rle run kernels, 8 bin_centers, 9 median_chain, 10..14 the stack block's
 * rle keys / mark / emit / runs / pack, 15 the workspace clears of the batched entry points), the summed milliseconds and the
 * number of launches since the last read.  Both arrays have EMP_PROFILE_STAGES entries (host). */
int emp_profile_enable(int on);
int emp_profile_read(double* ms_per_stage, int* launches_per_stage);

/* Bytes of workspace ONE tile of H x W needs when at most k_cap centers (or, for emp_merge,
 * instance ids up to k_cap) and n_things thing classes are in play. */
size_t emp_workspace_bytes(int H, int W, int k_cap, int n_things);

/* find_instance_center — empanada/inference/postprocess.py:38-76.
 * Fused threshold + k x k stride-1 max-pool NMS + ordered (row-major) stream compaction; there is
 * no top-k in the reference (every NMS peak is kept, in row-major order).
 *   hm        (H,W) float32 heat-map
 *   ctr_out   (cap,2) int64 rows (y,x), may be NULL; rows beyond cap are dropped
 *   status[EMP_ST_K] receives K. */
int emp_find_centers(const float* hm, int H, int W, float threshold, int nms_kernel,
                     int64_t* ctr_out, int cap, void* ws, size_t ws_bytes, void* stream);

/* group_pixels (+ chunked_pixel_grouping) — postprocess.py:118-169, :78-116.
 * Offset-shifted nearest-center argmin over ALL pixels, exact fp32 distance
 * sqrt_rn(fma(dx,dx,rn(dy*dy))), ties to the lowest center index, 1e5 sentinel when K > chunksize.
 *   ctr (K,2) int64 (y,x); off (2,H,W) float32 (ch0 = dy, ch1 = dx)
 *   ids_out (H,W) int64, or int32 if ids_i32 != 0 (the Render engines' coarse id map). */
int emp_group_pixels(const int64_t* ctr, int K, const float* off, int H, int W, float step,
                     int chunksize, void* ids_out, int ids_i32, void* ws, size_t ws_bytes,
                     void* stream);

/* get_instance_segmentation — postprocess.py:171-221: centers, then nearest-center ids on
 * thing pixels only (0 elsewhere).
 *   sem (H,W) int64; ins_out (H,W) int64; ctr_out as in emp_find_centers. */
int emp_instance_segmentation(const int64_t* sem, const float* hm, const float* off, int H, int W,
                              const int64_t* thing_list /* host */, int n_things, float threshold,
                              int nms_kernel, int64_t* ins_out, int64_t* ctr_out, int cap,
                              int k_cap, void* ws, size_t ws_bytes, void* stream);

/* merge_semantic_and_instance — postprocess.py:223-296 (majority vote per instance, per-class
 * renumbering in ascending id order, stuff-area filter).
 *   sem, ins, pan_out (H,W) int64; instance ids must lie in [0, max_id]. */
int emp_merge(const int64_t* sem, const int64_t* ins, int H, int W, int64_t label_divisor,
              const int64_t* thing_list /* host */, int n_things, int64_t stuff_area,
              int64_t void_label, int64_t max_id, int64_t* pan_out, void* ws, size_t ws_bytes,
              void* stream);

/* Render-engine merge — engines.py:277-292 with get_instance_cells' nearest upsample
 * (engines.py:257-275) folded in: ins[Y,X] = thing(sem[Y,X]) ? coarse_ids[Y>>shift, X>>shift] : 0.
 *   sem (H,W) int64 or uint8 (sem_u8 != 0); coarse_ids (hc,wc) int32 with ids in [0, max_id]. */
int emp_merge_coarse(const void* sem, int sem_u8, const int32_t* coarse_ids, int hc, int wc,
                     int shift, int H, int W, int64_t label_divisor,
                     const int64_t* thing_list /* host */, int n_things, int64_t stuff_area,
                     int64_t void_label, int64_t max_id, const int32_t* k_dev /* device, may be NULL:
                     number of ids actually in use (<= max_id), bounds the label-LUT build */,
                     int64_t* pan_out, void* ws, size_t ws_bytes, void* stream);

/* get_instance_cells before the upsample — engines.py:257-272: centers of the (coarse) heat-map,
 * then group_pixels(step) over ALL pixels, without a host round trip for K.
 *   hm (h,w) f32; off (2,h,w) f32; ids_out (h,w) int32 (all 0 when there is no center);
 *   status[EMP_ST_K] = K (a device pointer to it is ws + 0). */
int emp_coarse_ids(const float* hm, const float* off, int h, int w, float threshold, int nms_kernel,
                   float step, int32_t* ids_out, int k_cap, void* ws, size_t ws_bytes, void* stream);

/* get_panoptic_segmentation — postprocess.py:298-356, fused, B tiles per call:
 * centers -> nearest-center ids on thing pixels -> votes -> label LUT -> panoptic map.
 *   sem (B,H,W) int64 or uint8; hm (B,H,W) f32; off (B,2,H,W) f32; pan_out (B,H,W) int64
 *   ctr_out (B,cap,2) int64 or NULL;  ws: B workspaces of emp_workspace_bytes(H,W,k_cap,n) each.
 *   Per tile status[EMP_ST_K] = K, status[EMP_ST_FLAGS] = flags. */
int emp_panoptic_batched(int B, const void* sem, int sem_u8, const float* hm, const float* off,
                         int H, int W, const int64_t* thing_list /* host */, int n_things,
                         int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                         float threshold, int nms_kernel, int64_t* pan_out, int64_t* ctr_out,
                         int cap, int k_cap, void* ws, size_t ws_bytes_per_tile, void* stream);

/* Same pipeline with HOST buffers (pinned or pageable): tiles are streamed host->device, processed
 * and streamed back on internal streams (copy/compute overlap); this is the end-to-end entry the
 * benchmark's e2e figure times.  dev_scratch: device buffer of emp_host_scratch_bytes(). Blocks
 * until the batch is complete.  k_out (host, B int32) receives K per tile, flags_out the flags.
 * The link is the bound of this entry, so sem is narrowed to one byte per pixel on the host (worker
 * threads, pinned staging owned by the library) before it crosses; a tile with a class id outside
 * [0, 255] travels as int64.  EMP_HOST_THREADS sets the worker count (0: no narrowing; default: the CPUs of
 * the process's affinity mask — divided by the visible GPUs when the mask is the whole machine — at most 8);
 * the workers live as long as the library.  One pipeline (streams, staging) exists per device and serves one
 * caller at a time; on any error the call returns only after every copy it enqueued has finished.
 * emp_host_sem_bytes_per_px(): what the LAST call sent per sem pixel on average (1.0 .. 8.0), for byte
 * accounting.  emp_panoptic_batched_host_u8 is the same entry for callers that already hold the class map as
 * bytes (nothing to narrow). */
size_t emp_host_scratch_bytes(int H, int W, int k_cap, int n_things);
double emp_host_sem_bytes_per_px(void);
int emp_panoptic_batched_host(int B, const int64_t* sem_h, const float* hm_h, const float* off_h,
                              int H, int W, const int64_t* thing_list, int n_things,
                              int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                              float threshold, int nms_kernel, int64_t* pan_out_h, int32_t* k_out,
                              int32_t* flags_out, int k_cap, void* dev_scratch,
                              size_t dev_scratch_bytes);
int emp_panoptic_batched_host_u8(int B, const uint8_t* sem8_h, const float* hm_h, const float* off_h,
                                 int H, int W, const int64_t* thing_list, int n_things,
                                 int64_t label_divisor, int64_t stuff_area, int64_t void_label,
                                 float threshold, int nms_kernel, int64_t* pan_out_h, int32_t* k_out,
                                 int32_t* flags_out, int k_cap, void* dev_scratch,
                                 size_t dev_scratch_bytes);

/* _MedianQueue.get_median + _harden_seg — engines.py:59-66, :114-121.
 * Median (middle order statistic) over ks odd planes of (C,H,W) float32, optionally written back
 * (the reference stores it into the queued entry, which makes the filter recursive), then
 * hardened: C > 1 first-argmax over C, C == 1 `>= thr`.  ks == 1 hardens planes[0] directly.
 *   planes: HOST array of ks device pointers;  median_out (C,H,W) f32 or NULL;
 *   sem_out (H,W) int64, or uint8 if sem_u8 != 0, or NULL. */
int emp_median_harden(const float* const* planes /* host array */, int ks, int C, int H, int W,
                      float confidence_thr, float* median_out, void* sem_out, int sem_u8,
                      void* stream);

/* _MedianQueue + _harden_seg over a whole z-block — engines.py:47-90 (enqueue / get_next / end), :114-121.
 * The queue's filter is recursive: m_z = median(m_{z-mid} .. m_{z-1}, s_z .. s_{z+mid}) for mid <= z < depth - mid, and the
 * first / last mid slices of the stack pass through raw.  One launch (per channel) walks the block [z0, z0 + n) in z with
 * that state in registers: every raw plane is read once, only the hardened class byte is written.
 *   planes_dev   DEVICE array of n_planes pointers: raw (C,H,W) f32 probabilities of slices z0, z0+1, ... — the block and
 *                its look-ahead halo, n_planes >= min(n + mid, depth - z0)
 *   carry_in_dev DEVICE array of mid pointers: the filtered planes z0-mid .. z0-1 (ignored / may be NULL when z0 == 0)
 *   carry_out_dev DEVICE array of mid pointers or NULL: receives the filtered planes z0+n-mid .. z0+n-1 (n >= mid)
 *   sem8_out     (n, sem8_stride) uint8 class maps: C == 1: p >= confidence_thr; C > 1: first arg-max (NaN counts as max)
 *   best_scratch (n, hw) f32, only for C > 1
 *   need         optional: coarse (n, stride) uint8 maps, zeroed by the caller, in which the kernel marks every cell
 *                (y >> shift, x >> shift) that holds a pixel of a thing class — the only cells whose nearest-center id
 *                get_panoptic_seg ever looks at (engines.py:283-285); emp_stack_block skips the search elsewhere
 * A NaN in a window gives a NaN median, as torch.median does. */
typedef struct emp_need_map {
    uint8_t* map;                   /* device (n, stride) */
    size_t stride;
    int32_t W, shift, wc;           /* plane width, log2 of the cell size, cells per coarse row */
    uint64_t thing_bits;            /* bit c set <=> class c (< 64) is a thing class */
} emp_need_map;
int emp_median_chain(const float* const* planes_dev, int n, int n_planes, int z0, int depth, int ks, int C, size_t hw,
                     const float* const* carry_in_dev, float confidence_thr, uint8_t* sem8_out, size_t sem8_stride,
                     float* best_scratch, float* const* carry_out_dev, const emp_need_map* need /* host, may be NULL */,
                     void* stream);

/* The same block (C == 1) re-run from a corrected carry: per pixel, both chains (old carry, new carry) advance together
 * until their states agree bit for bit; only that prefix of sem8 is rewritten.  Pixels still apart at the block's end
 * store the new state into carry_out_dev and set *changed (device int32, zeroed by the caller) — the z-sharded stack
 * starts every rank's chain from a guessed carry and repairs it once the true one arrives (inference/stack.py). */
int emp_median_chain_repair(const float* const* planes_dev, int n, int n_planes, int z0, int depth, int ks, size_t hw,
                            const float* const* carry_old_dev, const float* const* carry_new_dev, float confidence_thr,
                            uint8_t* sem8, size_t sem8_stride, float* const* carry_out_dev, int32_t* changed,
                            const emp_need_map* need /* host, may be NULL */, void* stream);

/* pan_seg_to_rle_seg — empanada/inference/rle.py:26-86 (+ connected_components :18-24,
 * array_utils.rle_encode array_utils.py:209-235).  One pass over the pixels extracts row-runs
 * (maximal horizontal segments of one selected value); everything after that works on runs:
 * run-based 8-connected union-find for thing classes when force_connected (root = first run in
 * raster order, so components number themselves in raster-first order), instance slots in the
 * reference's output order (class order of `labels`, then ascending instance label), run merging
 * across row ends (a run continues from column W-1 into column 0 of the next row), boxes.
 *   pan (H,W) int64;  labels / thing_list: host arrays; labels >= 0, 0 < label_divisor <= 2^22.
 * Results (device, capacity-bounded; true counts in status[EMP_ST_NROWRUNS/NRUNS/NINST], and
 * EMP_FLAG_RLE_OVERFLOW is set if a capacity was exceeded — retry with the counts):
 *   runs_out  (run_cap,3) int64: start (flat index), length, instance slot — ascending start
 *   inst_out  (inst_cap,8) int64: class label, instance label, y0, x0, y1, x1, n runs, 0 —
 *             slot order IS the reference's dict order; a stable group-by-slot of runs_out gives
 *             each instance's starts / runs arrays. */
size_t emp_rle_workspace_bytes(int H, int W, int run_cap, int n_labels, int64_t label_divisor);
int emp_rle(const int64_t* pan, int H, int W, const int64_t* labels /* host */, int n_labels,
            int64_t label_divisor, const int64_t* thing_list /* host */, int n_things,
            int force_connected, int64_t* runs_out, int run_cap, int64_t* inst_out, int inst_cap,
            void* ws, size_t ws_bytes, void* stream);

/* B z-slices of the stack path behind the median queue, one launch per kernel for the whole block and no host
 * synchronisation: get_instance_cells on the coarse maps (engines.py:257-272) -> get_panoptic_seg with the nearest
 * upsample folded in (:274-292) -> crop to the unpadded size (:305-307) -> pan_seg_to_rle_seg's tables (rle.py:26-86),
 * cut straight from the 16-bit code map + label LUT (the int64 label map is never written).
 *   sem8 (B, sem8_stride) uint8 hardened classes (emp_median_chain); hm (B, hm_stride) f32 (h,w) heat-maps;
 *   off (B, off_stride) f32 (2,h,w) offsets, with (h << shift, w << shift) >= (H, W)
 *   need  optional (B, need_stride) uint8 (h,w) maps from emp_median_chain: the nearest-center search runs only in cells
 *         marked there (NULL: everywhere, as the reference does — the ids differ only where nobody reads them)
 *   scratch   device buffer of emp_stack_block_scratch_bytes(), 256-byte aligned, reusable by the next block
 *   packed_out  device int64[emp_stack_block_packed_words()], everything the host needs in ONE contiguous prefix:
 *       [0] B  [1] R = row-runs of all slices  [2] I = instances of all slices  [3] EMP_BLK_INST_WORDS
 *       [EMP_BLK_HDR_MAXLAB + i]  largest (label - labels[i] * label_divisor) in the block
 *       per slice b at EMP_BLK_HDR_WORDS + EMP_BLK_SLICE_WORDS * b:
 *           n instances, first instance row, first run row, row-runs, OR of the EMP_FLAG_* bits, K (centers found)
 *       starts[R'], lengths[R'] as INT32 (R' = R rounded up to even, so R' / 2 words each; flat indices into the
 *           crop_h x crop_w map, which is < 2^31 pixels; grouped by instance, ascending inside one)
 *       I instance rows of EMP_BLK_INST_WORDS int64: class label, instance label, y0, x0, y1, x1, n runs, first run
 *           row (into starts / lengths), area — slices in order, instances in the reference's dict order
 *   runs3_out  optional device (B, run_cap, 3) int64: (start, length, instance slot) of every row-run in ascending
 *              start order — the layout emp_rle_pair_overlaps / emp_fill_runs consume. */
typedef struct emp_stack_cfg {
    int32_t H, W, h, w;                 /* padded plane, coarse maps */
    int32_t shift;                      /* log2 of the coarse -> full upsampling */
    int32_t nms_kernel, k_cap;
    int32_t crop_h, crop_w;
    int32_t n_things, n_labels, force_connected;
    int32_t run_cap, inst_cap;          /* row-runs / instances per slice */
    float nms_threshold, step;
    int64_t label_divisor, stuff_area, void_label;
    const int64_t* thing_list;          /* host */
    const int64_t* labels;              /* host */
} emp_stack_cfg;
#define EMP_BLK_HDR_MAXLAB 4
#define EMP_BLK_MAXLAB_OVERFLOW (1ll << 62)   /* emp_stack_blocks: maxlab_all[0] is raised to this if any slice overflowed a table */
#define EMP_BLK_HDR_WORDS (4 + EMP_MAX_LABELS)
#define EMP_BLK_SLICE_WORDS 6
#define EMP_BLK_INST_WORDS 9
size_t emp_stack_block_scratch_bytes(const emp_stack_cfg* cfg, int B);
size_t emp_stack_block_packed_words(const emp_stack_cfg* cfg, int B);
int emp_stack_block(const emp_stack_cfg* cfg, int B, const uint8_t* sem8, size_t sem8_stride, const float* hm,
                    size_t hm_stride, const float* off, size_t off_stride, const uint8_t* need, size_t need_stride,
                    void* scratch, size_t scratch_bytes, int64_t* packed_out, size_t packed_words, int64_t* runs3_out,
                    void* stream);

/* Cross-slice matcher support — empanada/inference/matcher.py:136-232 (rle_matcher) with
 * array_utils.rle_intersection :371-403 / rle_iou :405-429 / rle_ioa :431-449: pixel overlaps between
 * the instances of consecutive slices, from the run tables emp_rle wrote.  For pair p = (slice p,
 * slice p+1) every overlapping pair of runs contributes one row (p, slot in slice p, slot in slice
 * p+1, overlap in pixels); summing rows with equal (p, slot, slot) gives the intersection the
 * reference computes per instance pair (IoU = inter / (area_a + area_b - inter), IoA = inter / area_b).
 *   runs    (n_slices, run_stride, 3) int64 rows (start, length, slot); the first n_runs[s] rows of
 *           slice s in ascending start order and disjoint (emp_rle's runs_out)
 *   n_runs  device int32[n_slices];  max_runs >= every n_runs[s] (sizes the grid)
 *   out     (cap, 4) int32 rows, unordered, 16-byte aligned;  count: device int32, rows found — if it
 *           exceeds cap the surplus rows were dropped: retry with cap >= count. */
int emp_rle_pair_overlaps(const int64_t* runs, size_t run_stride, const int32_t* n_runs, int n_slices,
                          int max_runs, int32_t* out, int cap, int32_t* count, void* stream);

/* Orthoplane consensus support — consensus.object_iou_graph (consensus.py:233-287): overlaps between
 * the objects of two trackers.  runs_a / runs_b: (n, 3) int64 rows (start, length, slot) in ascending
 * start order; objects of one list may overlap each other; lmax_a = longest run of A.  Output rows
 * (0, slot_a, slot_b, overlap) as for emp_rle_pair_overlaps; count must be re-checked against cap. */
int emp_rle_list_overlaps(const int64_t* runs_a, int n_a, int64_t lmax_a, const int64_t* runs_b, int n_b,
                          int32_t* out, int cap, int32_t* count, void* stream);

/* Dense fill of a z-block — array_utils.numpy_fill_instances (array_utils.py:725-736) over the run
 * tables in HBM: every run (start, length, slot) of slice s paints labels[s][slot] over
 * [start, start + length) of plane s of `out`; voxels no run covers keep their value; a negative label
 * skips the run.
 *   runs / n_runs / max_runs as for emp_rle_pair_overlaps;  labels (n_slices, label_stride) int64;
 *   out (n_slices, plane) of 4-byte (uint32) or 8-byte (int64) elements. */
int emp_fill_runs(const int64_t* runs, size_t run_stride, const int32_t* n_runs, int n_slices, int max_runs,
                  const int64_t* labels, size_t label_stride, void* out, int elem_bytes, size_t plane, void* stream);

/* A whole z-block through emp_stack_block, SB slices at a time, in ONE call: block j = slices [j * SB, ...) writes its
 * packed tables to packed_all + j * packed_words, and the first host_words words of them are copied to the PINNED host
 * buffer host_out + j * host_stride on copy_stream while the next block runs on stream; then header word 0 of the block
 * (its slice count, never 0) is copied to host_flags[j] (pinned, zeroed by the caller): copies on one stream land in
 * order, so a host thread that reads host_flags[j] != 0 may parse block j.  If the header says the block holds more than
 * host_words words, the caller fetches the rest from packed_all.  maxlab_all (device, EMP_MAX_LABELS words, zeroed by the
 * caller; may be NULL) accumulates the per-class maxima of the blocks' headers — what the z-sharded stack all-gathers;
 * a slice whose EMP_FLAG_K_OVERFLOW / EMP_FLAG_RLE_OVERFLOW bit is set raises maxlab_all[0] to EMP_BLK_MAXLAB_OVERFLOW. */
int emp_stack_blocks(const emp_stack_cfg* cfg, int n, int SB, const uint8_t* sem8, size_t sem8_stride, const float* hm,
                     size_t hm_stride, const float* off, size_t off_stride, const uint8_t* need, size_t need_stride,
                     void* scratch, size_t scratch_bytes, int64_t* packed_all, size_t packed_words, int64_t* runs3_all,
                     int64_t* maxlab_all, int64_t* host_out /* pinned host */, size_t host_stride, size_t host_words,
                     int64_t* host_flags /* pinned host */, void* stream, void* copy_stream);

/* Slices of a (D,H,W) volume resident in HBM — array_utils.take (array_utils.py:6-23) as data/volume_dataset.py:37-53
 * calls it inside the stack / orthoplane loops of scripts/pdl_inference3d.py:110-176: the n slices i0 .. i0+n-1 along
 * `axis` as one contiguous (n, A, B) batch, (A,B) = (H,W), (D,W), (D,H) for axis 0, 1, 2.  Elements of 1 byte (uint8
 * images) or 4 bytes.  The axis-2 (yz) gather reads runs of n neighbouring slices and transposes them through shared
 * memory, so taking slices in batches of >= 32 keeps it sector-efficient. */
int emp_take_slices(const void* vol, int elem_bytes, int D, int H, int W, int axis, int i0, int n, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMPANADA_B200_H */
