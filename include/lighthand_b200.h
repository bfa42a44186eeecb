/*
 * lighthand_b200.h -- C ABI of liblighthand_b200.so (sm_100a).
 *
 * B200-native implementation of the LightHand heatmap pipeline
 *     crop/warp + normalise -> Gaussian targets -> JointsMSELoss fwd/bwd -> argmax decode -> PCK/AUC/EPE
 * The reference (leejeongho3214/LightHand) is pure Python and has no FFI; the interface each
 * entry point replaces is the Python function cited next to it (paths relative to the reference
 * repo root).  INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - Plain C: pointers, sizes, scalars, two POD structs.  No torch / C++ types.
 *   - "dev" pointers are device pointers owned by the caller; nothing is allocated or freed by
 *     the lh_* kernels entry points (the lh_pipeline_* context owns its own buffers).
 *   - Every kernel entry point is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - Return value: 0 success; <0 argument error (LH_E*); >0 a cudaError_t from the launch.
 *   - Re-entrant; no global mutable state; safe from several host threads on distinct streams.
 *   - There is no CPU fallback anywhere behind this ABI.
 */
#ifndef LIGHTHAND_B200_H
#define LIGHTHAND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LH_ABI_VERSION 4

#define LH_OK            0
#define LH_EINVAL       (-1)  /* bad size / enum value */
#define LH_ENULL        (-2)  /* required pointer is NULL */
#define LH_EALIGN       (-3)  /* pointer / pitch alignment requirement not met */
#define LH_EUNSUPPORTED (-4)  /* geometry outside what the kernels implement */
#define LH_ENOMEM       (-5)  /* workspace too small / allocation failed (pipeline ctx) */
#define LH_ENODEVICE    (-6)  /* no sm_100 device */
#define LH_ENCCL        (-7)  /* NCCL missing (libnccl.so.2 could not be loaded) or an NCCL call failed */
#define LH_EJPEG        (-8)  /* nvJPEG missing (libnvjpeg.so.12), backend unavailable, or a bit-stream it cannot decode */

typedef void* lh_stream_t; /* cudaStream_t */

int lh_abi_version(void);
const char* lh_strerror(int code);
/* device properties the host side needs for grid sizing / reporting */
int lh_device_query(int device, int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem);

/* ------------------------------------------------------------------------------------------
 * K1  warp + normalise.
 * Replaces: generate_patch_image            src/utils/preprocessing.py:144-162
 *           (cv2.warpAffine INTER_LINEAR, BORDER_CONSTANT 0; flip src/utils/preprocessing.py:153-155)
 *           colour gain + clip              src/utils/preprocessing.py:81
 *           ToTensor / Resize / Normalize   src/tools/dataset.py:147-157
 * ------------------------------------------------------------------------------------------ */
#define LH_INTERP_CV2_U8   0 /* cv2 fixed-point grid, uint8 rounding ((sum p*w + 2^14)>>15)     */
#define LH_INTERP_CV2_F32  1 /* cv2 fixed-point grid, float32 weights (reference feeds float32) */
#define LH_INTERP_TORCH    2 /* torch bilinear resize, align_corners=False, edge clamp          */
#define LH_CHAN_RGB 0
#define LH_CHAN_BGR 1        /* src/datasets/frei_dataloader.py:338 returns channels (2,1,0)    */
#define LH_LAYOUT_CHW 0      /* network input  [N,3,out_h,out_w]                                */
#define LH_LAYOUT_HWC 1      /* image patch    [N,out_h,out_w,3]  (generate_patch_image result) */

/* Normalisation table for LH_INTERP_CV2_U8: lut[i][c][v] = ((clip(v*gain[i][c],0,255))/255 - mean[c])/std[c]
 * evaluated with IEEE float32/float64 ops in torch's order (bit-exact).  chan_gain dev [N,3] double or
 * NULL (then one table, N ignored).  mean_std host [6] = mean[3],std[3]; NULL -> identity table
 * (value 0..255 after gain/clip, no /255).  lut dev [(chan_gain?N:1)][3][256] float.
 *
 * lh_warp_normalize, LH_INTERP_CV2_U8: the warped uint8 pixel is always cv2's, bit for bit.  With `lut` the
 * float result is the table entry (bit-identical to ToTensor + Normalize); with lut == NULL it is one fused
 * multiply-add v*a + b, a = 1/(255 std), b = -mean/std (<= 2 ulp from the former; mean_std == NULL: the raw
 * 0..255 value).  chan_gain requires the table.  LH_INTERP_TORCH computes torch's bilinear resize
 * (align_corners = False, no antialias) to ~3 ulp (exact operation order when flip / chan_gain is given). */
int lh_norm_lut_build(const double* chan_gain, int N, const float* mean_std, float* lut, lh_stream_t stream);

int lh_warp_normalize(const uint8_t* src,        /* dev N HWC RGB uint8 images                   */
                      int64_t src_img_stride,    /* bytes between images                          */
                      int src_h, int src_w, int src_pitch /* bytes per row */,
                      const double* fwd_affine,  /* dev [N,2,3] src->dst as cv2.warpAffine takes it; unused for LH_INTERP_TORCH */
                      const uint8_t* flip,       /* dev [N] 0/1 or NULL                           */
                      const double* chan_gain,   /* dev [N,3] or NULL                             */
                      const float* mean_std,     /* host [6] or NULL (identity)                   */
                      const float* lut,          /* dev table from lh_norm_lut_build, or NULL     */
                      int interp, int chan_order, int out_layout,
                      float* dst, int N, int out_h, int out_w, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2  Gaussian targets.
 * Replaces: CustomDataset.generate_target   src/tools/dataset.py:165-212  (style 0)
 *           GenerateHeatmap.__call__        src/utils/dataset_loader.py:22-53 (style 1; input = joints/4)
 * gauss_table: dev side*side float32, computed by the caller exactly as the reference does with
 * numpy (bit-exact targets).  style 0: side = 2*tmp+1.  style 1: side = 6*sigma+3, sigma = H/64 integer.
 * target_weight dev [N,J] or NULL.
 * use_vis = 0: the reference's behaviour -- target_weight = 1 where the Gaussian window meets the map, else 0
 * (computed and discarded at src/tools/dataset.py:171,185,212; visibility never consulted).
 * use_vis = 1 (joint_stride >= 3): upstream-HRNet semantics the north star names -- target_weight = visibility
 * (third joint column) * in-bounds, and the Gaussian is drawn only where that is > 0.5 (oracle.restate.generate_target
 * use_vis=True; parity unpinned: not in the reference).
 * ------------------------------------------------------------------------------------------ */
#define LH_TARGET_HRNET     0
#define LH_TARGET_HOURGLASS 1
int lh_gen_targets(const float* joints, int joint_stride /*floats per joint: 2 or 3*/,
                   int N, int J, int H, int W, float stride_x, float stride_y,
                   int style, int use_vis, const float* gauss_table, int table_side,
                   float* target, float* target_weight, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3  JointsMSELoss forward + backward.   Replaces src/utils/loss.py:306-325 (+ autograd backward).
 * loss = 0.5/(N*J*HW) * sum((pred*w - target*w)^2), w = weight[n,j] or 1 (the reference ignores
 * target_weight: pass NULL).  grad (nullable) = d loss / d pred.  loss_f64 / loss_f32: dev [1], nullable.
 * ws: dev workspace of lh_heatmap_ws_bytes(N,J) bytes (per-CTA partial slots + the words through which the pieces of
 * a split map meet), zero-initialised ONCE by the caller (the
 * kernels leave it zeroed again).
 * ------------------------------------------------------------------------------------------ */
size_t lh_heatmap_ws_bytes(int64_t N, int J);
int lh_jmse_loss(const float* pred, const float* target, const float* weight,
                 int64_t N, int J, int H, int W,
                 double* loss_f64, float* loss_f32, float* grad,
                 void* ws, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  decode.   Replaces get_max_preds src/utils/loss.py:327-355 and the callers' "* 4"
 * (src/utils/method.py:173-178).  First-index tie-break, NaN counts as maximum, coordinates zeroed
 * where maxval <= 0.  subpixel / inv_affine: upstream-HRNet quarter-pixel shift and inverse crop
 * matrix (not in the reference; see oracle.restate.decode_final).
 * coords dev [N,J,2]; maxvals dev [N,J] (nullable); argidx dev [N,J] int32 (nullable).
 * ------------------------------------------------------------------------------------------ */
int lh_decode(const float* heat, int64_t N, int J, int H, int W, float out_scale, int subpixel,
              const double* inv_affine /*dev [N,2,3] or NULL*/,
              float* coords, float* maxvals, int32_t* argidx, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2+K3+K4(+K5) fused post-backbone pass: pred is read ONCE; target is generated on the fly from the
 * joints (and written if `target_out` != NULL), loss/grad and decode come out of the same pass.  With
 * metric_counts / metric_sums set, PCK_2d_loss + EPE_train of the decoded joints against `joints`
 * (src/utils/loss.py:50-67,116-148; all joints, no visibility) are folded into the accumulators of
 * lh_pck_epe_batch by the last CTA of the same launch.
 * ------------------------------------------------------------------------------------------ */
#define LH_PCK_PROPORTION 0
#define LH_PCK_MM         1
typedef struct lh_heatmap_args {
    const float* pred;          /* dev [N,J,H,W]                                           */
    const float* joints;        /* dev [N,J,joint_stride]  (input-image pixel coordinates) */
    const float* gauss_table;   /* dev side*side                                           */
    const float* target_in;     /* dev [N,J,H,W] or NULL; if set, used instead of joints   */
    const double* inv_affine;   /* dev [N,2,3] or NULL                                     */
    float* target_out;          /* dev [N,J,H,W] or NULL                                   */
    float* target_weight_out;   /* dev [N,J] or NULL                                       */
    float* grad;                /* dev [N,J,H,W] or NULL                                   */
    double* loss_f64;           /* dev [1] or NULL                                         */
    double* loss_meter;         /* dev [2] or NULL: += {loss*N, N}  (AverageMeter.update)  */
    float* loss_f32;            /* dev [1] or NULL                                         */
    float* coords;              /* dev [N,J,2] or NULL (no decode)                         */
    float* maxvals;             /* dev [N,J] or NULL                                       */
    int32_t* argidx;            /* dev [N,J] or NULL                                       */
    void* ws;                   /* dev workspace, lh_heatmap_ws_bytes(N,J)                 */
    unsigned long long* metric_counts; /* dev u64[4] or NULL, see lh_pck_epe_batch          */
    double* metric_sums;        /* dev f64[4] or NULL                                      */
    int64_t N;
    int32_t J, H, W;
    int32_t joint_stride;
    int32_t table_side;
    int32_t style;              /* LH_TARGET_*                                             */
    int32_t use_target_weight;  /* 0 = reference behaviour (weight ignored by the loss)    */
    int32_t subpixel;
    float stride_x, stride_y;
    float out_scale;
    float pck_thr;              /* threshold as float32: T ('proportion') or T*3.78 ('mm') */
    int32_t pck_mode;           /* LH_PCK_*                                                */
    int32_t use_vis;            /* 1: target_weight = visibility * in-bounds (see lh_gen_targets) */
} lh_heatmap_args;
int lh_heatmap_fused(const lh_heatmap_args* args, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5  metrics.
 * lh_pck_epe_batch replaces PCK_2d_loss / PCK_2d_loss_visible (src/utils/loss.py:83-148) and
 * EPE_train / EPE (src/utils/loss.py:28-67) for one batch; results ACCUMULATE into
 *   counts dev u64[4] = { n_wrong (d > T), n_total, epe_count, n_images }
 *   sums   dev f64[4] = { epe_sum, sum(loss_b * n_b), sum(n_b), reserved }   ([1],[2] via lh_heatmap_args.loss_meter)
 * so a whole epoch needs one device->host read (AverageMeter, src/utils/metric_logger.py:7-28);
 * PCK average = (n_total - n_wrong) / n_total, EPE average = epe_sum / epe_count.
 * thr_f32: the threshold already rounded the way torch does it: (float)T for 'proportion',
 * (float)(T*3.78) for 'mm'.
 * ------------------------------------------------------------------------------------------ */
int lh_pck_epe_batch(const float* pred_xy /*dev [N,J,2]*/, const float* gt /*dev [N,J,gt_stride]*/, int gt_stride,
                     int64_t N, int J, float thr_f32, int mode, int visible,
                     unsigned long long* counts, double* sums, lh_stream_t stream);

/* lh_pck_curve_accum replaces the counting core of pred_eval / pred_test
 * (src/utils/argparser.py:326-438): float64 distances, strict '<', per-category histograms.
 *   thr_sorted dev f64[K] ascending.  hist dev u64[n_cat][K+1]: bin k counts distances whose first
 *   satisfied threshold is k (bin K: none / NaN) -- count(d < thr[k]) = prefix sum.
 *   n_valid dev u64[n_cat] (joints passing the visibility mask), n_all dev u64[n_cat] (all joints),
 *   sum_d dev f64[n_cat] (pixel distance over ALL joints, the EPE numerator).
 *   bbox_diag dev f32[N] or NULL (then computed as pred_store does, src/utils/argparser.py:268-272).
 *   category dev i32[N] or NULL (all 0).  method 0 = pixels/mm thresholds, 1 = pckb (d / bbox).
 *   use_vis: keep joints whose gt[..., gt_stride-1] == 1 (pred_eval) or all joints (pred_test). */
int lh_pck_curve_accum(const float* pred_xy, const float* gt, int gt_stride,
                       const float* bbox_diag, const int32_t* category, int n_cat,
                       int64_t N, int J, const double* thr_sorted, int K, int method, int use_vis,
                       unsigned long long* hist, unsigned long long* n_valid,
                       unsigned long long* n_all, double* sum_d, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Batched keypoint transform.  Replaces the keypoint half of augmentation()  src/utils/preprocessing.py:83-90:
 *   if flipped: x = orig_w - x - 1 (float32), left / right joints (and their validity) swapped;
 *   (x, y) = trans_point2d((x, y), trans) = float64 dot of the float32 2x3 crop matrix with (x, y, 1)
 *            (src/utils/preprocessing.py:207-210), stored back as float32;
 *   joint_valid *= (0 <= x < out_w) * (0 <= y < out_h).
 * joints dev [N,J,joint_stride] (x, y[, visibility]); valid_in dev [N,J] or NULL (then the visibility column when
 * joint_stride >= 3, else 1); trans dev [N,2,3] double holding the float32 matrix values; flip dev [N] or NULL;
 * swap dev int32 [J] or NULL: swap[j] = the joint whose data slot j receives on a flip (an involution; identity
 * for a single hand).  joints_out dev [N,J,out_stride] (x, y[, valid]); valid_out dev [N,J] or NULL.
 * ------------------------------------------------------------------------------------------ */
int lh_transform_joints(const float* joints, int joint_stride, const float* valid_in, const double* trans,
                        const uint8_t* flip, int orig_w, const int32_t* swap, int64_t N, int J, int out_w, int out_h,
                        float* joints_out, int out_stride, float* valid_out, lh_stream_t stream);

/* bounding-box diagonal of the ground-truth joints of each sample, float32, exactly as the per-sample Python loops do it
 * (builtin max / min; src/utils/argparser.py:268-272, src/utils/loss.py:122-127).  gt dev [N,J,gt_stride] -> out dev [N]. */
int lh_bbox_diag(const float* gt, int gt_stride, int64_t N, int J, float* out, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6  metric all-reduce: the path's only collective.  The batch is data-partitioned, one
 * process per GPU; ranks exchange nothing but the SUM of the additive accumulators above -- the state of the
 * reference's AverageMeters (src/utils/metric_logger.py:7-28, src/utils/method.py:241-250) and of pred_eval's
 * counting loops (src/utils/argparser.py:347-378); src/utils/comm.py (never initialised in the reference) is the
 * role it takes.  NCCL over NVLink 5 / NVSwitch, in place, stream-ordered (enqueue it right after the last K5 on
 * the same stream: no host synchronisation), CUDA-graph capturable.  Integer counts are order-independent:
 * N ranks give the 1-rank result bit for bit; float64 sums differ by reduction order only.
 *   lh_comm_unique_id   rank 0 makes the 128-byte id and ships it to the other ranks by any means (MPI, sockets,
 *                       the torch.distributed store: lighthand_b200/distributed.py)
 *   lh_comm_init        collective over all ranks (ncclCommInitRank on `device`)
 *   lh_metrics_allreduce        counts u64[n_counts] and sums f64[n_sums] as ONE grouped NCCL operation
 *   lh_metrics_allreduce_multi  the same for any number of u64 / f64 buffers (curve histograms of an evaluation
 *                               sweep next to the loop counters): one grouped operation, no packing copies
 * NCCL is bound at run time; without libnccl.so.2 these four return LH_ENCCL and everything else works.
 * ------------------------------------------------------------------------------------------ */
#define LH_COMM_ID_BYTES 128
int lh_comm_unique_id(void* id_out /* host, LH_COMM_ID_BYTES */);
int lh_comm_init(const void* id, int world, int rank, int device, void** comm_out);
int lh_comm_destroy(void* comm);
int lh_metrics_allreduce(void* comm, unsigned long long* counts, int n_counts, double* sums, int n_sums,
                         lh_stream_t stream);
int lh_metrics_allreduce_multi(void* comm, int n_bufs, void* const* bufs /* dev */, const size_t* elems,
                               const int* is_f64 /* 0: uint64, 1: float64 */, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FreiHAND image path.  Replaces, batched:
 *   crop / myimrotate / myimresize        src/utils/image_ops.py:25-56, 97-131
 *   HandMeshTSVDataset.rgb_processing     src/datasets/frei_dataloader.py:134-147   (flip is always 0, :107)
 *   Resize((256,256)) + Normalize + [(2,1,0)]   src/datasets/frei_dataloader.py:279-283, 338
 * crop()'s integer window bookkeeping (transform(..., invert=1) via torch.inverse, int truncation, rotation
 * padding; image_ops.py:99-106) is host logic and arrives per sample in lh_frei_desc
 * (lighthand_b200/utils/image_ops.py fills it exactly as the reference computes it).
 * Accuracy: the rotation stage is bit-identical to cv2 (fixed-point grid, exact float32 sums); the cv2.resize
 * stage floors its coordinates in double like OpenCV and blends in float32: <= 1e-4 on the 0..255 scale
 * against the reference's float64 (bar for this path: 1 LSB).
 * ------------------------------------------------------------------------------------------ */
typedef struct lh_frei_desc {
    int32_t ulx, uly;       /* window origin in the source image (rotation padding included)            */
    int32_t cw, ch;         /* canvas size br - ul (rotation padding included)                           */
    int32_t pad;            /* rotation padding (image_ops.py:103); ignored unless rotated               */
    int32_t rotated;        /* rot != 0                                                                  */
    double iM[6];           /* rotated: cv2.getRotationMatrix2D(((cw-1)/2, (ch-1)/2), rot, 1.0) inverted in double the
                               way cv2.warpAffine inverts it (dst -> src, row-major)                       */
    double scale_x, scale_y;/* cv2.resize's 1.0 / ((double)res / iw), 1.0 / ((double)res / ih) with
                               iw = cw - 2 pad, ih = ch - 2 pad (pad counted only when rotated)            */
    float pn[3];            /* per-channel pixel noise (frei_dataloader.py:118)                          */
    int32_t reserved;
    int64_t scratch_off;    /* float offset of this sample's rotated canvas [(ch-2pad),(cw-2pad),3] in `scratch` */
} lh_frei_desc;

/* rgb_processing for N images: out dev [N,3,res,res] float32 in 0..1 (CHW, /255, noise applied).
 * desc dev [N]; rot_index dev int32 [n_rot] = indices of the samples with rotated != 0 (NULL: every sample is
 * visited and the flag decides); scratch dev float (sum of the rotated canvases) or NULL when no sample is rotated;
 * max_rot_w / max_rot_h = max over rotated samples of cw-2pad / ch-2pad (both 0: no sample is rotated). */
int lh_frei_rgb_processing(const uint8_t* src, int64_t src_img_stride, int src_h, int src_w, int src_pitch,
                           const lh_frei_desc* desc, int N, int res, const int32_t* rot_index, int n_rot,
                           float* scratch, int max_rot_w, int max_rot_h, float* out, lh_stream_t stream);

/* The two calls above + below in ONE pass over the output (round 2): rgb_processing's res x res float image stays in
 * shared memory (32 x 32 output tiles; 308 MB less HBM traffic per 256-image batch), results bit-identical to
 * lh_frei_rgb_processing followed by lh_resize_normalize_f32.  out dev [N,3,out_h,out_w]; LH_EUNSUPPORTED when
 * res / out > 0.9375 (224 -> 256 is 0.875): use the two calls then. */
int lh_frei_fused(const uint8_t* src, int64_t src_img_stride, int src_h, int src_w, int src_pitch,
                  const lh_frei_desc* desc, int N, int res, const int32_t* rot_index, int n_rot,
                  float* scratch, int max_rot_w, int max_rot_h, const float* mean_std, int chan_order,
                  float* out, int out_h, int out_w, lh_stream_t stream);

/* transforms.Resize((out_h,out_w)) on a float CHW batch (bilinear, align_corners=False, no antialias; ATen's
 * float32 operation order, within ~1 ulp of torch) + Normalize as one multiply-add (mean_std host [6] or NULL;
 * <= 2 ulp from (x - mean) / std) + channel order.
 * src dev [N,3,src_h,src_w]; dst dev [N,3,out_h,out_w]. */
int lh_resize_normalize_f32(const float* src, int N, int src_h, int src_w, const float* mean_std, int chan_order,
                            float* dst, int out_h, int out_w, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ColorJitter + Normalize on float CHW images in 0..1 (SURVEY 8 a-2 augmentation branch / f-2).
 * Replaces transforms.ColorJitter(0.5,0.5,0.5,0.5) -> transforms.Normalize  src/tools/dataset.py:139-145
 * (torchvision's tensor implementation: _blend, rgb_to_grayscale, _rgb2hsv/_hsv2rgb, same float32 operation order).
 * RNG contract: the draw stays on the host in torchvision's order (torch.randperm(4), then uniform draws for
 * brightness, contrast, saturation, hue: lighthand_b200/tools/dataset.py color_jitter_params); the device applies it.
 * ------------------------------------------------------------------------------------------ */
#define LH_JITTER_PARTS 32   /* partial sums per image of the contrast step's mean grey level */
typedef struct lh_jitter_params {
    int32_t order[4];       /* torchvision fn ids in application order: 0 brightness 1 contrast 2 saturation 3 hue; -1 skips */
    float factor[4];        /* indexed by fn id: brightness, contrast, saturation factors; hue shift in [-0.5, 0.5]            */
    float one_minus[4];     /* float(1.0 - factor) as Python computes it (double subtraction, then float32); unused for hue    */
    int32_t apply;          /* 0: this image is only normalised (idx >= len * ratio_of_aug, src/tools/dataset.py:134)          */
} lh_jitter_params;
/* src dev [N,3,H,W]; params dev [N]; mean_std host [6] or NULL; dst dev [N,3,H,W] (may alias src);
 * ws dev double [N * LH_JITTER_PARTS]. */
int lh_color_jitter_normalize(const float* src, const lh_jitter_params* params, int N, int H, int W,
                              const float* mean_std, float* dst, double* ws, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f-4  JPEG decode on the device: N baseline JPEG bit-streams (host memory) -> dst dev
 * [N, h, w, 3] RGB uint8 (HWC: what lh_warp_normalize / lh_frei_rgb_processing read), asynchronous on `stream`.
 * Replaces img_from_base64 (src/utils/tsv_file_ops.py:21-28: cv2.imdecode) + the BGR2RGB of get_image
 * (src/datasets/frei_dataloader.py:195-204); the base64 / TSV side is lighthand_b200/utils/tsv_file.py.
 * Decoder = nvJPEG, bound at run time; not bit-identical to libjpeg-turbo (IDCT / chroma upsampling accuracy is not
 * part of the JPEG standard): the tolerance is stated in DESIGN.md and tested in tests/test_gpu_jpeg.py.
 * backend: nvJPEG's (DEFAULT, HYBRID = Huffman on the CPU, GPU_HYBRID = Huffman on the GPU, HARDWARE = the NVJPG engine).
 * cpu_threads: with HYBRID and cpu_threads > 1 the host Huffman stage runs on that many worker threads (nvJPEG's decoupled
 * API, one decoder state per thread, images handed out dynamically; the call returns when the batch is decoded); ignored
 * by the other backends.  The pixels do not depend on the thread count.  One lh_jpeg serves one caller at a time.
 * Every bit-stream of a batch must have the size w x h (else LH_EINVAL).  The data / sizes arrays and the bit-streams
 * must stay valid until the stream has passed the call.
 * ------------------------------------------------------------------------------------------ */
#define LH_JPEG_BACKEND_DEFAULT    0
#define LH_JPEG_BACKEND_HYBRID     1
#define LH_JPEG_BACKEND_GPU_HYBRID 2
#define LH_JPEG_BACKEND_HARDWARE   3
typedef struct lh_jpeg lh_jpeg;
int lh_jpeg_create(int device, int backend, int cpu_threads, lh_jpeg** out);
void lh_jpeg_destroy(lh_jpeg* j);
int lh_jpeg_decode_batch(lh_jpeg* j, const uint8_t* const* data, const size_t* sizes, int n, int w, int h,
                         uint8_t* dst, int64_t dst_img_stride, int dst_pitch, lh_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer pipeline context (what a non-torch host binds; also the e2e benchmark path).
 * Owns device buffers, pinned staging and three streams; one step is
 *   data stage   H2D -> K1 -> input[slot]                       (copy + data streams)
 *   loss stage   [backbone] -> fused(K2,K3,K4,K5) -> D2H        (loss stream)
 * two kernel launches; the data stage of step i+1 overlaps the loss stage of step i, as the data loader of the
 * reference's loop runs ahead of the model (src/utils/method.py:160-183).
 * ------------------------------------------------------------------------------------------ */
typedef struct lh_pipeline_cfg {
    int32_t device;
    int32_t batch;              /* N */
    int32_t src_h, src_w;       /* uint8 RGB HWC source size */
    int32_t out_h, out_w;       /* network input size */
    int32_t J, H, W;            /* joints, heat-map size */
    int32_t interp;             /* LH_INTERP_* */
    int32_t train;              /* 1: write grad */
    int32_t materialize_target; /* 1: write the target heat-maps (what the reference's dataset returns) */
    int32_t subpixel;
    int32_t exact_norm;         /* 1: table normalisation (bit-identical to torch), 0: one FMA (<= 2 ulp) */
    float sigma;
    float pck_T;                /* in-loop PCK threshold, 'proportion' (src/utils/method.py:243) */
    float mean_std[6];
} lh_pipeline_cfg;
typedef struct lh_pipeline lh_pipeline;

typedef struct lh_pipeline_io {
    /* host inputs (pageable or pinned) */
    const uint8_t* images;      /* [N,src_h,src_w,3] */
    const double* fwd_affine;   /* [N,2,3] */
    const uint8_t* flip;        /* [N] or NULL */
    const float* joints;        /* [N,J,2] ground truth in network-input pixels */
    const float* pred;          /* host [N,J,H,W] or NULL: use the device-resident pred (lh_pipeline_pred_dev) */
    /* host outputs (nullable) */
    float* loss;                /* [1] */
    float* coords;              /* [N,J,2] */
    float* maxvals;             /* [N,J] */
    unsigned long long* counts; /* [4] running totals, see lh_pck_epe_batch */
    double* sums;               /* [4] */
} lh_pipeline_io;

int lh_pipeline_create(const lh_pipeline_cfg* cfg, const float* gauss_table_host, int table_side, lh_pipeline** out);
void lh_pipeline_destroy(lh_pipeline* p);
/* device-resident buffers of the context (backbone input / output live here) */
float* lh_pipeline_input_dev(lh_pipeline* p);            /* [N,3,out_h,out_w] normalised images of the LAST submitted slot */
float* lh_pipeline_input_dev_slot(lh_pipeline* p, int slot);   /* each of the two slots owns its input buffer        */
float* lh_pipeline_pred_dev(lh_pipeline* p);    /* [N,J,H,W] where the backbone writes   */
float* lh_pipeline_grad_dev(lh_pipeline* p);    /* [N,J,H,W] d loss / d pred             */
float* lh_pipeline_target_dev(lh_pipeline* p);  /* [N,J,H,W] or NULL                     */
/* one step from host buffers: copies in, runs the kernels, copies results out, returns after the
 * outputs are valid (synchronous w.r.t. the host). */
int lh_pipeline_step_host(lh_pipeline* p, const lh_pipeline_io* io);
/* asynchronous pair (two slots, double-buffered): submit step i+1 before waiting for step i so its
 * host->device copy and its K1 overlap step i's loss stage.  The io pointers must stay valid until the wait. */
int lh_pipeline_submit(lh_pipeline* p, const lh_pipeline_io* io, int* slot_out);
int lh_pipeline_wait(lh_pipeline* p, int slot);
/* the two halves of lh_pipeline_submit, for a host that runs the backbone between them:
 *   lh_pipeline_submit_inputs(p, io, &slot);                 data stage; the loss stream now waits for K1
 *   <enqueue the backbone on lh_pipeline_loss_stream(p): reads lh_pipeline_input_dev_slot(p, slot),
 *    writes lh_pipeline_pred_dev(p)>
 *   lh_pipeline_submit_loss(p, slot);                        loss stage + result copies
 *   lh_pipeline_wait(p, slot); */
int lh_pipeline_submit_inputs(lh_pipeline* p, const lh_pipeline_io* io, int* slot_out);
int lh_pipeline_submit_loss(lh_pipeline* p, int slot);
/* data stage from COMPRESSED input (f-4): `data` / `sizes` = batch JPEG bit-streams of src_w x src_h (host memory; must
 * stay valid until the slot is waited for); io->images is ignored; the context creates its lh_jpeg with `jpeg_backend`
 * on first use.  PCIe then carries the compressed bytes. */
int lh_pipeline_submit_inputs_jpeg(lh_pipeline* p, const uint8_t* const* data, const size_t* sizes, int jpeg_backend,
                                   const lh_pipeline_io* io, int* slot_out);
void* lh_pipeline_loss_stream(lh_pipeline* p);   /* cudaStream_t */
/* multi-GPU: SUM the running accumulators over `comm` (lh_comm_init), ordered after the last submitted loss stage;
 * lh_pipeline_read_metrics = synchronous read of counts u64[4] / sums f64[4] (see lh_pipeline_io) */
int lh_pipeline_allreduce_metrics(lh_pipeline* p, void* comm);
int lh_pipeline_read_metrics(lh_pipeline* p, unsigned long long* counts4, double* sums4);
int lh_pipeline_reset_metrics(lh_pipeline* p);
/* number of kernel launches one step issues (for accounting) */
int lh_pipeline_launches_per_step(const lh_pipeline* p);
/* ------------------------------------------------------------------------------------------
 * Stage ring: stream / event plumbing of the pipelined DEVICE-resident form (what HeatmapPipeline.data_stage / loss_stage
 * and PreparedStep.launch_data / launch_loss run on).  The ring owns a data stream and the events of two slots; the
 * caller owns the two network-input buffers.  K1 of batch i+1 (data stream) then overlaps whatever the caller enqueues
 * on its own stream for batch i (backbone, fused loss pass) -- the reference's loop shape (src/utils/method.py:160-183).
 *   begin_data(slot, ...)  data stream waits for the inputs -- wait_mode 0: nothing (resident inputs), 1: `wait_event`
 *                          (a cudaEvent_t), 2: everything enqueued so far on `after_stream` -- and for the slot's
 *                          previous consumer;   end_data(slot) records "input ready"
 *   begin_loss(slot, s)    stream s waits for "input ready";   end_loss(slot, s) records "slot consumed" on s
 *   lh_stage_ring_data / lh_stage_ring_loss = begin + lh_warp_normalize / lh_heatmap_fused + end in ONE call
 * ------------------------------------------------------------------------------------------ */
typedef struct lh_warp_args {     /* the arguments of lh_warp_normalize */
    const uint8_t* src; int64_t src_img_stride; int32_t src_h, src_w, src_pitch;
    const double* fwd_affine; const uint8_t* flip; const double* chan_gain;
    float mean_std[6]; int32_t has_mean_std;
    const float* lut; int32_t interp, chan_order, out_layout;
    float* dst; int32_t N, out_h, out_w;
} lh_warp_args;
typedef struct lh_stage_ring lh_stage_ring;
int lh_stage_ring_create(int device, lh_stage_ring** out);
void lh_stage_ring_destroy(lh_stage_ring* r);
void* lh_stage_ring_data_stream(lh_stage_ring* r);   /* cudaStream_t */
int lh_stage_ring_begin_data(lh_stage_ring* r, int slot, void* wait_event, void* after_stream, int wait_mode);
int lh_stage_ring_end_data(lh_stage_ring* r, int slot);
int lh_stage_ring_begin_loss(lh_stage_ring* r, int slot, lh_stream_t stream);
int lh_stage_ring_end_loss(lh_stage_ring* r, int slot, lh_stream_t stream);
int lh_stage_ring_data(lh_stage_ring* r, int slot, const lh_warp_args* a);
int lh_stage_ring_loss(lh_stage_ring* r, int slot, const lh_heatmap_args* a, lh_stream_t stream);

/* pinned host memory so the copies above run at full DMA rate */
void* lh_host_alloc(size_t bytes);
void lh_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTHAND_B200_H */
