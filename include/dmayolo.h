/*
 * dmayolo.h — C-ABI of libdmayolo.so: hand-written sm_100a kernels for the DMA-YOLO
 * detection-forward hot path (custom blocks + Conv/BN/SiLU + Detect decode + NMS).
 *
 * The reference (Yaling-Li/DMA-YOLO) is pure PyTorch and has no FFI for this path
 * (SURVEY.md F1); every entry point below therefore cites the *Python* reference
 * symbol (file:line under the reference tree) whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers + sizes only; no torch types.  All pointers are DEVICE pointers
 *     unless the field name ends in `_host`.
 *   - activations: bf16, NHWC ("channels_last"): element (n,h,w,c) of a tensor that lives
 *     in a channel slice of a wider slab is at  base + ((n*H + h)*W + w)*ld + c,
 *     `ld` = pixel stride in ELEMENTS (= channels of the slab).  ld and channel offsets
 *     are multiples of 8 so every pixel row is 16-byte aligned.
 *   - ownership: the caller allocates every input / output / workspace buffer.  The
 *     library never allocates or frees device memory and keeps no mutable global state
 *     except the launch counter.
 *   - all work is enqueued on the passed stream; no implicit synchronisation.
 *   - return value: 0 = success, <0 = DMAY_E* (bad argument / unsupported shape),
 *     >0 = cudaError_t.  Nothing throws across the boundary.
 *   - structs contain only `const void*`, `void*`, `int`, `long long`, `float`, `double`
 *     fields (one per line) so the Python side can parse this header into ctypes.
 */
#ifndef DMAYOLO_H_
#define DMAYOLO_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DMAY_OK 0
#define DMAY_EINVAL -1      /* null pointer / non-positive size / misaligned pointer */
#define DMAY_EUNSUPPORTED -2 /* shape outside what the kernel supports */
#define DMAY_EDRIVER -3     /* cuTensorMapEncode* entry point missing or failed */
#define DMAY_ETOOBIG -4     /* workspace / capacity too small */

#define DMAY_ACT_NONE 0
#define DMAY_ACT_SILU 1
#define DMAY_ACT_HARDSWISH 2
#define DMAY_ACT_SIGMOID 3
#define DMAY_ACT_GELU 4      /* exact (erf) GELU, nn.GELU default */

#define DMAY_DT_BF16 0
#define DMAY_DT_F32 1
#define DMAY_DT_F16 2
#define DMAY_DT_U8 3

/* opaque CUDA stream handle (cudaStream_t) */
typedef void* dmay_stream_t;

/* ---- library introspection ------------------------------------------------------------ */
int dmay_version(void);
/* number of kernels this library has launched since load (bench.py `gpu_launches`) */
long long dmay_launch_count(void);
/* human readable text for a return code */
const char* dmay_strerror(int code);

/* ---- a1/a2: Conv + folded BN + activation (+ residual) as tcgen05 implicit GEMM --------
 * replaces models/common.py:50-77 Conv.forward / forward_fuse (dup models/cspcm.py:11-23),
 * BN fold utils/torch_utils.py:198-218, Bottleneck residual models/common.py:136-137,
 * and (via ldy / channel-slice outputs) the torch.cat in C3.forward models/common.py:181-182.
 *   y[m, co] = act( scale[co] * sum_{r,s,ci} x[n, p*stride-pad+r, q*stride-pad+s, ci] * w[co,r,s,ci] + bias[co] ) (+ residual[m,co])
 *   m = (n*Ho + p)*Wo + q.
 * w is packed [Cout_pad][kh][kw][Cin] bf16 (Cin multiple of 16, Cout_pad multiple of 16).
 * mode gate (gate_x != NULL): y = (scale*acc+bias) * sigmoid(gate_x[m,co] + gate_k[n, hs, ws, co])
 *   with nearest index hs=min(floor(p*gate_sh),gHk-1) — SCConv.forward models/common.py:1308-1316.
 * res_op: 0 = the residual operand is added after the activation (Bottleneck); 1 = act must be NONE and the result is
 *   MULTIPLIED by the residual operand: y = (scale*acc + bias) * residual  (GnConv recursive gating, common.py:1344).
 * x1 / x2 (Cin1 / Cin2 > 0; 1x1, stride 1, pad 0 only): VIRTUAL channel concat -- the input is [x | x1 | x2] over the same
 *   pixels, each part with its own base pointer and pixel pitch (ldx / ldx1 / ldx2), channel counts in multiples of 64;
 *   Cin stays the TOTAL and w the weights of the concatenated input.  Replaces the torch.cat of Concat.forward
 *   (models/common.py:656-664) and of AdConcat2/3.forward (models/common.py:1003-1008, 1021-1026; the BiFPN weights are
 *   folded into the columns of w by the caller) in front of a 1x1 Conv / C3.cv1 | cv2.
 * pre (fp32 [N, preH, preW, ldpre], act SiLU, bf16 output, no residual / gate): y = silu(scale * (acc + pre[n, hs, ws, co])
 *   + bias) with the nearest source pixel of (p, q) -- a 1x1 layer commutes with nn.Upsample(nearest), so the columns
 *   of an up-sampled concat part can be summed at ITS resolution and enter here.
 * pool4_out (bf16 [N, Ho/4, Wo/4, ldpool4]; plain SiLU 3x3 stride-1 layers with resident weights only, else DMAY_EUNSUPPORTED):
 *   additionally writes AvgPool2d(4, 4) of y -- the k2 branch input of a following SCConv (models/common.py:1281-1287), whose
 *   pooling kernel would re-read the whole map.  Bit-identical to dmay_avgpool(y, 4).
 * block_n: 0 = auto, >0 = force the N tile, -2 = 2-CTA cluster multicast of the weight tile (experiment).
 * flags (tuning / A-B switches, 0 = auto): bit0 = never use the halo path (3x3 s1 p1 input patch loaded once
 *   per channel chunk, taps read shifted windows), bit1 = force it where legal, bit2 = never keep the weight
 *   set resident in shared memory in halo mode, bit3 / bit4 = force 8 / 16 epilogue warps,
 *   bit5 = never split the epilogue warps into two alternate-tile groups (narrow tiles),
 *   bit6 = two TMEM accumulator buffers instead of 512 / block_n,
 *   bit7 = launch without programmatic dependent launch (the kernel's preamble then waits for the previous kernel),
 *   bit8 / bit9 = force / forbid the CTA-pair mode (cta_group::2: a 2-CTA cluster computes one 256 x block_n tile,
 *   each CTA loading its own 128 rows of A and half of the weight tile), bit10 / bit11 = never / always (where legal) keep
 *   the weights resident in shared memory on the plain (non-halo) path, bit12 = never keep weight halves resident in the
 *   CTA-pair halo mode, bit13 = 16 epilogue warps with the residual / gate operand tile sharing the output staging tile
 *   for every staged operand (automatic only for CTA-pair layers with resident weights), bit14 = fp32 outputs of 1x1 convs
 *   (Detect heads) with direct per-lane stores instead of the staged tile + TMA store, bit15 = streamed halo mode with up
 *   to four input patches in flight (default: two, the rest of shared memory holds weight stages), bit16 = issue the nine
 *   taps of a resident-weight 3x3 layer with the per-tap loop instead of the one-block form. */
typedef struct dmay_conv_params {
  const void* x;
  const void* w;
  const void* scale;
  const void* bias;
  const void* residual;
  const void* gate_x;
  const void* gate_k;
  void* y;
  int N;
  int H;
  int W;
  int Cin;
  int ldx;
  int Cout;
  int Cout_pad;
  int kh;
  int kw;
  int stride;
  int pad;
  int Ho;
  int Wo;
  int ldy;
  int ldr;
  int ldgx;
  int gHk;
  int gWk;
  int act;
  int out_dtype;
  int block_n;
  int num_sms;
  int flags;
  int res_op;
  const void* x1;
  const void* x2;
  int Cin1;
  int Cin2;
  int ldx1;
  int ldx2;
  const void* pre;
  int ldpre;
  int preH;
  int preW;
  void* pool4_out;
  int ldpool4;
} dmay_conv_params;
int dmay_conv_bn_act(const dmay_conv_params* p, dmay_stream_t stream);
/* launch-plan cache of dmay_conv_bn_act (mode / tile decisions + encoded CUtensorMaps, keyed by the parameter struct and the
 * device): what = 0 -> hits, 1 -> misses, 2 -> entries.  Immutable entries; no device memory is owned. */
long long dmay_conv_plan_stats(int what);

/* ---- input prep: NCHW {f32,f16,bf16,u8} image -> NHWC bf16 ------------------------------
 * replaces val.py:199-202 (`img.float()/255`) + the layout change the first conv needs.
 * spd=1 writes the 2x2 pixel-unshuffled form [N,H/2,W/2,Cpad] with channel (dy*2+dx)*C+c
 * (stem 6x6 s2 p2 conv == 3x3 s1 p1 conv on it); channels >= real are zero.  mul scales
 * (1/255 for u8 callers that skip the division). */
typedef struct dmay_prep_params {
  const void* x;
  void* y;
  int N;
  int C;
  int H;
  int W;
  int Cpad;
  int spd;
  int in_dtype;
  float mul;
} dmay_prep_params;
int dmay_input_prep(const dmay_prep_params* p, dmay_stream_t stream);

/* NHWC bf16 (slice, ld) <-> NCHW contiguous {f32,bf16,f16}: glue for modules that stay in
 * torch ops.  dir=0: nhwc->nchw, dir=1: nchw->nhwc. */
typedef struct dmay_layout_params {
  const void* x;
  void* y;
  int N;
  int C;
  int H;
  int W;
  int ld;
  int dtype;
  int dir;
} dmay_layout_params;
int dmay_layout_convert(const dmay_layout_params* p, dmay_stream_t stream);

/* ---- a4: space_to_depth (SPD-Conv), models/common.py:1457-1458 (same: SM 1466, Focus 94)
 * y[n,h,w,(dy+2*dx)*C + c] = x[n,2h+dy,2w+dx,c]   (bit-exact copy) */
typedef struct dmay_spd_params {
  const void* x;
  void* y;
  int N;
  int H;
  int W;
  int C;
  int ldx;
  int ldy;
} dmay_spd_params;
int dmay_spd(const dmay_spd_params* p, dmay_stream_t stream);

/* ---- a6: AdConcat2/3 (+ fused preceding nn.Upsample nearest), Concat -------------------
 * models/common.py:1003-1008,1021-1026 (weight = w/(sum w + 1e-4) is computed by the
 * caller in fp32 and passed as w0..w2), models/common.py:656-664 (Concat: weights 1).
 * y[n,h,w, off_i + c] = bf16( wi * x_i[n, h>>up_i, w>>up_i, c] ),  i < n_in.
 * H,W are OUTPUT dims; input i has dims (H>>up_i, W>>up_i). */
typedef struct dmay_adconcat_params {
  const void* x0;
  const void* x1;
  const void* x2;
  void* y;
  int n_in;
  int N;
  int H;
  int W;
  int C0;
  int C1;
  int C2;
  int ld0;
  int ld1;
  int ld2;
  int up0;
  int up1;
  int up2;
  int ldy;
  float w0;
  float w1;
  float w2;
} dmay_adconcat_params;
int dmay_adconcat(const dmay_adconcat_params* p, dmay_stream_t stream);

/* Adapt_Add2 models/common.py:1040-1045: y = silu(w0*x0 + w1*x1) (+ w2*x2 for Adapt_Add3
 * after its shared 1x1 conv, models/common.py:1047-1061).  Same layout rules. */
typedef struct dmay_adaptadd_params {
  const void* x0;
  const void* x1;
  const void* x2;
  void* y;
  int n_in;
  long long npix;
  int C;
  int ld0;
  int ld1;
  int ld2;
  int ldy;
  float w0;
  float w1;
  float w2;
} dmay_adaptadd_params;
int dmay_adaptadd(const dmay_adaptadd_params* p, dmay_stream_t stream);

/* nn.Upsample(None, 2, 'nearest') stand-alone (yaml glue), bit-exact copy;
 * residual-free elementwise add  y = a + b  (Bottleneck fallback when not fused). */
typedef struct dmay_upsample_params {
  const void* x;
  void* y;
  int N;
  int H;
  int W;
  int C;
  int ldx;
  int ldy;
  int factor;
} dmay_upsample_params;
int dmay_upsample_nearest(const dmay_upsample_params* p, dmay_stream_t stream);

/* ---- a7: cascaded 5x5 s1 p2 max-pools of SPPF / SPPFCSPC -------------------------------
 * models/common.py:1272-1274 and 252-258:  y1=mp_k(x), y2=mp_k(y1), y3=mp_k(y2)
 * (== windows k, 2k-1, 3k-2 with -inf padding).  x and y1..y3 may be channel slices of one
 * slab (the 4C concat input of cv5/cv2): separate base pointers, common ld.  Bit-exact. */
typedef struct dmay_sppf_params {
  const void* x;
  void* y1;
  void* y2;
  void* y3;
  int N;
  int H;
  int W;
  int C;
  int ldx;
  int ldy;
  int k;
} dmay_sppf_params;
int dmay_sppf_pool3(const dmay_sppf_params* p, dmay_stream_t stream);

/* generic max-pool kxk stride 1 pad k/2 (SPP / SPPCSPC branches, models/common.py:212-227,1237-1255) */
typedef struct dmay_maxpool_params {
  const void* x;
  void* y;
  int N;
  int H;
  int W;
  int C;
  int ldx;
  int ldy;
  int k;
} dmay_maxpool_params;
int dmay_maxpool_s1(const dmay_maxpool_params* p, dmay_stream_t stream);

/* ---- a5: SCConv pieces, models/common.py:1279-1316 --------------------------------------
 * avgpool: nn.AvgPool2d(r, r) (floor), fp32 accumulate -> bf16.
 * gate:    y = k3 * sigmoid(x + k2[n, min(floor(h*Hk/H),Hk-1), min(floor(w*Wk/W),Wk-1), c])
 *          (F.interpolate default 'nearest', index math in fp32 like ATen). */
typedef struct dmay_avgpool_params {
  const void* x;
  void* y;
  int N;
  int H;
  int W;
  int C;
  int ldx;
  int ldy;
  int r;
} dmay_avgpool_params;
int dmay_avgpool(const dmay_avgpool_params* p, dmay_stream_t stream);

typedef struct dmay_scgate_params {
  const void* x;
  const void* k3;
  const void* k2;
  void* y;
  int N;
  int H;
  int W;
  int C;
  int Hk;
  int Wk;
  int ldx;
  int ld3;
  int ld2;
  int ldy;
} dmay_scgate_params;
int dmay_scconv_gate(const dmay_scgate_params* p, dmay_stream_t stream);

/* ---- a3: CoordAtt, models/common.py:1183-1207 -------------------------------------------
 * step 1 (pool):  pooled[n, h, c]   = mean_w x[n,h,w,c]        (h <  H)
 *                 pooled[n, H+w, c] = mean_h x[n,h,w,c]        (fp32, [N, H+W, C])
 * step 2 (mlp):   y = hardswish(s1 * (W1 . pooled + b1) + t1)  (conv1 bias b1, BN folded into s1/t1);
 *                 gates[n,p,:] = sigmoid(Wh . y + bh) for p < H, sigmoid(Ww . y + bw) for p >= H,
 *                 fp32 [N,H+W,Cout].  Weight operands are passed TRANSPOSED, fp32:
 *                 w1 = W1^T [C][Cm],  wh = Wh^T [Cm][Cout],  ww = Ww^T [Cm][Cout]  (Cout == C).
 * step 3 (apply): out = (x * a_w[n,w,c]) * a_h[n,h,c]  -> bf16
 * dmay_coordatt runs the three steps on the stream.
 * Fast path (two launches; taken when `ws` is given and a 64-channel plane fits in shared memory): pool + partial
 * hidden layer per (image, 64-channel group) with the image's last CTA finishing y, then gates + apply per group.
 * For Cm <= 32 (every reference model) the two small dense layers run on mma.sync with fp32 = bf16 hi + lo operand
 * splitting (three products, fp32 accumulate: ~2^-16 relative), the second launch is a programmatic dependent launch.
 * ws: workspace of dmay_coordatt_ws() bytes, ZERO-initialised before its first use (it holds per-image arrival
 * tickets that the kernel resets itself; one workspace per stream).  pooled / gates may then be NULL (they are
 * only written when given). */
typedef struct dmay_coordatt_params {
  const void* x;
  void* y;
  void* pooled;
  void* gates;
  const void* w1;
  const void* b1;
  const void* s1;
  const void* t1;
  const void* wh;
  const void* bh;
  const void* ww;
  const void* bw;
  int N;
  int H;
  int W;
  int C;
  int Cm;
  int ldx;
  int ldy;
  int num_sms;
  void* ws;
  long long ws_bytes;
} dmay_coordatt_params;
long long dmay_coordatt_ws(int N, int H, int W, int C, int Cm);
int dmay_coordatt(const dmay_coordatt_params* p, dmay_stream_t stream);

/* ---- a8: Detect grid/anchor decode, models/yolo.py:81-101 -------------------------------
 * logits: fp32 NHWC [N, ny, nx, ld] per level, channel = a*row_pitch + o  (1x1 conv + bias output; row_pitch = 0 means no,
 *         a head GEMM may pad each anchor's no outputs to a multiple of 4 floats, see dmay_nms_filter_fused).
 * dense:  pred[n, row0 + (a*ny + y)*nx + x, :] = decode(sigmoid(logit)) — rows ordered
 *         (level, anchor, y, x), out [N, rows_total, no] fp32. anchors_px = anchors*stride. */
typedef struct dmay_decode_params {
  const void* logits;
  void* pred;
  int N;
  int ny;
  int nx;
  int na;
  int no;
  int ld;
  int row0;
  int rows_total;
  float stride;
  float aw0;
  float ah0;
  float aw1;
  float ah1;
  float aw2;
  float ah2;
  float aw3;
  float ah3;
  float aw4;
  float ah4;
  int row_pitch;
} dmay_decode_params;
int dmay_detect_decode(const dmay_decode_params* p, dmay_stream_t stream);

/* ---- a9: batched class-aware NMS, utils/general.py:633-725 (+ torchvision.ops.nms) ------
 * Candidate generation ("filter"), three launches, order-preserving:
 *   count : per (image, row-block) number of candidates -> blk_counts[N*nblk]
 *   scan  : per image exclusive scan -> blk_offsets, img_counts[N], img_offsets[N+1]
 *   write : candidates in reference row-major order:
 *           keys[g]   = (img << 32) | ~bits(conf)      (ascending key == descending score)
 *           cand[g,:] = x1,y1,x2,y2,conf,cls  (fp32; xyxy from xywh exactly as xywh2xyxy
 *                       utils/general.py:539-546), g = img_offsets[img] + local index
 * source: either dense `pred` [N, R, 5+nc] fp32 (levels = 0) or the Detect logits of up to
 * 5 levels (fused decode + confidence filter; nothing dense is materialised).
 * multi_label=0: best class only (first max), keep conf > thr;  =1: every (row,class) with
 * obj*cls > thr, row-major.  class_mask: optional [nc] u8 keep-list (`classes=` argument). */
typedef struct dmay_filter_params {
  const void* pred;
  const void* lv_logits0;
  const void* lv_logits1;
  const void* lv_logits2;
  const void* lv_logits3;
  const void* lv_logits4;
  const void* lv_meta;
  const void* class_mask;
  void* blk_counts;
  void* blk_offsets;
  void* img_counts;
  void* img_offsets;
  void* keys;
  void* cand;
  int N;
  int R;
  int nc;
  int levels;
  int multi_label;
  int rows_per_block;
  int phase;
  long long capacity;
  float conf_thres;
  int row_pitch;
} dmay_filter_params;
int dmay_nms_filter(const dmay_filter_params* p, dmay_stream_t stream);

/* Single-pass fused variant (Detect logits source only): decode + confidence filter + ORDER-PRESERVING
 * compaction of the whole batch with ONE read of the logits.  A CTA owns 64 consecutive pixels of one
 * anchor plane, in the reference row order (image, level, anchor, y, x); its candidates are placed by a
 * decoupled look-back scan over the tiles.  Outputs as dmay_nms_filter (keys, cand in reference order),
 * plus img_offsets[N+1] (i64; [N] = total candidates of the batch) and img_counts[N] (i32).
 * ws: caller-ZEROED workspace of dmay_nms_filter_fused_ws() bytes (ticket + one status word per tile).
 * Candidates beyond `capacity` are counted but not written: the caller compares img_offsets[N] with
 * capacity and repeats the call with larger buffers.  lv_meta_host: HOST array of `levels` x 16 words
 * {row0, ny, nx, ld, na, stride(f32), anchor_px[10](f32)}.
 * keys_tmp / cand_tmp (optional, `capacity` entries each like keys / cand): when given (and nc <= 96) the tiles do not
 * order themselves with the look-back; each reserves its run in the temporary buffers with one atomic, and a scan over
 * the per-tile counts plus a gather put the runs in reference order (three launches, same outputs).
 * row_pitch (0 = 5 + nc): floats between the rows of two anchors inside a pixel's channel vector.  A head GEMM that pads
 *   every anchor's 5 + nc outputs to a multiple of 4 floats (zero weight rows) makes the rows 16-byte aligned: they are
 *   then staged with 16-byte copies and scanned four logits per shared-memory load (ld and the base pointers must be
 *   16-byte multiples).
 * dense = 1: the source is a DENSE prediction [N, R, 5 + nc] (utils/general.py:633 input, already decoded) passed as
 *   lv_logits0 with one level {row0 0, ny 1, nx R, ld 5 + nc, na 1}: same kernels, values used as they are.
 * bin_thr (dense multi-label only, optional): per-image key-bin threshold of dmay_nms_dense_prethreshold.
 * per_image_regions = 1 (reserve mode): the temporary buffers are split into N regions of capacity / N candidates and every
 *   image reserves in its own region with its own counter (one contended atomic address per image instead of one per batch).
 *   An image then overflows when ITS count exceeds capacity / N: the caller checks max_i (img_offsets[i+1] - img_offsets[i])
 *   against capacity / N as well as img_offsets[N] against capacity. */
typedef struct dmay_filter_fused_params {
  const void* lv_logits0;
  const void* lv_logits1;
  const void* lv_logits2;
  const void* lv_logits3;
  const void* lv_logits4;
  const void* lv_meta_host;
  const void* class_mask;
  void* ws;
  void* img_counts;
  void* img_offsets;
  void* keys;
  void* cand;
  long long ws_bytes;
  int N;
  int nc;
  int levels;
  int multi_label;
  long long capacity;
  float conf_thres;
  void* keys_tmp;
  void* cand_tmp;
  int row_pitch;
  int dense;
  const void* bin_thr;
  int per_image_regions;
  void* hist;
  int prethr_k;
} dmay_filter_fused_params;
long long dmay_nms_filter_fused_ws(const void* lv_meta_host, int levels, int N);
/* Pre-selection for candidate-rich Detect logits (padded anchor rows, multi_label): utils/general.py:702-703 keeps the max_nms
 * best candidates of an image, a saturated head yields many times that (cfg-4a / cfg-4b at 0.001: 1.7 M per image).  One more pass
 * over the logits bins every candidate's exact confidence by a monotone 2048-bin function of its score key (64 bins per binade below 1.0; hist: caller-ZEROED i32 [N, 2048])
 * and writes bin_thr[img] = the bin holding the prethr_k-th best key; dmay_nms_filter_fused with the same parameters and that
 * bin_thr then writes only candidates of bins <= bin_thr: a superset of the top prethr_k, ties included, in candidate order --
 * the detections are unchanged. */
int dmay_nms_fused_prethreshold(const dmay_filter_fused_params* p, dmay_stream_t stream);

/* Pre-selection for candidate-dense MULTI-LABEL dense predictions (utils/general.py:702-703 keeps the max_nms best candidates
 * of an image; a dense prediction can expand to many times that).  Per image: a 2048-bin histogram (key_bin: 64 bins per binade below 1.0) of every
 * candidate's score key (~bits(conf)), then bin_thr[img] = the bin that holds the K-th best key (2047 when the image has at most
 * K candidates).  dmay_nms_filter_fused (dense = 1, bin_thr given) then writes only candidates whose key bin is <= bin_thr:
 * a superset of the top K, ties included, in candidate order -- the detections are unchanged.
 * pred fp32 [N, R, 5 + nc]; hist: caller-ZEROED i32 [N, 2048]; bin_thr i32 [N]. */
typedef struct dmay_prethr_params {
  const void* pred;
  const void* class_mask;
  void* hist;
  void* bin_thr;
  int N;
  int R;
  int nc;
  int K;
  float conf_thres;
} dmay_prethr_params;
int dmay_nms_dense_prethreshold(const dmay_prethr_params* p, dmay_stream_t stream);
int dmay_nms_filter_fused(const dmay_filter_fused_params* p, dmay_stream_t stream);

/* stable sort of candidate keys (payload = candidate index).  CUB radix sort over the
 * (img, ~score) composite key: equal scores keep ascending candidate index like
 * torchvision's stable descending sort.  ws_bytes query: call with ws == NULL. */
typedef struct dmay_sort_params {
  const void* keys_in;
  void* keys_out;
  void* idx_out;
  void* ws;
  long long ws_bytes;
  long long n;
  int img_bits;
  const void* vals_in;
} dmay_sort_params;
long long dmay_nms_sort_ws(long long n, int img_bits);
int dmay_nms_sort(const dmay_sort_params* p, dmay_stream_t stream);

/* exact pre-selection of the K = max_nms best candidates per image (utils/general.py:702-703) BEFORE the sort:
 * per image a 3-level radix select finds the K-th smallest score key, then every key below it and the first
 * (K - #below) keys equal to it are compacted in candidate order (so the stable sort that follows breaks ties exactly
 * as a full sort would).  keys: u64 [total] as written by the filter, images contiguous; outputs: keys_out u64 and
 * idx_out u32 (candidate indices) [sum_i min(count_i, K)], counts_out i32 [N], offsets_out i64 [N+1].
 * dmay_nms_sort then takes idx_out as `vals_in` (payload) instead of generating positions. */
typedef struct dmay_topk_params {
  const void* keys;
  const void* img_counts;
  const void* img_offsets;
  void* keys_out;
  void* idx_out;
  void* counts_out;
  void* offsets_out;
  int N;
  int K;
} dmay_topk_params;
int dmay_nms_topk_select(const dmay_topk_params* p, dmay_stream_t stream);

/* Sync-free guard for capacity-bounded candidate buffers (CUDA-graph replay of the NMS chain, where the host cannot
 * re-size the buffers after the filter): offsets_out[i] = min(img_offsets[i], capacity), counts_out[i] =
 * offsets_out[i+1] - offsets_out[i].  The filter counts every candidate but writes only those below `capacity`; with
 * the clamped counts no later kernel reads past it.  img_offsets[N] (the true total) is left untouched: the caller
 * compares it with `capacity` after the replay and repeats the step with larger buffers when it overflowed. */
typedef struct dmay_clamp_params {
  const void* img_offsets;
  void* offsets_out;
  void* counts_out;
  int N;
  long long capacity;
} dmay_clamp_params;
int dmay_nms_clamp_offsets(const dmay_clamp_params* p, dmay_stream_t stream);

/* greedy NMS over each image's sorted candidates, stopping after max_det keeps.
 * IoU exactly as torchvision CPU nms_kernel: fp32 inter/(areaA+areaB-inter), compared in
 * double against iou_thres; boxes offset by cls*max_wh in fp32 first (agnostic -> 0).
 * Only the first min(count, max_nms) sorted candidates of an image take part
 * (utils/general.py:702-703).  out [N, max_det, 6] fp32, out_counts [N] int32. */
typedef struct dmay_nms_params {
  const void* cand;
  const void* sorted_idx;
  const void* img_counts;
  const void* img_offsets;
  void* out;
  void* out_counts;
  int N;
  int max_det;
  int max_nms;
  int agnostic;
  float max_wh;
  double iou_thres;
} dmay_nms_params;
int dmay_nms_greedy(const dmay_nms_params* p, dmay_stream_t stream);

/* ---- 8f-1: SwinTransformerLayer pieces (models/common.py:452-634; C3STR in cfg-3) ----------------------------
 * The four Linear layers run as 1x1 convolutions on dmay_conv_bn_act (act = DMAY_ACT_GELU for mlp.fc1).
 * dmay_layernorm: nn.LayerNorm over the C channels of every pixel (biased variance), bf16 in / out, fp32 affine. */
typedef struct dmay_layernorm_params {
  const void* x;
  void* y;
  const void* gamma;
  const void* beta;
  long long npix;
  int C;
  int ldx;
  int ldy;
  float eps;
} dmay_layernorm_params;
int dmay_layernorm(const dmay_layernorm_params* p, dmay_stream_t stream);

/* dmay_window_attention: WindowAttention.forward (models/common.py:483-515) together with the layer's padding, cyclic
 * shift, window_partition / window_reverse and crop (models/common.py:603-627) as index arithmetic.
 *   qkv  [N, H, W, ldq] bf16, channel = which*C + head*32 + d (output of the qkv Linear, no bias)
 *   out  [N, H, W, ldo] bf16, channel = head*32 + d           (input of the proj Linear)
 *   rel_bias [heads][64][64] fp32: relative_position_bias_table gathered through relative_position_index
 *   mask [nW][64][64] fp32 (0 / -100) for shifted layers, NULL otherwise; nW ordered as the reference's
 *        window_partition of its (h, w) = (our W, our H) frame
 * window must be 8 and C == heads * 32 (what C3STR builds: SwinTransformerBlock(c_, c_, c_//32, n)).
 * variant 0: one warp per (image, window, head) on mma.sync bf16 tensor instructions (default);
 * variant 1: scalar fp32 kernel, one thread per query token (kept as the in-library cross-check). */
typedef struct dmay_winattn_params {
  const void* qkv;
  void* out;
  const void* rel_bias;
  const void* mask;
  int N;
  int H;
  int W;
  int C;
  int heads;
  int window;
  int shift;
  int ldq;
  int ldo;
  float scale;
  int variant;
} dmay_winattn_params;
int dmay_window_attention(const dmay_winattn_params* p, dmay_stream_t stream);

/* ---- 8f-2: HorBlock / GnConv pieces (models/common.py:1318-1426; C3HB in spdconv.yaml) ----------------------------
 * GEMM parts (proj_in, pws chain with the gating product as epilogue — dmay_conv_params.res_op = 1 —, proj_out, MLP) run
 * on dmay_conv_bn_act.  dmay_dwconv7: 7x7 depth-wise conv (pad 3) + bias, times `scale`, over Cd channels starting at
 * `x` (the caller passes the base already offset to the first `abc` channel); w is tap-major [49][Cd] fp32.  Output
 * segment i (input channels [seg_start_i, seg_start_{i+1})) is written at channel offset seg_out_i of y. */
typedef struct dmay_dwconv7_params {
  const void* x;
  const void* w;
  const void* bias;
  void* y;
  int N;
  int H;
  int W;
  int Cd;
  int ldx;
  int ldy;
  float scale;
  int n_seg;
  int seg_start0;
  int seg_start1;
  int seg_start2;
  int seg_start3;
  int seg_start4;
  int seg_out0;
  int seg_out1;
  int seg_out2;
  int seg_out3;
  int seg_out4;
} dmay_dwconv7_params;
int dmay_dwconv7(const dmay_dwconv7_params* p, dmay_stream_t stream);

/* y[pix, 0:d] = a[pix, 0:d] * b[pix, 0:d]  (bf16; d even) — GnConv's first gating step pwa * dw_list[0]. */
typedef struct dmay_mulch_params {
  const void* a;
  const void* b;
  void* y;
  long long npix;
  int d;
  int lda;
  int ldb;
  int ldy;
} dmay_mulch_params;
int dmay_mul_channels(const dmay_mulch_params* p, dmay_stream_t stream);

/* y = x + gamma[c] * g  (bf16 activations, fp32 per-channel gamma) — HorBlock layer scale + residual. */
typedef struct dmay_axpych_params {
  const void* x;
  const void* g;
  const void* gamma;
  void* y;
  long long npix;
  int C;
  int ldx;
  int ldg;
  int ldy;
} dmay_axpych_params;
int dmay_axpy_channels(const dmay_axpych_params* p, dmay_stream_t stream);

/* ---- measurement helpers ---------------------------------------------------------------
 * plain vectorised copy (roofline calibration inside bench.py) and L2 flush (memset-like
 * write of a scratch buffer larger than L2). */
typedef struct dmay_copy_params {
  const void* src;
  void* dst;
  long long bytes;
} dmay_copy_params;
int dmay_copy(const dmay_copy_params* p, dmay_stream_t stream);

/* ---- 8f-4: TDetect eval tail, models/detect_t.py:46-58, 81-102 -------------------------------------------------
 * One level per call.  box: fp32 NHWC [N, ny, nx, ld_box], channel = side * reg_max + bin (the 4 * reg_max outputs of the
 * cv2 head); cls: fp32 NHWC [N, ny, nx, ld_cls] (nc class logits).  y: fp32 [N, 4 + nc, A] (A = anchor points of all
 * levels, this level's start at a0, row-major (y, x)):
 *   d_s = sum_j j * softmax_j(box[s * reg_max + j])   (DFL.forward: softmax over the bins, frozen conv with weights 0..15)
 *   (x1, y1) = (gx + .5, gy + .5) - (d_0, d_1), (x2, y2) = (gx + .5, gy + .5) + (d_2, d_3)      (dist2bbox, xywh=True)
 *   y[0:4] = ((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1) * stride,   y[4 + c] = sigmoid(cls[c]). */
typedef struct dmay_dfl_params {
  const void* box;
  const void* cls;
  void* y;
  int N;
  int ny;
  int nx;
  int nc;
  int reg_max;
  int ld_box;
  int ld_cls;
  int a0;
  long long A;
  float stride;
} dmay_dfl_params;
int dmay_dfl_decode(const dmay_dfl_params* p, dmay_stream_t stream);

/* ---- 8f-3: val.py:62-83 process_batch for a whole batch, after scale_coords / clip_coords (utils/general.py:605-630) ----
 * det: fp32 [B, max_det, 6] (x1, y1, x2, y2, conf, cls) as dmay_nms_greedy writes it, counts i32 [B].
 * labels: fp32 [L, 5] (cls, x1, y1, x2, y2) in ORIGINAL-image pixels, images contiguous; lab_off i32 [B + 1].
 * geom: fp32 [B, 5] = (gain, pad_x, pad_y, h0, w0) of scale_coords (img1 -> img0).  iouv: fp32 [niou] IoU levels.
 * correct: u8 [B, max_det, niou]: detection j is the matched detection of a label of its class at level t:
 *   every detection takes the label of its class with the highest IoU >= iouv[0]; every label keeps the first (most
 *   confident) detection that took it; correct = (that IoU >= iouv[t]).  IoU in fp32 exactly as utils/metrics.py:254-276.
 * predn (optional): fp32 [B, max_det, 6], the rescaled detections (what val.py stores as `predn`).
 * max_labels: an upper bound of the labels of one image (shared-memory sizing). */
typedef struct dmay_match_params {
  const void* det;
  const void* counts;
  const void* labels;
  const void* lab_off;
  const void* geom;
  const void* iouv;
  void* correct;
  void* predn;
  int B;
  int max_det;
  int niou;
  int max_labels;
  int single_cls;
} dmay_match_params;
int dmay_val_match(const dmay_match_params* p, dmay_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DMAYOLO_H_ */
