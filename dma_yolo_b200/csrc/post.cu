// Rows 8f-3 / 8f-4 of the scope table: the pieces that follow the heads.
//
//   dfl_decode_kernel  TDetect's eval tail (models/detect_t.py:46-58, 81-102): softmax over the 16 distribution bins of each
//                      box side, expectation of the bin index (DFL's frozen 1x1 conv with weights 0..15), dist2bbox around
//                      the cell centre, * stride, and sigmoid of the class logits -> y[b, 4 + nc, A].  One pass over the
//                      fp32 head logits; replaces softmax + cuDNN conv + cat + sigmoid in eager torch.
//   val_match_kernel   val.py:62-83 `process_batch` for a whole batch on the device, with `scale_coords` / `clip_coords`
//                      (utils/general.py:605-630) applied to the detections first: correct[b, j, t] = detection j of image b
//                      is the matched detection of some label at IoU level t.  One CTA per image.
#include "common.cuh"

namespace dmay {

// ---- TDetect / DFL ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dfl_decode_kernel(const float* __restrict__ box, const float* __restrict__ cls,
                                                         float* __restrict__ y, int ny, int nx, int nc, int reg_max, int ld_box,
                                                         int ld_cls, int a0, long long A, float stride) {
  const int npix = ny * nx;
  const int tiles = (npix + 127) / 128;
  const int n = blockIdx.x / tiles, t = blockIdx.x - n * tiles;
  const int pix = t * 128 + threadIdx.x;
  if (pix >= npix) return;
  const int gy = pix / nx, gx = pix - gy * nx;
  const float* bp = box + ((long long)n * npix + pix) * ld_box;
  float d[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float* l = bp + s * reg_max;
    float m = -INFINITY;
    for (int j = 0; j < reg_max; ++j) m = fmaxf(m, l[j]);
    float den = 0.f;
    for (int j = 0; j < reg_max; ++j) den += expf(l[j] - m);
    // softmax(1) first, then the frozen conv: sum_j j * p_j in bin order (detect_t.py:100-102)
    float e = 0.f;
    for (int j = 0; j < reg_max; ++j) e = __fmaf_rn((float)j, __fdiv_rn(expf(l[j] - m), den), e);
    d[s] = e;
  }
  // dist2bbox(xywh=True) around the cell centre, then * stride (detect_t.py:56, 81-90)
  const float ax = (float)gx + 0.5f, ay = (float)gy + 0.5f;
  const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
  float* yp = y + (long long)n * (4 + nc) * A + a0 + pix;
  yp[0] = __fmul_rn(__fdiv_rn(x1 + x2, 2.f), stride);
  yp[A] = __fmul_rn(__fdiv_rn(y1 + y2, 2.f), stride);
  yp[2 * A] = __fmul_rn(x2 - x1, stride);
  yp[3 * A] = __fmul_rn(y2 - y1, stride);
  const float* cp = cls + ((long long)n * npix + pix) * ld_cls;
  for (int c = 0; c < nc; ++c) yp[(long long)(4 + c) * A] = sigmoid_acc(cp[c]);
}

// ---- val.py process_batch ----------------------------------------------------------------------------------------------
constexpr int kMatchThreads = 256;

__global__ void __launch_bounds__(kMatchThreads) val_match_kernel(const float* __restrict__ det, const int* __restrict__ counts,
                                                                 const float* __restrict__ labels, const int* __restrict__ lab_off,
                                                                 const float* __restrict__ geom, const float* __restrict__ iouv,
                                                                 unsigned char* __restrict__ correct, float* __restrict__ predn_out,
                                                                 int max_det, int niou, int single_cls) {
  extern __shared__ int sm_i[];
  int* best_lab = sm_i;                               // [max_det]
  float* best_iou = reinterpret_cast<float*>(sm_i + max_det);   // [max_det]
  int* win = sm_i + 2 * max_det;                      // [nl]
  const int b = blockIdx.x;
  const int n = min(counts[b], max_det);
  const int l0 = lab_off[b], nl = lab_off[b + 1] - l0;
  const float gain = geom[b * 5 + 0], padx = geom[b * 5 + 1], pady = geom[b * 5 + 2], h0 = geom[b * 5 + 3], w0 = geom[b * 5 + 4];
  for (int i = threadIdx.x; i < nl; i += kMatchThreads) win[i] = 0x7fffffff;
  __syncthreads();
  for (int j = threadIdx.x; j < max_det; j += kMatchThreads) {
    int bl = -1;
    float bi = 0.f;
    if (j < n) {
      const float* d = det + ((long long)b * max_det + j) * 6;
      // scale_coords: subtract the letterbox padding, divide by the gain, clip to the original image (general.py:605-630)
      float x1 = __fdiv_rn(__fsub_rn(d[0], padx), gain), y1 = __fdiv_rn(__fsub_rn(d[1], pady), gain);
      float x2 = __fdiv_rn(__fsub_rn(d[2], padx), gain), y2 = __fdiv_rn(__fsub_rn(d[3], pady), gain);
      x1 = fminf(fmaxf(x1, 0.f), w0); x2 = fminf(fmaxf(x2, 0.f), w0);
      y1 = fminf(fmaxf(y1, 0.f), h0); y2 = fminf(fmaxf(y2, 0.f), h0);
      if (predn_out != nullptr) {
        float* o = predn_out + ((long long)b * max_det + j) * 6;
        o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = d[4]; o[5] = single_cls ? 0.f : d[5];
      }
      const float dc = single_cls ? 0.f : d[5];
      const float area2 = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
      for (int i = 0; i < nl; ++i) {
        const float* L = labels + (long long)(l0 + i) * 5;      // cls, x1, y1, x2, y2 (already in original-image pixels)
        if (L[0] != dc) continue;
        // utils/metrics.py:267-276, box1 = labels, box2 = detections
        const float area1 = __fmul_rn(__fsub_rn(L[3], L[1]), __fsub_rn(L[4], L[2]));
        const float iw = fmaxf(__fsub_rn(fminf(L[3], x2), fmaxf(L[1], x1)), 0.f);
        const float ih = fmaxf(__fsub_rn(fminf(L[4], y2), fmaxf(L[2], y1)), 0.f);
        const float inter = __fmul_rn(iw, ih);
        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area1, area2), inter));
        // descending-IoU order decides the label of a detection; equal IoUs: the later label (what a stable ascending
        // argsort read backwards gives -- the reference's order of ties is implementation-defined)
        if (iou >= iouv[0] && (bl < 0 || iou >= bi)) {
          bl = i;
          bi = iou;
        }
      }
    }
    best_lab[j] = bl;
    best_iou[j] = bi;
    if (bl >= 0) atomicMin(&win[bl], j);    // per label: the first (most confident) detection that chose it
  }
  __syncthreads();
  for (int j = threadIdx.x; j < max_det; j += kMatchThreads) {
    const int bl = best_lab[j];
    const bool hit = bl >= 0 && win[bl] == j;
    const float bi = best_iou[j];
    unsigned char* c = correct + ((long long)b * max_det + j) * niou;
    for (int t = 0; t < niou; ++t) c[t] = (hit && bi >= iouv[t]) ? 1 : 0;
  }
}

}  // namespace dmay

using namespace dmay;

extern "C" int dmay_dfl_decode(const dmay_dfl_params* p, dmay_stream_t stream) {
  if (!p || !p->box || !p->cls || !p->y) return DMAY_EINVAL;
  if (p->N <= 0 || p->ny <= 0 || p->nx <= 0 || p->nc <= 0 || p->reg_max <= 0 || p->A <= 0 || p->a0 < 0) return DMAY_EINVAL;
  if (p->ld_box < 4 * p->reg_max || p->ld_cls < p->nc || p->reg_max > 64) return DMAY_EUNSUPPORTED;
  if ((long long)p->a0 + (long long)p->ny * p->nx > p->A) return DMAY_EINVAL;
  const long long tiles = ((long long)p->ny * p->nx + 127) / 128;
  if (tiles * p->N > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  dfl_decode_kernel<<<(int)(tiles * p->N), 128, 0, (cudaStream_t)stream>>>((const float*)p->box, (const float*)p->cls, (float*)p->y,
                                                                          p->ny, p->nx, p->nc, p->reg_max, p->ld_box, p->ld_cls,
                                                                          p->a0, p->A, p->stride);
  return finish_launch();
}

extern "C" int dmay_val_match(const dmay_match_params* p, dmay_stream_t stream) {
  if (!p || !p->det || !p->counts || !p->labels || !p->lab_off || !p->geom || !p->iouv || !p->correct) return DMAY_EINVAL;
  if (p->B <= 0 || p->max_det <= 0 || p->niou <= 0 || p->niou > 32 || p->max_labels < 0) return DMAY_EINVAL;
  const size_t smem = ((size_t)2 * p->max_det + (size_t)p->max_labels) * 4;
  if (smem > 200 * 1024) return DMAY_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(val_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  val_match_kernel<<<p->B, kMatchThreads, smem, (cudaStream_t)stream>>>((const float*)p->det, (const int*)p->counts,
                                                                        (const float*)p->labels, (const int*)p->lab_off,
                                                                        (const float*)p->geom, (const float*)p->iouv,
                                                                        (unsigned char*)p->correct, (float*)p->predn, p->max_det,
                                                                        p->niou, p->single_cls);
  return finish_launch();
}
