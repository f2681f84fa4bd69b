// a8 + a9: Detect decode (models/yolo.py:81-101) and batched class-aware NMS
// (utils/general.py:633-725 + torchvision.ops.nms CPU semantics), all on device, whole batch per launch.
//
//   filter  : order-preserving candidate generation.  Dense [N,R,5+nc] fp32 prediction: count -> scan -> write.
//             Raw Detect logits (decode fused with the confidence filter, the dense [N,R,no] tensor is never
//             materialised, ONE read of the logits): thread-per-row tiles that reserve their runs + a scan over the
//             tile counts + a gather (nc <= 96), or the single-launch tile look-back forms.
//   select  : exact top-max_nms per image (3-level radix select + ordered compaction) when an image exceeds max_nms.
//   sort    : one stable radix sort of (image, ~score) keys over the whole batch (CUB).
//   greedy  : one CTA per image walks its sorted candidates in chunks of 128, testing each chunk
//             against the kept list (<= max_det boxes in shared memory), resolving the chunk's internal
//             order with a 128x128 bitmask and one warp's serial scan, and stops at max_det keeps.
//
// Arithmetic contract (bit-exact keep indices): every fp32 operation that the reference performs on
// boxes / scores is issued with explicit round-to-nearest intrinsics so nvcc cannot contract a
// multiply into a following add (no FMA), IoU uses IEEE division and is compared in double.
#include <stdlib.h>
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <cstring>

namespace dmay {

// Decode sigmoid: __expf + fast reciprocal (rel. error ~1e-6, far inside the 1e-4 decode tolerance; the
// accurate expf + IEEE division form cost 4x more ALU on 137 M logits per batch).  The fused filter and the
// dense decode kernel use THIS function, so both paths produce bit-identical candidates.
__device__ __forceinline__ float sigmoid_dec(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct LevelMeta {  // 16 x 4 bytes per level, see ops.py `_level_meta`
  int row0, ny, nx, ld, na;
  float stride;
  float anchor[10];  // (w,h) px for up to 5 anchors
};

struct RowSrc {
  const float* pred;         // dense source or nullptr
  const float* logits[5];    // fused source
  const LevelMeta* meta;
  int levels;
  int R, nc;
  int pitch;                 // floats per anchor row inside a pixel's channel vector (0 = 5 + nc)
};

// Loads one prediction row for the whole warp: returns obj and xywh (valid in all lanes); class
// confidences are fetched by `cls_at`.  Dense rows are read as-is; fused rows are decoded from logits
// exactly in the reference's operation order: xy = (s*2 - 0.5 + g) * stride ; wh = (s*2)^2 * anchor_px.
struct RowView {
  const float* base;  // points at element 0 of the row (dense) or logit 0 (fused)
  bool fused;
  float x, y, w, h, obj;
  __device__ __forceinline__ float cls_at(int c) const {
    float v = base[5 + c];
    return fused ? sigmoid_dec(v) : v;
  }
};

__device__ __forceinline__ RowView load_row(const RowSrc& s, int img, int r) {
  RowView v;
  if (s.levels == 0) {
    v.fused = false;
    v.base = s.pred + ((long long)img * s.R + r) * (5 + s.nc);
    v.x = v.base[0]; v.y = v.base[1]; v.w = v.base[2]; v.h = v.base[3]; v.obj = v.base[4];
    return v;
  }
  v.fused = true;
  int l = 0;
  while (l + 1 < s.levels && r >= s.meta[l + 1].row0) ++l;
  const LevelMeta m = s.meta[l];
  int rr = r - m.row0;
  const int gx = rr % m.nx;
  rr /= m.nx;
  const int gy = rr % m.ny;
  const int a = rr / m.ny;
  const int no = 5 + s.nc;
  v.base = s.logits[l] + (((long long)img * m.ny + gy) * m.nx + gx) * m.ld + a * (s.pitch > 0 ? s.pitch : no);
  const float sx = sigmoid_dec(v.base[0]), sy = sigmoid_dec(v.base[1]);
  const float sw = sigmoid_dec(v.base[2]), sh = sigmoid_dec(v.base[3]);
  v.obj = sigmoid_dec(v.base[4]);
  v.x = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.f), 0.5f), (float)gx), m.stride);
  v.y = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.f), 0.5f), (float)gy), m.stride);
  const float tw = __fmul_rn(sw, 2.f), th = __fmul_rn(sh, 2.f);
  v.w = __fmul_rn(__fmul_rn(tw, tw), m.anchor[2 * a]);
  v.h = __fmul_rn(__fmul_rn(th, th), m.anchor[2 * a + 1]);
  return v;
}

// Per-row candidate enumeration by one warp.  Calls emit(k, conf, cls) for the k-th candidate of the
// row in class order (lane-cooperative: emit is invoked by the lane that owns the class); returns the
// row's candidate count (uniform across the warp).
template <bool WRITE, typename Emit>
__device__ __forceinline__ int row_candidates(const RowView& v, int nc, bool multi_label, float thr,
                                              const unsigned char* __restrict__ class_mask, int lane, Emit emit) {
  if (!(v.obj > thr)) return 0;
  if (multi_label) {
    int total = 0;
    for (int c0 = 0; c0 < nc; c0 += 32) {
      const int c = c0 + lane;
      bool pass = false;
      float conf = 0.f;
      if (c < nc) {
        conf = __fmul_rn(v.cls_at(c), v.obj);
        pass = conf > thr && (class_mask == nullptr || class_mask[c]);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      if (WRITE && pass) emit(total + __popc(bal & ((1u << lane) - 1u)), conf, c);
      total += __popc(bal);
    }
    return total;
  }
  // best class: first maximum (torch.max on CPU), NaN anywhere -> max is NaN -> `conf > thr` false
  float best = -INFINITY;
  int bi = 0x7fffffff;
  bool has_nan = false;
  for (int c = lane; c < nc; c += 32) {
    const float conf = __fmul_rn(v.cls_at(c), v.obj);
    if (conf != conf) has_nan = true;
    if (conf > best) {  // strictly greater keeps the first index within this lane's ascending scan
      best = conf;
      bi = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  has_nan = __any_sync(0xffffffffu, has_nan);
  if (has_nan || !(best > thr) || bi >= nc) return 0;
  if (class_mask != nullptr && !class_mask[bi]) return 0;
  if (WRITE && lane == 0) emit(0, best, bi);
  return 1;
}

constexpr int kFilterThreads = 256;
constexpr int kFilterWarps = kFilterThreads / 32;

__global__ void __launch_bounds__(kFilterThreads) filter_count_kernel(RowSrc src, const unsigned char* class_mask,
                                                                      int* __restrict__ blk_counts, int rows_per_block,
                                                                      int nblk, int multi_label, float thr) {
  __shared__ int wsum[kFilterWarps];
  const int img = blockIdx.x / nblk, blk = blockIdx.x % nblk;
  const int r0 = blk * rows_per_block, r1 = min(r0 + rows_per_block, src.R);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int local = 0;
  for (int r = r0 + warp; r < r1; r += kFilterWarps) {
    RowView v = load_row(src, img, r);
    local += row_candidates<false>(v, src.nc, multi_label != 0, thr, class_mask, lane, [](int, float, int) {});
  }
  if (lane == 0) wsum[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < kFilterWarps; ++i) t += wsum[i];
    blk_counts[blockIdx.x] = t;
  }
}

// one CTA per image: exclusive scan of its nblk block counts
__global__ void __launch_bounds__(1024) filter_scan_kernel(const int* __restrict__ blk_counts,
                                                           int* __restrict__ blk_offsets, int* __restrict__ img_counts,
                                                           int nblk) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s, chunk_tot_s;
  const int img = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? blk_counts[(long long)img * nblk + i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int w = lane < nwarps ? warp_tot[lane] : 0;
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_tot[lane] = winc - w;  // exclusive warp offsets
      if (lane == 31) chunk_tot_s = winc;
    }
    __syncthreads();
    if (i < nblk) blk_offsets[(long long)img * nblk + i] = carry_s + warp_tot[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += chunk_tot_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) img_counts[img] = carry_s;
}

// single CTA: exclusive scan of img_counts -> img_offsets[N+1] (serial over N; N is a batch size)
__global__ void img_scan_kernel(const int* __restrict__ img_counts, long long* __restrict__ img_offsets, int N) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < N; ++i) {
      img_offsets[i] = t;
      t += img_counts[i];
    }
    img_offsets[N] = t;
  }
}

__global__ void __launch_bounds__(kFilterThreads) filter_write_kernel(
    RowSrc src, const unsigned char* class_mask, const int* __restrict__ blk_offsets,
    const long long* __restrict__ img_offsets, unsigned long long* __restrict__ keys, float* __restrict__ cand,
    int rows_per_block, int nblk, int multi_label, float thr, long long capacity) {
  extern __shared__ int row_off[];  // [rows_per_block + 1]
  const int img = blockIdx.x / nblk, blk = blockIdx.x % nblk;
  const int r0 = blk * rows_per_block, r1 = min(r0 + rows_per_block, src.R);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = r1 - r0;
  // pass 1: per-row counts
  for (int r = r0 + warp; r < r1; r += kFilterWarps) {
    RowView v = load_row(src, img, r);
    int c = row_candidates<false>(v, src.nc, multi_label != 0, thr, class_mask, lane, [](int, float, int) {});
    if (lane == 0) row_off[r - r0] = c;
  }
  __syncthreads();
  // exclusive scan over <= rows_per_block rows by warp 0 (chunks of 32 with carry)
  if (warp == 0) {
    int carry = 0;
    for (int base = 0; base < nrows; base += 32) {
      const int i = base + lane;
      const int v = i < nrows ? row_off[i] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (i < nrows) row_off[i] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  const long long gbase = img_offsets[img] + blk_offsets[(long long)img * nblk + blk];
  // pass 2: write
  for (int r = r0 + warp; r < r1; r += kFilterWarps) {
    RowView v = load_row(src, img, r);
    const long long g0 = gbase + row_off[r - r0];
    // xywh2xyxy, utils/general.py:539-546: x - w/2 etc. (w/2 is exact in fp32)
    const float hw = __fmul_rn(v.w, 0.5f), hh = __fmul_rn(v.h, 0.5f);
    const float x1 = __fsub_rn(v.x, hw), y1 = __fsub_rn(v.y, hh), x2 = __fadd_rn(v.x, hw), y2 = __fadd_rn(v.y, hh);
    row_candidates<true>(v, src.nc, multi_label != 0, thr, class_mask, lane, [&](int k, float conf, int cls) {
      const long long g = g0 + k;
      if (g < capacity) {
        float* c = cand + g * 6;
        c[0] = x1; c[1] = y1; c[2] = x2; c[3] = y2; c[4] = conf; c[5] = (float)cls;
        keys[g] = ((unsigned long long)(unsigned)img << 32) | (unsigned long long)(~__float_as_uint(conf));
      }
    });
  }
}

// ---- fused single-pass filter (Detect logits source) ---------------------------------------------------
// One CTA = one tile of P consecutive pixels of ONE anchor of one (image, level), taken in the reference row
// order (image, level, anchor, y, x — models/yolo.py:101-103), so "tile order" == "candidate order".  The
// tile's 5+nc logits per pixel are staged in shared memory (each logit is read from HBM once; a pixel's 1 KB
// is shared by the na tiles of its anchors), a warp per row decodes the box and turns the class logits into
// confidences in place, and the CTA's candidates are placed by a decoupled look-back scan over the tiles
// (status word = flag | count; tile ids come from an atomic ticket so every predecessor is already running).
// Result: order-preserving compaction of the whole batch in ONE pass, no count/scan/write re-reads.
constexpr int kFuseThreads = 256;
constexpr int kFuseWarps = kFuseThreads / 32;
constexpr int kFuseP = 64;                   // rows (pixels of one anchor) per tile

struct FuseArgs {
  const float* logits[5];
  LevelMeta meta[5];
  int tile0[6];        // first tile (within an image) of each level, [levels] = tiles per image
  int tpa[5];          // tiles per anchor plane
  int levels, nc, multi_label, N;
  float thr;
  long long capacity;
  int pitch;           // floats between the rows of two anchors inside a pixel's channel vector (5 + nc, or padded to 4n)
  const int* bin_thr;  // dense multi-label source: per-image score-key bin threshold of the pre-selection (nullptr: none)
  int* hist;           // HIST instantiation of the rows kernel: caller-zeroed [N][2048] score-key histogram (see dense_hist_kernel)
  unsigned long long* img_ctr;   // reserve mode: one reservation counter per image (nullptr: one counter for the batch)
  long long region;              // ... and the size of an image's region of the temporary buffers (capacity / N)
};

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive prefix of `count` over all tiles before `tile` (called by one full warp; 32 predecessors per poll —
// wider windows were measured slower in both filter kernels: 256-wide 0.71 vs 0.58 ms in the warp-per-row one,
// 128-wide 0.39 / 0.35 (with back-off) vs 0.34 ms in the thread-per-row one.  What a tile waits for is the slowest of
// its ~700 in-flight predecessors to publish a count, not the walk over their status words)
__device__ __forceinline__ long long tile_lookback(unsigned long long* status, int tile, int count, int lane) {
  constexpr unsigned long long kAgg = 1ull << 62, kIncl = 2ull << 62, kVal = (1ull << 62) - 1;
  if (tile == 0) {
    if (lane == 0) st_status(status, kIncl | (unsigned long long)count);
    return 0;
  }
  if (lane == 0) st_status(status + tile, kAgg | (unsigned long long)count);
  long long excl = 0;
  int j = tile - 1;
  while (true) {
    const int idx = j - lane;
    const unsigned long long w = idx >= 0 ? ld_status(status + idx) : kIncl;   // before tile 0: inclusive prefix 0
    const unsigned f = (unsigned)(w >> 62);
    const unsigned incl = __ballot_sync(0xffffffffu, f == 2u), notready = __ballot_sync(0xffffffffu, f == 0u);
    if (incl) {
      const int first = __ffs(incl) - 1;                       // nearest predecessor with an inclusive prefix
      const unsigned need = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
      if (notready & need) continue;                           // somebody closer is not published yet: poll again
      long long v = lane <= first ? (long long)(w & kVal) : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      excl += v;
      break;
    }
    if (notready) continue;
    long long v = (long long)(w & kVal);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    j -= 32;
  }
  if (lane == 0) st_status(status + tile, kIncl | (unsigned long long)(excl + count));
  return excl;
}

__global__ void __launch_bounds__(kFuseThreads) filter_fused_kernel(const __grid_constant__ FuseArgs fa,
                                                                    const unsigned char* __restrict__ class_mask,
                                                                    unsigned* __restrict__ ticket,
                                                                    unsigned long long* __restrict__ status,
                                                                    long long* __restrict__ img_offsets,
                                                                    unsigned long long* __restrict__ keys,
                                                                    float* __restrict__ cand) {
  extern __shared__ float tile[];     // [P][no]
  __shared__ int row_cnt[kFuseP];
  __shared__ long long base_s;
  __shared__ int tile_s;
  if (threadIdx.x == 0) tile_s = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile_id = tile_s;
  const int tiles_img = fa.tile0[fa.levels];
  const int img = tile_id / tiles_img;
  int t = tile_id - img * tiles_img;
  int l = 0;
  while (l + 1 < fa.levels && t >= fa.tile0[l + 1]) ++l;
  t -= fa.tile0[l];
  const LevelMeta& m = fa.meta[l];
  const int a = t / fa.tpa[l], ti = t - a * fa.tpa[l];
  const int npix = m.ny * m.nx;
  const int p0 = ti * kFuseP;
  const int np = min(kFuseP, npix - p0);
  const int nc = fa.nc, no = 5 + nc, ld = m.ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float thr = fa.thr;
  const float* gsrc = fa.logits[l] + ((long long)img * npix + p0) * ld + a * fa.pitch;

  // ---- stage + phase 1 (the warp that loads a row also processes it: only warp-level sync needed) ----
  // (rows start on 4-byte boundaries only — a*no floats into the pixel — hence 4-byte cp.async: every lane's copy is
  // in flight at once, no register staging)
  for (int row = warp; row < np; row += kFuseWarps) {
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(tile + row * no);
    const float* g = gsrc + (long long)row * ld;
    for (int c = lane; c < no; c += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst + c * 4), "l"(g + c) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
  for (int row = warp; row < np; row += kFuseWarps) {
    float* s = tile + row * no;
    const float hv = lane < 5 ? sigmoid_dec(s[lane]) : 0.f;
    const float sx = __shfl_sync(0xffffffffu, hv, 0), sy = __shfl_sync(0xffffffffu, hv, 1);
    const float sw = __shfl_sync(0xffffffffu, hv, 2), sh = __shfl_sync(0xffffffffu, hv, 3);
    const float obj = __shfl_sync(0xffffffffu, hv, 4);
    int cnt = 0;
    if (obj > thr) {
      if (fa.multi_label) {
        for (int c0 = 0; c0 < nc; c0 += 32) {
          const int c = c0 + lane;
          bool pass = false;
          if (c < nc) {
            const float conf = __fmul_rn(sigmoid_dec(s[5 + c]), obj);
            pass = conf > thr && (class_mask == nullptr || class_mask[c]);
            s[5 + c] = pass ? conf : -1.f;
          }
          cnt += __popc(__ballot_sync(0xffffffffu, pass));
        }
      } else {
        // best class: first maximum (torch.max on CPU); NaN anywhere -> no candidate
        float best = -INFINITY;
        int bi = 0x7fffffff;
        bool has_nan = false;
        for (int c = lane; c < nc; c += 32) {
          const float conf = __fmul_rn(sigmoid_dec(s[5 + c]), obj);
          if (conf != conf) has_nan = true;
          if (conf > best) {
            best = conf;
            bi = c;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
          }
        }
        has_nan = __any_sync(0xffffffffu, has_nan);
        const bool keep = !has_nan && best > thr && bi < nc && (class_mask == nullptr || class_mask[bi]);
        __syncwarp();
        if (keep && lane == 0) {
          s[4] = best;
          s[5] = (float)bi;
        }
        cnt = keep ? 1 : 0;
      }
      if (cnt > 0 && lane == 0) {
        const int pix = p0 + row;
        const int gy = pix / m.nx, gx = pix - gy * m.nx;
        // models/yolo.py:91-97 operation order, then xywh2xyxy (utils/general.py:539-546)
        const float x = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.f), 0.5f), (float)gx), m.stride);
        const float y = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.f), 0.5f), (float)gy), m.stride);
        const float tw = __fmul_rn(sw, 2.f), th = __fmul_rn(sh, 2.f);
        const float w = __fmul_rn(__fmul_rn(tw, tw), m.anchor[2 * a]);
        const float h = __fmul_rn(__fmul_rn(th, th), m.anchor[2 * a + 1]);
        const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
        s[0] = __fsub_rn(x, hw);
        s[1] = __fsub_rn(y, hh);
        s[2] = __fadd_rn(x, hw);
        s[3] = __fadd_rn(y, hh);
      }
    }
    if (lane == 0) row_cnt[row] = cnt;
  }
  __syncthreads();
  // ---- exclusive scan of the row counts, then this tile's place among all tiles (warp 0) ----
  if (warp == 0) {
    int carry = 0;
    for (int b0 = 0; b0 < np; b0 += 32) {
      const int i = b0 + lane;
      const int v = i < np ? row_cnt[i] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      if (i < np) row_cnt[i] = (v > 0) ? (carry + inc - v) : -1;   // offset, or -1 = empty row
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    const long long excl = tile_lookback(status, tile_id, carry, lane);
    if (lane == 0) {
      base_s = excl;
      if (tile_id == img * tiles_img) img_offsets[img] = excl;                          // first tile of an image
      if (tile_id == fa.N * tiles_img - 1) img_offsets[fa.N] = excl + carry;            // last tile: batch total
    }
  }
  __syncthreads();
  // ---- phase 2: write ----
  const long long base = base_s;
  for (int row = warp; row < np; row += kFuseWarps) {
    const int off = row_cnt[row];
    if (off < 0) continue;
    const float* s = tile + row * no;
    const float x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
    long long g0 = base + off;
    if (fa.multi_label) {
      for (int c0 = 0; c0 < nc; c0 += 32) {
        const int c = c0 + lane;
        const float conf = c < nc ? s[5 + c] : -1.f;
        const bool pass = conf > thr;                           // failed classes were overwritten with -1
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (pass) {
          const long long g = g0 + __popc(bal & ((1u << lane) - 1u));
          if (g < fa.capacity) {
            float2* cd = reinterpret_cast<float2*>(cand + g * 6);   // 24-byte rows: three aligned 8-byte stores
            cd[0] = make_float2(x1, y1);
            cd[1] = make_float2(x2, y2);
            cd[2] = make_float2(conf, (float)c);
            keys[g] = ((unsigned long long)(unsigned)img << 32) | (unsigned long long)(~__float_as_uint(conf));
          }
        }
        g0 += __popc(bal);
      }
    } else if (lane == 0 && g0 < fa.capacity) {
      float* cd = cand + g0 * 6;
      cd[0] = x1; cd[1] = y1; cd[2] = x2; cd[3] = y2; cd[4] = s[4]; cd[5] = s[5];
      keys[g0] = ((unsigned long long)(unsigned)img << 32) | (unsigned long long)(~__float_as_uint(s[4]));
    }
  }
}

// ---- fused single-pass filter, one THREAD per row (nc <= 96) -------------------------------------------------
// ncu of the warp-per-row kernel above: 86 thread-instructions per logit (a 32-lane warp walks one 85-logit row: three
// passes with the last one two-thirds empty, five shuffles, lane-0-only box decode, a second pass to re-test every
// class).  Here a thread owns a row.  Multi-label rows do not evaluate sigmoid(c)*obj for every class: since that
// product is monotonic in the class logit c, a class can only pass when c > logit(thr / obj); the kernel compares the
// raw logits against that bound minus a safety margin (the bound is clamped, so every class whose exact test could
// succeed is still evaluated) and runs the reference arithmetic -- fl(fl(sigmoid(c)) * obj) > thr -- only on those.
// The passing classes are remembered as a 96-bit mask, so the write pass touches nothing else.  Candidate sets, order
// and values are identical to the kernel above (tests compare both with the CPU restatement).
// Build-time A/B switches of the rows kernel (tools/build_variant.sh + tools/ab_filter.sh, same box, cfg-2, us per launch):
//   128-row tiles, CTA barrier after staging, loads consumed one by one ............ 174
//   + DMAY_FILTER_GROUPED_LDS (six LDS.128 issued back to back) .................. 164
//   + DMAY_FILTER_WARP_PRIVATE (a warp stages the rows it scans, __syncwarp only) . 158
//   + DMAY_FILTER_ROWS 96 / 64 (35 / 23.5 KB tiles: 6 / 9 resident CTAs per SM) ... 145 / 138   <- default
//   one survivor loop over the 96-bit set with two sigmoid chains per trip ....... 195 (removed)
#ifndef DMAY_FILTER_WARP_PRIVATE
#define DMAY_FILTER_WARP_PRIVATE 1
#endif
#ifndef DMAY_FILTER_GRP
#define DMAY_FILTER_GRP 6
#endif
#ifndef DMAY_FILTER_ROWS
#define DMAY_FILTER_ROWS 64
#endif
#ifndef DMAY_FILTER_GROUPED_LDS
#define DMAY_FILTER_GROUPED_LDS 1
#endif
constexpr int kRowsThreads = DMAY_FILTER_ROWS;   // rows (pixels of one anchor) per tile; threads per CTA = kRowsThreads * TPR
constexpr int kHistBins = 2048;     // score-key bins of the top-max_nms pre-selection
// Bin of a confidence: monotone (non-decreasing) in the sort key ~bits(conf), i.e. better scores -> lower bins.  The key's top 15
// bits (sign, exponent, 6 mantissa bits) minus those of conf = 1.0, clamped: confidences in (2^-21, 1] spread over 64 bins per binade
// (~640 bins for a 0.001 threshold).  The first form -- the top 11 bits of the key, 4 bins per binade -- put cfg-4b's 55 M candidates
// into ~40 bins: shared-memory atomics serialised on them, and the threshold bin alone held several times max_nms.
__device__ __forceinline__ unsigned key_bin(float conf) {
  const unsigned k = (~__float_as_uint(conf)) >> 17, base = (~0x3F800000u) >> 17;
  return min(k > base ? k - base : 0u, (unsigned)(kHistBins - 1));
}
constexpr int kRowsWide = 128;      // tile of the unpadded / dense layouts
constexpr int kRowsMaxNc = 96;
constexpr bool kAnchorFastest = true;

// RESERVE = false: tiles take a ticket and place their candidates with the decoupled look-back (final, ordered buffers).
// RESERVE = true : no ordering inside this kernel -- a tile reserves its run in the TEMPORARY buffers with one atomicAdd
//                  (`status` then holds tile_base[] / tile_cnt[]), tile_scan_kernel + tile_gather_kernel put the runs
//                  in order afterwards.  ncu of the look-back form: 46 % of the warp samples wait for the slowest of
//                  the ~700 in-flight predecessor tiles to publish a count; the gather costs 2 x 32 B per candidate.
// KIND 0: Detect logits, anchor rows of 5 + nc floats (only 4-byte aligned: 4-byte cp.async staging, scalar LDS).
// KIND 1: Detect logits whose anchor rows are padded to a multiple of 4 floats by the head GEMM (zero weight rows; the
//         layout is ours): rows are 16-byte aligned, staged with 16-byte cp.async into shared rows of pitch + 4 words (an
//         odd number of 16-byte chunks: conflict-free LDS.128), and the multi-label scan reads four logits per load.
// KIND 2: dense prediction [N, R, 5 + nc] (already decoded: utils/general.py's input when the caller holds a tensor):
//         values are used as they are (no sigmoid, no grid decode); the tile is one contiguous run of memory.
// TPR = threads per row.  The warps in flight are bounded by shared memory (11.8 KB of staged logits per 32 rows: 16 warps
// per SM with one thread per row -- ncu: 23 % occupancy, 28 % issue activity, the scan is a chain of dependent LDS).  With
// TPR = 2 (KIND 1) two adjacent threads share a row: thread 2r scans the 16-byte chunks [1, 13), thread 2r + 1 the rest
// (an offset of 12 chunks keeps the eight LDS.128 of a quarter-warp on distinct banks), twice the warps per staged byte.
// Candidate order is unchanged: the CTA scan runs over (row, half) in thread order, and the lower half holds the lower classes.
// HIST (KIND 1, multi-label): no candidate is written -- the exact confidences of the passing (row, class) pairs are binned
// by key_bin() of their score key into fa.hist[img][2048] (per-CTA shared histogram, one global atomic per non-empty
// bin): first pass of the top-max_nms pre-selection for candidate-rich Detect logits (dmay_nms_fused_prethreshold).
template <bool RESERVE, int KIND, int TPR = 1, int ROWS = kRowsThreads, bool HIST = false>
__global__ void __launch_bounds__(ROWS * TPR) filter_fused_rows_kernel(const __grid_constant__ FuseArgs fa,
                                                                         const unsigned char* __restrict__ class_mask,
                                                                         unsigned* __restrict__ ticket,
                                                                         unsigned long long* __restrict__ status,
                                                                         long long* __restrict__ img_offsets,
                                                                         unsigned long long* __restrict__ keys,
                                                                         float* __restrict__ cand) {
  extern __shared__ float tile[];     // [rows][no]
  constexpr int kThreads = ROWS * TPR;
  __shared__ int warp_tot[kThreads / 32];
  __shared__ long long base_s;
  __shared__ int tile_s;
  if (!RESERVE) {
    if (threadIdx.x == 0) tile_s = (int)atomicAdd(ticket, 1u);
    __syncthreads();
  }
  const int tiles_img = fa.tile0[fa.levels];
  int tile_id = RESERVE ? (int)blockIdx.x : tile_s;
  const int img = tile_id / tiles_img;
  int t = tile_id - img * tiles_img;
  int l = 0;
  while (l + 1 < fa.levels && t >= fa.tile0[l + 1]) ++l;
  t -= fa.tile0[l];
  const LevelMeta& m = fa.meta[l];
  int a = t / fa.tpa[l], ti = t - a * fa.tpa[l];
  if (RESERVE && kAnchorFastest) {
    // reservation does not care which CTA handles which tile: consecutive CTAs take the anchors of ONE pixel range, so the
    // interleaved thirds of those pixels' channel vectors are fetched at about the same time
    ti = t / m.na;
    a = t - ti * m.na;
    tile_id = img * tiles_img + fa.tile0[l] + a * fa.tpa[l] + ti;    // logical tile (reference order) for the scan / gather
  }
  const int npix = m.ny * m.nx;
  const int p0 = ti * ROWS;
  const int np = min(ROWS, npix - p0);
  const int nc = fa.nc, no = 5 + nc, ld = m.ld;
  const int pitch = KIND == 1 ? fa.pitch : no;              // global floats per anchor row
  const int srow = KIND == 1 ? fa.pitch + 4 : no;           // shared-memory words per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float thr = fa.thr;
  int* hist_s = reinterpret_cast<int*>(tile + ROWS * srow);   // HIST: 2048 bins behind the staged tile
  if (HIST)
    for (int i = threadIdx.x; i < kHistBins; i += kThreads) hist_s[i] = 0;
  {
    const float* gsrc = fa.logits[l] + ((long long)img * npix + p0) * ld + a * pitch;
    const uint32_t tile_sm = (uint32_t)__cvta_generic_to_shared(tile);
    bool staged = false;
    if (KIND == 1) {   // a warp copies whole rows, one 16-byte chunk per lane
      const int cpr = pitch >> 2;
#if DMAY_FILTER_WARP_PRIVATE
      // a warp stages exactly the rows its own lanes scan (TPR == 1: rows [32 warp, 32 warp + 32)): no CTA barrier between the
      // copies and the scan, the four warps of a tile drift apart and one warp's loads overlap another's arithmetic
      const int row_first = TPR == 1 ? warp * 32 : warp;
      const int row_last = TPR == 1 ? min(np, warp * 32 + 32) : np;
      const int row_inc = TPR == 1 ? 1 : kThreads / 32;
#else
      const int row_first = warp, row_last = np, row_inc = kThreads / 32;
#endif
      uint32_t sdst = tile_sm + (uint32_t)(row_first * srow) * 4u + (uint32_t)lane * 16u;
      const float* g = gsrc + (long long)row_first * ld + lane * 4;
      const uint32_t sstep = (uint32_t)row_inc * (uint32_t)srow * 4u;
      const long long gstep = (long long)row_inc * ld;
#pragma unroll 4
      for (int row = row_first; row < row_last; row += row_inc) {
        if (lane < cpr) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(g) : "memory");
        if (lane + 32 < cpr) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + 512u), "l"(g + 128) : "memory");
        sdst += sstep;
        g += gstep;
      }
      staged = true;
    } else if (KIND == 2) {   // the tile is contiguous: flat 16-byte copies when base and length allow
      const int words = np * no;
      if (((reinterpret_cast<uintptr_t>(gsrc) & 15u) == 0) && (words & 3) == 0) {
        for (int i = threadIdx.x; i < (words >> 2); i += kThreads)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile_sm + (uint32_t)i * 16u), "l"(gsrc + i * 4) : "memory");
        staged = true;
      }
    }
    if (!staged) {   // a warp copies whole rows (coalesced along the row; 4-byte cp.async, rows are only 4-byte aligned)
      // per row: nfull unconditional copies (lane, lane + 32, ...) and one partial; addresses advance by constants
      const int nfull = no >> 5, rem = no & 31;
      const uint32_t rem_off = (uint32_t)nfull * 128u;
      uint32_t sdst = tile_sm + (uint32_t)(warp * srow + lane) * 4u;
      const float* g = gsrc + (long long)warp * ld + lane;
      const uint32_t sstep = (uint32_t)(kThreads / 32) * (uint32_t)srow * 4u;
      const long long gstep = (long long)(kThreads / 32) * ld;
      for (int row = warp; row < np; row += kThreads / 32) {
        // no <= 5 + kRowsMaxNc = 101: at most three full 32-lane copies; straight-line code with uniform predicates
        if (nfull > 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(g) : "memory");
        if (nfull > 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst + 128u), "l"(g + 32) : "memory");
        if (nfull > 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst + 256u), "l"(g + 64) : "memory");
        if (lane < rem) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst + rem_off), "l"(g + nfull * 32) : "memory");
        sdst += sstep;
        g += gstep;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
#if DMAY_FILTER_WARP_PRIVATE
  if (KIND == 1 && TPR == 1 && !HIST) __syncwarp();
  else __syncthreads();
#else
  __syncthreads();
#endif

  const int row = TPR == 1 ? (int)threadIdx.x : (int)(threadIdx.x >> 1);
  const int half = TPR == 1 ? 0 : (int)(threadIdx.x & 1);
  float* s = tile + row * srow;   // KIND 0 / 2: row stride 5 + nc words, conflict-free whenever it is odd (nc = 80, 10, ...)
  int cnt = 0;
  uint32_t pm[4] = {0u, 0u, 0u, 0u};  // passing classes (bit = class; KIND 1 indexes words 0..3 statically)
  float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f, bconf = 0.f;
  int bcls = 0;
  float bc_best = -INFINITY;       // best-class mode: this thread's (half) row maximum, combined after the branch
  int bc_bi = 0x7fffffff;
  bool bc_nan = false, bc_valid = false;
  if (row < np) {
    const float obj = KIND == 2 ? s[4] : sigmoid_dec(s[4]);
    if (obj > thr) {
      if (fa.multi_label) {
        if (KIND == 2) {
          // dense rows hold the class confidences themselves: the reference arithmetic is one multiply (general.py:677)
          const unsigned bt = fa.bin_thr != nullptr ? (unsigned)fa.bin_thr[img] : 0xFFFFFFFFu;   // see dense_hist_kernel
#pragma unroll
          for (int w = 0; w < 3; ++w) {
            uint32_t mk = 0u;
            const int cend = min(32, nc - w * 32);
            float* sc = s + 5 + w * 32;
#pragma unroll 8
            for (int cc = 0; cc < cend; ++cc) {
              const float conf = __fmul_rn(sc[cc], obj);
              if (conf > thr && key_bin(conf) <= bt && (class_mask == nullptr || class_mask[w * 32 + cc])) {
                sc[cc] = conf;
                mk |= 1u << cc;
                ++cnt;
              }
            }
            pm[w] = mk;
          }
        } else {
          // conservative pre-filter on the raw logit (see the header comment); exact test only for survivors
          const float q = (thr / obj) * (1.0f - 4e-6f);
          float t_lo = -INFINITY;
          if (q > 0.f) t_lo = q < 1.f ? fminf(__logf(q / (1.0f - q)) - 0.02f, 10.0f) : 10.0f;
          if (KIND == 1) {
            const unsigned bt1 = (!HIST && fa.bin_thr != nullptr) ? (unsigned)fa.bin_thr[img] : 0xFFFFFFFFu;
            // four logits per LDS.128; chunk k holds words 4k .. 4k+3 of the row, class c sits at word 5 + c
            const float4* s4 = reinterpret_cast<const float4*>(s);
            const int nchunk = (no + 3) >> 2;
            const int k_lo = (TPR == 2 && half == 1) ? min(13, nchunk) : 1;
            const int k_hi = (TPR == 2 && half == 0) ? min(13, nchunk) : nchunk;
            // pass 1, branch-free: which class logits clear the conservative bound.  (Testing and evaluating in one loop made
            // a warp run the ~40-instruction exact path for every class that ANY of its 32 rows cleared -- 60 % of the classes
            // at a 2-3 % candidate rate -- a ~10 us dependent chain per tile that no amount of warps, staging width, tile order
            // or load overlap moved; see profiles/r2_notes.md section 6.)
            uint32_t pre[4] = {0u, 0u, 0u, 0u};
#if DMAY_FILTER_GROUPED_LDS
            // the LDS.128 of a group are issued back to back (ptxas kept 32 registers and consumed every load at once)
            constexpr int kGrp = DMAY_FILTER_GRP;
#pragma unroll
            for (int k0 = 1; k0 < (5 + kRowsMaxNc + 3) / 4; k0 += kGrp) {
              float4 v4[kGrp];
#pragma unroll
              for (int u = 0; u < kGrp; ++u) {
                const int k = k0 + u;
                v4[u] = (k >= k_lo && k < k_hi) ? s4[k] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
              }
#pragma unroll
              for (int u = 0; u < kGrp; ++u) {
                const int k = k0 + u;
                const float vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int c = 4 * k + e - 5;
                  if (c >= 0 && c < kRowsMaxNc) pre[c >> 5] |= ((c < nc && vv[e] > t_lo) ? 1u : 0u) << (c & 31);
                }
              }
            }
#else
#pragma unroll
            for (int k = 1; k < (5 + kRowsMaxNc + 3) / 4; ++k) {
              if (k >= k_lo && k < k_hi) {
                const float4 v4 = s4[k];
                const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int c = 4 * k + e - 5;
                  if (c >= 0 && c < nc) pre[c >> 5] |= (vv[e] > t_lo ? 1u : 0u) << (c & 31);
                }
              }
            }
#endif
            // pass 2: the reference arithmetic for the survivors only (iterations = the busiest row of the warp)
#pragma unroll
            for (int w = 0; w < 3; ++w) {
              uint32_t mk = pre[w];
              while (mk) {
                const int c = w * 32 + __ffs(mk) - 1;
                mk &= mk - 1;
                const float conf = __fmul_rn(sigmoid_dec(s[5 + c]), obj);
                if (conf > thr && (class_mask == nullptr || class_mask[c])) {
                  const unsigned kbin = key_bin(conf);
                  if (HIST) {
                    atomicAdd(&hist_s[kbin], 1);
                  } else if (kbin <= bt1) {   // pre-selection (when given): only what the top-max_nms selection can keep
                    s[5 + c] = conf;
                    pm[w] |= 1u << (c & 31);
                    ++cnt;
                  }
                }
              }
            }
          } else {
#pragma unroll
            for (int w = 0; w < 3; ++w) {   // 32 classes per mask word (static register indexing)
              uint32_t pre = 0u, mk = 0u;
              const int cend = min(32, nc - w * 32);
              float* sc = s + 5 + w * 32;
#pragma unroll 8
              for (int cc = 0; cc < cend; ++cc) pre |= (sc[cc] > t_lo ? 1u : 0u) << cc;   // pass 1, branch-free (see KIND 1)
              while (pre) {                                                               // pass 2: survivors only
                const int cc = __ffs(pre) - 1;
                pre &= pre - 1;
                const float conf = __fmul_rn(sigmoid_dec(sc[cc]), obj);
                if (conf > thr && (class_mask == nullptr || class_mask[w * 32 + cc])) {
                  sc[cc] = conf;
                  mk |= 1u << cc;
                  ++cnt;
                }
              }
              pm[w] = mk;
            }
          }
        }
      } else {
        // best class: first maximum (torch.max on CPU); NaN anywhere -> no candidate
        float best = -INFINITY;
        int bi = 0x7fffffff;
        bool has_nan = false;
        const int c_lo = (TPR == 2 && half == 1) ? (nc + 1) / 2 : 0, c_hi = (TPR == 2 && half == 0) ? (nc + 1) / 2 : nc;
        for (int c = c_lo; c < c_hi; ++c) {
          const float conf = __fmul_rn(KIND == 2 ? s[5 + c] : sigmoid_dec(s[5 + c]), obj);
          if (conf != conf) has_nan = true;
          if (conf > best) {
            best = conf;
            bi = c;
          }
        }
        bc_best = best;
        bc_bi = bi;
        bc_nan = has_nan;
        bc_valid = true;
      }
    }
  }
  if (!fa.multi_label) {   // warp-uniform: the two halves of a row meet here outside any divergent region
    if (TPR == 2) {        // first maximum over the two halves of the row
      const float ob = __shfl_xor_sync(0xffffffffu, bc_best, 1);
      const int oi = __shfl_xor_sync(0xffffffffu, bc_bi, 1);
      const bool on = __shfl_xor_sync(0xffffffffu, bc_nan ? 1 : 0, 1) != 0;
      if (ob > bc_best || (ob == bc_best && oi < bc_bi)) {
        bc_best = ob;
        bc_bi = oi;
      }
      bc_nan = bc_nan || on;
    }
    if (bc_valid && half == 0 && !bc_nan && bc_best > thr && bc_bi < nc && (class_mask == nullptr || class_mask[bc_bi])) {
      cnt = 1;
      bconf = bc_best;
      bcls = bc_bi;
    }
  }
  if (cnt > 0) {
    float x, y, w, h;
    if (KIND == 2) {
      x = s[0]; y = s[1]; w = s[2]; h = s[3];
    } else {
      const int pix = p0 + row;
      const int gy = pix / m.nx, gx = pix - gy * m.nx;
      const float sx = sigmoid_dec(s[0]), sy = sigmoid_dec(s[1]), sw = sigmoid_dec(s[2]), sh = sigmoid_dec(s[3]);
      // models/yolo.py:91-97 operation order
      x = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.f), 0.5f), (float)gx), m.stride);
      y = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.f), 0.5f), (float)gy), m.stride);
      const float tw = __fmul_rn(sw, 2.f), th = __fmul_rn(sh, 2.f);
      w = __fmul_rn(__fmul_rn(tw, tw), m.anchor[2 * a]);
      h = __fmul_rn(__fmul_rn(th, th), m.anchor[2 * a + 1]);
    }
    // xywh2xyxy (utils/general.py:539-546)
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    x1 = __fsub_rn(x, hw);
    y1 = __fsub_rn(y, hh);
    x2 = __fadd_rn(x, hw);
    y2 = __fadd_rn(y, hh);
  }
  if (HIST) {   // flush the tile's histogram; nothing else to do
    __syncthreads();
    for (int i = threadIdx.x; i < kHistBins; i += kThreads)
      if (hist_s[i]) atomicAdd(&fa.hist[(long long)img * kHistBins + i], hist_s[i]);
    return;
  }
  // ---- exclusive scan of the row counts over the CTA, then this tile's place among all tiles (warp 0) ----
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const int v = warp_tot[w];
    if (w < warp) before += v;
    total += v;
  }
  if (RESERVE) {
    if (threadIdx.x == 0) {
      const long long ntiles = (long long)fa.N * tiles_img;
      long long* tile_base = reinterpret_cast<long long*>(status);
      int* tile_cnt = reinterpret_cast<int*>(tile_base + ntiles);
      // one counter per image: 12,864 same-address atomics with a return value were the pacing item of this kernel
      long long b = 0;
      if (total > 0)
        b = fa.img_ctr != nullptr ? (long long)img * fa.region + (long long)atomicAdd(fa.img_ctr + img, (unsigned long long)total)
                                  : (long long)atomicAdd(reinterpret_cast<unsigned long long*>(ticket), (unsigned long long)total);
      tile_base[tile_id] = b;
      tile_cnt[tile_id] = total;
      base_s = b;
    }
  } else if (warp == 0) {
    const long long excl = tile_lookback(status, tile_id, total, lane);
    if (lane == 0) {
      base_s = excl;
      if (tile_id == img * tiles_img) img_offsets[img] = excl;                          // first tile of an image
      if (tile_id == fa.N * tiles_img - 1) img_offsets[fa.N] = excl + total;            // last tile: batch total
    }
  }
  __syncthreads();
  if (cnt == 0) return;
  long long g = base_s + before + inc - cnt;
  const long long g_lim = (RESERVE && fa.img_ctr != nullptr) ? (long long)(img + 1) * fa.region : fa.capacity;   // never past the image's region
  const unsigned long long img_hi = (unsigned long long)(unsigned)img << 32;
  if (fa.multi_label) {
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      uint32_t mk = pm[w];
      while (mk) {
        const int c = w * 32 + __ffs(mk) - 1;
        mk &= mk - 1;
        if (g < g_lim) {
          const float conf = s[5 + c];
          float2* cd = reinterpret_cast<float2*>(cand + g * 6);   // 24-byte rows: three aligned 8-byte stores
          cd[0] = make_float2(x1, y1);
          cd[1] = make_float2(x2, y2);
          cd[2] = make_float2(conf, (float)c);
          keys[g] = img_hi | (unsigned long long)(~__float_as_uint(conf));
        }
        ++g;
      }
    }
  } else if (g < g_lim) {
    float2* cd = reinterpret_cast<float2*>(cand + g * 6);
    cd[0] = make_float2(x1, y1);
    cd[1] = make_float2(x2, y2);
    cd[2] = make_float2(bconf, (float)bcls);
    keys[g] = img_hi | (unsigned long long)(~__float_as_uint(bconf));
  }
}

// ---- persistent form of the reserve-mode kernel -------------------------------------------------------------------------
// The one-tile-per-CTA kernel runs in lock-step waves: every CTA loads (DRAM busy, SMs idle), then scans and writes (SMs busy,
// DRAM idle); ncu: 2.5 TB/s, and neither more warps per row nor another tile order moved it.  Here a CTA is persistent (grid =
// resident CTAs), walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... and keeps TWO tile buffers: the cp.async loads of the
// next tile are in flight while the current one is scanned.  With two threads per row the two resident CTAs of an SM still
// run 16 warps.  Same tiles, same reservation scheme, same outputs as filter_fused_rows_kernel<true, KIND, TPR>.
template <int KIND, int TPR>
__global__ void __launch_bounds__(kRowsThreads * TPR) filter_rows_persistent_kernel(const __grid_constant__ FuseArgs fa,
                                                                                   const unsigned char* __restrict__ class_mask,
                                                                                   unsigned* __restrict__ ticket,
                                                                                   unsigned long long* __restrict__ status,
                                                                                   long long* __restrict__ img_offsets,
                                                                                   unsigned long long* __restrict__ keys,
                                                                                   float* __restrict__ cand) {
  constexpr bool RESERVE = true;
  constexpr int kThreads = kRowsThreads * TPR;
  extern __shared__ float tile_all[];     // 2 x [rows][srow]
  __shared__ int warp_tot[kThreads / 32];
  __shared__ long long base_s;
  const int nc = fa.nc, no = 5 + nc;
  const int pitch = KIND == 1 ? fa.pitch : no;
  const int srow = KIND == 1 ? fa.pitch + 4 : no;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float thr = fa.thr;
  const int tiles_img = fa.tile0[fa.levels];
  const int ntiles_all = fa.N * tiles_img;
  const int buf_words = kRowsThreads * srow;
  struct TileIdx {
    int img, l, a, p0, np, npix, tile_id;
  };
  auto decode = [&](int phys) {
    TileIdx r;
    r.img = phys / tiles_img;
    int t = phys - r.img * tiles_img;
    int l = 0;
    while (l + 1 < fa.levels && t >= fa.tile0[l + 1]) ++l;
    t -= fa.tile0[l];
    const int na = fa.meta[l].na;
    const int ti = t / na;                 // anchor fastest: neighbouring CTAs share a pixel range
    r.a = t - ti * na;
    r.l = l;
    r.tile_id = r.img * tiles_img + fa.tile0[l] + r.a * fa.tpa[l] + ti;
    r.npix = fa.meta[l].ny * fa.meta[l].nx;
    r.p0 = ti * kRowsThreads;
    r.np = min(kRowsThreads, r.npix - r.p0);
    return r;
  };
  auto stage = [&](const TileIdx& ti_, float* tile) {   // issue (do not wait for) the copies of one tile
    const int img = ti_.img, l = ti_.l, a = ti_.a, p0 = ti_.p0, np = ti_.np, npix = ti_.npix, ld = fa.meta[l].ld;
  {
    const float* gsrc = fa.logits[l] + ((long long)img * npix + p0) * ld + a * pitch;
    const uint32_t tile_sm = (uint32_t)__cvta_generic_to_shared(tile);
    bool staged = false;
    if (KIND == 1) {   // a warp copies whole rows, one 16-byte chunk per lane
      const int cpr = pitch >> 2;
      uint32_t sdst = tile_sm + (uint32_t)(warp * srow) * 4u + (uint32_t)lane * 16u;
      const float* g = gsrc + (long long)warp * ld + lane * 4;
      const uint32_t sstep = (uint32_t)(kThreads / 32) * (uint32_t)srow * 4u;
      const long long gstep = (long long)(kThreads / 32) * ld;
#pragma unroll 4
      for (int row = warp; row < np; row += kThreads / 32) {
        if (lane < cpr) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(g) : "memory");
        if (lane + 32 < cpr) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + 512u), "l"(g + 128) : "memory");
        sdst += sstep;
        g += gstep;
      }
      staged = true;
    } else if (KIND == 2) {   // the tile is contiguous: flat 16-byte copies when base and length allow
      const int words = np * no;
      if (((reinterpret_cast<uintptr_t>(gsrc) & 15u) == 0) && (words & 3) == 0) {
        for (int i = threadIdx.x; i < (words >> 2); i += kThreads)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile_sm + (uint32_t)i * 16u), "l"(gsrc + i * 4) : "memory");
        staged = true;
      }
    }
    if (!staged) {   // a warp copies whole rows (coalesced along the row; 4-byte cp.async, rows are only 4-byte aligned)
      // per row: nfull unconditional copies (lane, lane + 32, ...) and one partial; addresses advance by constants
      const int nfull = no >> 5, rem = no & 31;
      const uint32_t rem_off = (uint32_t)nfull * 128u;
      uint32_t sdst = tile_sm + (uint32_t)(warp * srow + lane) * 4u;
      const float* g = gsrc + (long long)warp * ld + lane;
      const uint32_t sstep = (uint32_t)(kThreads / 32) * (uint32_t)srow * 4u;
      const long long gstep = (long long)(kThreads / 32) * ld;
      for (int row = warp; row < np; row += kThreads / 32) {
        // no <= 5 + kRowsMaxNc = 101: at most three full 32-lane copies; straight-line code with uniform predicates
        if (nfull > 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(g) : "memory");
        if (nfull > 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst + 128u), "l"(g + 32) : "memory");
        if (nfull > 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst + 256u), "l"(g + 64) : "memory");
        if (lane < rem) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst + rem_off), "l"(g + nfull * 32) : "memory");
        sdst += sstep;
        g += gstep;
      }
    }
  }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int phys = (int)blockIdx.x;
  if (phys < ntiles_all) stage(decode(phys), tile_all);
  for (int it = 0; phys < ntiles_all; ++it, phys += (int)gridDim.x) {
    float* tile = tile_all + (it & 1) * buf_words;
    if (phys + (int)gridDim.x < ntiles_all) {
      stage(decode(phys + (int)gridDim.x), tile_all + ((it + 1) & 1) * buf_words);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const TileIdx cur = decode(phys);
    const LevelMeta& m = fa.meta[cur.l];
    const int img = cur.img, a = cur.a, p0 = cur.p0, np = cur.np, tile_id = cur.tile_id;

  const int row = TPR == 1 ? (int)threadIdx.x : (int)(threadIdx.x >> 1);
  const int half = TPR == 1 ? 0 : (int)(threadIdx.x & 1);
  float* s = tile + row * srow;   // KIND 0 / 2: row stride 5 + nc words, conflict-free whenever it is odd (nc = 80, 10, ...)
  int cnt = 0;
  uint32_t pm[4] = {0u, 0u, 0u, 0u};  // passing classes (bit = class; KIND 1 indexes words 0..3 statically)
  float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f, bconf = 0.f;
  int bcls = 0;
  float bc_best = -INFINITY;       // best-class mode: this thread's (half) row maximum, combined after the branch
  int bc_bi = 0x7fffffff;
  bool bc_nan = false, bc_valid = false;
  if (row < np) {
    const float obj = KIND == 2 ? s[4] : sigmoid_dec(s[4]);
    if (obj > thr) {
      if (fa.multi_label) {
        if (KIND == 2) {
          // dense rows hold the class confidences themselves: the reference arithmetic is one multiply (general.py:677)
          const unsigned bt = fa.bin_thr != nullptr ? (unsigned)fa.bin_thr[img] : 0xFFFFFFFFu;   // see dense_hist_kernel
#pragma unroll
          for (int w = 0; w < 3; ++w) {
            uint32_t mk = 0u;
            const int cend = min(32, nc - w * 32);
            float* sc = s + 5 + w * 32;
#pragma unroll 8
            for (int cc = 0; cc < cend; ++cc) {
              const float conf = __fmul_rn(sc[cc], obj);
              if (conf > thr && key_bin(conf) <= bt && (class_mask == nullptr || class_mask[w * 32 + cc])) {
                sc[cc] = conf;
                mk |= 1u << cc;
                ++cnt;
              }
            }
            pm[w] = mk;
          }
        } else {
          // conservative pre-filter on the raw logit (see the header comment); exact test only for survivors
          const float q = (thr / obj) * (1.0f - 4e-6f);
          float t_lo = -INFINITY;
          if (q > 0.f) t_lo = q < 1.f ? fminf(__logf(q / (1.0f - q)) - 0.02f, 10.0f) : 10.0f;
          if (KIND == 1) {
            // four logits per LDS.128; chunk k holds words 4k .. 4k+3 of the row, class c sits at word 5 + c
            const float4* s4 = reinterpret_cast<const float4*>(s);
            const int nchunk = (no + 3) >> 2;
            const int k_lo = (TPR == 2 && half == 1) ? min(13, nchunk) : 1;
            const int k_hi = (TPR == 2 && half == 0) ? min(13, nchunk) : nchunk;
            // pass 1, branch-free: which class logits clear the conservative bound.  (Testing and evaluating in one loop made
            // a warp run the ~40-instruction exact path for every class that ANY of its 32 rows cleared -- 60 % of the classes
            // at a 2-3 % candidate rate -- a ~10 us dependent chain per tile that no amount of warps, staging width, tile order
            // or load overlap moved; see profiles/r2_notes.md section 6.)
            uint32_t pre[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 1; k < (5 + kRowsMaxNc + 3) / 4; ++k) {
              if (k >= k_lo && k < k_hi) {
                const float4 v4 = s4[k];
                const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int c = 4 * k + e - 5;
                  if (c >= 0 && c < nc) pre[c >> 5] |= (vv[e] > t_lo ? 1u : 0u) << (c & 31);
                }
              }
            }
            // pass 2: the reference arithmetic for the survivors only (iterations = the busiest row of the warp)
#pragma unroll
            for (int w = 0; w < 3; ++w) {
              uint32_t mk = pre[w];
              while (mk) {
                const int c = w * 32 + __ffs(mk) - 1;
                mk &= mk - 1;
                const float conf = __fmul_rn(sigmoid_dec(s[5 + c]), obj);
                if (conf > thr && (class_mask == nullptr || class_mask[c])) {
                  s[5 + c] = conf;
                  pm[w] |= 1u << (c & 31);
                  ++cnt;
                }
              }
            }
          } else {
#pragma unroll
            for (int w = 0; w < 3; ++w) {   // 32 classes per mask word (static register indexing)
              uint32_t pre = 0u, mk = 0u;
              const int cend = min(32, nc - w * 32);
              float* sc = s + 5 + w * 32;
#pragma unroll 8
              for (int cc = 0; cc < cend; ++cc) pre |= (sc[cc] > t_lo ? 1u : 0u) << cc;   // pass 1, branch-free (see KIND 1)
              while (pre) {                                                               // pass 2: survivors only
                const int cc = __ffs(pre) - 1;
                pre &= pre - 1;
                const float conf = __fmul_rn(sigmoid_dec(sc[cc]), obj);
                if (conf > thr && (class_mask == nullptr || class_mask[w * 32 + cc])) {
                  sc[cc] = conf;
                  mk |= 1u << cc;
                  ++cnt;
                }
              }
              pm[w] = mk;
            }
          }
        }
      } else {
        // best class: first maximum (torch.max on CPU); NaN anywhere -> no candidate
        float best = -INFINITY;
        int bi = 0x7fffffff;
        bool has_nan = false;
        const int c_lo = (TPR == 2 && half == 1) ? (nc + 1) / 2 : 0, c_hi = (TPR == 2 && half == 0) ? (nc + 1) / 2 : nc;
        for (int c = c_lo; c < c_hi; ++c) {
          const float conf = __fmul_rn(KIND == 2 ? s[5 + c] : sigmoid_dec(s[5 + c]), obj);
          if (conf != conf) has_nan = true;
          if (conf > best) {
            best = conf;
            bi = c;
          }
        }
        bc_best = best;
        bc_bi = bi;
        bc_nan = has_nan;
        bc_valid = true;
      }
    }
  }
  if (!fa.multi_label) {   // warp-uniform: the two halves of a row meet here outside any divergent region
    if (TPR == 2) {        // first maximum over the two halves of the row
      const float ob = __shfl_xor_sync(0xffffffffu, bc_best, 1);
      const int oi = __shfl_xor_sync(0xffffffffu, bc_bi, 1);
      const bool on = __shfl_xor_sync(0xffffffffu, bc_nan ? 1 : 0, 1) != 0;
      if (ob > bc_best || (ob == bc_best && oi < bc_bi)) {
        bc_best = ob;
        bc_bi = oi;
      }
      bc_nan = bc_nan || on;
    }
    if (bc_valid && half == 0 && !bc_nan && bc_best > thr && bc_bi < nc && (class_mask == nullptr || class_mask[bc_bi])) {
      cnt = 1;
      bconf = bc_best;
      bcls = bc_bi;
    }
  }
  if (cnt > 0) {
    float x, y, w, h;
    if (KIND == 2) {
      x = s[0]; y = s[1]; w = s[2]; h = s[3];
    } else {
      const int pix = p0 + row;
      const int gy = pix / m.nx, gx = pix - gy * m.nx;
      const float sx = sigmoid_dec(s[0]), sy = sigmoid_dec(s[1]), sw = sigmoid_dec(s[2]), sh = sigmoid_dec(s[3]);
      // models/yolo.py:91-97 operation order
      x = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.f), 0.5f), (float)gx), m.stride);
      y = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.f), 0.5f), (float)gy), m.stride);
      const float tw = __fmul_rn(sw, 2.f), th = __fmul_rn(sh, 2.f);
      w = __fmul_rn(__fmul_rn(tw, tw), m.anchor[2 * a]);
      h = __fmul_rn(__fmul_rn(th, th), m.anchor[2 * a + 1]);
    }
    // xywh2xyxy (utils/general.py:539-546)
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    x1 = __fsub_rn(x, hw);
    y1 = __fsub_rn(y, hh);
    x2 = __fadd_rn(x, hw);
    y2 = __fadd_rn(y, hh);
  }
  // ---- exclusive scan of the row counts over the CTA, then this tile's place among all tiles (warp 0) ----
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const int v = warp_tot[w];
    if (w < warp) before += v;
    total += v;
  }
  if (RESERVE) {
    if (threadIdx.x == 0) {
      const long long ntiles = (long long)fa.N * tiles_img;
      long long* tile_base = reinterpret_cast<long long*>(status);
      int* tile_cnt = reinterpret_cast<int*>(tile_base + ntiles);
      // one counter per image: 12,864 same-address atomics with a return value were the pacing item of this kernel
      long long b = 0;
      if (total > 0)
        b = fa.img_ctr != nullptr ? (long long)img * fa.region + (long long)atomicAdd(fa.img_ctr + img, (unsigned long long)total)
                                  : (long long)atomicAdd(reinterpret_cast<unsigned long long*>(ticket), (unsigned long long)total);
      tile_base[tile_id] = b;
      tile_cnt[tile_id] = total;
      base_s = b;
    }
  } else if (warp == 0) {
    const long long excl = tile_lookback(status, tile_id, total, lane);
    if (lane == 0) {
      base_s = excl;
      if (tile_id == img * tiles_img) img_offsets[img] = excl;                          // first tile of an image
      if (tile_id == fa.N * tiles_img - 1) img_offsets[fa.N] = excl + total;            // last tile: batch total
    }
  }
  __syncthreads();
  if (cnt != 0) {
  long long g = base_s + before + inc - cnt;
  const long long g_lim = (RESERVE && fa.img_ctr != nullptr) ? (long long)(img + 1) * fa.region : fa.capacity;   // never past the image's region
  const unsigned long long img_hi = (unsigned long long)(unsigned)img << 32;
  if (fa.multi_label) {
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      uint32_t mk = pm[w];
      while (mk) {
        const int c = w * 32 + __ffs(mk) - 1;
        mk &= mk - 1;
        if (g < g_lim) {
          const float conf = s[5 + c];
          float2* cd = reinterpret_cast<float2*>(cand + g * 6);   // 24-byte rows: three aligned 8-byte stores
          cd[0] = make_float2(x1, y1);
          cd[1] = make_float2(x2, y2);
          cd[2] = make_float2(conf, (float)c);
          keys[g] = img_hi | (unsigned long long)(~__float_as_uint(conf));
        }
        ++g;
      }
    }
  } else if (g < g_lim) {
    float2* cd = reinterpret_cast<float2*>(cand + g * 6);
    cd[0] = make_float2(x1, y1);
    cd[1] = make_float2(x2, y2);
    cd[2] = make_float2(bconf, (float)bcls);
    keys[g] = img_hi | (unsigned long long)(~__float_as_uint(bconf));
  }
  }
    __syncthreads();   // every thread is done with this buffer, warp_tot and base_s before they are reused
  }
}

// ---- pre-selection for candidate-dense multi-label predictions --------------------------------------------------------
// utils/general.py:702-703 keeps only the max_nms = 30000 best candidates of an image.  A dense multi-label prediction can
// expand to 10x that (BASELINE cfg-5: 25,200 rows x 10 classes = 252 k candidates per image, 2 GB of candidate records per
// batch of 256 for 0.39 GB of input), and seven eighths of what the filter writes is thrown away by the top-K selection.
// Two cheap passes over the INPUT avoid that: a histogram of key_bin() of every candidate's score key per image
// (dense_hist_kernel), the bin b* that contains the K-th best key (hist_threshold_kernel), and a filter that only writes
// candidates whose key bin is <= b*.  Everything the exact top-K selection can keep (keys <= the K-th key, ties included)
// has a bin <= b*, and the survivors keep their candidate order, so the selection / sort / greedy stages see a superset of
// the top K in the same relative order and return the same detections bit for bit.
__global__ void __launch_bounds__(256) dense_hist_kernel(const float* __restrict__ pred, const unsigned char* __restrict__ class_mask,
                                                         int* __restrict__ hist, int R, int nc, float thr, int rows_per_cta) {
  __shared__ int h[kHistBins];
  const int img = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, R);
  for (int i = threadIdx.x; i < kHistBins; i += 256) h[i] = 0;
  __syncthreads();
  const int no = 5 + nc;
  for (int r = r0 + threadIdx.x; r < r1; r += 256) {
    const float* row = pred + ((long long)img * R + r) * no;
    const float obj = row[4];
    if (!(obj > thr)) continue;
    for (int c = 0; c < nc; ++c) {
      const float conf = __fmul_rn(row[5 + c], obj);
      if (conf > thr && (class_mask == nullptr || class_mask[c])) atomicAdd(&h[key_bin(conf)], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHistBins; i += 256)
    if (h[i]) atomicAdd(&hist[(long long)img * kHistBins + i], h[i]);
}

// bin_thr[img] = first bin whose inclusive count reaches K (all bins when the image has at most K candidates)
__global__ void __launch_bounds__(1024) hist_threshold_kernel(const int* __restrict__ hist, int* __restrict__ bin_thr, int K) {
  __shared__ int warp_tot[32];
  __shared__ int found;
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) found = kHistBins - 1;
  const int h0 = hist[(long long)img * kHistBins + 2 * tid], h1 = hist[(long long)img * kHistBins + 2 * tid + 1];
  int inc = h0 + h1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int v = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    warp_tot[lane] = v;   // inclusive over warps
  }
  __syncthreads();
  const int excl = inc - (h0 + h1) + (warp > 0 ? warp_tot[warp - 1] : 0);
  if (excl < K && excl + h0 >= K) found = 2 * tid;              // exactly one thread matches (prefix sums are monotone)
  else if (excl + h0 < K && excl + h0 + h1 >= K) found = 2 * tid + 1;
  __syncthreads();
  if (tid == 0) bin_thr[img] = found;
}

// exclusive scan of the per-tile candidate counts (reference tile order == candidate order) -> tile_off[], img_offsets[]
__global__ void __launch_bounds__(1024) tile_scan_kernel(const int* __restrict__ tile_cnt, long long* __restrict__ tile_off,
                                                         long long* __restrict__ img_offsets, int* __restrict__ img_counts,
                                                         long long ntiles, int tiles_img, int N) {
  __shared__ long long warp_sum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ntiles < 2^31 (checked by the caller): 32-bit indices, and the image boundary test is a running counter -- the 64-bit
  // modulo per tile made this single CTA 23 us for the 25.7 k tiles of cfg-2
  const int nt = (int)ntiles, per = (nt + 1023) / 1024;
  const int t0 = min(tid * per, nt), t1 = min(t0 + per, nt);
  long long local = 0;
#pragma unroll 4
  for (int t = t0; t < t1; ++t) local += tile_cnt[t];
  long long inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    long long v = warp_sum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    warp_sum[lane] = v;   // inclusive over warps
  }
  __syncthreads();
  long long run = inc - local + (warp > 0 ? warp_sum[warp - 1] : 0);
  int img = t0 / tiles_img, in_img = t0 - img * tiles_img;
  for (int t = t0; t < t1; ++t) {
    tile_off[t] = run;
    if (in_img == 0) img_offsets[img] = run;
    if (++in_img == tiles_img) {
      in_img = 0;
      ++img;
    }
    run += tile_cnt[t];
  }
  if (tid == 1023) img_offsets[N] = warp_sum[31];
  __syncthreads();
  __threadfence_block();
  // per-image counts from the offsets just written (block-local visibility: same CTA)
  for (int i = tid; i < N; i += 1024) {
    const long long a0 = img_offsets[i], a1 = img_offsets[i + 1];
    img_counts[i] = (int)(a1 - a0);
  }
}

// Multi-CTA form of the scan above: CUB's single-pass decoupled look-back scan over the per-tile counts (int -> 64-bit offsets),
// then the image boundaries.  The one-CTA kernel is 24 us for the 25.7 k tiles of cfg-2 and 264 us for the 278 k of cfg-4b.
struct TileCntToLL {
  __host__ __device__ __forceinline__ long long operator()(const int& v) const { return (long long)v; }
};
typedef cub::TransformInputIterator<long long, TileCntToLL, const int*> TileCntIter;
static size_t tile_scan_temp_bytes(long long ntiles) {
  size_t bytes = 0;
  TileCntIter it((const int*)nullptr, TileCntToLL());
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (long long*)nullptr, (int)ntiles, (cudaStream_t)0);
  return (bytes + 255) & ~(size_t)255;
}
__global__ void __launch_bounds__(256) tile_img_offsets_kernel(const long long* __restrict__ tile_off, const int* __restrict__ tile_cnt,
                                                               long long* __restrict__ img_offsets, int* __restrict__ img_counts,
                                                               long long ntiles, int tiles_img, int N) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  const long long a = tile_off[(long long)i * tiles_img];
  const long long b = i + 1 < N ? tile_off[(long long)(i + 1) * tiles_img] : tile_off[ntiles - 1] + tile_cnt[ntiles - 1];
  img_offsets[i] = a;
  img_counts[i] = (int)(b - a);
  if (i == N - 1) img_offsets[N] = b;
}

// move every tile's run from its reserved place in the temporary buffers to its ordered place
__global__ void __launch_bounds__(128) tile_gather_kernel(const long long* __restrict__ tile_base, const int* __restrict__ tile_cnt,
                                                          const long long* __restrict__ tile_off,
                                                          const unsigned long long* __restrict__ keys_tmp,
                                                          const float* __restrict__ cand_tmp, unsigned long long* __restrict__ keys,
                                                          float* __restrict__ cand, long long capacity) {
  const long long t = blockIdx.x;
  const int cnt = tile_cnt[t];
  if (cnt == 0) return;
  const long long src = tile_base[t], dst = tile_off[t];
  for (int i = threadIdx.x; i < cnt; i += 128) {
    if (src + i >= capacity || dst + i >= capacity) break;   // overflow: the caller repeats with larger buffers
    const float2* cs = reinterpret_cast<const float2*>(cand_tmp + (src + i) * 6);
    float2* cd = reinterpret_cast<float2*>(cand + (dst + i) * 6);
    const float2 v0 = cs[0], v1 = cs[1], v2 = cs[2];
    const unsigned long long k = keys_tmp[src + i];
    cd[0] = v0;
    cd[1] = v1;
    cd[2] = v2;
    keys[dst + i] = k;
  }
}

__global__ void img_counts_kernel(const long long* __restrict__ img_offsets, int* __restrict__ img_counts, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) img_counts[i] = (int)(img_offsets[i + 1] - img_offsets[i]);
}

// ---- exact top-K pre-selection (utils/general.py:702-703: only the max_nms = 30000 best candidates of an image enter
// torchvision.ops.nms) ---------------------------------------------------------------------------------------------
// With val-style thresholds an image has several 100 k candidates; sorting all of them only to look at the first
// 30000 made the radix sort the second most expensive kernel of the step.  A 4-CTA cluster per image finds the K-th smallest
// key with a 3-level radix select (11 + 11 + 10 bits of the score key; warp-aggregated shared-memory histograms, so
// the saturated case "330 k candidates with conf == 1.0" costs one atomic per warp, not per candidate) and then
// compacts, IN CANDIDATE ORDER, every key below it plus the first (K - #below) keys equal to it.  The stable sort
// that follows therefore sees exactly the candidates — and the tie order — the full sort would have put first.
constexpr int kSelThreads = 1024;
constexpr int kSelPer = 8;   // candidates per thread per chunk in the compaction pass

__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int& total) {   // all kSelThreads threads call
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();                 // warp_tot may still be read from the previous call
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = warp_tot[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;
    if (lane == 31) warp_tot[32] = winc;
  }
  __syncthreads();
  total = warp_tot[32];
  return warp_tot[warp] + inc - v;
}

// first bin whose inclusive prefix reaches `want` (bins in ascending order); nbins <= 2 * kSelThreads
__device__ __forceinline__ void find_bin(const int* hist, int nbins, int want, int* warp_tot, int* res /*[2]*/) {
  const int b0 = 2 * threadIdx.x, b1 = b0 + 1;
  const int h0 = b0 < nbins ? hist[b0] : 0, h1 = b1 < nbins ? hist[b1] : 0;
  int total;
  const int excl = block_excl_scan(h0 + h1, warp_tot, total);
  if (excl < want && excl + h0 >= want) { res[0] = b0; res[1] = excl; }
  else if (excl + h0 < want && excl + h0 + h1 >= want) { res[0] = b1; res[1] = excl + h0; }
  __syncthreads();
}

// A cluster of kSelCluster CTAs serves one image: each CTA histograms / compacts one contiguous slice of the image's
// keys, the per-CTA histograms are merged through distributed shared memory (every CTA reads its peers' bins and
// finds the same bin, so nothing has to be broadcast), and the slices' "keys below" / "keys equal" counts — which
// fall out of the local histograms — give each CTA its output offset and its share of the tie quota.
constexpr int kSelCluster = 4;

__device__ __forceinline__ int ld_dsmem_s32(const int* local_ptr, unsigned rank) {
  unsigned local = (unsigned)__cvta_generic_to_shared(local_ptr), remote;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kSelThreads) topk_select_kernel(const unsigned long long* __restrict__ keys,
                                                                  const int* __restrict__ img_counts,
                                                                  const long long* __restrict__ img_offsets,
                                                                  unsigned long long* __restrict__ keys_out,
                                                                  unsigned* __restrict__ idx_out, int* __restrict__ counts_out,
                                                                  long long* __restrict__ offsets_out, int N, int K) {
  __shared__ int hist[2048];        // this CTA's slice
  __shared__ int merged[2048];      // sum over the cluster
  __shared__ int warp_tot[33];
  __shared__ int res[2];
  __shared__ int slice_info[2];     // keys of this slice below / equal to the K-th key (read by the peers)
  __shared__ long long coff_s;
  unsigned crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int img = blockIdx.x / kSelCluster;
  const long long off = img_offsets[img];
  const int cnt = img_counts[img];
  if (threadIdx.x == 0) {
    long long c = 0;
    for (int j = 0; j < img; ++j) c += min(img_counts[j], K);
    coff_s = c;
    if (crank == 0) {
      counts_out[img] = min(cnt, K);
      offsets_out[img] = c;
      if (img == N - 1) offsets_out[N] = c + min(cnt, K);
    }
  }
  __syncthreads();
  const long long coff = coff_s;
  // slice of this CTA (multiples of kSelPer so that the compaction's per-thread runs stay aligned)
  const int per = ((cnt + kSelCluster - 1) / kSelCluster + kSelPer - 1) / kSelPer * kSelPer;
  const int s0 = min((int)crank * per, cnt), s1 = min(s0 + per, cnt);
  const int scnt = s1 - s0;
  const unsigned long long* kp = keys + off + s0;
  if (cnt <= K) {   // everything survives: plain copy (uniform across the cluster: no barrier is skipped by a subset)
    for (int i = threadIdx.x; i < scnt; i += kSelThreads) {
      keys_out[coff + s0 + i] = kp[i];
      idx_out[coff + s0 + i] = (unsigned)(off + s0 + i);
    }
    return;
  }
  // ---- 3-level radix select of the K-th smallest 32-bit score key ----
  unsigned prefix = 0;      // bits fixed so far (value of the selected bins)
  int want = K;             // rank still to find inside the current prefix
  int n_less = 0;           // keys (whole image) strictly below the final K-th key
  int my_less = 0;          // same, this slice only
#pragma unroll 1
  for (int level = 0; level < 3; ++level) {
    const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
    const int nbins = level == 2 ? 1024 : 2048;
    const unsigned pmask = level == 0 ? 0u : (level == 1 ? 0xFFE00000u : 0xFFFFFC00u);
    for (int i = threadIdx.x; i < 2048; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    // 8 independent loads in flight per thread (the CTA streams its slice: latency, not bandwidth, is the limit)
    for (int c0 = 0; c0 < scnt; c0 += kSelThreads * kSelPer) {
      unsigned k32[kSelPer];
#pragma unroll
      for (int u = 0; u < kSelPer; ++u) {
        const int i = c0 + u * kSelThreads + threadIdx.x;
        k32[u] = i < scnt ? (unsigned)kp[i] : 0u;
      }
#pragma unroll
      for (int u = 0; u < kSelPer; ++u) {
        const int i = c0 + u * kSelThreads + threadIdx.x;
        unsigned bin = 0xFFFFFFFFu;
        if (i < scnt && (k32[u] & pmask) == prefix) bin = (k32[u] >> shift) & (unsigned)(nbins - 1);
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin != 0xFFFFFFFFu && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], __popc(peers));
      }
    }
    cluster_barrier();                                     // every slice's histogram is complete
    for (int i = threadIdx.x; i < nbins; i += kSelThreads) {
      int t = 0;
#pragma unroll
      for (unsigned r = 0; r < (unsigned)kSelCluster; ++r) t += ld_dsmem_s32(&hist[i], r);
      merged[i] = t;
    }
    __syncthreads();
    find_bin(merged, nbins, want, warp_tot, res);
    // this slice's keys below the selected bin at this level
    {
      int part = 0;
      for (int i = threadIdx.x; i < res[0]; i += kSelThreads) part += hist[i];
      int tot;
      block_excl_scan(part, warp_tot, tot);
      my_less += tot;
    }
    if (level == 2 && threadIdx.x == 0) slice_info[1] = hist[res[0]];   // keys of this slice equal to the K-th key
    prefix |= (unsigned)res[0] << shift;
    n_less += res[1];
    want -= res[1];
    cluster_barrier();                                     // peers are done reading this level's histogram
  }
  const unsigned kstar = prefix;
  const int quota = K - n_less;     // keys equal to kstar that are kept: the first `quota` in candidate order
  if (threadIdx.x == 0) slice_info[0] = my_less;
  cluster_barrier();
  // offsets of this slice: what the slices before it keep
  int run_eq = 0, run_keep = 0;
  for (unsigned r = 0; r < crank; ++r) {
    const int less_r = ld_dsmem_s32(&slice_info[0], r), eq_r = ld_dsmem_s32(&slice_info[1], r);
    const int keep_eq = max(0, min(eq_r, quota - run_eq));
    run_keep += less_r + keep_eq;
    run_eq += eq_r;
  }
  cluster_barrier();                                       // nobody exits (or reuses smem) while a peer still reads it
  // ---- ordered compaction of this slice ----
  const int chunk = kSelThreads * kSelPer;
#pragma unroll 1
  for (int c0 = 0; c0 < scnt; c0 += chunk) {
    const int i0 = c0 + threadIdx.x * kSelPer;
    unsigned long long kk[kSelPer];
    int neq = 0;
#pragma unroll
    for (int j = 0; j < kSelPer; ++j) {
      kk[j] = i0 + j < scnt ? kp[i0 + j] : 0xFFFFFFFFFFFFFFFFull;
      neq += (i0 + j < scnt && (unsigned)kk[j] == kstar) ? 1 : 0;
    }
    int tot_eq;
    int eq_rank = run_eq + block_excl_scan(neq, warp_tot, tot_eq);
    unsigned keep_mask = 0;
    int nkeep = 0;
#pragma unroll
    for (int j = 0; j < kSelPer; ++j) {
      if (i0 + j < scnt) {
        const unsigned k32 = (unsigned)kk[j];
        bool keep = k32 < kstar;
        if (k32 == kstar) {
          keep = eq_rank < quota;
          ++eq_rank;
        }
        if (keep) {
          keep_mask |= 1u << j;
          ++nkeep;
        }
      }
    }
    int tot_keep;
    int pos = run_keep + block_excl_scan(nkeep, warp_tot, tot_keep);
#pragma unroll
    for (int j = 0; j < kSelPer; ++j) {
      if (keep_mask & (1u << j)) {
        keys_out[coff + pos] = kk[j];
        idx_out[coff + pos] = (unsigned)(off + s0 + i0 + j);
        ++pos;
      }
    }
    run_eq += tot_eq;
    run_keep += tot_keep;
  }
}

__global__ void clamp_offsets_kernel(const long long* __restrict__ img_offsets, long long* __restrict__ offsets_out,
                                     int* __restrict__ counts_out, int N, long long capacity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > N) return;
  const long long a = img_offsets[i] < capacity ? img_offsets[i] : capacity;
  offsets_out[i] = a;
  if (i < N) {
    const long long b = img_offsets[i + 1] < capacity ? img_offsets[i + 1] : capacity;
    counts_out[i] = (int)(b - a);
  }
}

__global__ void iota_kernel(unsigned* __restrict__ idx, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    idx[i] = (unsigned)i;
}

// ---- greedy NMS -------------------------------------------------------------------------------------
// torchvision/csrc/ops/cpu/nms_kernel.cpp semantics.
__device__ __forceinline__ bool suppresses(float ix1, float iy1, float ix2, float iy2, float iarea, float jx1, float jy1,
                                           float jx2, float jy2, float jarea, double thr) {
  const float xx1 = ix1 < jx1 ? jx1 : ix1;  // std::max(ix1, x1[j])
  const float xx2 = jx2 < ix2 ? jx2 : ix2;  // std::min(ix2, x2[j])
  const float dw = __fsub_rn(xx2, xx1);
  const float w = 0.f < dw ? dw : 0.f;  // std::max(0, xx2 - xx1)
  const float yy1 = iy1 < jy1 ? jy1 : iy1;
  const float yy2 = jy2 < iy2 ? jy2 : iy2;
  const float dh = __fsub_rn(yy2, yy1);
  const float h = 0.f < dh ? dh : 0.f;
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
  return (double)ovr > thr;
}

constexpr int kNmsThreads = 512;
constexpr int kChunk = 128;

__global__ void __launch_bounds__(kNmsThreads) nms_greedy_kernel(const float* __restrict__ cand,
                                                                 const unsigned* __restrict__ sorted_idx,
                                                                 const int* __restrict__ img_counts,
                                                                 const long long* __restrict__ img_offsets,
                                                                 float* __restrict__ out, int* __restrict__ out_counts,
                                                                 int max_det, int max_nms, float max_wh, int agnostic,
                                                                 double thr) {
  extern __shared__ float sm[];
  float* kept = sm;                                   // [max_det][5]  offset box + area
  float* cb = kept + (size_t)max_det * 5;             // [kChunk][5]   chunk offset boxes + area
  float* craw = cb + kChunk * 5;                      // [kChunk][6]   chunk raw rows
  unsigned* mask = reinterpret_cast<unsigned*>(craw + kChunk * 6);  // [kChunk][4] intra-chunk suppression bits
  int* dead = reinterpret_cast<int*>(mask + kChunk * 4);            // [kChunk]
  __shared__ int K_s;
  __shared__ int newk[kChunk];
  __shared__ int n_new_s;

  const int img = blockIdx.x;
  const long long off = img_offsets[img];
  int n = img_counts[img];
  if (n > max_nms) n = max_nms;
  if (threadIdx.x == 0) K_s = 0;
  __syncthreads();

  for (int c0 = 0; c0 < n; c0 += kChunk) {
    const int cn = min(kChunk, n - c0);
    const int K = K_s;
    if (K >= max_det) break;
    // 1. load the chunk
    if (threadIdx.x < kChunk) {
      const int t = threadIdx.x;
      dead[t] = t < cn ? 0 : 1;
      mask[t * 4 + 0] = mask[t * 4 + 1] = mask[t * 4 + 2] = mask[t * 4 + 3] = 0u;
      if (t < cn) {
        const float* r = cand + ((long long)sorted_idx[off + c0 + t]) * 6;
        float b[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          b[j] = r[j];
          craw[t * 6 + j] = b[j];
        }
        // boxes + cls * (0 if agnostic else max_wh), utils/general.py:706-707
        const float c = __fmul_rn(b[5], agnostic ? 0.f : max_wh);
        const float x1 = __fadd_rn(b[0], c), y1 = __fadd_rn(b[1], c), x2 = __fadd_rn(b[2], c), y2 = __fadd_rn(b[3], c);
        cb[t * 5 + 0] = x1; cb[t * 5 + 1] = y1; cb[t * 5 + 2] = x2; cb[t * 5 + 3] = y2;
        cb[t * 5 + 4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
      }
    }
    __syncthreads();
    // 2. chunk vs kept list: 4 threads per candidate, each strides the kept list by 4
    {
      const int t = threadIdx.x & (kChunk - 1), sub = threadIdx.x / kChunk;  // sub in 0..3
      if (t < cn) {
        const float x1 = cb[t * 5], y1 = cb[t * 5 + 1], x2 = cb[t * 5 + 2], y2 = cb[t * 5 + 3], ar = cb[t * 5 + 4];
        bool d = false;
        for (int k = sub; k < K && !d; k += kNmsThreads / kChunk) {
          const float* kb = kept + k * 5;
          d = suppresses(kb[0], kb[1], kb[2], kb[3], kb[4], x1, y1, x2, y2, ar, thr);
        }
        if (d) dead[t] = 1;
      }
    }
    // 3a. intra-chunk pair bits: thread (i, word) computes which j in [32*word, 32*word+32), j > i, box i suppresses
    {
      const int i = threadIdx.x >> 2, wd = threadIdx.x & 3;
      if (i < cn) {
        const float x1 = cb[i * 5], y1 = cb[i * 5 + 1], x2 = cb[i * 5 + 2], y2 = cb[i * 5 + 3], ar = cb[i * 5 + 4];
        unsigned bits = 0u;
        const int j0 = wd * 32;
        for (int jj = 0; jj < 32; ++jj) {
          const int j = j0 + jj;
          if (j > i && j < cn) {
            const float* jb = cb + j * 5;
            if (suppresses(x1, y1, x2, y2, ar, jb[0], jb[1], jb[2], jb[3], jb[4], thr)) bits |= 1u << jj;
          }
        }
        mask[i * 4 + wd] = bits;
      }
    }
    __syncthreads();
    // 3b. serial resolution by one thread (<= 128 steps of 4-word bit ops)
    if (threadIdx.x == 0) {
      unsigned rem[4] = {0u, 0u, 0u, 0u};
      int k = K, nn = 0;
      for (int i = 0; i < cn && k < max_det; ++i) {
        if (dead[i]) continue;
        if ((rem[i >> 5] >> (i & 31)) & 1u) continue;
        newk[nn++] = i;
        ++k;
        rem[0] |= mask[i * 4 + 0]; rem[1] |= mask[i * 4 + 1]; rem[2] |= mask[i * 4 + 2]; rem[3] |= mask[i * 4 + 3];
      }
      n_new_s = nn;
    }
    __syncthreads();
    // 4. append the new keeps (kept list + output rows), in order
    const int nn = n_new_s;
    if (threadIdx.x < nn) {
      const int i = newk[threadIdx.x];
      const int k = K + threadIdx.x;
#pragma unroll
      for (int j = 0; j < 5; ++j) kept[k * 5 + j] = cb[i * 5 + j];
      float* o = out + ((long long)img * max_det + k) * 6;
#pragma unroll
      for (int j = 0; j < 6; ++j) o[j] = craw[i * 6 + j];
    }
    __syncthreads();
    if (threadIdx.x == 0) K_s = K + nn;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_counts[img] = K_s;
}

// ---- dense Detect decode (API-compat materialisation of `pred`) ------------------------------------------
__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ logits, float* __restrict__ pred, int N,
                                                     int ny, int nx, int na, int no, int ld, int pitch, int row0, int rows_total,
                                                     float stride, float aw0,
                                                     float ah0, float aw1, float ah1, float aw2, float ah2, float aw3,
                                                     float ah3, float aw4, float ah4) {
  const float aw[5] = {aw0, aw1, aw2, aw3, aw4}, ah[5] = {ah0, ah1, ah2, ah3, ah4};
  const long long items = (long long)N * na * ny * nx * no;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    int o = (int)(i % no);
    long long t = i / no;
    int gx = (int)(t % nx);
    t /= nx;
    int gy = (int)(t % ny);
    t /= ny;
    int a = (int)(t % na);
    int n = (int)(t / na);
    const float s = sigmoid_dec(logits[(((long long)n * ny + gy) * nx + gx) * ld + a * pitch + o]);
    float v = s;
    if (o == 0) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.f), 0.5f), (float)gx), stride);
    else if (o == 1) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.f), 0.5f), (float)gy), stride);
    else if (o == 2) { float q = __fmul_rn(s, 2.f); v = __fmul_rn(__fmul_rn(q, q), aw[a]); }
    else if (o == 3) { float q = __fmul_rn(s, 2.f); v = __fmul_rn(__fmul_rn(q, q), ah[a]); }
    pred[((long long)n * rows_total + row0 + ((long long)a * ny + gy) * nx + gx) * no + o] = v;
  }
}

}  // namespace dmay

using namespace dmay;

extern "C" {

int dmay_detect_decode(const dmay_decode_params* p, dmay_stream_t stream) {
  if (!p || !p->logits || !p->pred) return DMAY_EINVAL;
  if (p->N <= 0 || p->ny <= 0 || p->nx <= 0 || p->na <= 0 || p->no <= 5) return DMAY_EINVAL;
  const int pitch = p->row_pitch > 0 ? p->row_pitch : p->no;
  if (p->na > 5 || pitch < p->no || p->ld < p->na * pitch) return DMAY_EUNSUPPORTED;
  long long items = (long long)p->N * p->na * p->ny * p->nx * p->no;
  decode_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float*)p->logits, (float*)p->pred, p->N, p->ny, p->nx, p->na, p->no, p->ld, pitch, p->row0, p->rows_total,
      p->stride, p->aw0, p->ah0, p->aw1, p->ah1, p->aw2, p->ah2, p->aw3, p->ah3, p->aw4, p->ah4);
  return finish_launch();
}

int dmay_nms_filter(const dmay_filter_params* p, dmay_stream_t stream) {
  if (!p || p->N <= 0 || p->R <= 0 || p->nc <= 0) return DMAY_EINVAL;
  if (!p->blk_counts || !p->blk_offsets || !p->img_counts || !p->img_offsets) return DMAY_EINVAL;
  if (p->levels < 0 || p->levels > 5) return DMAY_EUNSUPPORTED;
  if (p->levels == 0 && !p->pred) return DMAY_EINVAL;
  if (p->levels > 0 && (!p->lv_meta || !p->lv_logits0)) return DMAY_EINVAL;
  const int rpb = p->rows_per_block > 0 ? p->rows_per_block : 128;
  if (rpb > 1024) return DMAY_EUNSUPPORTED;
  const int nblk = (p->R + rpb - 1) / rpb;
  const long long grid = (long long)p->N * nblk;
  if (grid > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  RowSrc src;
  src.pred = (const float*)p->pred;
  src.logits[0] = (const float*)p->lv_logits0; src.logits[1] = (const float*)p->lv_logits1;
  src.logits[2] = (const float*)p->lv_logits2; src.logits[3] = (const float*)p->lv_logits3;
  src.logits[4] = (const float*)p->lv_logits4;
  src.meta = (const LevelMeta*)p->lv_meta;
  src.levels = p->levels;
  src.R = p->R;
  src.nc = p->nc;
  src.pitch = p->row_pitch;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned char* cm = (const unsigned char*)p->class_mask;
  if (p->phase == 0) {
    filter_count_kernel<<<(int)grid, kFilterThreads, 0, s>>>(src, cm, (int*)p->blk_counts, rpb, nblk, p->multi_label,
                                                             p->conf_thres);
    filter_scan_kernel<<<p->N, 1024, 0, s>>>((const int*)p->blk_counts, (int*)p->blk_offsets, (int*)p->img_counts, nblk);
    img_scan_kernel<<<1, 32, 0, s>>>((const int*)p->img_counts, (long long*)p->img_offsets, p->N);
    return finish_launch(3);
  }
  if (!p->keys || !p->cand || p->capacity <= 0) return DMAY_EINVAL;
  filter_write_kernel<<<(int)grid, kFilterThreads, (rpb + 1) * sizeof(int), s>>>(
      src, cm, (const int*)p->blk_offsets, (const long long*)p->img_offsets, (unsigned long long*)p->keys,
      (float*)p->cand, rpb, nblk, p->multi_label, p->conf_thres, p->capacity);
  return finish_launch();
}

static long long fused_tiles_per_image(const LevelMeta* hm, int levels, FuseArgs* fa, int P = kFuseP) {
  long long tiles = 0;
  for (int l = 0; l < levels; ++l) {
    const int npix = hm[l].ny * hm[l].nx;
    const int tpa = (npix + P - 1) / P;
    if (fa) {
      fa->tpa[l] = tpa;
      fa->tile0[l] = (int)tiles;
    }
    tiles += (long long)hm[l].na * tpa;
  }
  if (fa) fa->tile0[levels] = (int)tiles;
  return tiles;
}

long long dmay_nms_filter_fused_ws(const void* lv_meta_host, int levels, int N) {
  if (!lv_meta_host || levels <= 0 || levels > 5 || N <= 0) return DMAY_EINVAL;
  const long long tiles = (long long)N * fused_tiles_per_image((const LevelMeta*)lv_meta_host, levels, nullptr);
  // status words, or tile base / count / offset; per-image reservation counters; temporary storage of the multi-CTA tile scan
  return 16 + 24 * tiles + 8LL * N + 8 + 256 + (long long)tile_scan_temp_bytes(tiles);
}

int dmay_nms_filter_fused(const dmay_filter_fused_params* p, dmay_stream_t stream) {
  if (!p || p->N <= 0 || p->nc <= 0 || p->levels <= 0 || p->levels > 5) return DMAY_EINVAL;
  if (!p->lv_meta_host || !p->lv_logits0 || !p->ws || !p->img_counts || !p->img_offsets) return DMAY_EINVAL;
  if (!p->keys || !p->cand || p->capacity <= 0) return DMAY_EINVAL;
  FuseArgs fa;
  memset(&fa, 0, sizeof(fa));
  const void* lg[5] = {p->lv_logits0, p->lv_logits1, p->lv_logits2, p->lv_logits3, p->lv_logits4};
  const LevelMeta* hm = (const LevelMeta*)p->lv_meta_host;
  const int no = 5 + p->nc;
  const int pitch = p->row_pitch > 0 ? p->row_pitch : no;
  if (pitch < no) return DMAY_EINVAL;
  if (p->dense && (p->levels != 1 || hm[0].na != 1 || pitch != no)) return DMAY_EINVAL;
  const bool padded = !p->dense && pitch != no;
  long long rows = 0;
  for (int l = 0; l < p->levels; ++l) {
    const LevelMeta& m = hm[l];
    if (!lg[l] || (reinterpret_cast<uintptr_t>(lg[l]) & 3u)) return DMAY_EINVAL;
    if (m.na <= 0 || m.na > 5 || m.ny <= 0 || m.nx <= 0) return DMAY_EINVAL;
    if (m.ld < m.na * pitch) return DMAY_EUNSUPPORTED;
    if (padded && ((pitch & 3) || (m.ld & 3) || (reinterpret_cast<uintptr_t>(lg[l]) & 15u))) return DMAY_EUNSUPPORTED;
    if (m.row0 != rows) return DMAY_EINVAL;
    fa.logits[l] = (const float*)lg[l];
    fa.meta[l] = m;
    rows += (long long)m.na * m.ny * m.nx;
  }
  // thread-per-row kernel for nc <= 96 (DMAY_FILTER_ROWS=0 keeps the warp-per-row one, for A/B runs and the tests)
  static const bool no_rows = [] { const char* e = getenv("DMAY_FILTER_ROWS"); return e && e[0] == '0'; }();
  const int srow = padded ? pitch + 4 : no;
  // rows per tile: 64 for the padded Detect logits (352-byte rows: 23.5 KB tiles, nine CTAs per SM), 128 otherwise (the dense
  // rows of cfg-5 are 60 bytes: same-box, 64-row tiles 1.46 ms per batch of 256 against 1.35 ms)
  const int rows_tile = padded ? kRowsThreads : kRowsWide;
  const bool rows_kernel = (!no_rows || padded || p->dense) && p->nc <= kRowsMaxNc && (size_t)rows_tile * srow * sizeof(float) <= 100 * 1024;
  if ((padded || p->dense) && !rows_kernel) return DMAY_EUNSUPPORTED;   // those layouts exist for the thread-per-row kernel only
  const int P = rows_kernel ? rows_tile : kFuseP;
  const long long tiles = (long long)p->N * fused_tiles_per_image(hm, p->levels, &fa, P);
  if (tiles > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  if (p->ws_bytes < 16 + 8 * tiles) return DMAY_ETOOBIG;
  const size_t smem = (size_t)P * srow * sizeof(float);
  if (smem > 200 * 1024) return DMAY_EUNSUPPORTED;
  fa.levels = p->levels;
  fa.nc = p->nc;
  fa.multi_label = p->multi_label;
  fa.N = p->N;
  fa.thr = p->conf_thres;
  fa.capacity = p->capacity;
  fa.pitch = pitch;
  fa.bin_thr = (p->dense || padded) ? (const int*)p->bin_thr : nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* ticket = (unsigned*)p->ws;
  unsigned long long* status = (unsigned long long*)((char*)p->ws + 16);
  static const bool no_reserve = [] { const char* e = getenv("DMAY_FILTER_RESERVE"); return e && e[0] == '0'; }();
  const bool reserve = rows_kernel && p->keys_tmp != nullptr && p->cand_tmp != nullptr && !no_reserve;
  const int kind = p->dense ? 2 : (padded ? 1 : 0);
  if (rows_kernel) {
    // two threads per row for the padded-logits layout (DMAY_FILTER_TPR=1: one thread per row, A/B)
    static const bool tpr2 = [] { const char* e = getenv("DMAY_FILTER_TPR"); return e && e[0] == '2'; }();
    const int tpr = (kind == 1 && tpr2) ? 2 : 1;   // measured: 1 thread per row 176 us, 2 threads per row 281 us (r5y)
    void (*kern)(const FuseArgs, const unsigned char*, unsigned*, unsigned long long*, long long*, unsigned long long*, float*) =
        reserve ? (kind == 0 ? filter_fused_rows_kernel<true, 0, 1, kRowsWide> : kind == 2 ? filter_fused_rows_kernel<true, 2, 1, kRowsWide>
                   : tpr == 2 ? filter_fused_rows_kernel<true, 1, 2> : filter_fused_rows_kernel<true, 1>)
                : (kind == 0 ? filter_fused_rows_kernel<false, 0, 1, kRowsWide> : kind == 2 ? filter_fused_rows_kernel<false, 2, 1, kRowsWide>
                   : tpr == 2 ? filter_fused_rows_kernel<false, 1, 2> : filter_fused_rows_kernel<false, 1>);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    if (reserve) {
      // reserve + scan + gather: no ordering dependency between the tiles of the big kernel
      if (p->ws_bytes < 16 + 20 * tiles) return DMAY_ETOOBIG;
      long long* tile_base = (long long*)status;
      int* tile_cnt = (int*)(tile_base + tiles);
      long long* tile_off = (long long*)((char*)p->ws + 16 + ((12 * tiles + 7) & ~7LL));
      if (p->ws_bytes < 16 + ((12 * tiles + 7) & ~7LL) + 8 * tiles) return DMAY_ETOOBIG;
      const long long ctr_off = (16 + ((12 * tiles + 7) & ~7LL) + 8 * tiles + 7) & ~7LL;
      if (p->per_image_regions && p->ws_bytes >= ctr_off + 8LL * p->N && p->capacity / p->N > 0) {
        fa.img_ctr = (unsigned long long*)((char*)p->ws + ctr_off);
        fa.region = p->capacity / p->N;
      }
      static const int persist = [] { const char* e = getenv("DMAY_FILTER_PERSIST"); return e ? atoi(e) : 0; }();
      if (persist && kind == 1 && 2 * smem <= 110 * 1024) {
        // persistent CTAs, two tile buffers each (see filter_rows_persistent_kernel)
        void (*pk)(const FuseArgs, const unsigned char*, unsigned*, unsigned long long*, long long*, unsigned long long*, float*) =
            tpr == 2 ? filter_rows_persistent_kernel<1, 2> : filter_rows_persistent_kernel<1, 1>;
        const size_t smem2 = 2 * smem;
        cudaError_t e = cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return (int)e;
        const long long slots = (long long)sm_count() * (227 * 1024 / (long long)(smem2 + 1024));
        pk<<<(int)(tiles < slots ? tiles : slots), kRowsThreads * tpr, smem2, s>>>(fa, (const unsigned char*)p->class_mask, ticket, status,
                                                                                  (long long*)p->img_offsets,
                                                                                  (unsigned long long*)p->keys_tmp, (float*)p->cand_tmp);
      } else {
      kern<<<(int)tiles, P * tpr, smem, s>>>(fa, (const unsigned char*)p->class_mask, ticket, status, (long long*)p->img_offsets,
                                                        (unsigned long long*)p->keys_tmp, (float*)p->cand_tmp);
      }
      // scan of the per-tile counts: CUB's multi-CTA scan when the workspace holds its temporary storage, else one CTA
      const long long scan_off = (ctr_off + 8LL * p->N + 8 + 255) & ~255LL;
      size_t scan_bytes = tile_scan_temp_bytes(tiles);
      static const bool one_cta_scan = [] { const char* e = getenv("DMAY_TILE_SCAN_CUB"); return e && e[0] == '0'; }();
      if (!one_cta_scan && tiles >= 4096 && p->ws_bytes >= scan_off + (long long)scan_bytes &&
          ((reinterpret_cast<uintptr_t>(p->ws) + scan_off) & 255) == 0) {
        TileCntIter it((const int*)tile_cnt, TileCntToLL());
        cudaError_t e = cub::DeviceScan::ExclusiveSum((char*)p->ws + scan_off, scan_bytes, it, tile_off, (int)tiles, s);
        if (e != cudaSuccess) return (int)e;
        tile_img_offsets_kernel<<<(p->N + 255) / 256, 256, 0, s>>>(tile_off, tile_cnt, (long long*)p->img_offsets, (int*)p->img_counts,
                                                                   tiles, (int)(tiles / p->N), p->N);
      } else {
        tile_scan_kernel<<<1, 1024, 0, s>>>(tile_cnt, tile_off, (long long*)p->img_offsets, (int*)p->img_counts, tiles,
                                            (int)(tiles / p->N), p->N);
      }
      tile_gather_kernel<<<(int)tiles, 128, 0, s>>>(tile_base, tile_cnt, tile_off, (const unsigned long long*)p->keys_tmp,
                                                    (const float*)p->cand_tmp, (unsigned long long*)p->keys, (float*)p->cand,
                                                    p->capacity);
      return finish_launch(3);
    }
    kern<<<(int)tiles, P * tpr, smem, s>>>(fa, (const unsigned char*)p->class_mask, ticket, status, (long long*)p->img_offsets,
                                                      (unsigned long long*)p->keys, (float*)p->cand);
  } else {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(filter_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    filter_fused_kernel<<<(int)tiles, kFuseThreads, smem, s>>>(fa, (const unsigned char*)p->class_mask, ticket, status,
                                                             (long long*)p->img_offsets, (unsigned long long*)p->keys,
                                                             (float*)p->cand);
  }
  img_counts_kernel<<<(p->N + 255) / 256, 256, 0, s>>>((const long long*)p->img_offsets, (int*)p->img_counts, p->N);
  return finish_launch(2);
}

// Pre-selection for candidate-rich Detect LOGITS (padded anchor rows, multi-label): histogram pass of the rows kernel over the
// same tiles, then the per-image bin threshold.  The following dmay_nms_filter_fused (same parameters + bin_thr) writes only
// candidates whose key bin is <= bin_thr[img] -- a superset of the top K, ties included, in candidate order.
int dmay_nms_fused_prethreshold(const dmay_filter_fused_params* p, dmay_stream_t stream) {
  if (!p || p->N <= 0 || p->nc <= 0 || p->levels <= 0 || p->levels > 5) return DMAY_EINVAL;
  if (!p->lv_meta_host || !p->lv_logits0 || !p->hist || !p->bin_thr || p->prethr_k <= 0) return DMAY_EINVAL;
  if (p->dense || !p->multi_label || p->nc > kRowsMaxNc) return DMAY_EUNSUPPORTED;
  FuseArgs fa;
  memset(&fa, 0, sizeof(fa));
  const void* lg[5] = {p->lv_logits0, p->lv_logits1, p->lv_logits2, p->lv_logits3, p->lv_logits4};
  const LevelMeta* hm = (const LevelMeta*)p->lv_meta_host;
  const int no = 5 + p->nc, pitch = p->row_pitch;
  if (pitch <= no || (pitch & 3)) return DMAY_EUNSUPPORTED;   // the padded-row layout only
  long long rows = 0;
  for (int l = 0; l < p->levels; ++l) {
    const LevelMeta& m = hm[l];
    if (!lg[l] || (reinterpret_cast<uintptr_t>(lg[l]) & 15u) || (m.ld & 3)) return DMAY_EUNSUPPORTED;
    if (m.na <= 0 || m.na > 5 || m.ny <= 0 || m.nx <= 0 || m.ld < m.na * pitch || m.row0 != rows) return DMAY_EINVAL;
    fa.logits[l] = (const float*)lg[l];
    fa.meta[l] = m;
    rows += (long long)m.na * m.ny * m.nx;
  }
  const int srow = pitch + 4;
  const long long tiles = (long long)p->N * fused_tiles_per_image(hm, p->levels, &fa, kRowsThreads);
  if (tiles > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  const size_t smem = (size_t)kRowsThreads * srow * sizeof(float) + kHistBins * sizeof(int);
  if (smem > 200 * 1024) return DMAY_EUNSUPPORTED;
  fa.levels = p->levels;
  fa.nc = p->nc;
  fa.multi_label = 1;
  fa.N = p->N;
  fa.thr = p->conf_thres;
  fa.pitch = pitch;
  fa.hist = (int*)p->hist;
  cudaStream_t s = (cudaStream_t)stream;
  auto kern = filter_fused_rows_kernel<true, 1, 1, kRowsThreads, true>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<(int)tiles, kRowsThreads, smem, s>>>(fa, (const unsigned char*)p->class_mask, nullptr, nullptr, nullptr, nullptr, nullptr);
  hist_threshold_kernel<<<p->N, 1024, 0, s>>>((const int*)p->hist, (int*)p->bin_thr, p->prethr_k);
  return finish_launch(2);
}

int dmay_nms_dense_prethreshold(const dmay_prethr_params* p, dmay_stream_t stream) {
  if (!p || !p->pred || !p->hist || !p->bin_thr || p->N <= 0 || p->R <= 0 || p->nc <= 0 || p->K <= 0) return DMAY_EINVAL;
  if (p->N > 65535) return DMAY_EUNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int rows_per_cta = 2048;
  dim3 grid((p->R + rows_per_cta - 1) / rows_per_cta, p->N);
  dense_hist_kernel<<<grid, 256, 0, s>>>((const float*)p->pred, (const unsigned char*)p->class_mask, (int*)p->hist, p->R, p->nc,
                                         p->conf_thres, rows_per_cta);
  hist_threshold_kernel<<<p->N, 1024, 0, s>>>((const int*)p->hist, (int*)p->bin_thr, p->K);
  return finish_launch(2);
}

static size_t cub_ws_bytes(long long n, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const unsigned*)nullptr, (unsigned*)nullptr, n, 0, end_bit, (cudaStream_t)0);
  return (bytes + 255) & ~(size_t)255;
}

long long dmay_nms_sort_ws(long long n, int img_bits) {
  if (n <= 0) return 256;
  return (long long)cub_ws_bytes(n, 32 + img_bits) + (long long)((n * 4 + 255) & ~255LL);
}

int dmay_nms_sort(const dmay_sort_params* p, dmay_stream_t stream) {
  if (!p || !p->keys_in || !p->keys_out || !p->idx_out || !p->ws || p->n <= 0) return DMAY_EINVAL;
  if (p->img_bits < 0 || p->img_bits > 31) return DMAY_EINVAL;
  const int end_bit = 32 + p->img_bits;
  size_t cub_bytes = cub_ws_bytes(p->n, end_bit);
  if ((long long)cub_bytes + p->n * 4 > p->ws_bytes) return DMAY_ETOOBIG;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned* vals = (const unsigned*)p->vals_in;
  if (vals == nullptr) {   // payload = position in keys_in
    unsigned* idx_in = reinterpret_cast<unsigned*>((char*)p->ws + cub_bytes);
    iota_kernel<<<grid_for(p->n, 256), 256, 0, s>>>(idx_in, p->n);
    vals = idx_in;
  }
  cudaError_t e = cub::DeviceRadixSort::SortPairs(p->ws, cub_bytes, (const unsigned long long*)p->keys_in,
                                                  (unsigned long long*)p->keys_out, vals, (unsigned*)p->idx_out, p->n, 0,
                                                  end_bit, s);
  if (e != cudaSuccess) return (int)e;
  return finish_launch(2);
}

int dmay_nms_topk_select(const dmay_topk_params* p, dmay_stream_t stream) {
  if (!p || !p->keys || !p->img_counts || !p->img_offsets || !p->keys_out || !p->idx_out || !p->counts_out || !p->offsets_out)
    return DMAY_EINVAL;
  if (p->N <= 0 || p->K <= 0) return DMAY_EINVAL;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p->N * kSelCluster);
  cfg.blockDim = dim3(kSelThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kSelCluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, topk_select_kernel, (const unsigned long long*)p->keys, (const int*)p->img_counts,
                                     (const long long*)p->img_offsets, (unsigned long long*)p->keys_out, (unsigned*)p->idx_out,
                                     (int*)p->counts_out, (long long*)p->offsets_out, p->N, p->K);
  if (e != cudaSuccess) return (int)e;
  return finish_launch();
}

int dmay_nms_clamp_offsets(const dmay_clamp_params* p, dmay_stream_t stream) {
  if (!p || !p->img_offsets || !p->offsets_out || !p->counts_out || p->N <= 0 || p->capacity <= 0) return DMAY_EINVAL;
  if (p->img_offsets == p->offsets_out) return DMAY_EINVAL;   // out of place: threads read their neighbour's input
  clamp_offsets_kernel<<<(p->N + 256) / 256, 256, 0, (cudaStream_t)stream>>>((const long long*)p->img_offsets, (long long*)p->offsets_out,
                                                                          (int*)p->counts_out, p->N, p->capacity);
  return finish_launch();
}

int dmay_nms_greedy(const dmay_nms_params* p, dmay_stream_t stream) {
  if (!p || !p->cand || !p->sorted_idx || !p->img_counts || !p->img_offsets || !p->out || !p->out_counts) return DMAY_EINVAL;
  if (p->N <= 0 || p->max_det <= 0 || p->max_nms <= 0) return DMAY_EINVAL;
  const size_t smem = ((size_t)p->max_det * 5 + kChunk * 5 + kChunk * 6 + kChunk * 4 + kChunk) * 4;
  if (smem > 220 * 1024) return DMAY_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  nms_greedy_kernel<<<p->N, kNmsThreads, smem, (cudaStream_t)stream>>>(
      (const float*)p->cand, (const unsigned*)p->sorted_idx, (const int*)p->img_counts,
      (const long long*)p->img_offsets, (float*)p->out, (int*)p->out_counts, p->max_det, p->max_nms, p->max_wh,
      p->agnostic, p->iou_thres);
  return finish_launch();
}

}  // extern "C"
