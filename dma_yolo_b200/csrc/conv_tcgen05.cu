// a1/a2: Conv + folded BN + activation (+ residual / SCConv gate) as a persistent, warp-specialised
// tcgen05 implicit GEMM for sm_100a.
//
//   D[m, co] = sum_{tap, ci} A[m, (tap, ci)] * W[co, (tap, ci)]         m = (n*Ho + p)*Wo + q
//
//   A operand : activations, NHWC bf16.  3x3 / strided convs use TMA *im2col* tensor maps
//               (cp.async.bulk.tensor.4d...im2col): one instruction loads the 128 output pixels of a
//               tile for one filter tap (r,s) and one 16/32/64-channel chunk, zero-filling the padding
//               halo, straight into the 128B/64B/32B-swizzled K-major layout tcgen05 reads.  1x1 stride-1
//               convs use a plain 2-D tiled map over [M, Cin].
//   B operand : packed weights [Cout_pad][kh*kw*Cin] bf16, 2-D tiled map, same swizzle.
//   D         : fp32 accumulators in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of
//               tile i overlaps the MMAs of tile i+1.
//   epilogue  : tcgen05.ld 32x32b (thread == output pixel) -> scale*acc + bias (folded BatchNorm,
//               utils/torch_utils.py:198-218) -> SiLU / Hardswish -> (+ residual, Bottleneck
//               models/common.py:136-137) or (* sigmoid(x + up(k2)), SCConv models/common.py:1310-1314)
//               -> bf16 (or fp32 for the Detect head) -> 16-byte stores into a channel slice of the
//               destination slab (ldy), which is how torch.cat in C3/SPPFCSPC disappears.
//
// Warp roles (256 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lane quadrant = warp % 4).
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> epilogue),
// static persistent tile scheduler (tile = blockIdx.x + i*gridDim.x; n-tile fastest so CTAs running
// at the same time share the activation tile in L2 and the weights stay L2-resident).
#include "common.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>

namespace dmay {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {   // see mbar_arrive_pair_relaxed
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost arrival (bad descriptor, wrong byte count) traps after ~2 s instead of
// hanging the GPU.  The slow path is out of line to keep the role loops small.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// The slow path suspends in hardware (try_wait with a time hint: the warp wakes as soon as the phase completes) instead
// of spinning: ncu showed half of the conv kernels' executed instructions in this loop, competing with the epilogue
// warps for issue slots.  The watchdog clock is read once per 256 wake-ups.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (unsigned it = 1;; ++it) {
    if (mbar_try_wait_hint(bar, parity, 20000u)) return;
    if ((it & 255u) == 0u && clock64() - t0 > 4000000000LL) {
      printf("dmayolo conv: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const void* tmap, uint32_t bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// B tile slice multicast to every CTA of the cluster (same smem offset, same mbarrier offset in each)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmap_prefetch(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask` (stage release across the cluster)
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// ---- CTA-pair (cta_group::2) variants: one 256-row MMA across the two SMs of a TPC --------------------------------
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> the pair's even CTA
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// TMA loads issued by either CTA of the pair; the transaction bytes are credited to the EVEN CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c, int w, int h, int n,
                                                        uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(tmap), "r"(bar & kPeerMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// epilogue hand-back: arrive on the even CTA's barrier (local for the even CTA, remote for its peer)
__device__ __forceinline__ void mbar_arrive_pair(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
// TMEM hand-back of the epilogue warps in CTA-pair mode.  Nothing in memory is published by this arrive: the accumulator
// has been copied to registers (tcgen05.wait::ld) and tcgen05.fence::before_thread_sync orders those loads before it, so
// the arrive itself can be relaxed.  The release form costs a MEMBAR.ALL + ERRBAR per item that also waits for the warp's
// outstanding shared / global traffic: 26 % of the epilogue warps' samples on the 128-channel pair layers (ncu r5i).
__device__ __forceinline__ void mbar_arrive_pair_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMA tile stores (shared -> global), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
#ifndef DMAY_EPI_F2
#define DMAY_EPI_F2 1
#endif
constexpr int BLOCK_M = 128;
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 16;                      // up to four warps per TMEM lane quadrant (8 or 16 per launch)
constexpr int kConvThreads = 128 + kEpiWarps * 32; // warps 0-3: TMA / MMA / TMEM alloc / spare

// epilogue flavours (compile-time: keeps the hot loop branch-free and small enough for the I-cache)
enum EpiMode : int {
  EPI_SILU = 0,        // bf16( silu(s*acc + b) )
  EPI_SILU_RES = 1,    // bf16( silu(s*acc + b) + residual )
  EPI_LINEAR = 2,      // bf16( s*acc + b )
  EPI_GATE = 3,        // bf16( (s*acc + b) * sigmoid(gate_x + up(gate_k)) )
  EPI_LINEAR_F32 = 4,  // fp32( s*acc + b )                       (Detect head logits)
  EPI_GENERIC = 5,     // runtime activation / optional residual / either dtype (rare layers)
  EPI_LINEAR_RES = 6,  // bf16( s*acc + b + residual )              (Swin proj / fc2: Linear + shortcut)
  EPI_GELU = 7,        // bf16( gelu(s*acc + b) )                   (Swin mlp.fc1, exact erf GELU)
  EPI_LINEAR_MUL = 8,  // bf16( (s*acc + b) * operand )             (GnConv gating: pws_i(x) * dw_{i+1})
  EPI_SILU_PRE = 9,    // bf16( silu(s*(acc + up(pre)) + b) )        (1x1 layer over [Up(x0) | x1]: W0.x0 comes in at low
                       //                                             resolution as an fp32 partial sum, see ops.VCat)
};

struct __align__(64) ConvArgs {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmY;   // bf16 output tile store: box = 32 channels x 32 rows (halo mode: 32 ch x 8 x 4 pixels)
  CUtensorMap tmR;   // residual / gate_x tile load, same geometry
  CUtensorMap tmA2;  // 1x1 layers over a VIRTUAL channel concat: channel chunks [seg1, seg2) come from a second tensor,
  CUtensorMap tmA3;  // chunks >= seg2 from a third one (same pixels, own pointer / row pitch); seg1 == 0: one source
  int seg1, seg2;
  int opnd_stage;    // 1: the epilogue stages a residual / gate_x tile per item
  int flags;         // tuning / A-B switches of the parameter struct (see include/dmayolo.h)
  int sb_floats;     // staged scale / bias entries (columns covered by all n-tiles, <= kMaxCout)
  int epi_warps;     // 8 or 16 (blockDim = 128 + 32 * epi_warps): 16 where the epilogue, not the MMA, paces the tile
  int epi_groups;    // 2: narrow tiles (block_n <= 64) — the 16 warps form two groups that take alternate tiles (one TMEM
                     //    accumulator buffer each), because 8 warps already cover the 64 columns
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* gate_x;
  const __nv_bfloat16* gate_k;
  __nv_bfloat16* pool_out;   // T9 images: 4x4 average of the OUTPUT, [N, Ho/4, Wo/4, ldpool] (the AvgPool2d(4) of a following SCConv)
  int ldpool;
  const float* pre;    // EPI_SILU_PRE: fp32 [N, gHk, gWk, ldgk] partial pre-activation sums, nearest-upsampled to Ho x Wo
  void* y;
  int M, n_store, Cout_pad;
  int num_m_tiles, num_n_tiles;
  int taps, kw, c_chunks, CK, Cin;
  int conv_stride, pad, Ho, Wo;
  int im2col;
  int block_n, acc_stride, tmem_cols, stages;
  int nacc;          // accumulator buffers in TMEM (2..8): the MMA of tile j+nacc waits for the epilogue's TMEM loads of tile j
  int subs, total_subs;      // sub-tiles (one tap x one CK-channel chunk) per pipeline stage / per tile
  int cs;                    // cluster size: the B tile is loaded in `cs` row slices, each multicast to all CTAs
  int pair;                  // 1: cta_group::2 — the 2 CTAs of a cluster form one 256 x block_n tile (each loads its own 128
                             //    rows of A and HALF of the weight tile; the even CTA issues the MMAs for both)
  // halo mode (3x3 stride-1 pad-1): an output tile is a 16x8 pixel patch; its 18 x pitch input patch is loaded
  // ONCE per channel chunk and the 9 taps read shifted windows of it (descriptor start = +(r*pitch+s) rows,
  // SBO = one patch row) instead of 9 im2col loads of the same pixels from L2.
  int halo, halo_pitch, tiles_x, tiles_y, n_abuf;
  uint32_t a_halo_bytes, a_halo_tx, halo_sbo_enc;
  // weights-resident variant of halo mode: the whole [block_n x 9*Cin] weight set (<= ~96 KB) is loaded into
  // smem once per CTA, so the steady state streams input patches only (deep patch ring, no B pipeline)
  int b_resident;
  int stage_tile;    // bytes of one epilogue warp's staging tile: 2048 (32 rows x 32 bf16) or 4096 (32 rows x 32 fp32)
  uint32_t bres_bytes;
  int ldy, ldr, ldgx, ldgk, gHk, gWk;
  float g_sh, g_sw;
  int act, out_f32;
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t idesc;
  uint32_t sbo_enc, layout_type;
  // magic-number division (q = umulhi(n, mul) >> shr, mul == 0 <=> divisor 1) for the per-tile / per-item index
  // decompositions: the generic integer division is ~25 instructions + 2 XU ops each, and the epilogue of the
  // small-channel layers is issue- and XU-bound
  uint32_t fd_nn[2], fd_pi[2], fd_tx[2], fd_hw[2], fd_wo[2];
};

// smem carve-up (after manual 1024-byte alignment):
//   [stages][A tile | B tile]   stage_bytes each
//   full[8], empty[8], tmem_full[2], tmem_empty[2], afull[4], aempty[4] mbarriers, tmem base ptr, scale[256], bias[256]
//   (halo mode: [n_abuf halo buffers] precede the stages, which then hold B tiles only)
constexpr int kMaxABuf = 8;
constexpr double kHaloMinEff = 0.8;   // fraction of a map's pixels among the pixels of its 16x8 patches
constexpr int kHaloTH = 16, kHaloTW = 8;
constexpr int kMaxAcc = 8;                                                     // TMEM accumulator ring (512 columns / block_n)
constexpr uint32_t kNumBars = 2 * kMaxStages + 2 * kMaxAcc + 2 * kMaxABuf + kEpiWarps;   // + one operand barrier per epilogue warp
constexpr int kMaxCout = 2048;                                                 // scale/bias staged once per CTA
// tail (1024-aligned, after the pipeline stages):
//   [kEpiWarps x 2 KB] output staging (thread == row writes 4 x 16 B, 64-byte-swizzled, a TMA store drains it)
//   [kEpiWarps x 2 KB] operand staging (residual / gate_x tile of the warp's NEXT item, TMA-loaded) — only when used
//   barriers, TMEM base slot, scale[kMaxCout], bias[kMaxCout]
__host__ __device__ constexpr uint32_t tail_bytes(int epi_warps, bool operand_stage, uint32_t sb_floats, uint32_t tile = 2048u) {
  return (uint32_t)epi_warps * tile * (operand_stage ? 2u : 1u) + kNumBars * 8 + 32 + 2 * sb_floats * 4;
}

// one elected lane of a converged warp (the warp stays converged, so operands live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Four consecutive k-steps (one 64-channel tap) from ONE asm block: descriptor low words advance by 2 (32 bytes >> 4)
// inside the block, so ptxas sees the invariant operands (TMEM address, instruction descriptor, descriptor high words)
// once per tap instead of once per MMA (the per-MMA form re-creates them in uniform registers each time: four R2UR).
template <bool PAIR>
__device__ __forceinline__ void umma_bf16_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
#define DMAY_MMA4(CG)                                                                    \
  asm volatile(                                                                           \
      "{\n\t.reg .pred p, t;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"             \
      "setp.ne.b32 p, %6, 0;\n\t"                                                         \
      "setp.eq.u32 t, %6, %6;\n\t"                                                        \
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"                                \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %5, p;\n\t"                  \
      "add.u32 al, %1, 2;\n\tadd.u32 bl, %3, 2;\n\t"                                      \
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"                                \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %5, t;\n\t"                  \
      "add.u32 al, %1, 4;\n\tadd.u32 bl, %3, 4;\n\t"                                      \
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"                                \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %5, t;\n\t"                  \
      "add.u32 al, %1, 6;\n\tadd.u32 bl, %3, 6;\n\t"                                      \
      "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"                                \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %5, t;\n\t}" ::"r"(tmem_d),   \
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)            \
      : "memory")
  if (PAIR) DMAY_MMA4("2");
  else DMAY_MMA4("1");
#undef DMAY_MMA4
}
// All nine taps of one input patch (3x3 / stride 1, resident weights) in ONE asm block: KK MMAs per tap (KK = CK / 16).
// ncu (r5g, 16->64 and 64->64 3x3 at 320x320): the issuing warp executed ~40 instructions per tap -- loop control, the
// tap's address arithmetic (an IMAD fed by a constant-bank load), predicate and descriptor set-up -- ~2700 clk per 128 x 64
// tile for 290 (stem) / 1150 (64 channels) clk of tensor work, and every other warp of the CTA waited for it.  Here the
// tap offsets are immediates ((r * 10 + s) patch rows, t weight tiles), the invariant operands are named once, and the
// block is straight-line code.
//   a_lo / b_lo : descriptor low words of tap 0 (start address >> 4 | LBO); row16 = bytes of one patch pixel >> 4;
//   b_step      : bytes of one tap's weight tile >> 4.
#define DMAY_T9_MMA(CG, P) "mov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\ttcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %5, " P ";\n\t"
#define DMAY_T9_NEXT "add.u32 al, al, 2;\n\tadd.u32 bl, bl, 2;\n\t"
#define DMAY_T9_KK1(CG, P) DMAY_T9_MMA(CG, P)
#define DMAY_T9_KK2(CG, P) DMAY_T9_MMA(CG, P) DMAY_T9_NEXT DMAY_T9_MMA(CG, "t")
#define DMAY_T9_KK4(CG, P) DMAY_T9_MMA(CG, P) DMAY_T9_NEXT DMAY_T9_MMA(CG, "t") DMAY_T9_NEXT DMAY_T9_MMA(CG, "t") DMAY_T9_NEXT DMAY_T9_MMA(CG, "t")
#define DMAY_T9_TAP(KKM, CG, T, AOFF, P) "mad.lo.u32 al, %7, " #AOFF ", %1;\n\tmad.lo.u32 bl, %8, " #T ", %3;\n\t" KKM(CG, P)
#define DMAY_T9_ALL(KKM, CG)                                                                                              \
  asm volatile("{\n\t.reg .pred p, t;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"                                       \
               "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 t, %6, %6;\n\t"                                                       \
               DMAY_T9_TAP(KKM, CG, 0, 0, "p") DMAY_T9_TAP(KKM, CG, 1, 1, "t") DMAY_T9_TAP(KKM, CG, 2, 2, "t")             \
               DMAY_T9_TAP(KKM, CG, 3, 10, "t") DMAY_T9_TAP(KKM, CG, 4, 11, "t") DMAY_T9_TAP(KKM, CG, 5, 12, "t")          \
               DMAY_T9_TAP(KKM, CG, 6, 20, "t") DMAY_T9_TAP(KKM, CG, 7, 21, "t") DMAY_T9_TAP(KKM, CG, 8, 22, "t") "}"      \
               ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(row16), "r"(b_step) \
               : "memory")
template <bool PAIR>
__device__ __forceinline__ void umma_taps9(int kk_n, uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate, uint32_t row16, uint32_t b_step) {
  static_assert(kHaloTW + 2 == 10, "tap offsets below assume a 10-pixel patch pitch");
  if (kk_n == 4) {
    if (PAIR) DMAY_T9_ALL(DMAY_T9_KK4, "2");
    else DMAY_T9_ALL(DMAY_T9_KK4, "1");
  } else if (kk_n == 2) {
    if (PAIR) DMAY_T9_ALL(DMAY_T9_KK2, "2");
    else DMAY_T9_ALL(DMAY_T9_KK2, "1");
  } else {
    if (PAIR) DMAY_T9_ALL(DMAY_T9_KK1, "2");
    else DMAY_T9_ALL(DMAY_T9_KK1, "1");
  }
}
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_enc, uint32_t layout_type) {
  // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout[61,64)
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)sbo_enc << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout_type << 61);
}

__device__ __forceinline__ int fdiv(int n, const uint32_t (&fd)[2]) {
  return fd[0] ? (int)(__umulhi((uint32_t)n, fd[0]) >> fd[1]) : n;
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// silu(x) = x*sigmoid(x) = h + h*tanh(h), h = x/2   (one MUFU; rel. error 2^-11, below bf16 resolution)
__device__ __forceinline__ float silu_t(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
__device__ __forceinline__ float sigmoid_t(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

// bf16 modes: one 8-column group of one output row -> packed bf16x8 (stored later, coalesced, via the warp's
// shared-memory transpose stage).  aux0 / aux1 = residual (or gate_x) / gate_k values of the same 8 columns.
template <int MODE>
__device__ __forceinline__ uint4 epi_compute8(const uint32_t* r, const float* sc, const float* bi, const uint4& aux0,
                                              const uint4& aux1) {
  // scale / bias live in shared memory, 16-byte aligned: four broadcast LDS.128 per 8 columns
  const float4 s0 = reinterpret_cast<const float4*>(sc)[0], s1 = reinterpret_cast<const float4*>(sc)[1];
  const float4 b0 = reinterpret_cast<const float4*>(bi)[0], b1 = reinterpret_cast<const float4*>(bi)[1];
  const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float f[8];
#if DMAY_EPI_F2
  // packed fp32x2 FMAs (same IEEE roundings as the scalar form: outputs bit-identical): 4 FFMA2 for scale / bias and 4 for
  // h + h * tanh(h) instead of 8 + 8 FFMA per 8 columns.  Same-box A/B (cfg-2): 12.33 / 12.34 -> 12.16 / 12.26 ms per step.
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f32x2_t h2 = f2_fma(f2_pack(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])), f2_pack(s[2 * k], s[2 * k + 1]),
                        f2_pack(b[2 * k], b[2 * k + 1]));
    if (MODE == EPI_SILU || MODE == EPI_SILU_RES || MODE == EPI_SILU_PRE) {
      // the staged scale / bias of these modes are pre-halved (exact), so h2 is already h = x/2: silu(x) = h + h*tanh(h)
      float h0, h1;
      f2_unpack(h2, h0, h1);
      h2 = f2_fma(h2, f2_pack(tanh_approx(h0), tanh_approx(h1)), h2);
    }
    f2_unpack(h2, f[2 * k], f[2 * k + 1]);
  }
#else
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = fmaf(__uint_as_float(r[j]), s[j], b[j]);
  if (MODE == EPI_SILU || MODE == EPI_SILU_RES || MODE == EPI_SILU_PRE) {
    // the staged scale / bias of these modes are pre-halved (exact), so f is already h = x/2: silu(x) = h + h*tanh(h)
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], tanh_approx(f[j]), f[j]);
  }
#endif
  if (MODE == EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.5f * f[j] * (1.0f + erff(f[j] * 0.70710678118654752f));
  }
  if (MODE == EPI_SILU_RES || MODE == EPI_LINEAR_RES) {
    float rs[8];
    unpack8(aux0, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += rs[j];
  }
  if (MODE == EPI_LINEAR_MUL) {
    float rs[8];
    unpack8(aux0, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= rs[j];
  }
  if (MODE == EPI_GATE) {
    float gx[8], gk[8];
    unpack8(aux0, gx);
    unpack8(aux1, gk);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= sigmoid_t(gx[j] + gk[j]);
  }
  return pack8(f);
}

// One 8-column group of one output row: registers r[0..7] (fp32 accumulators) -> global.
template <int MODE>
__device__ __forceinline__ void epi_store8(const ConvArgs& a, const uint32_t* r, const float* sc, const float* bi,
                                           long long row, int col, const uint4& aux0, const uint4& aux1) {
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = fmaf(__uint_as_float(r[j]), sc[j], bi[j]);
  if (MODE == EPI_SILU || MODE == EPI_SILU_RES) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], tanh_approx(f[j]), f[j]);   // staged scale / bias are pre-halved
  }
  if (MODE == EPI_SILU_RES) {
    float rs[8];
    unpack8(aux0, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += rs[j];
  }
  if (MODE == EPI_GATE) {
    float gx[8], gk[8];
    unpack8(aux0, gx);
    unpack8(aux1, gk);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= sigmoid_t(gx[j] + gk[j]);
  }
  if (MODE == EPI_GENERIC) {
    if (a.act != DMAY_ACT_NONE) {
#pragma unroll 1
      for (int j = 0; j < 8; ++j) f[j] = apply_act(f[j], a.act);
    }
    if (a.residual != nullptr) {
      float rs[8];
      unpack8(ld_nc16(a.residual + row * a.ldr + col), rs);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += rs[j];
    }
  }
  if (MODE == EPI_LINEAR_F32 || (MODE == EPI_GENERIC && a.out_f32)) {
    float* yo = reinterpret_cast<float*>(a.y) + row * a.ldy + col;
    *reinterpret_cast<float4*>(yo) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(yo + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    st16(reinterpret_cast<__nv_bfloat16*>(a.y) + row * a.ldy + col, pack8(f));
  }
}

// virtual channel concat (1x1 layers): which tensor map / which column the channel chunk `cc` of the K loop reads
__device__ __forceinline__ const CUtensorMap* seg_map(const ConvArgs& a, int cc) {
  return (a.seg1 == 0 || cc < a.seg1) ? &a.tmA : (cc < a.seg2 ? &a.tmA2 : &a.tmA3);
}
__device__ __forceinline__ int seg_col(const ConvArgs& a, int cc) {
  return ((a.seg1 == 0 || cc < a.seg1) ? cc : (cc < a.seg2 ? cc - a.seg1 : cc - a.seg2)) * a.CK;
}

// PAIR is a template parameter (not a runtime flag): a kernel image that contains cta_group::2 instructions can only be
// launched as a cluster (error 912 otherwise), so the single-CTA and the CTA-pair tile shapes are separate images.
// T9 (plain SiLU layers with resident 3x3 weights only): the nine taps of a patch are issued as one straight-line asm block
// (umma_taps9).  A separate image, because with those blocks in the function ptxas spills more in EVERY instantiation's
// epilogue (residual / gate: 116 / 204 bytes instead of 32 / 84; plain: 20 instead of 0), which cost the other layers
// 4-45 % in the network (same-box A/B r5k-r5m) -- they keep the image without the blocks.
template <int MODE, bool PAIR, bool T9 = false>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bres0 = base;                                           // resident weights (b_resident only)
  const uint32_t abuf0 = base + a.bres_bytes;                            // halo ring (halo mode only)
  const uint32_t stage0 = abuf0 + (a.halo ? a.n_abuf * a.a_halo_bytes : 0u);
  const uint32_t data_bytes = (stage0 - base) + a.stages * a.stage_bytes;
  const uint32_t tail = base + data_bytes;                               // 1024-aligned (all data regions are)
  uint8_t* tail_ptr = base_ptr + data_bytes;
  const uint32_t epi_stage_bytes = (uint32_t)a.epi_warps * (uint32_t)a.stage_tile;
  const uint32_t bars_off = epi_stage_bytes * (a.opnd_stage == 1 ? 2u : 1u);   // opnd_stage 2: operand tile shares the output tile
  const uint32_t full_bar = tail + bars_off, empty_bar = full_bar + kMaxStages * 8;
  const uint32_t tfull_bar = full_bar + 2 * kMaxStages * 8, tempty_bar = tfull_bar + kMaxAcc * 8;
  const uint32_t afull_bar = tempty_bar + kMaxAcc * 8, aempty_bar = afull_bar + kMaxABuf * 8;
  const uint32_t opnd_bar = aempty_bar + kMaxABuf * 8;                   // [kEpiWarps]
  const uint32_t tmem_slot = full_bar + kNumBars * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tail_ptr + bars_off + kNumBars * 8);
  float* s_scale = reinterpret_cast<float*>(tail_ptr + ((bars_off + kNumBars * 8 + 16 + 15) & ~15u));
  float* s_bias = s_scale + a.sb_floats;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile schedule: a cluster walks "super tiles" = cs consecutive m-tiles of one n-tile (so its CTAs share
  // the B tile); n-tile fastest, so clusters running side by side share the activation rows in L2.
  const int cs = a.cs;
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / cs, num_clusters = gridDim.x / cs;
  const int m_groups = (a.num_m_tiles + cs - 1) / cs;
  const int total_super = m_groups * a.num_n_tiles;
  const int k_iters = (a.total_subs + a.subs - 1) / a.subs;
  const uint16_t mc_mask = (uint16_t)((1u << cs) - 1u);

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&a.tmA);
    tmap_prefetch(&a.tmB);
    tmap_prefetch(&a.tmY);
    if (a.opnd_stage) tmap_prefetch(&a.tmR);
    if (a.seg1) {
      tmap_prefetch(&a.tmA2);
      if (a.seg2 < a.c_chunks) tmap_prefetch(&a.tmA3);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(full_bar + i * 8, 1);
      mbar_init(empty_bar + i * 8, PAIR ? 1 : cs);   // multicast mode: every CTA of the cluster releases the stage
    }
    for (int i = 0; i < kMaxAcc; ++i) {
      mbar_init(tfull_bar + i * 8, 1);
      mbar_init(tempty_bar + i * 8, (uint32_t)(a.epi_warps / a.epi_groups) * 32u * (PAIR ? 2u : 1u));
    }
    for (int i = 0; i < kMaxABuf; ++i) {
      mbar_init(afull_bar + i * 8, 1);
      mbar_init(aempty_bar + i * 8, 1);
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(opnd_bar + i * 8, 1);
    fence_barrier_init();
  }
  // folded-BN scale / bias of every output channel, staged once (Cout_pad <= kMaxCout, checked on the host)
  // (SiLU modes keep them halved -- exact -- so the epilogue's FMA yields x/2 directly, see epi_compute8)
  const float sb_mul = (MODE == EPI_SILU || MODE == EPI_SILU_RES || MODE == EPI_SILU_PRE) ? 0.5f : 1.0f;
  for (int i = threadIdx.x; i < a.sb_floats; i += blockDim.x) {
    const bool in = i < a.Cout_pad;
    s_scale[i] = in ? sb_mul * a.scale[i] : 0.f;
    s_bias[i] = in ? sb_mul * a.bias[i] : 0.f;
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc2(tmem_slot, (uint32_t)a.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // peers' barriers are initialised before anything multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch, folded-BN constants) ran
  // while the previous kernel of the stream was still draining; activations are only touched after this point.
  // The next kernel may start its own preamble as soon as every CTA of this grid has passed here (or exited).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the schedule (uniform), one elected lane issues =====
    int stage = 0, ab = 0;
    uint32_t phase = 0, aphase = 0;
    const int HoWo = a.Ho * a.Wo;
    if (a.b_resident == 1) {
      // weights once (barrier full[0]), then one input patch per (tile, channel chunk)
      // (CTA pair: each CTA keeps HALF of the weight set -- its block_n / 2 output channels -- and both halves are
      //  credited to the even CTA's barrier; the 128-channel 3x3 layers re-streamed 147 KB of weights per tile before)
      if (elect_one()) {
        if (!PAIR || crank == 0) mbar_expect_tx(full_bar, (PAIR ? 2u : 1u) * (uint32_t)a.total_subs * a.b_bytes);
        int g = 0;
#pragma unroll 1
        for (int cc = 0; cc < a.c_chunks; ++cc)
#pragma unroll 1
          for (int tap = 0; tap < a.taps; ++tap, ++g) {
            if (PAIR) tma_load_2d_pair(bres0 + g * a.b_bytes, &a.tmB, full_bar, tap * a.Cin + cc * a.CK, (int)crank * (a.block_n / 2));
            else tma_load_2d(bres0 + g * a.b_bytes, &a.tmB, full_bar, tap * a.Cin + cc * a.CK, 0);
          }
      }
      __syncwarp();
      const int per_img = a.tiles_x * a.tiles_y;
#pragma unroll 1
      for (int st = cluster_id; st < total_super; st += num_clusters) {
        const int m_tile = PAIR ? st * 2 + (int)crank : st;   // beyond the last tile (odd count): the loads zero-fill
        const int n_img = fdiv(m_tile, a.fd_pi);
        const int rem = m_tile - n_img * per_img;
        const int ty = fdiv(rem, a.fd_tx), tx = rem - ty * a.tiles_x;
#pragma unroll 1
        for (int cc = 0; cc < a.c_chunks; ++cc) {
          mbar_wait(aempty_bar + ab * 8, aphase ^ 1u);
          if (elect_one()) {
            if (PAIR) {
              if (crank == 0) mbar_expect_tx(afull_bar + ab * 8, 2u * a.a_halo_tx);
              tma_load_4d_pair(abuf0 + ab * a.a_halo_bytes, &a.tmA, afull_bar + ab * 8, cc * a.CK, tx * kHaloTW - 1,
                               ty * kHaloTH - 1, n_img);
            } else {
              mbar_expect_tx(afull_bar + ab * 8, a.a_halo_tx);
              tma_load_4d(abuf0 + ab * a.a_halo_bytes, &a.tmA, afull_bar + ab * 8, cc * a.CK, tx * kHaloTW - 1,
                          ty * kHaloTH - 1, n_img);
            }
          }
          __syncwarp();
          if (++ab == a.n_abuf) {
            ab = 0;
            aphase ^= 1u;
          }
        }
      }
    } else {
    const int b_rows = a.block_n / cs;                       // rows of the B tile this CTA fetches
    const uint32_t b_slice = (uint32_t)(b_rows * a.CK * 2);
    if (a.b_resident == 2) {
      // plain (non-halo) weights-resident mode: the whole [block_n x K] weight set is fetched ONCE per CTA (single
      // n-tile, <= ~130 KB), the pipeline then streams activation tiles only.  For the shallow-K 1x1 layers the weight
      // tile was 2/3 of the L2 -> SM traffic of every tile.
      if (elect_one()) {
        mbar_expect_tx(afull_bar, (uint32_t)a.total_subs * a.b_bytes);
#pragma unroll 1
        for (int g = 0; g < a.total_subs; ++g) tma_load_2d(bres0 + g * a.b_bytes, &a.tmB, afull_bar, g * a.CK, 0);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int st = cluster_id; st < total_super; st += num_clusters) {
      const int sq = fdiv(st, a.fd_nn);
      const int n_tile = st - sq * a.num_n_tiles, m_tile = sq * cs + (int)crank;
      const int m0 = m_tile * BLOCK_M, n0 = n_tile * a.block_n;   // m0 >= M for a padding tile: loads zero-fill
      int n_img = 0, h0 = 0, w0 = 0;
      if (a.halo) {
        const int per_img = a.tiles_x * a.tiles_y;
        n_img = fdiv(m_tile, a.fd_pi);
        const int rem = m_tile - n_img * per_img;
        const int ty = fdiv(rem, a.fd_tx), tx = rem - ty * a.tiles_x;
        h0 = ty * kHaloTH - 1;
        w0 = tx * kHaloTW - 1;
      } else if (a.im2col) {
        n_img = fdiv(m0, a.fd_hw);
        const int rem = m0 - n_img * HoWo;
        const int p = fdiv(rem, a.fd_wo), q = rem - p * a.Wo;
        h0 = p * a.conv_stride - a.pad;
        w0 = q * a.conv_stride - a.pad;
      }
      // running (tap, channel chunk) of the next sub-tile: no divisions in the loop.
      // im2col order: tap outer, chunk inner.  halo order: chunk outer, tap inner (one patch serves 9 taps).
      int tap = 0, cc = 0, r = 0, s = 0;
#pragma unroll 1
      for (int it = 0; it < k_iters; ++it) {
        const int nsub = min(a.subs, a.total_subs - it * a.subs);
        const bool leader = elect_one();
        if (a.halo && tap == 0) {   // first tap of a channel chunk: fetch the input patch once
          mbar_wait(aempty_bar + ab * 8, aphase ^ 1u);
          if (leader) {
            if (PAIR) {   // both CTAs' patches are credited to the even CTA's barrier
              if (crank == 0) mbar_expect_tx(afull_bar + ab * 8, 2u * a.a_halo_tx);
              tma_load_4d_pair(abuf0 + ab * a.a_halo_bytes, &a.tmA, afull_bar + ab * 8, cc * a.CK, w0, h0, n_img);
            } else {
              mbar_expect_tx(afull_bar + ab * 8, a.a_halo_tx);
              tma_load_4d(abuf0 + ab * a.a_halo_bytes, &a.tmA, afull_bar + ab * 8, cc * a.CK, w0, h0, n_img);
            }
          }
          if (++ab == a.n_abuf) {
            ab = 0;
            aphase ^= 1u;
          }
        }
        mbar_wait(empty_bar + stage * 8, phase ^ 1u);
        const uint32_t fb = full_bar + stage * 8;
        const uint32_t sa = stage0 + stage * a.stage_bytes, sb = sa + (a.halo ? 0u : a.subs * a.a_bytes);
        if (PAIR) {
          // the even CTA's barrier collects the bytes of both CTAs (its own expect may come after the peer's first
          // complete_tx of the phase: the transaction count is signed)
          if (leader && crank == 0) mbar_expect_tx(fb, 2u * (uint32_t)nsub * ((a.halo ? 0u : a.a_bytes) + a.b_bytes));
        } else if (leader) {
          mbar_expect_tx(fb, (uint32_t)nsub * ((a.halo ? 0u : a.a_bytes) + (a.b_resident == 2 ? 0u : a.b_bytes)));
        }
#pragma unroll 1
        for (int j = 0; j < nsub; ++j) {
          if (leader) {
            if (PAIR) {
              if (a.halo) {
              } else if (a.im2col)
                tma_load_im2col_4d_pair(sa + j * a.a_bytes, &a.tmA, fb, cc * a.CK, w0, h0, n_img, (uint16_t)s, (uint16_t)r);
              else
                tma_load_2d_pair(sa + j * a.a_bytes, seg_map(a, cc), fb, seg_col(a, cc), m0);
              tma_load_2d_pair(sb + j * a.b_bytes, &a.tmB, fb, tap * a.Cin + cc * a.CK, n0 + (int)crank * (a.block_n / 2));
            } else if (a.halo) {
            } else if (a.im2col)
              tma_load_im2col_4d(sa + j * a.a_bytes, &a.tmA, fb, cc * a.CK, w0, h0, n_img, (uint16_t)s, (uint16_t)r);
            else
              tma_load_2d(sa + j * a.a_bytes, seg_map(a, cc), fb, seg_col(a, cc), m0);
            if (PAIR || a.b_resident == 2) {
            } else if (cs > 1)
              tma_load_2d_mc(sb + j * a.b_bytes + crank * b_slice, &a.tmB, fb, tap * a.Cin + cc * a.CK,
                             n0 + (int)crank * b_rows, mc_mask);
            else
              tma_load_2d(sb + j * a.b_bytes, &a.tmB, fb, tap * a.Cin + cc * a.CK, n0);
          }
          if (a.halo) {
            if (++tap == a.taps) {
              tap = 0;
              ++cc;
            }
          } else if (++cc == a.c_chunks) {
            cc = 0;
            ++tap;
            if (++s == a.kw) {
              s = 0;
              ++r;
            }
          }
        }
        __syncwarp();
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    }  // !b_resident
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, elect.sync around the tcgen05 instructions =====
    int stage = 0, ab = 0;
    uint32_t phase = 0, aphase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int kk_n = a.CK / 16;
    // descriptor halves: lo = start>>4 | LBO(1)<<16 ; hi = SBO | version 1 (bit 46) | layout (bits 61..63)
    const uint32_t desc_hi = a.sbo_enc | (1u << 14) | (a.layout_type << 29);
    const uint32_t desc_hi_halo = a.halo_sbo_enc | (1u << 14) | (a.layout_type << 29);   // SBO = one patch row
    const uint32_t row_bytes = (uint32_t)a.CK * 2u;
    const uint32_t a_step = a.a_bytes >> 4, b_step = a.b_bytes >> 4;
    if (a.b_resident == 1) {
      if (!(PAIR && crank != 0)) {   // CTA pair: the even CTA issues every MMA
      mbar_wait(full_bar, 0);   // resident weights have landed (both halves in pair mode)
      // one-block issue of the nine taps (T9 images; flags bit16 selects the plain image: A/B).  Same-box A/B r5h: stem
      // 16->64 @320 0.521 -> 0.280 ms, 128->128 plain (CTA pairs) 0.3425 -> 0.3016 ms @160, 0.0926 -> 0.0812 @80.
      constexpr bool taps9 = T9;
#pragma unroll 1
      for (int st = cluster_id; st < total_super; st += num_clusters) {
        mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * a.acc_stride);
        uint32_t accum = 0;
        uint32_t b_lo = ((bres0 >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll 1
        for (int cc = 0; cc < a.c_chunks; ++cc) {
          mbar_wait(afull_bar + ab * 8, aphase);
          tc_fence_after();
          const uint32_t a_patch = abuf0 + ab * a.a_halo_bytes;
          if (elect_one()) {
            const uint32_t hlo0 = ((a_patch >> 4) & 0x3FFFu) | (1u << 16);
            if constexpr (taps9) {
              // one straight-line block for the nine taps (see umma_taps9)
              umma_taps9<PAIR>(kk_n, tmem_d, hlo0, desc_hi_halo, b_lo, desc_hi, a.idesc, accum, row_bytes >> 4, b_step);
            } else {
#pragma unroll 1
            for (int r = 0; r < 3; ++r)
#pragma unroll 1
              for (int s2 = 0; s2 < 3; ++s2) {
                const uint32_t start = a_patch + (uint32_t)(r * a.halo_pitch + s2) * row_bytes;
                const uint32_t hlo = ((start >> 4) & 0x3FFFu) | (1u << 16);
                // (single-CTA tiles only: on the CTA-pair 128-channel layers the tighter issue measured the same for the
                //  plain / gate epilogues and 14 % slower for the residual one, same box A/B r4k)
                if (!PAIR && kk_n == 4) {
                  umma_bf16_x4<PAIR>(tmem_d, hlo, desc_hi_halo, b_lo + (uint32_t)(r * 3 + s2) * b_step, desc_hi, a.idesc, accum);
                  accum = 1;
                } else {
#pragma unroll 2
                  for (int kk = 0; kk < kk_n; ++kk) {
                    const uint32_t bl = b_lo + (uint32_t)(r * 3 + s2) * b_step + kk * 2;
                    if (PAIR) umma_bf16_2(tmem_d, pack64(hlo + kk * 2, desc_hi_halo), pack64(bl, desc_hi), a.idesc, accum);
                    else umma_bf16(tmem_d, pack64(hlo + kk * 2, desc_hi_halo), pack64(bl, desc_hi), a.idesc, accum);
                    accum = 1;
                  }
                }
              }
            }
            if (PAIR) umma_commit_2(aempty_bar + ab * 8);
            else umma_commit(aempty_bar + ab * 8);
          }
          b_lo += 9u * b_step;   // (uniform) running descriptor of the next channel chunk's weights
          accum = 1;
          __syncwarp();
          if (++ab == a.n_abuf) {
            ab = 0;
            aphase ^= 1u;
          }
        }
        if (elect_one()) {
          if (PAIR) umma_commit_2(tfull_bar + acc * 8);
          else umma_commit(tfull_bar + acc * 8);
        }
        __syncwarp();
        if (++acc == a.nacc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      }
    } else if (PAIR && crank != 0) {
      // CTA pair: the even CTA issues every MMA (they read both CTAs' shared memory and write both CTAs' TMEM)
    } else {
    if (a.b_resident == 2) mbar_wait(afull_bar, 0);   // resident weights have landed
#pragma unroll 1
    for (int st = cluster_id; st < total_super; st += num_clusters) {
      mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * a.acc_stride);
      uint32_t accum = 0;
      int tap = 0;                       // halo mode: tap of the next sub-tile (chunk outer, tap inner)
      uint32_t a_patch = 0;
#pragma unroll 1
      for (int it = 0; it < k_iters; ++it) {
        const int nsub = min(a.subs, a.total_subs - it * a.subs);
        if (a.halo && tap == 0) {
          mbar_wait(afull_bar + ab * 8, aphase);
          a_patch = abuf0 + ab * a.a_halo_bytes;
        }
        mbar_wait(full_bar + stage * 8, phase);
        tc_fence_after();
        const uint32_t sa = stage0 + stage * a.stage_bytes, sb = sa + (a.halo ? 0u : a.subs * a.a_bytes);
        const bool last_of_chunk = a.halo && (tap + nsub == a.taps);
        if (elect_one()) {
          uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16), b_lo = ((sb >> 4) & 0x3FFFu) | (1u << 16);
          if (a.b_resident == 2) b_lo = (((bres0 + (uint32_t)(it * a.subs) * a.b_bytes) >> 4) & 0x3FFFu) | (1u << 16);
          if (a.halo) {
#pragma unroll 1
            for (int j = 0; j < nsub; ++j) {
              const int t = tap + j;
              const int r = t / 3, s = t - 3 * r;
              const uint32_t start = a_patch + (uint32_t)(r * a.halo_pitch + s) * row_bytes;   // shifted window
              const uint32_t hlo = ((start >> 4) & 0x3FFFu) | (1u << 16);
              const uint32_t hhi = desc_hi_halo;   // base-offset field stays 0: UMMA swizzles on absolute smem address bits
#pragma unroll 4
              for (int kk = 0; kk < kk_n; ++kk) {
                if (PAIR) umma_bf16_2(tmem_d, pack64(hlo + kk * 2, hhi), pack64(b_lo + kk * 2, desc_hi), a.idesc, accum);
                else umma_bf16(tmem_d, pack64(hlo + kk * 2, hhi), pack64(b_lo + kk * 2, desc_hi), a.idesc, accum);
                accum = 1;
              }
              b_lo += b_step;
            }
          } else {
#pragma unroll 1
            for (int j = 0; j < nsub; ++j) {
#pragma unroll 4
              for (int kk = 0; kk < kk_n; ++kk) {  // advance 16 elements = 32 bytes (>>4 = 2) inside the swizzle atom
                if (PAIR) umma_bf16_2(tmem_d, pack64(a_lo + kk * 2, desc_hi), pack64(b_lo + kk * 2, desc_hi), a.idesc, accum);
                else umma_bf16(tmem_d, pack64(a_lo + kk * 2, desc_hi), pack64(b_lo + kk * 2, desc_hi), a.idesc, accum);
                accum = 1;
              }
              a_lo += a_step;
              b_lo += b_step;
            }
          }
          if (PAIR) umma_commit_2(empty_bar + stage * 8);
          else if (cs > 1) umma_commit_mc(empty_bar + stage * 8, mc_mask);
          else umma_commit(empty_bar + stage * 8);
          if (last_of_chunk) {   // all 9 taps of this patch have been issued
            if (PAIR) umma_commit_2(aempty_bar + ab * 8);
            else umma_commit(aempty_bar + ab * 8);
          }
        }
        accum = 1;
        if (a.halo) {
          tap += nsub;
          if (tap == a.taps) {
            tap = 0;
            if (++ab == a.n_abuf) {
              ab = 0;
              aphase ^= 1u;
            }
          }
        }
        __syncwarp();
        if (++stage == a.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) {
        if (PAIR) umma_commit_2(tfull_bar + acc * 8);
        else umma_commit(tfull_bar + acc * 8);
      }
      __syncwarp();
      if (++acc == a.nacc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    }  // !b_resident
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps; warp w reads TMEM lane quadrant w%4 (hardware rule) and owns the 32-column chunks
    // c0 = sub*32 + k*128 (sub = (w-4)/4) of every tile of this CTA, as a flat sequence of (tile, chunk) items.
    // Four epilogue warps per scheduler instead of two: the per-item instruction stream is a chain of short
    // dependencies (TMEM load -> FMA -> MUFU -> pack -> smem), so throughput comes from warps, not from ILP.
    //   output  : thread == row packs 32 bf16 into the warp's 2 KB staging tile (64-byte swizzle) and ONE TMA
    //             store drains it — bounds (M tail, image border in halo mode, Cout) are clipped by the tensor map.
    //   operand : the residual / gate_x tile of the warp's NEXT item is TMA-loaded into a second 2 KB tile while
    //             the current item is processed (no registers held across items).
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int grp = a.epi_groups == 2 ? (ew >> 3) : 0;          // tile parity this warp serves (two-group mode)
    const int wpg = a.epi_warps / a.epi_groups;                 // warps per group
    const int sub = (ew >> 2) & (wpg / 4 - 1);
    const int row_in_tile = quad * 32 + lane;
    int acc = grp;                                              // ring slot of this warp's next tile (tile j -> j % nacc)
    uint32_t acc_phase = 0;
    const int HoWo = a.Ho * a.Wo;
    const int cstride = wpg * 8;                                                       // columns between a warp's chunks
    const int cpw = a.block_n > sub * 32 ? (a.block_n - sub * 32 + cstride - 1) / cstride : 0;   // chunks per tile of this warp
    const int cta_tiles = cluster_id < total_super ? (total_super - cluster_id + num_clusters - 1) / num_clusters : 0;
    const int my_tiles = a.epi_groups == 2 ? (cta_tiles + 1 - grp) / 2 : cta_tiles;
    const int items = my_tiles * (cpw > 0 ? cpw : 1);                                  // cpw == 0: one "empty" item per tile

    constexpr bool kStaged = (MODE == EPI_SILU || MODE == EPI_SILU_RES || MODE == EPI_LINEAR || MODE == EPI_GATE ||
                              MODE == EPI_LINEAR_RES || MODE == EPI_GELU || MODE == EPI_LINEAR_MUL || MODE == EPI_SILU_PRE);
    constexpr bool kOpnd = (MODE == EPI_SILU_RES || MODE == EPI_GATE || MODE == EPI_LINEAR_RES || MODE == EPI_LINEAR_MUL);
    const uint32_t out_stage = tail + (uint32_t)ew * (uint32_t)a.stage_tile;
    uint4* out_ptr = reinterpret_cast<uint4*>(tail_ptr + ew * a.stage_tile);
    // opnd_stage 2 (CTA-pair layers with resident weights: no room for a second tile per warp): the operand tile lands in
    // the warp's OUTPUT staging tile -- it is consumed into registers before the output is packed, and the next one is
    // requested only after the store has finished reading the tile.  With 16 epilogue warps (one item per warp and tile)
    // that request completes while the warp waits for its next accumulator.
    const bool op_shared = a.opnd_stage == 2;
    const uint32_t op_stage = op_shared ? out_stage : tail + epi_stage_bytes + (uint32_t)ew * 2048u;
    const uint4* op_ptr = op_shared ? reinterpret_cast<const uint4*>(tail_ptr + ew * a.stage_tile)
                                    : reinterpret_cast<const uint4*>(tail_ptr + epi_stage_bytes + ew * 2048);
    const uint32_t my_opnd_bar = opnd_bar + (uint32_t)ew * 8u;
    uint32_t opnd_phase = 0;
    auto sidx = [](int r, int j) { return r * 4 + (j ^ ((r >> 1) & 3)); };   // 64-byte rows, TMA SWIZZLE_64B pattern

    // (a relaxed arrive is legal here and measured 9-14 % faster on isolated CTA-pair residual / gate layers, r5i, but
    //  3-6 % slower on the same layers inside the network, r5m: the release form stays)
    auto tmem_hand_back = [&](uint32_t bar) {
      if (PAIR) mbar_arrive_pair(bar);
      else mbar_arrive(bar);
    };
    struct Item {
      int row;         // this thread's output row (pixel index), -1: not stored (beyond M / outside the image)
      const __nv_bfloat16* gk_row;   // EPI_GATE: this row's gate_k pixel; EPI_SILU_PRE: its fp32 `pre` pixel (reinterpreted)
      int n0, c0, width;
      int t1, t2, t3;  // TMA coordinates of the warp's 32-row box: 2-D {col, t1}; halo {col, t1 = w, t2 = h, t3 = n}
      bool first, last;
    };
    // (tile, chunk) of the next item to build: advanced incrementally, no division by `per`
    int mk_ti = 0, mk_ch = 0;
    const int per = cpw > 0 ? cpw : 1;
    auto make_item = [&](Item& it) {
      const int ti = mk_ti, ch = mk_ch;
      if (++mk_ch == per) {
        mk_ch = 0;
        ++mk_ti;
      }
      const int st = cluster_id + (a.epi_groups == 2 ? 2 * ti + grp : ti) * num_clusters;
      const int sq = fdiv(st, a.fd_nn);
      const int n_tile = st - sq * a.num_n_tiles, m_tile = sq * cs + (int)crank;
      it.n0 = n_tile * a.block_n;
      it.c0 = sub * 32 + ch * cstride;
      it.width = cpw > 0 ? (a.block_n - it.c0 >= 32 ? 32 : 16) : 0;
      it.first = ch == 0;
      it.last = ch == per - 1;
      const long long row64 = (long long)m_tile * BLOCK_M + row_in_tile;
      it.row = (int)row64;
      bool valid = row64 < a.M;
      it.t1 = m_tile * BLOCK_M + quad * 32;
      it.t2 = it.t3 = 0;
      int hp = 0, hq = 0, hn = 0;
      if (a.halo) {   // 16x8 pixel patch: row i of the tile is pixel (ty*16 + i/8, tx*8 + i%8)
        const int per_img = a.tiles_x * a.tiles_y;
        hn = fdiv(m_tile, a.fd_pi);
        const int rem = m_tile - hn * per_img;
        const int ty = fdiv(rem, a.fd_tx), tx = rem - ty * a.tiles_x;
        hp = ty * kHaloTH + (row_in_tile >> 3);
        hq = tx * kHaloTW + (row_in_tile & 7);
        valid = m_tile < a.num_m_tiles && hp < a.Ho && hq < a.Wo;
        it.row = (hn * a.Ho + hp) * a.Wo + hq;
        it.t1 = tx * kHaloTW;
        it.t2 = ty * kHaloTH + quad * 4;
        it.t3 = hn;
      }
      it.gk_row = nullptr;
      if ((MODE == EPI_GATE || MODE == EPI_SILU_PRE) && valid) {
        int n_img, p, q;
        if (a.halo) {
          n_img = hn; p = hp; q = hq;
        } else {
          n_img = fdiv(it.row, a.fd_hw);
          const int rem = it.row - n_img * HoWo;
          p = fdiv(rem, a.fd_wo);
          q = rem - p * a.Wo;
        }
        const int hs = nearest_src(p, a.gHk, a.Ho, a.g_sh), ws = nearest_src(q, a.gWk, a.Wo, a.g_sw);
        const long long src_pix = ((long long)n_img * a.gHk + hs) * a.gWk + ws;
        it.gk_row = MODE == EPI_GATE ? a.gate_k + src_pix * a.ldgk
                                     : reinterpret_cast<const __nv_bfloat16*>(a.pre + src_pix * a.ldgk);
      }
      if (!valid) it.row = -1;
    };
    // residual / gate_x tile of `it` -> the warp's operand staging tile (lane 0 issues; 2048 bytes always land:
    // out-of-bounds elements are zero-filled)
    auto issue_opnd = [&](const Item& it) {
      if (lane == 0) {
        if (op_shared) tma_store_wait_read();   // the previous store of this warp still reads the shared tile
        mbar_expect_tx(my_opnd_bar, 2048u);
        if (a.halo) tma_load_4d(op_stage, &a.tmR, my_opnd_bar, it.n0 + it.c0, it.t1, it.t2, it.t3);
        else tma_load_2d(op_stage, &a.tmR, my_opnd_bar, it.n0 + it.c0, it.t1);
      }
    };

    Item cur, nxt;
    if (items > 0) {
      make_item(cur);
      if (kOpnd && cpw > 0) issue_opnd(cur);
    }
#pragma unroll 1
    for (int idx = 0; idx < items; ++idx) {
      const bool has_next = idx + 1 < items;
      if (has_next) {
        make_item(nxt);
        // the gate_k row of the next item: one line towards L1 now, so that its loads (an L2 round trip each,
        // 25 % of this kernel's warp samples in ncu) hit when the item is processed
        if (MODE == EPI_GATE && nxt.gk_row != nullptr)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt.gk_row + nxt.n0 + nxt.c0));
        if (MODE == EPI_SILU_PRE && nxt.gk_row != nullptr)   // the 128 bytes (32 fp32 columns) of the next item's partial sums
          asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float*>(nxt.gk_row) + nxt.n0 + nxt.c0));
      }
      if (cur.first) {
        mbar_wait(tfull_bar + acc * 8, acc_phase);
        tc_fence_after();
      }
      if (cur.width > 0) {
        const uint32_t taddr = tmem_base + (uint32_t)(acc * a.acc_stride) + ((uint32_t)(quad * 32) << 16) + cur.c0;
        const float* sc = s_scale + cur.n0 + cur.c0;
        const float* bi = s_bias + cur.n0 + cur.c0;
        uint32_t r[32];
        if (cur.width == 32) tmem_ld32(taddr, r);
        else tmem_ld16(taddr, r);
        // fp32 outputs (Detect heads) are staged too when the launch provides 4 KB tiles: a lane's direct 32-byte stores
        // land 1 KB apart (256 -> 255 @80: 40 % of the HBM rate), the 128-byte-swizzled tile leaves as one TMA store
        const bool f32_staged = MODE == EPI_LINEAR_F32 && a.stage_tile == 4096;
        if (kStaged || f32_staged) {
          uint4 own[4];
          if (kOpnd) {
            mbar_wait(my_opnd_bar, opnd_phase);
            opnd_phase ^= 1u;
#pragma unroll
            for (int j = 0; j < 4; ++j) own[j] = op_ptr[sidx(lane, j)];
          }
          tmem_ld_wait();
          if (cur.last) {   // the accumulators are in registers: hand the TMEM buffer back before the math and the store
            tc_fence_before();
            tmem_hand_back(tempty_bar + acc * 8);
          }
          if (lane == 0) tma_store_wait_read();   // the previous store of this warp has drained the staging tile
          __syncwarp();
          if (MODE == EPI_LINEAR_F32) {
            // 128-byte rows (32 fp32), TMA SWIZZLE_128B pattern: 16-byte chunk j of row r sits at chunk j ^ (r & 7)
            float4* o4 = reinterpret_cast<float4*>(out_ptr);
#pragma unroll
            for (int v8 = 0; v8 < 4; ++v8) {
              const float4 s0 = reinterpret_cast<const float4*>(sc + v8 * 8)[0], s1 = reinterpret_cast<const float4*>(sc + v8 * 8)[1];
              const float4 b0 = reinterpret_cast<const float4*>(bi + v8 * 8)[0], b1 = reinterpret_cast<const float4*>(bi + v8 * 8)[1];
              const uint32_t* rr = r + v8 * 8;
              o4[lane * 8 + ((v8 * 2) ^ (lane & 7))] =
                  make_float4(fmaf(__uint_as_float(rr[0]), s0.x, b0.x), fmaf(__uint_as_float(rr[1]), s0.y, b0.y),
                              fmaf(__uint_as_float(rr[2]), s0.z, b0.z), fmaf(__uint_as_float(rr[3]), s0.w, b0.w));
              o4[lane * 8 + ((v8 * 2 + 1) ^ (lane & 7))] =
                  make_float4(fmaf(__uint_as_float(rr[4]), s1.x, b1.x), fmaf(__uint_as_float(rr[5]), s1.y, b1.y),
                              fmaf(__uint_as_float(rr[6]), s1.z, b1.z), fmaf(__uint_as_float(rr[7]), s1.w, b1.w));
            }
          } else {
#pragma unroll
          for (int v8 = 0; v8 < 4; ++v8) {
            uint4 gk = make_uint4(0, 0, 0, 0);
            if (MODE == EPI_SILU_PRE) {   // acc += W0.x0 of the source pixel (two adjacent rows share it: one request)
              const int c = cur.n0 + cur.c0 + v8 * 8;
              if (cur.row >= 0 && v8 * 8 < cur.width && c + 8 <= a.n_store) {
                const float4* pp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(cur.gk_row) + c);
                const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
                uint32_t* rr = r + v8 * 8;
                rr[0] = __float_as_uint(__uint_as_float(rr[0]) + p0.x);
                rr[1] = __float_as_uint(__uint_as_float(rr[1]) + p0.y);
                rr[2] = __float_as_uint(__uint_as_float(rr[2]) + p0.z);
                rr[3] = __float_as_uint(__uint_as_float(rr[3]) + p0.w);
                rr[4] = __float_as_uint(__uint_as_float(rr[4]) + p1.x);
                rr[5] = __float_as_uint(__uint_as_float(rr[5]) + p1.y);
                rr[6] = __float_as_uint(__uint_as_float(rr[6]) + p1.z);
                rr[7] = __float_as_uint(__uint_as_float(rr[7]) + p1.w);
              }
            }
            if (MODE == EPI_GATE) {
              const int c = cur.n0 + cur.c0 + v8 * 8;
              if (cur.row >= 0 && v8 * 8 < cur.width && c + 8 <= a.n_store) gk = ld16(cur.gk_row + c);
            }
            out_ptr[sidx(lane, v8)] = epi_compute8<MODE>(r + v8 * 8, sc + v8 * 8, bi + v8 * 8, own[v8], gk);
          }
          }
          fence_proxy_async();
          __syncwarp();   // every lane's staging writes are done, and its operand row has been consumed (own[] fed the math)
          if (T9 && MODE == EPI_SILU && a.pool_out != nullptr && cur.width == 32) {
            // AvgPool2d(4) of this layer's output as a by-product (SCConv k2 reads it: its kernel re-read the whole map).  The
            // warp's 32 rows are a 4 x 8 pixel block of the 16 x 8 patch = two 4 x 4 windows; lane = channel of the 32-channel
            // chunk, summed from the bf16 staging tile in (dy, dx) order, then * 1/16: bit-identical to avgpool_kernel.
            const int cj = lane >> 3, ce = lane & 7;
            const __nv_bfloat16* st = reinterpret_cast<const __nv_bfloat16*>(out_ptr);
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float v = __bfloat162float(st[sidx(r, cj) * 8 + ce]);
              if ((r & 7) < 4) sa += v;
              else sb += v;
            }
            const int c = cur.n0 + cur.c0 + lane;
            if (c < a.n_store && cur.t2 + 3 < a.Ho) {
              __nv_bfloat16* po = a.pool_out + (((long long)cur.t3 * (a.Ho >> 2) + (cur.t2 >> 2)) * (a.Wo >> 2) + (cur.t1 >> 2)) * a.ldpool + c;
              if (cur.t1 + 3 < a.Wo) po[0] = __float2bfloat16_rn(sa * 0.0625f);
              if (cur.t1 + 7 < a.Wo) po[a.ldpool] = __float2bfloat16_rn(sb * 0.0625f);
            }
          }
          if (lane == 0) {
            if (a.halo) tma_store_4d(&a.tmY, out_stage, cur.n0 + cur.c0, cur.t1, cur.t2, cur.t3);
            else tma_store_2d(&a.tmY, out_stage, cur.n0 + cur.c0, cur.t1);
            tma_store_commit();
          }
          if (kOpnd && has_next) issue_opnd(nxt);   // refill the operand tile for the next item
        } else {
          tmem_ld_wait();
          if (cur.last) {
            tc_fence_before();
            tmem_hand_back(tempty_bar + acc * 8);
          }
          if (cur.row >= 0) {
            const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int v8 = 0; v8 < 4; ++v8) {
              const int col = cur.n0 + cur.c0 + v8 * 8;
              if (v8 * 8 < cur.width && col + 8 <= a.n_store)
                epi_store8<MODE>(a, r + v8 * 8, sc + v8 * 8, bi + v8 * 8, (long long)cur.row, col, z, z);
            }
          }
        }
      }
      if (cur.last && cur.width == 0) {   // a warp without columns in this tile still takes part in the hand-back
        tc_fence_before();
        tmem_hand_back(tempty_bar + acc * 8);
      }
      if (cur.last) {
        acc += a.epi_groups;        // next tile of this warp (nacc is even, so a group keeps its slot parity)
        if (acc >= a.nacc) {
          acc -= a.nacc;
          acc_phase ^= 1u;
        }
      }
      cur = nxt;
    }
    if ((kStaged || (MODE == EPI_LINEAR_F32 && a.stage_tile == 4096)) && lane == 0) tma_store_wait_all();   // staging tiles stay valid until every store has read them
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into it / arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)a.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map encoding through the driver entry points (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static bool load_driver_fns() {
  if (g_encode_tiled && g_encode_im2col) return true;
  void* f1 = nullptr;
  void* f2 = nullptr;
  cudaDriverEntryPointQueryResult q1, q2;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1) != cudaSuccess || !f1) return false;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2) != cudaSuccess || !f2) return false;
  cudaDriverGetVersion(&g_driver_version);
  g_encode_tiled = (EncodeTiledFn)f1;
  g_encode_im2col = (EncodeIm2colFn)f2;
  return true;
}

static void find_divisor(uint32_t (&fd)[2], int d) {   // q = umulhi(n, mul) >> shr for 0 <= n < 2^31 (CUTLASS FastDivmod)
  if (d <= 1) {
    fd[0] = 0;
    fd[1] = 0;
    return;
  }
  int lg = 0;
  while ((1ll << lg) < d) ++lg;
  const int p = 31 + lg;
  fd[0] = (uint32_t)(((1ull << p) + (unsigned)d - 1) / (unsigned)d);
  fd[1] = (uint32_t)(p - 32);
}

static int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// A launch plan: everything conv_plan derives from the parameter struct (mode / tile / pipeline decisions and the
// encoded CUtensorMaps).  Plans are immutable and cached per (device, parameter struct) — SURVEY.md 8b "immutable
// CUtensorMap caches keyed by pointer + shape" — so a steady-state step re-encodes nothing on the host.
struct ConvPlan {
  ConvArgs a;
  int mode;
  int t9;
  int grid;
  int pdl;
  size_t smem;
};

static int conv_plan(const dmay_conv_params* p, ConvPlan& pl) {
  if (!p || !p->x || !p->w || !p->scale || !p->bias || !p->y) return DMAY_EINVAL;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Cin <= 0 || p->Cout <= 0 || p->kh <= 0 || p->kw <= 0 || p->stride <= 0 ||
      p->pad < 0 || p->Ho <= 0 || p->Wo <= 0)
    return DMAY_EINVAL;
  if (!aligned16(p->x) || !aligned16(p->w) || !aligned16(p->y)) return DMAY_EINVAL;
  if (p->Cin & 15) return DMAY_EUNSUPPORTED;
  if (p->Cout_pad < p->Cout || (p->Cout_pad & 15) || p->Cout_pad > kMaxCout) return DMAY_EUNSUPPORTED;
  if (p->ldx < p->Cin - p->Cin1 - p->Cin2 || (p->ldx & 7) || (p->ldy & 3)) return DMAY_EUNSUPPORTED;
  const bool out_f32 = p->out_dtype == DMAY_DT_F32;
  if (p->out_dtype != DMAY_DT_F32 && p->out_dtype != DMAY_DT_BF16) return DMAY_EUNSUPPORTED;
  if (p->Cout & 7) return DMAY_EUNSUPPORTED;  // store width: whole 16-byte vectors (caller pads the slab)
  if (!out_f32 && (p->ldy & 7)) return DMAY_EUNSUPPORTED;
  if (p->ldy < p->Cout) return DMAY_EINVAL;
  if (p->residual && (!aligned16(p->residual) || (p->ldr & 7) || p->ldr < p->Cout)) return DMAY_EINVAL;
  if (p->gate_x && (!p->gate_k || !aligned16(p->gate_x) || !aligned16(p->gate_k) || (p->ldgx & 7) || p->gHk <= 0 || p->gWk <= 0))
    return DMAY_EINVAL;
  if (p->pre && (p->gate_x || p->residual || !aligned16(p->pre) || (p->ldpre & 3) || p->ldpre < p->Cout || p->preH <= 0 ||
                 p->preW <= 0 || p->act != DMAY_ACT_SILU || p->out_dtype != DMAY_DT_BF16))
    return DMAY_EINVAL;
  if ((p->Ho != (p->H + 2 * p->pad - p->kh) / p->stride + 1) || (p->Wo != (p->W + 2 * p->pad - p->kw) / p->stride + 1))
    return DMAY_EINVAL;
  if (p->pad > 127 || p->kh > 64 || p->kw > 64 || p->stride > 8) return DMAY_EUNSUPPORTED;
  // virtual channel concat: up to three sources over the same pixels (1x1 / stride 1 only), 64-channel granularity
  const int cin0 = p->Cin - p->Cin1 - p->Cin2;
  if (p->Cin1 < 0 || p->Cin2 < 0 || (p->Cin2 > 0 && p->Cin1 == 0)) return DMAY_EINVAL;
  if (p->Cin1 > 0) {
    if (!(p->kh == 1 && p->kw == 1 && p->stride == 1 && p->pad == 0)) return DMAY_EUNSUPPORTED;
    if (cin0 <= 0 || (cin0 & 63) || (p->Cin1 & 63) || (p->Cin2 & 63)) return DMAY_EUNSUPPORTED;
    if (!p->x1 || !aligned16(p->x1) || p->ldx1 < p->Cin1 || (p->ldx1 & 7) || p->ldx < cin0) return DMAY_EINVAL;
    if (p->Cin2 > 0 && (!p->x2 || !aligned16(p->x2) || p->ldx2 < p->Cin2 || (p->ldx2 & 7))) return DMAY_EINVAL;
  }
  if (!load_driver_fns()) return DMAY_EDRIVER;

  ConvArgs& a = pl.a;
  memset(&a, 0, sizeof(a));
  const long long M = (long long)p->N * p->Ho * p->Wo;
  if (M > 0x7fffff00LL) return DMAY_EUNSUPPORTED;
  a.M = (int)M;
  a.CK = (p->Cin % 64 == 0) ? 64 : (p->Cin % 32 == 0) ? 32 : 16;
  const CUtensorMapSwizzle swz = a.CK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (a.CK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  a.layout_type = a.CK == 64 ? 2u : (a.CK == 32 ? 4u : 6u);
  a.sbo_enc = (uint32_t)(8 * a.CK * 2) >> 4;  // 8 rows of CK bf16
  a.Cin = p->Cin;
  a.c_chunks = p->Cin / a.CK;
  a.taps = p->kh * p->kw;
  a.kw = p->kw;
  a.conv_stride = p->stride;
  a.pad = p->pad;
  a.Ho = p->Ho;
  a.Wo = p->Wo;
  a.im2col = !(p->kh == 1 && p->kw == 1 && p->stride == 1 && p->pad == 0);
  int bn = p->block_n > 0 ? p->block_n : (p->Cout_pad < 256 ? p->Cout_pad : 256);
  int mode;
  if (p->pre) mode = EPI_SILU_PRE;
  else if (p->gate_x) mode = (!out_f32 && !p->residual) ? EPI_GATE : -1;
  else if (out_f32) mode = (p->act == DMAY_ACT_NONE && !p->residual) ? EPI_LINEAR_F32 : EPI_GENERIC;
  else if (p->act == DMAY_ACT_SILU) mode = p->residual ? EPI_SILU_RES : EPI_SILU;
  else if (p->res_op == 1) mode = (p->act == DMAY_ACT_NONE && p->residual) ? EPI_LINEAR_MUL : -1;
  else if (p->act == DMAY_ACT_NONE) mode = p->residual ? EPI_LINEAR_RES : EPI_LINEAR;
  else if (p->act == DMAY_ACT_GELU && !p->residual) mode = EPI_GELU;
  else mode = EPI_GENERIC;
  if (mode < 0) return DMAY_EUNSUPPORTED;
  const bool staged = mode == EPI_SILU || mode == EPI_SILU_RES || mode == EPI_LINEAR || mode == EPI_GATE ||
                      mode == EPI_LINEAR_RES || mode == EPI_GELU || mode == EPI_LINEAR_MUL || mode == EPI_SILU_PRE;
  a.opnd_stage = (mode == EPI_SILU_RES || mode == EPI_GATE || mode == EPI_LINEAR_RES || mode == EPI_LINEAR_MUL) ? 1 : 0;
  a.sb_floats = ((p->Cout_pad + bn - 1) / bn) * bn;
  if (a.sb_floats > kMaxCout) return DMAY_EUNSUPPORTED;
  // epilogue warps, measured (profiles/r1_conv_notes.md): 16 warps win only where a tile's MMA is a few hundred
  // clocks (1x1 convs with K <= 256: +10..30 %); from K = 576 up 8 warps are as fast or faster, and their smaller
  // staging area leaves one more pipeline stage
  a.epi_warps = ((long long)p->kh * p->kw * p->Cin <= 384) ? 16 : 8;
  if (bn <= 64 && !(p->flags & 32)) a.epi_warps = 16;   // narrow tiles: two groups of 8 warps on alternate tiles
  if (p->flags & 8) a.epi_warps = 8;
  if (p->flags & 16) a.epi_warps = 16;
  a.epi_groups = (a.epi_warps == 16 && bn <= 64 && !(p->flags & 32)) ? 2 : 1;
  // fp32 output of a 1x1 conv (Detect heads): staged 4 KB tiles + TMA store (flags bit14 = keep the direct stores)
  const bool f32_staged = mode == EPI_LINEAR_F32 && p->kh == 1 && p->kw == 1 && p->stride == 1 && !(p->flags & 16384) &&
                          (p->ldy % 4) == 0 && aligned16(p->y);
  a.stage_tile = f32_staged ? 4096 : 2048;
  if ((p->flags & 8192) && a.opnd_stage && staged && bn > 64) {   // A/B: shared operand tile + 16 warps for every staged operand
    a.epi_warps = 16;
    a.opnd_stage = 2;
  }
  uint32_t kTailBytes = tail_bytes(a.epi_warps, a.opnd_stage == 1, (uint32_t)a.sb_floats, (uint32_t)a.stage_tile);

  if ((bn & 15) || bn > 256 || bn < 16) return DMAY_EUNSUPPORTED;
  a.block_n = bn;
  a.acc_stride = pow2ceil(bn) < 32 ? 32 : pow2ceil(bn);
  a.nacc = 512 / a.acc_stride;
  if (a.nacc > kMaxAcc) a.nacc = kMaxAcc;
  if (p->flags & 64) a.nacc = 2;
  a.tmem_cols = a.nacc * a.acc_stride;
  a.num_m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M);
  a.num_n_tiles = (p->Cout_pad + bn - 1) / bn;
  a.n_store = p->Cout;
  a.Cout_pad = p->Cout_pad;
  // ---- CTA-pair mode (cta_group::2): 256 x block_n tiles, each CTA of a 2-CTA cluster loads its own 128 rows of A and
  // HALF of the weight tile.  For the deep-K layers a 128 x 256 tile is L2->SM bound (85 flop per byte fetched,
  // profiles/r1_conv_notes.md section 1); the pair fetches (128 + 128) x K per CTA for the same flops: 128 flop/B.
  const int sms_q = p->num_sms > 0 ? p->num_sms : sm_count();
  const long long m_tiles_q = (M + BLOCK_M - 1) / BLOCK_M;
  {
    const int tx_ = (p->Wo + kHaloTW - 1) / kHaloTW, ty_ = (p->Ho + kHaloTH - 1) / kHaloTH;
    const double eff_ = (double)p->Ho * p->Wo / ((double)tx_ * kHaloTW * ty_ * kHaloTH);
    const bool halo_auto = p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && !(p->flags & 1) &&
                           ((p->flags & 2) || (eff_ >= kHaloMinEff && (long long)p->N * tx_ * ty_ >= 2LL * sms_q));
    const bool legal = p->block_n != -2 && a.CK == 64 && bn >= 64 && (bn % 32) == 0 && m_tiles_q >= 2 && sms_q >= 2 &&
                       mode != EPI_GENERIC;
    // halo + pair (128-channel 3x3 layers): per CTA one input patch + HALF of the 9-tap weight set per chunk
    // (measured: 64->128 3x3 s2 @160, K = 576: 0.390 -> 0.334 ms; 1x1 layers with K <= 512 lose, see r1_conv_notes.md)
    const long long Ktot = (long long)a.taps * p->Cin;
    const bool want = m_tiles_q >= sms_q && ((Ktot >= 1024 && (bn == 256 || (halo_auto && bn >= 128))) ||
                                             (a.taps >= 9 && Ktot >= 512 && bn >= 128));
    a.pair = legal && ((p->flags & 256) || (want && !(p->flags & 512))) ? 1 : 0;
  }
  a.a_bytes = (uint32_t)(BLOCK_M * a.CK * 2);
  a.b_bytes = (uint32_t)((a.pair ? bn / 2 : bn) * a.CK * 2);
  a.total_subs = a.taps * a.c_chunks;
  // several sub-tiles per stage when they are small (one mbarrier round trip per stage, not per 16/32-channel tap)
  int subs = (int)((56u * 1024u) / (a.a_bytes + a.b_bytes));
  if (subs < 1) subs = 1;
  if (subs > a.total_subs) subs = a.total_subs;
  if (subs > 16) subs = 16;
  a.subs = subs;
  // B multicast over a 2-CTA cluster whenever there are at least two m-tiles and the slice keeps its alignment
  // Measured on B200 (profiles/r1_conv_notes.md): the 2-CTA multicast is bit-correct but 20-30 % SLOWER than
  // independent CTAs (cluster lock-step stalls; L2 already de-duplicates concurrent requests for the same
  // weight tile), so it is opt-in: block_n == -2 requests it.
  int cs = (p->block_n == -2) ? 2 : 1;
  if ((bn % (8 * cs)) || (((bn / cs) * a.CK * 2) % 1024 && a.CK == 64) || m_tiles_q < 2 || sms_q < 2) cs = 1;
  if (a.CK != 64 && (((bn / cs) * a.CK * 2) % (a.CK == 32 ? 512 : 256))) cs = 1;
  if (a.pair) cs = 2;
  a.cs = cs;
  // ---- halo mode decision: 3x3 / stride 1 / pad 1 on maps that 16x8 patches tile well, enough tiles to fill the chip
  const int tiles_x = (p->Wo + kHaloTW - 1) / kHaloTW, tiles_y = (p->Ho + kHaloTH - 1) / kHaloTH;
  const double patch_eff = (double)p->Ho * p->Wo / ((double)tiles_x * kHaloTW * tiles_y * kHaloTH);
  bool halo = p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && !(p->flags & 1) && (cs == 1 || a.pair);
  if (halo && !(p->flags & 2)) {
    // every 3x3 s1 layer is L2->SM bound with im2col operands (nine times the input bytes per tile); the patch form wins
    // as long as the 16x8 patches cover the map reasonably.  Measured (r3u, batch 64): 256->256 @80 1344 -> 1675 TF/s,
    // 256->256 @40 (83 % cover) 1167 -> 1276, 512->512 @40 1382 -> 1500; 20x20 maps (52 % cover) lose: 1199 -> 825.
    halo = patch_eff >= kHaloMinEff && (long long)p->N * tiles_x * tiles_y >= 2LL * sms_q;
  }
  uint32_t halo_bytes_total = 0;
  if (halo) {
    a.halo = 1;
    a.halo_pitch = kHaloTW + 2;
    a.tiles_x = tiles_x;
    a.tiles_y = tiles_y;
    a.num_m_tiles = p->N * tiles_x * tiles_y;
    a.a_halo_tx = (uint32_t)(a.halo_pitch * (kHaloTH + 2) * a.CK * 2);
    a.a_halo_bytes = (a.a_halo_tx + 1023u) & ~1023u;
    a.halo_sbo_enc = (uint32_t)(a.halo_pitch * a.CK * 2) >> 4;
    const uint32_t smem_avail = 227u * 1024u - 1024u - kTailBytes - 64u;
    const uint32_t w_bytes = (uint32_t)a.total_subs * a.b_bytes;           // all taps x chunks of this n-tile
    // resident weights: the whole n-tile's taps x chunks (pair mode: this CTA's half) next to >= 2 input patches
    const bool res_fits = a.pair ? (uint64_t)((w_bytes + 1023u) & ~1023u) + 2ull * a.a_halo_bytes <= smem_avail && !(p->flags & 4096)
                                 : w_bytes <= 100u * 1024u;
    if (a.num_n_tiles == 1 && res_fits && !(p->flags & 4)) {
      a.b_resident = 1;
      a.bres_bytes = (w_bytes + 1023u) & ~1023u;
      uint32_t avail_r = smem_avail;
      // pair + resident: with the weights on chip the tile's MMA time no longer covers two items per epilogue warp
      // (residual / gate modes: 1072 / 949 vs 1426 TF/s plain at 128->128 @80), so run 16 epilogue warps on 16 single
      // staging tiles (operand tile shared with the output tile, see the kernel) whenever that still leaves two patches
      if (a.pair && staged && a.epi_warps == 8 && !(p->flags & 8)) {
        const uint32_t tail16 = tail_bytes(16, false, (uint32_t)a.sb_floats);
        const uint32_t avail16 = 227u * 1024u - 1024u - tail16 - 64u;
        if ((uint64_t)a.bres_bytes + 2ull * a.a_halo_bytes <= avail16) {
          a.epi_warps = 16;
          if (a.opnd_stage) a.opnd_stage = 2;
          kTailBytes = tail16;
          avail_r = avail16;
        }
      }
      int nb = (int)((avail_r - a.bres_bytes) / a.a_halo_bytes);
      a.n_abuf = nb > kMaxABuf ? kMaxABuf : nb;
      if (a.n_abuf < 2) return DMAY_EUNSUPPORTED;
      subs = 1;
      a.subs = 1;
      a.stage_bytes = 1024u;   // no B pipeline
    } else {
      // stages hold B tiles only; sub-tiles per stage must divide the 9 taps (a stage never straddles a chunk).
      // Patch ring of 4 (3, 2 when shared memory is short) and the widest stage that still leaves >= 3 stages.
      // Measured: one mbarrier round trip per tap (subs = 1) costs 40 % on the 128-channel layers, so the widest
      // stage wins even when only two stages and a shorter patch ring fit.
      bool ok = false;
      for (int sb : {9, 3, 1}) {
        const uint32_t stg = ((uint32_t)sb * a.b_bytes + 1023u) & ~1023u;
        if (stg > 48u * 1024u && sb > 1) continue;
        // two input patches are enough (one per ~2.3 k clk of MMAs); every further KB goes to the weight stages, which
        // are what the L2->SM-bound streamed layers wait for (r4y: 256->256 @40 residual 1163 -> 1295 TF/s, @80 plain
        // 1665 -> 1809, 512->512 @40 plain 1500 -> 1610).  flags bit15 = the old preference (up to four patches).
        const int nab_hi = (p->flags & 32768) ? 4 : 2;
        for (int nab = nab_hi; nab >= 2 && !ok; --nab)
          if ((uint64_t)nab * a.a_halo_bytes + 2ull * stg <= smem_avail) {
            a.n_abuf = nab;
            subs = sb;
            a.stage_bytes = stg;
            ok = true;
          }
        if (ok) break;
      }
      if (!ok) return DMAY_EUNSUPPORTED;
      a.subs = subs;
    }
    halo_bytes_total = a.bres_bytes + (uint32_t)a.n_abuf * a.a_halo_bytes;
  } else {
    const uint32_t w_bytes = (uint32_t)a.total_subs * a.b_bytes;
    const uint32_t avail = 227u * 1024u - 1024u - kTailBytes - 64u;
    if (!a.pair && a.cs == 1 && a.num_n_tiles == 1 && w_bytes <= 132u * 1024u && !(p->flags & 1024) &&
        (p->flags & 2048) && avail >   // opt-in: measured no faster than streaming (profiles/r1_conv_notes.md section 4)
        ((w_bytes + 1023u) & ~1023u) + 2u * a.a_bytes) {
      a.b_resident = 2;
      a.bres_bytes = (w_bytes + 1023u) & ~1023u;
      const uint32_t left = avail - a.bres_bytes;
      subs = (int)(left / 3u / a.a_bytes);            // aim for >= 3 stages
      if (subs < 1) subs = 1;
      if (subs > a.total_subs) subs = a.total_subs;
      if (subs > 4) subs = 4;
      a.subs = subs;
      a.stage_bytes = ((uint32_t)subs * a.a_bytes + 1023u) & ~1023u;
      halo_bytes_total = a.bres_bytes;
    } else {
      a.stage_bytes = ((uint32_t)subs * (a.a_bytes + a.b_bytes) + 1023u) & ~1023u;
    }
  }
  const uint32_t budget = 227u * 1024u - 1024u - kTailBytes - 64u - halo_bytes_total;
  int stages = (int)(budget / a.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (a.b_resident == 1) stages = 1;
  else if (stages < 2) return DMAY_EUNSUPPORTED;
  a.stages = stages;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=BF16 [7,10), b=BF16 [10,13),
  // both K-major, N>>3 at [17,23), M>>4 at [24,29)
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) |
            ((uint32_t)((a.pair ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);
  a.scale = (const float*)p->scale;
  a.bias = (const float*)p->bias;
  a.residual = (const __nv_bfloat16*)p->residual;
  a.gate_x = (const __nv_bfloat16*)p->gate_x;
  a.gate_k = (const __nv_bfloat16*)p->gate_k;
  a.pre = (const float*)p->pre;
  a.y = p->y;
  a.ldy = p->ldy;
  a.ldr = p->ldr;
  a.ldgx = p->ldgx;
  a.ldgk = p->Cout;  // k2 output is a dense [N,Hk,Wk,C] tensor
  a.gHk = p->gHk;
  a.gWk = p->gWk;
  if (p->pre) {   // the partial sums use the gate_k geometry fields of the kernel arguments
    a.ldgk = p->ldpre;
    a.gHk = p->preH;
    a.gWk = p->preW;
  }
  a.g_sh = a.gHk > 0 ? (float)a.gHk / (float)p->Ho : 0.f;
  a.g_sw = a.gWk > 0 ? (float)a.gWk / (float)p->Wo : 0.f;
  a.act = p->act;
  a.flags = p->flags;
  a.out_f32 = out_f32 ? 1 : 0;
  find_divisor(a.fd_nn, a.num_n_tiles);
  find_divisor(a.fd_pi, a.halo ? a.tiles_x * a.tiles_y : 1);
  find_divisor(a.fd_tx, a.halo ? a.tiles_x : 1);
  find_divisor(a.fd_hw, p->Ho * p->Wo);
  find_divisor(a.fd_wo, p->Wo);

  // ---- tensor maps ----
  CUresult r;
  if (a.halo) {
    cuuint64_t gdim[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    cuuint64_t gstr[3] = {(cuuint64_t)p->ldx * 2, (cuuint64_t)p->W * p->ldx * 2, (cuuint64_t)p->H * p->W * p->ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)a.CK, (cuuint32_t)a.halo_pitch, (cuuint32_t)(kHaloTH + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = g_encode_tiled(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->x), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DMAY_EDRIVER;
  } else if (a.im2col) {
    cuuint64_t gdim[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    cuuint64_t gstr[3] = {(cuuint64_t)p->ldx * 2, (cuuint64_t)p->W * p->ldx * 2, (cuuint64_t)p->H * p->W * p->ldx * 2};
    int lower[2] = {-p->pad, -p->pad};
    int upper[2] = {p->pad - (p->kw - 1), p->pad - (p->kh - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)p->stride, (cuuint32_t)p->stride, 1};
    r = g_encode_im2col(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->x), gdim, gstr, lower, upper,
                        (cuuint32_t)a.CK, (cuuint32_t)BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DMAY_EDRIVER;
    // Same driver work-around CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): drivers <= 13.1
    // set a descriptor bit that breaks im2col loads from tensors smaller than 128 KiB.
    const unsigned long long bytes = (unsigned long long)p->N * p->H * p->W * p->ldx * 2ull;
    if (g_driver_version <= 13010 && bytes < 131072ull) reinterpret_cast<uint64_t*>(&a.tmA)[1] &= ~(1ull << 21);
  } else {
    auto encode_flat = [&](CUtensorMap* tm, const void* basep, int c, int ld) -> bool {
      cuuint64_t gdim[2] = {(cuuint64_t)c, (cuuint64_t)M};
      cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
      cuuint32_t box[2] = {(cuuint32_t)a.CK, (cuuint32_t)BLOCK_M};
      cuuint32_t estr[2] = {1, 1};
      return g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(basep), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!encode_flat(&a.tmA, p->x, cin0, p->ldx)) return DMAY_EDRIVER;
    if (p->Cin1 > 0) {
      a.seg1 = cin0 / a.CK;
      a.seg2 = (cin0 + p->Cin1) / a.CK;   // == c_chunks when there is no third source
      if (!encode_flat(&a.tmA2, p->x1, p->Cin1, p->ldx1)) return DMAY_EDRIVER;
      if (p->Cin2 > 0 && !encode_flat(&a.tmA3, p->x2, p->Cin2, p->ldx2)) return DMAY_EDRIVER;
    }
  }
  {
    const cuuint64_t ktot = (cuuint64_t)a.taps * p->Cin;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)p->Cout_pad};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)a.CK, (cuuint32_t)(bn / a.cs)};
    cuuint32_t estr[2] = {1, 1};
    r = g_encode_tiled(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->w), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DMAY_EDRIVER;
  }

  if (staged) {
    // output tile store / operand tile load: 32 channels x 32 rows (halo: 32 ch x 8 x 4 pixels), 64-byte swizzle.
    // The maps are clipped to [Cout] channels, so a slab neighbour is never touched, and to M rows / the image.
    auto encode_epi = [&](CUtensorMap* tm, const void* basep, int ld) -> bool {
      CUresult rr;
      if (a.halo) {
        cuuint64_t gdim[4] = {(cuuint64_t)p->Cout, (cuuint64_t)p->Wo, (cuuint64_t)p->Ho, (cuuint64_t)p->N};
        cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)p->Wo * ld * 2, (cuuint64_t)p->Ho * p->Wo * ld * 2};
        cuuint32_t box[4] = {32, (cuuint32_t)kHaloTW, 4, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        rr = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(basep), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint64_t gdim[2] = {(cuuint64_t)p->Cout, (cuuint64_t)M};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
        cuuint32_t box[2] = {32, 32};
        cuuint32_t estr[2] = {1, 1};
        rr = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(basep), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      return rr == CUDA_SUCCESS;
    };
    if (!encode_epi(&a.tmY, p->y, p->ldy)) return DMAY_EDRIVER;
    if (a.opnd_stage) {
      const void* src = mode == EPI_GATE ? p->gate_x : p->residual;
      const int ld = mode == EPI_GATE ? p->ldgx : p->ldr;
      if (!encode_epi(&a.tmR, src, ld)) return DMAY_EDRIVER;
    }
  } else if (f32_staged) {
    cuuint64_t gdim[2] = {(cuuint64_t)p->Cout, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)p->ldy * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    r = g_encode_tiled(&a.tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p->y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DMAY_EDRIVER;
  } else {
    a.tmY = a.tmB;   // never used by the direct-store modes; keeps the descriptor prefetch well-defined
  }

  const size_t smem = 1024 + (size_t)halo_bytes_total + (size_t)a.stages * a.stage_bytes + kTailBytes;
  const int sms = p->num_sms > 0 ? p->num_sms : sm_count();
  const long long supers = (long long)((a.num_m_tiles + a.cs - 1) / a.cs) * a.num_n_tiles;
  const long long max_clusters = sms / a.cs;
  pl.grid = (int)((supers < max_clusters ? supers : max_clusters) * a.cs);
  pl.mode = mode;
  pl.t9 = (mode == EPI_SILU && a.b_resident == 1 && a.halo && !(p->flags & 65536)) ? 1 : 0;
  if (p->pool4_out) {   // only the T9 images carry the pooling epilogue; the caller falls back to dmay_avgpool otherwise
    if (!pl.t9 || (p->Ho & 3) || (p->Wo & 3) || (p->ldpool4 & 7) || p->ldpool4 < p->Cout || (bn & 31) || out_f32) return DMAY_EUNSUPPORTED;
    a.pool_out = (__nv_bfloat16*)p->pool4_out;
    a.ldpool = p->ldpool4;
  }
  pl.smem = smem;
  pl.pdl = (p->flags & 128) ? 0 : 1;
  return DMAY_OK;
}

static int conv_issue(const ConvPlan& pl, cudaStream_t stream) {
  const ConvArgs& a = pl.a;
  const int grid = pl.grid;
  const size_t smem = pl.smem;
#define DMAY_LAUNCH_MODE(MODE)                                                                                     \
  case MODE: {                                                                                                      \
    static std::atomic<unsigned long long> attr_mask{0};                                                            \
    if (first_time_on_device(attr_mask)) {                                                                          \
      cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           227 * 1024);                                                            \
      if (e != cudaSuccess) return (int)e;                                                                          \
      e = cudaFuncSetAttribute(conv_gemm_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return (int)e;                                                                          \
    }                                                                                                               \
    cudaLaunchConfig_t cfg = {};                                                                                    \
    cfg.gridDim = dim3(grid);                                                                                       \
    cfg.blockDim = dim3(128 + 32 * a.epi_warps);                                                                    \
    cfg.dynamicSmemBytes = smem;                                                                                    \
    cfg.stream = stream;                                                                                            \
    cudaLaunchAttribute at[2];                                                                                      \
    int nat = 0;                                                                                                    \
    if (pl.pdl) {                                                                                                   \
      at[nat].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                              \
      at[nat].val.programmaticStreamSerializationAllowed = 1;                                                       \
      ++nat;                                                                                                        \
    }                                                                                                               \
    if (a.cs > 1) {                                                                                                 \
      at[nat].id = cudaLaunchAttributeClusterDimension;                                                             \
      at[nat].val.clusterDim.x = a.cs;                                                                              \
      at[nat].val.clusterDim.y = 1;                                                                                 \
      at[nat].val.clusterDim.z = 1;                                                                                 \
      ++nat;                                                                                                        \
    }                                                                                                               \
    cfg.attrs = at;                                                                                                 \
    cfg.numAttrs = nat;                                                                                             \
    cudaError_t e = a.pair ? cudaLaunchKernelEx(&cfg, conv_gemm_kernel<MODE, true>, a)                              \
                           : cudaLaunchKernelEx(&cfg, conv_gemm_kernel<MODE, false>, a);                            \
    if (e != cudaSuccess) return (int)e;                                                                            \
    break;                                                                                                          \
  }
  if (pl.t9) {   // plain SiLU, resident 3x3 weights: the image with the one-block tap issue
    static std::atomic<unsigned long long> attr_mask9{0};
    if (first_time_on_device(attr_mask9)) {
      cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_SILU, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(conv_gemm_kernel<EPI_SILU, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128 + 32 * a.epi_warps);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int nat = 0;
    if (pl.pdl) {
      at[nat].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[nat].val.programmaticStreamSerializationAllowed = 1;
      ++nat;
    }
    if (a.cs > 1) {
      at[nat].id = cudaLaunchAttributeClusterDimension;
      at[nat].val.clusterDim.x = a.cs;
      at[nat].val.clusterDim.y = 1;
      at[nat].val.clusterDim.z = 1;
      ++nat;
    }
    cfg.attrs = at;
    cfg.numAttrs = nat;
    cudaError_t e = a.pair ? cudaLaunchKernelEx(&cfg, conv_gemm_kernel<EPI_SILU, true, true>, a)
                           : cudaLaunchKernelEx(&cfg, conv_gemm_kernel<EPI_SILU, false, true>, a);
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
  }
  switch (pl.mode) {
    DMAY_LAUNCH_MODE(EPI_SILU)
    DMAY_LAUNCH_MODE(EPI_SILU_RES)
    DMAY_LAUNCH_MODE(EPI_LINEAR)
    DMAY_LAUNCH_MODE(EPI_GATE)
    DMAY_LAUNCH_MODE(EPI_LINEAR_F32)
    DMAY_LAUNCH_MODE(EPI_GENERIC)
    DMAY_LAUNCH_MODE(EPI_LINEAR_RES)
    DMAY_LAUNCH_MODE(EPI_GELU)
    DMAY_LAUNCH_MODE(EPI_LINEAR_MUL)
    DMAY_LAUNCH_MODE(EPI_SILU_PRE)
  }
#undef DMAY_LAUNCH_MODE
  return finish_launch();
}

// ---- plan cache -----------------------------------------------------------------------------------------------------
// Key = the raw bytes of the parameter struct + the device ordinal.  The caching allocator of the host framework hands the
// same addresses back step after step, so after the first step of a (batch, H, W) every launch is a cache hit.  A plan
// holds no reference to device memory (tensor maps carry addresses, not ownership); a stale entry is merely never hit again.
struct PlanKey {
  dmay_conv_params p;
  int dev;
  int pad_;
};
struct PlanKeyHash {
  size_t operator()(const PlanKey& k) const {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(PlanKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
struct PlanKeyEq {
  bool operator()(const PlanKey& a, const PlanKey& b) const { return memcmp(&a, &b, sizeof(PlanKey)) == 0; }
};
static std::mutex g_plan_mu;
static std::unordered_map<PlanKey, std::unique_ptr<ConvPlan>, PlanKeyHash, PlanKeyEq> g_plans;
static std::atomic<long long> g_plan_hits{0}, g_plan_misses{0};
static const size_t kMaxPlans = 16384;

static int conv_launch(const dmay_conv_params* p, cudaStream_t stream) {
  if (!p) return DMAY_EINVAL;
  PlanKey key;
  memset(&key, 0, sizeof(key));
  key.p = *p;
  if (cudaGetDevice(&key.dev) != cudaSuccess) key.dev = -1;
  const ConvPlan* pl = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) pl = it->second.get();
  }
  if (pl) {
    g_plan_hits.fetch_add(1, std::memory_order_relaxed);
    return conv_issue(*pl, stream);
  }
  std::unique_ptr<ConvPlan> np(new ConvPlan);
  const int rc = conv_plan(p, *np);
  if (rc != DMAY_OK) return rc;
  g_plan_misses.fetch_add(1, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (g_plans.size() < kMaxPlans) {   // entries are never erased, so a pointer obtained under the lock stays valid
      auto ins = g_plans.emplace(key, std::move(np));
      pl = ins.first->second.get();
    }
  }
  return conv_issue(pl ? *pl : *np, stream);
}

}  // namespace dmay

extern "C" int dmay_conv_bn_act(const dmay_conv_params* p, dmay_stream_t stream) {
  return dmay::conv_launch(p, (cudaStream_t)stream);
}

extern "C" long long dmay_conv_plan_stats(int what) {
  return what == 0 ? dmay::g_plan_hits.load() : what == 1 ? dmay::g_plan_misses.load() : (long long)dmay::g_plans.size();
}
