// tcgen05 / TMEM implicit-GEMM 1-D convolution for sm_100a.
//
//   D[t, co] = sum_{tap, ci} A_tap[t, ci] * W_tap[co, ci]      (M = 128 positions, N = co tile, K = ci)
//
// One persistent CTA per SM (640 threads, 96 registers each), warp-specialised; every CTA owns a contiguous range of
// work items (sample-major) of 1 or 2 tiles of 128 positions:
// * TMA warp: raw fp32 activation boxes [16 channels x (128 + halo)] -> staging ring (zero fill outside the sequence).
// * Weight warp: the operand image pre-split into bf16 hi/lo and pre-arranged by vqvs_pack_conv_weights, resident in
//   shared memory for the whole CTA when it fits (<= 150 KB), else streamed per K block with cp.async.bulk.
// * Transform warps (8 + 1 for the halo rows): optional GroupNorm(+FiLM) FINALIZE from the producers' statistics at every
//   sample change, then per element affine -> erf-form GELU (packed fp32x2 approximation, <= 6.4e-7) -> pool / upsample /
//   concat select -> bf16 hi + lo split (or fp16), stored K-major, un-swizzled, as [chunk of 8 channels][row = position][16 B].  Rows are 16 B apart, so
//   the three conv taps are the SAME tile read through descriptors whose start address is shifted by tap*dilation rows --
//   the halo is staged once.
// * MMA warp: warp-uniform loop, elected lane issues tcgen05.mma (kind::f16 -> fp32 in TMEM).  Two operand formats per conv:
//   "bf16x3" = hi*hi + hi*lo + lo*hi (~2^-16 relative operand error, measured 1.5e-5 on a whole UNet forward vs 9.7e-4 for
//   single TF32 -- tools/precision_study.py), the second product reusing the A tile from the collector; for 32-channel N
//   tiles the weight rows are stacked [W_hi ; W_lo] so that two MMAs per tap give all four products; "fp16" = one product
//   (the deep levels, engine.conv_precision).  Accumulators are double-buffered in TMEM.
// * Epilogue warps (8): thread = time row.  tcgen05.ld -> + bias (+ lo half) (+ identity skip, prefetched) -> coalesced
//   stores along time -> GroupNorm (sum, sumsq) of the OUTPUT accumulated in registers across the CTA's tiles, one
//   accumulator per G-channel granule, reduced across lanes and flushed with fp64 atomics only when the sample changes.
// The optional 1x1 skip conv (reference models/unet.py:265-271) is extra K blocks over the raw input.
// The kernel exists in four KINDs (generic / LEAN / PLAIN / SIMPLE, see conv_umma_kernel) compiled as separate
// translation units; measurements behind the design choices: tools/mma_bench.cu, tools/alu_bench.cu, profiles/.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace vqvs {
namespace umma {

constexpr int TILE_M = 128;
constexpr int KBLK = 16;  // channels per MMA K block (kind::f16: K = 16)
// warp roles of the persistent CTA (one CTA per SM)
// (epilogue first: low warp ids; warpgroups 0-1 = epilogue, 2-4 = MMA / TMA / transform)
constexpr int EPI_WARP0 = 0;      // warps 0..7 <-> TMEM lane quarters (warp & 3)
constexpr int MMA_WARP = 8;       // single-thread tcgen05.mma issue
constexpr int TMA_W_WARP = 9;     // weight image (resident or streamed per K block)
constexpr int TMA_RAW_WARP = 10;  // raw fp32 activation boxes -> staging ring
constexpr int XFORM_WARP0 = 11;   // warps 11..19 transform: warp 11 the halo rows, warps 12..19 the 128 main rows
constexpr int XFORM_WARPS = 9;
constexpr int HALO_WARP = XFORM_WARP0;  // shares warpgroup 2 (small register budget) with the MMA / TMA warps
constexpr int EPI_WARPS = 8;      // two warps per lane quarter split the column chunks
constexpr int EPI_SPLIT = EPI_WARPS / 4;  // warps sharing a TMEM lane quarter take alternate 32-column chunks
constexpr int THREADS = (XFORM_WARP0 + XFORM_WARPS) * 32;
constexpr int SMEM_HEADER = 1024;  // 53 mbarriers + TMEM base holder (first 512 B), per-group (mean, rstd) of the fused GroupNorm finalize (last 512 B)
constexpr int SMEM_GROUP_TABLE = 512;  // byte offset of the 32 x (mean, rstd) fp64 pairs
constexpr int MAX_RAW_SLOTS = 8;
constexpr int MAX_B_SLOTS = 8;
constexpr int MAX_AB_SLOTS = 8;
constexpr int SIMPLE_BOXW = 136;  // staging pitch of every un-resampled conv with dilation 1 or 2 (128 + 2 * 4)

// Host-computed geometry shared by the packer and the kernel.
struct Geo {
  int n_tiles, n_tile;      // output-channel tiling (n_tile <= 256, multiple of 16)
  int nkb_main, nkb_skip;   // K blocks (16 channels) of the main taps / of the 1x1 skip
  int pad, rows;            // halo and operand rows (= 128 + 2*pad)
  int acc_cols, tmem_cols;  // TMEM columns of one accumulator / allocated (two accumulators)
  int epi_fast, epi_nch, epi_na;  // register-statistics epilogue: eligible / 32-column chunks per warp / accumulators per chunk
  int epi_w16;                    // 32-channel N tile: the two warps of a TMEM lane quarter take 16 columns each (per-channel statistics)
  int prec;                 // operand format: VQVS_PREC_BF16X3 (hi/lo split, three products) or VQVS_PREC_F16 (one fp16 product)
  int stack;                // 1: weight rows are [W_hi ; W_lo] (N = 2*n_tile): two MMAs per tap give all four
                            //    hi/lo products, the epilogue adds the two column halves
  int a_kb_bytes;           // operand bytes per K block: hi+lo, 2 chunks, `rows` rows of 16 B
  int b_unit_main;          // weight bytes per main K block: ksize taps x hi/lo x 2 chunks x n_tile rows x 16 B
  int b_unit_skip;          // weight bytes per skip K block
  long long per_tile_bytes; // packed weight image bytes per N tile
  // shared-memory plan (byte offsets from the dynamic smem base)
  int off_stat, off_ss, off_bias, off_w, off_raw, off_ab;
  int kbs;                        // K blocks per pipeline stage (2 when the rings still fit, else 1)
  int mt;                         // time tiles per work item: streamed weights are reused by 2 tiles (halves L2->smem traffic)
  int nbuf;                       // TMEM accumulator sets (2 = epilogue overlaps the next item's MMAs)
  int raw_kb_bytes;               // raw staging bytes of ONE K block (a slot holds kbs of them)
  int raw_slot_bytes, raw_slots;  // fp32 staging ring filled by TMA (0 slots in direct mode)
  int ab_slot_bytes, ab_slots;    // operand ring: the A tiles of one stage (mt * kbs K blocks)
  int b_slot_bytes, b_slots, off_b;  // streamed weights: their own ring, ONE K block per slot (0 slots when resident)
  int main_stages, skip_stages;
  int w_resident;                 // 1: the whole weight image of the N tile stays in smem for all tiles of the CTA
  int smem_bytes;
  // TMA activation boxes, in SOURCE coordinates relative to the tile origin
  int tma;                  // 1: raw activations arrive by cp.async.bulk.tensor into the staging ring
  int main_box_w, main_boxes, main_origin_mul, main_origin_off;  // box x0 = t0*mul/2 + off (mul: 1=up2, 2=none, 4=down2)
  int skip_box_w, skip_origin_mul;
  // persistent scheduling: tiles ordered (sample, n tile, t tile); each CTA owns a contiguous range
  int tiles_t, tiles_total, tiles_per_cta;
};

__host__ __device__ inline int round_up4(int v) { return (v + 3) & ~3; }
static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

static bool make_geo(int c_in, int c_out, int ksize, int dilation, int c_skip, int resize, int skip_resize, bool tma, Geo* g,
                     int prefer_mt = 0, int prec = 0) {
  if (c_in <= 0 || c_in % KBLK || c_out <= 0 || c_out % 16 || c_skip % KBLK) return false;
  g->n_tiles = (c_out + 255) / 256;
  if (c_out % g->n_tiles) return false;
  g->n_tile = c_out / g->n_tiles;
  if (g->n_tile % 16 || g->n_tile < 16) return false;
  g->nkb_main = c_in / KBLK;
  g->nkb_skip = c_skip / KBLK;
  g->pad = (ksize / 2) * dilation;
  g->rows = TILE_M + 2 * g->pad;
  g->prec = prec;
  const int parts = prec == VQVS_PREC_F16 ? 1 : 2;  // operand images per K block: bf16 hi + lo, or one fp16
  g->a_kb_bytes = g->rows * 32 * parts;
  g->b_unit_main = ksize * g->n_tile * 32 * parts;
  g->b_unit_skip = g->n_tile * 32 * parts;
  g->per_tile_bytes = (long long)g->nkb_main * g->b_unit_main + (long long)g->nkb_skip * g->b_unit_skip;
  // Stacking W_hi and W_lo along N replaces hi*hi + lo*hi + hi*lo (3 MMAs) by A_hi*[W_hi;W_lo] + A_lo*[W_hi;W_lo] (2 MMAs, which
  // also adds the lo*lo term).  Kept for 32-channel tiles; 64-channel tiles issue the three products with the A tile of the
  // second one reused from the collector: same tensor time (51 + 32 + 51 against 2 x 67 clocks), 25 % fewer multiply-adds --
  // the step runs under the board's power cap, and the freed power came back as +2.9 % SM clock / +2.6 % samples/s
  // (bench.py A/B on one box) -- and a 64-column accumulator the epilogue reads once.
  g->stack = (parts == 2 && g->n_tile == 32) ? 1 : 0;
  int cols = 32;
  while (cols < (g->stack ? 2 : 1) * g->n_tile) cols *= 2;
  g->acc_cols = cols;
  g->tmem_cols = 2 * cols;
  g->tma = tma ? 1 : 0;
  g->raw_slot_bytes = 0;
  g->main_box_w = g->main_boxes = g->main_origin_mul = g->main_origin_off = g->skip_box_w = g->skip_origin_mul = 0;
  if (tma) {
    const int off = round_up4(g->pad);
    if (resize == VQVS_RESIZE_NONE) {          // source [t0 - off, t0 + 128 + off)
      g->main_box_w = TILE_M + 2 * off; g->main_boxes = 1; g->main_origin_mul = 2; g->main_origin_off = -off;
    } else if (resize == VQVS_RESIZE_UP2) {    // conv [t0-pad, t0+128+pad) -> source [t0/2 - 4, t0/2 + 68)
      if (g->pad > 4) return false;
      g->main_box_w = TILE_M / 2 + 8; g->main_boxes = 1; g->main_origin_mul = 1; g->main_origin_off = -4;
    } else {                                   // source [2*t0 - 4, 2*t0 + 260): two boxes of 132
      if (g->pad > 2) return false;
      g->main_box_w = TILE_M + 4; g->main_boxes = 2; g->main_origin_mul = 4; g->main_origin_off = -4;
    }
    int widest = g->main_box_w * g->main_boxes;
    if (c_skip) {
      g->skip_box_w = skip_resize == VQVS_RESIZE_NONE ? TILE_M : skip_resize == VQVS_RESIZE_UP2 ? TILE_M / 2 : 2 * TILE_M;
      g->skip_origin_mul = skip_resize == VQVS_RESIZE_NONE ? 2 : skip_resize == VQVS_RESIZE_UP2 ? 1 : 4;
      if (g->skip_box_w > widest) widest = g->skip_box_w;
    }
    if (g->main_box_w > 256 || g->skip_box_w > 256) return false;
    g->raw_slot_bytes = KBLK * widest * 4;
  }
  // ---- shared-memory plan: one persistent CTA per SM -----------------------------------------
  g->raw_kb_bytes = g->raw_slot_bytes;
  const int budget = 225 * 1024 + 512;  // (of the 227 KB a CTA may opt into)
  int off = SMEM_HEADER;
  g->off_stat = off;
  g->off_ss = off;   off += c_in * 8;
  g->off_bias = off; off += g->n_tile * 4;
  off = align_up(off, 128);
  g->off_w = off;
  const long long w_img = g->per_tile_bytes;
  // Resident weights leave room for 2-K-block stages only once the image passes ~90 KB (128 -> 64: 98 KB, 192 -> 64: 147 KB);
  // streaming those through the weight ring with 4-K-block stages and one time tile per item measured 5-18 % faster
  // (profiles/r2_*: the transform warps' per-stage overhead is what the larger stage amortises).  Wider N tiles keep the
  // 150 KB limit (nothing between 90 and 150 KB occurs there).
  static const int resident_kb_env = getenv("VQVS_RESIDENT_KB") ? atoi(getenv("VQVS_RESIDENT_KB")) : 0;  // tuning aid
  const int resident_kb = resident_kb_env ? resident_kb_env : (g->n_tile <= 64 ? 90 : 150);
  g->w_resident = (g->n_tiles == 1 && w_img <= (long long)resident_kb * 1024) ? 1 : 0;
  const bool narrow_streamed = !g->w_resident && g->n_tile <= 64;  // prefers one time tile per item (see above)
  if (g->w_resident) off += (int)w_img;
  g->off_raw = off;
  const int left0 = budget - off;
  bool ok = false;
  static const int force_mt = getenv("VQVS_FORCE_MT") ? atoi(getenv("VQVS_FORCE_MT")) : 0;
  static const int force_kbs = getenv("VQVS_FORCE_KBS") ? atoi(getenv("VQVS_FORCE_KBS")) : 0;
  // stage sizes must be 1, 2 or 4 K blocks (the transform warps split a stage by powers of two)
  auto sizes_ok = [](int nkb, int kbs) { return nkb == 0 || kbs < 4 || (nkb % 4) != 3; };
  for (int cand = 0; cand < 6 && !ok; ++cand) {
    // streamed weights: prefer two time tiles per item (mt = 2); resident weights gain nothing from it
    const int mt = (!g->w_resident && cand < 3) ? 2 : 1;
    const int kbs = 4 >> (cand % 3);
    if (g->w_resident && cand < 3) continue;
    if (force_mt && mt != force_mt && !g->w_resident) continue;   // tuning aids (VQVS_FORCE_MT / VQVS_FORCE_KBS)
    if (!force_mt && prefer_mt && mt != prefer_mt && !g->w_resident) continue;
    if (!force_mt && !prefer_mt && narrow_streamed && mt != 1) continue;
    if (force_kbs && kbs != force_kbs) continue;
    if (!sizes_ok(g->nkb_main, kbs) || !sizes_ok(g->nkb_skip, kbs)) continue;
    if (mt * g->acc_cols > 512 || (mt > 1 && (g->n_tile & 31))) continue;
    // Streamed weights live in their own ring of single-K-block slots, so the A stage (kbs K blocks, shared by the mt time
    // tiles) can be as large as the transform warps like: with the weights inside the operand slot every layer with
    // C_out >= 128 ran one K block per stage, i.e. one row per transform thread per synchronisation (ncu: 29 instructions
    // per element against 17 for the 4-K-block stages of the 64-channel layers).
    const int ab_slot = kbs * mt * g->a_kb_bytes;
    const int raw_slot = mt * kbs * g->raw_kb_bytes;
    const int b_slot = g->w_resident ? 0 : (g->b_unit_main > g->b_unit_skip ? g->b_unit_main : g->b_unit_skip);
    const int min_ab = 2, min_raw = tma ? ((kbs == 4 || w_img > 100 * 1024) ? 2 : 3) : 0;
    int b_slots = 0;
    if (!g->w_resident) {
      b_slots = left0 - 3 * b_slot >= min_ab * ab_slot + min_raw * raw_slot ? 3 : 2;
      if (kbs == 1 && left0 - 4 * b_slot >= 3 * ab_slot + (min_raw + 1) * raw_slot) b_slots = 4;
    }
    const int left = left0 - b_slots * b_slot;
    if (left < min_ab * ab_slot + min_raw * raw_slot) continue;
    int ab = MAX_AB_SLOTS;
    while (ab > min_ab && left - ab * ab_slot < (tma ? (kbs == 4 ? 3 : 4) * raw_slot : 0)) --ab;
    if (left - ab * ab_slot < min_raw * raw_slot) continue;
    if (ab > 4) ab = 4;
    int raw = tma ? (left - ab * ab_slot) / raw_slot : 0;
    if (raw > MAX_RAW_SLOTS) raw = MAX_RAW_SLOTS;
    g->b_slot_bytes = b_slot;
    g->b_slots = b_slots;
    g->kbs = kbs;
    g->mt = mt;
    g->nbuf = (2 * mt * g->acc_cols <= 512) ? 2 : 1;
    g->tmem_cols = g->nbuf * mt * g->acc_cols;
    g->ab_slot_bytes = ab_slot;
    g->ab_slots = ab;
    g->raw_slot_bytes = raw_slot;
    g->raw_slots = raw;
    ok = true;
  }
  if (!ok) return false;
  off += g->raw_slots * g->raw_slot_bytes;
  g->off_ab = off;
  off += g->ab_slots * g->ab_slot_bytes;
  g->off_b = off;
  off += g->b_slots * g->b_slot_bytes;
  g->smem_bytes = off;
  g->main_stages = (g->nkb_main + g->kbs - 1) / g->kbs;
  g->skip_stages = (g->nkb_skip + g->kbs - 1) / g->kbs;
  g->tiles_t = g->tiles_total = g->tiles_per_cta = 0;  // filled at launch
  return true;
}

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Optional suspend-time hint (ns) for the probe of a waiting role.  Measured with 2 us and 20 us hints against none on
// one box (tools/op_profile.py, A/B builds): no difference beyond run-to-run noise, so the default is the plain form.
#ifndef VQVS_WAIT_HINT_NS
#define VQVS_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
#if VQVS_WAIT_HINT_NS > 0
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity), "r"((uint32_t)VQVS_WAIT_HINT_NS)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t holder, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor: 16-B rows at a 16-B pitch inside an 8-row
// core matrix, SBO between 8-row groups, LBO between the two 8-channel chunks of one K block.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc(int n, bool f16 = false) {  // a/b format: 1 = bf16, 0 = fp16
  const uint32_t fmt = f16 ? 0u : ((1u << 7) | (1u << 10));
  return (1u << 4) | fmt | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// ---- grouped MMA issue ---------------------------------------------------------------------------------------
// All tcgen05.mma of ONE K block (k = 3: three taps) behind a single elected branch.  Descriptors travel as 32-bit low
// words (start address + LBO; the callers advance them with warp-uniform 32-bit adds) plus one constant high word per
// operand, and only the first product of an accumulator carries a runtime accumulate flag.  ncu (profiles/r2_*) showed
// the issue loop itself -- ~17 instructions per MMA with a branch region, an R2UR and 64-bit adds for every tap -- to be
// the limit of the 64-channel layers (134 cycles per N = 128 MMA against the 67 the tensor pipe needs).
#define VQVS_MMA_(A, B, P) \
  "mov.b64 da, {" A ", %1};\n\tmov.b64 db, {" B ", %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, " P ";\n\t"
// A-operand collector hints (SASS: UTCHMMA gdesc[..].A_KEEP / .A_REUSE): a product that shares its A tile with the NEXT
// one keeps it in the tensor core's collector buffer ("fill"), the next one consumes it without fetching it from shared
// memory again ("lastuse").  hi*hi and hi*lo share A_hi, so one of the three A-tile reads of a tap disappears.  Measured
// (A/B builds): nothing for N = 128 tiles (those MMAs sit on the 64-clock compute floor either way), but it is what makes
// three separate products as cheap as two stacked ones for N = 64 (an N = 64 MMA is bound by its operand fetch: 51 clocks,
// 32 with A reused -- tools/mma_bench.cu), with a quarter fewer multiply-adds.  -DVQVS_NO_COLLECTOR issues plain MMAs.
#ifdef VQVS_NO_COLLECTOR
#define VQVS_MMA_FILL_(A, B, P) VQVS_MMA_(A, B, P)
#define VQVS_MMA_LAST_(A, B, P) VQVS_MMA_(A, B, P)
#else
#define VQVS_MMA_FILL_(A, B, P) \
  "mov.b64 da, {" A ", %1};\n\tmov.b64 db, {" B ", %2};\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], da, db, %3, " P ";\n\t"
#define VQVS_MMA_LAST_(A, B, P) \
  "mov.b64 da, {" A ", %1};\n\tmov.b64 db, {" B ", %2};\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], da, db, %3, " P ";\n\t"
#endif
#define VQVS_MMA_HEAD_ "{\n\t.reg .pred p, t;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, %4, %4;\n\t"
// stacked weight rows [W_hi ; W_lo]: (A_hi, A_lo) per tap
__device__ __forceinline__ void mma_group_stack3(uint32_t d, uint32_t ahi, uint32_t bhi, uint32_t idesc, uint32_t accf,
                                                 uint32_t a0, uint32_t a0l, uint32_t a1, uint32_t a1l, uint32_t a2, uint32_t a2l,
                                                 uint32_t b0, uint32_t b1, uint32_t b2) {
  asm volatile(VQVS_MMA_HEAD_
               VQVS_MMA_("%5", "%11", "p") VQVS_MMA_("%6", "%11", "t")
               VQVS_MMA_("%7", "%12", "t") VQVS_MMA_("%8", "%12", "t")
               VQVS_MMA_("%9", "%13", "t") VQVS_MMA_("%10", "%13", "t")
               "}" ::"r"(d), "r"(ahi), "r"(bhi), "r"(idesc), "r"(accf),
               "r"(a0), "r"(a0l), "r"(a1), "r"(a1l), "r"(a2), "r"(a2l), "r"(b0), "r"(b1), "r"(b2)
               : "memory");
}
__device__ __forceinline__ void mma_group_stack1(uint32_t d, uint32_t ahi, uint32_t bhi, uint32_t idesc, uint32_t accf,
                                                 uint32_t a0, uint32_t a0l, uint32_t b0) {
  asm volatile(VQVS_MMA_HEAD_ VQVS_MMA_("%5", "%7", "p") VQVS_MMA_("%6", "%7", "t") "}" ::"r"(d), "r"(ahi), "r"(bhi), "r"(idesc),
               "r"(accf), "r"(a0), "r"(a0l), "r"(b0)
               : "memory");
}
// separate W_hi / W_lo tiles: hi*hi + lo*hi + hi*lo per tap
__device__ __forceinline__ void mma_group_split3(uint32_t d, uint32_t ahi, uint32_t bhi, uint32_t idesc, uint32_t accf,
                                                 uint32_t a0, uint32_t a0l, uint32_t a1, uint32_t a1l, uint32_t a2, uint32_t a2l,
                                                 uint32_t b0, uint32_t b0l, uint32_t b1, uint32_t b1l, uint32_t b2, uint32_t b2l) {
  asm volatile(VQVS_MMA_HEAD_
               VQVS_MMA_FILL_("%5", "%11", "p") VQVS_MMA_LAST_("%5", "%12", "t") VQVS_MMA_("%6", "%11", "t")
               VQVS_MMA_FILL_("%7", "%13", "t") VQVS_MMA_LAST_("%7", "%14", "t") VQVS_MMA_("%8", "%13", "t")
               VQVS_MMA_FILL_("%9", "%15", "t") VQVS_MMA_LAST_("%9", "%16", "t") VQVS_MMA_("%10", "%15", "t")
               "}" ::"r"(d), "r"(ahi), "r"(bhi), "r"(idesc), "r"(accf),
               "r"(a0), "r"(a0l), "r"(a1), "r"(a1l), "r"(a2), "r"(a2l), "r"(b0), "r"(b0l), "r"(b1), "r"(b1l), "r"(b2), "r"(b2l)
               : "memory");
}
// VQVS_PREC_F16: one product per tap
__device__ __forceinline__ void mma_group_single3(uint32_t d, uint32_t ahi, uint32_t bhi, uint32_t idesc, uint32_t accf,
                                                  uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1, uint32_t b2) {
  asm volatile(VQVS_MMA_HEAD_ VQVS_MMA_("%5", "%8", "p") VQVS_MMA_("%6", "%9", "t") VQVS_MMA_("%7", "%10", "t") "}" ::"r"(d),
               "r"(ahi), "r"(bhi), "r"(idesc), "r"(accf), "r"(a0), "r"(a1), "r"(a2), "r"(b0), "r"(b1), "r"(b2)
               : "memory");
}
__device__ __forceinline__ void mma_group_single1(uint32_t d, uint32_t ahi, uint32_t bhi, uint32_t idesc, uint32_t accf,
                                                  uint32_t a0, uint32_t b0) {
  asm volatile(VQVS_MMA_HEAD_ VQVS_MMA_("%5", "%6", "p") "}" ::"r"(d), "r"(ahi), "r"(bhi), "r"(idesc), "r"(accf), "r"(a0), "r"(b0)
               : "memory");
}
__device__ __forceinline__ void mma_group_split1(uint32_t d, uint32_t ahi, uint32_t bhi, uint32_t idesc, uint32_t accf,
                                                 uint32_t a0, uint32_t a0l, uint32_t b0, uint32_t b0l) {
  asm volatile(VQVS_MMA_HEAD_ VQVS_MMA_FILL_("%5", "%7", "p") VQVS_MMA_LAST_("%5", "%8", "t") VQVS_MMA_("%6", "%7", "t") "}" ::"r"(d),
               "r"(ahi), "r"(bhi), "r"(idesc), "r"(accf), "r"(a0), "r"(a0l), "r"(b0), "r"(b0l)
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 8-column TMEM load WITHOUT the wait (pair with tmem_ld_wait): lets global loads be issued under its latency
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two 16-column loads (e.g. the hi and lo halves of a stacked accumulator) behind one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, float* v, float* w) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr0));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr1));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = __uint_as_float(r[i]);
    w[i] = __uint_as_float(q[i]);
  }
}

// ---------------------------------------------------------------------------
// prologue math
// ---------------------------------------------------------------------------
// rcp_approx / ex2_approx / gelu_as live in common.cuh (shared with the SIMT kernels)

// split 8 floats into bf16 hi / lo vectors (16 B each)
__device__ __forceinline__ void split8(const float* v, uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
    const float r0 = v[2 * i] - __uint_as_float(h[i] << 16);
    const float r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(r0, r1);
    l[i] = *reinterpret_cast<const uint32_t*>(&lb);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 8 floats -> one 16-B row of fp16 (VQVS_PREC_F16 operands)
__device__ __forceinline__ uint4 pack8h(const float* v) {
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hb = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
  }
  return make_uint4(h[0], h[1], h[2], h[3]);
}

__device__ __forceinline__ void store_rows(const float* v, uint8_t* a_hi, uint8_t* a_lo, int row, bool f16) {
  if (f16) {
    *reinterpret_cast<uint4*>(a_hi + row * 16) = pack8h(v);
    return;
  }
  uint4 hi, lo;
  split8(v, &hi, &lo);
  *reinterpret_cast<uint4*>(a_hi + row * 16) = hi;
  *reinterpret_cast<uint4*>(a_lo + row * 16) = lo;
}

// (scale, shift) for 8 consecutive channels: four float4, one per channel PAIR = {sc0, sc1, sh0, sh1}
// (pair-major so that the packed fp32x2 path reads {sc0, sc1} and {sh0, sh1} as 64-bit operands)
__device__ __forceinline__ void affine_gelu8(float* v, const float4* ss) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 p = ss[i];
    v[2 * i] = gelu_as(fmaf(v[2 * i], p.x, p.z));
    v[2 * i + 1] = gelu_as(fmaf(v[2 * i + 1], p.y, p.w));
  }
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2, two fp32 lanes per instruction) -------------------------
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t bcast2(float c) { return pack2(c, c); }

// erf GELU of 4 packed channel pairs.  VQVS_GELU_DEG = 0: the Abramowitz-Stegun 7.1.26 form (11 packed FP + 4 MUFU per pair).
// Default (5): GELU(y) = max(y, 0) - |y| * Phi(-|y|) with Phi(-x) = 2^P(x), P = the weighted-minimax degree-5 fit of
// log2 Phi(-x) (tools/fit_gelu_poly.py: |GELU error| <= 4.4e-7 exact, 6.4e-7 in this fp32 evaluation; monotone decreasing
// for all x, so large |y| flush to 0 through ex2(-inf)): 6 packed FP + 2 MUFU + 2 FMNMX per pair.  The transform warps
// are bound by the FMA and XU pipes they share with the epilogue (profiles/r2_ncu_stalls.txt), so the op count is
// what this buys.  Degree 6 (<= 1.5e-7) needs its argument clamped (positive leading coefficient).
#ifndef VQVS_GELU_DEG
#define VQVS_GELU_DEG 5
#endif
__device__ __forceinline__ void gelu4p(uint64_t* y) {
  uint64_t nay[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) nay[i] = y[i] | 0x8000000080000000ull;  // -|y|
#if VQVS_GELU_DEG == 0
  uint64_t t[4], q[4], e[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(nay[i], bcast2(-0.2316418882f), bcast2(1.0f));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2(t[i], a, b);
    t[i] = pack2(rcp_approx(a), rcp_approx(b));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(y[i], y[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(e[i], bcast2(-0.72134752044f));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2(e[i], a, b);
    e[i] = pack2(ex2_approx(a), ex2_approx(b));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(t[i], bcast2(0.5307027145f), bcast2(-0.7265760135f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], bcast2(0.7107068705f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], bcast2(-0.142248368f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], bcast2(0.127414796f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = mul2(q[i], t[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], e[i], bcast2(-0.5f));  // 0.5*erfc(|y|/sqrt2) - 0.5
  // gelu(y) = relu(y) - |y|*0.5*erfc(|y|/sqrt2) = 0.5*y + (-|y|)*(0.5*erfc(|y|/sqrt2) - 0.5): packed throughout (no scalar max)
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = fma2(nay[i], q[i], mul2(y[i], bcast2(0.5f)));
#else
  // Horner in n = -|y| (odd coefficients of P change sign), stage by stage across the 4 pairs so the chains interleave
#if VQVS_GELU_DEG == 6
  constexpr int NC = 7;
  const float c[NC] = {3.309271415e-05f, 7.692189538e-04f, 8.080714382e-03f, 5.341210216e-02f, -4.587709606e-01f, 1.151201725e+00f,
                       -9.999930859e-01f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // P6 turns upwards beyond |y| = 13.7: clamp (2^P(7.5) = 3e-14)
    float a, b;
    unpack2(nay[i], a, b);
    nay[i] = pack2(fmaxf(a, -7.5f), fmaxf(b, -7.5f));
  }
#else
  constexpr int NC = 6;
  const float c[NC] = {4.733092792e-04f, 7.084557321e-03f, 5.182738230e-02f, -4.599924386e-01f, 1.150787830e+00f, -1.000037670e+00f};
#endif
  uint64_t p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = fma2(nay[i], bcast2(c[0]), bcast2(c[1]));
#pragma unroll
  for (int k = 2; k < NC; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], nay[i], bcast2(c[k]));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2(p[i], a, b);
    p[i] = pack2(ex2_approx(a), ex2_approx(b));  // Phi(-|y|)
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2(y[i], a, b);
    y[i] = fma2(nay[i], p[i], pack2(fmaxf(a, 0.f), fmaxf(b, 0.f)));
  }
#endif
}

// 4 packed pairs (8 consecutive channels of one position) -> bf16 hi / lo operand rows (16 B each)
__device__ __forceinline__ void split4p(const uint64_t* y, uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2(y[i], a, b);
    const __nv_bfloat162 hb = __floats2bfloat162_rn(a, b);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t hf = pack2(__uint_as_float(h[i] << 16), __uint_as_float(h[i] & 0xffff0000u));
    const uint64_t r = fma2(hf, bcast2(-1.0f), y[i]);  // exact residual
    float a, b;
    unpack2(r, a, b);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(a, b);
    l[i] = *reinterpret_cast<const uint32_t*>(&lb);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct Src {
  const float* a;
  const float* b;
  int c_a, c_b, t_in, t_conv, resize;
};

// Direct mode (no TMA: lengths not a multiple of 4): 8 channels [c8, c8+8) at conv-input position tc.
__device__ __forceinline__ void produce_direct(const Src& s, int n, int c8, int tc, bool act, const float4* ss,
                                               uint8_t* a_hi, uint8_t* a_lo, int row, bool f16) {
  float v[8];
  if (tc < 0 || tc >= s.t_conv) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
  } else {
    const float* base = c8 < s.c_a ? s.a + ((size_t)n * s.c_a + c8) * s.t_in
                                   : s.b + ((size_t)n * s.c_b + (c8 - s.c_a)) * s.t_in;
    if (s.resize == VQVS_RESIZE_DOWN2) {
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = __ldg(base + (size_t)e * s.t_in + 2 * tc);
        w[e] = __ldg(base + (size_t)e * s.t_in + 2 * tc + 1);
      }
      if (act) {
        affine_gelu8(v, ss);
        affine_gelu8(w, ss);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.5f * (v[e] + w[e]);
    } else {
      const int ts = s.resize == VQVS_RESIZE_UP2 ? (tc >> 1) : tc;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(base + (size_t)e * s.t_in + ts);
      if (act) affine_gelu8(v, ss);
    }
  }
  store_rows(v, a_hi, a_lo, row, f16);
}

// Transposing butterfly: every lane holds v[0..31] (one row, 32 columns); afterwards lane l holds
// in v[0] the sum over the 32 lanes (rows) of column l.  31 shuffles instead of 160.
__device__ __forceinline__ float column_sums32(float* v, int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = upper ? v[i] : v[i + step];
      const float keep = upper ? v[i + step] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
// TMA tensor load of one [16 channels x box_w positions] fp32 box (SASS: UTMALDG); out-of-range
// positions are zero-filled by the hardware.
__device__ __forceinline__ void tma_box_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Per-tile description of one operand source (main taps or 1x1 skip) in TMA mode.
struct StageView {
  int rows;           // operand rows of this conv (row pitch of the hi/lo blocks)
  int box_w, boxes, x0, tcs, n_rows, t_src, t_conv, resize;
  bool act, f16;
  bool edge;  // some operand rows of this tile lie outside the sequence (first / last tile of a sample): they must read as zero
};

// Row-wise work of one thread in one stage: rows row_first, row_first + 32, ... (NIT of them) of ONE 8-channel
// chunk of ONE K block -> bf16 hi/lo operand rows.  The chunk's (scale, shift) are read once into registers as
// packed pairs; the next row's raw values are fetched before the current row is evaluated.
// raw / a point at the K block; ss at its 16 (scale, shift) pairs.
// BOXW != 0: the staging pitch is a compile-time constant (the channel offsets become immediates of the loads)
template <bool DOWN, int BOXW = 0>
__device__ __forceinline__ void load_row8(const StageView& v, const float* raw_c, int tc, float* x, float* w) {
  if (!DOWN) {
    const float* raw = raw_c + (tc - v.x0);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = raw[e * (BOXW ? BOXW : v.box_w)];
  } else {
    int col = 2 * tc - v.x0;
    const float* raw = raw_c;
    if (v.boxes == 2 && col >= v.box_w) {
      col -= v.box_w;
      raw += KBLK * v.box_w;
    }
    raw += col;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 pr = *reinterpret_cast<const float2*>(raw + e * v.box_w);
      x[e] = pr.x;
      w[e] = pr.y;
    }
  }
}

template <bool DOWN, bool ACT, int R, int BOXW = 0>
__device__ __forceinline__ void transform_row_group(const StageView& v, const float* raw_c, uint8_t* a_hi, uint8_t* a_lo,
                                                    const uint64_t* sc, const uint64_t* sh, int row0) {
  // R rows (32 apart) in flight at once: phase-major source order lets the scheduler interleave the GELU chains
  float x[R][8], w[R][8];
  uint64_t y[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r) load_row8<DOWN, BOXW>(v, raw_c, v.tcs + row0 + 32 * r, x[r], w[r]);
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) y[r][i] = pack2(x[r][2 * i], x[r][2 * i + 1]);
  if (ACT) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) y[r][i] = fma2(y[r][i], sc[i], sh[i]);
#pragma unroll
    for (int r = 0; r < R; ++r) gelu4p(y[r]);
  }
  if (DOWN) {
    uint64_t z[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) z[r][i] = pack2(w[r][2 * i], w[r][2 * i + 1]);
    if (ACT) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) z[r][i] = fma2(z[r][i], sc[i], sh[i]);
#pragma unroll
      for (int r = 0; r < R; ++r) gelu4p(z[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) y[r][i] = fma2(y[r][i], bcast2(0.5f), mul2(z[r][i], bcast2(0.5f)));
  }
  // Out-of-range positions read zero-filled (finite) staging memory, but GELU(shift) != 0: their operand rows must be zero.
  // Only the first and last tile of a sample have such rows, so the common path stores unconditionally (no per-row range
  // test, no divergent region) and an edge tile overwrites its out-of-range rows afterwards (same thread, program order).
  if (v.f16) {  // one fp16 operand row
    uint4 h[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint32_t q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a, b;
        unpack2(y[r][i], a, b);
        const __half2 hb = __floats2half2_rn(a, b);
        q[i] = *reinterpret_cast<const uint32_t*>(&hb);
      }
      h[r] = make_uint4(q[0], q[1], q[2], q[3]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) *reinterpret_cast<uint4*>(a_hi + (row0 + 32 * r) * 16) = h[r];
  } else {
    uint4 hi[R], lo[R];
#pragma unroll
    for (int r = 0; r < R; ++r) split4p(y[r], &hi[r], &lo[r]);
#pragma unroll
    for (int r = 0; r < R; ++r) {  // (all rows packed before the first store: a store's source registers are not recycled under it)
      *reinterpret_cast<uint4*>(a_hi + (row0 + 32 * r) * 16) = hi[r];
      *reinterpret_cast<uint4*>(a_lo + (row0 + 32 * r) * 16) = lo[r];
    }
  }
  if (v.edge) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + 32 * r, tc = v.tcs + row;
      if (tc < 0 || tc >= v.t_conv) {
        *reinterpret_cast<uint4*>(a_hi + row * 16) = make_uint4(0, 0, 0, 0);
        if (!v.f16) *reinterpret_cast<uint4*>(a_lo + row * 16) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

// The row loop is deliberately NOT unrolled beyond two rows and every mode is a template parameter: the persistent
// CTA runs four different role loops at once, and ncu showed instruction-fetch stalls (34 % of all warp samples) as
// the top stall reason when these loops were unrolled into tens of KB of code.
template <bool DOWN, bool ACT, int BOXW = 0>
__device__ __forceinline__ void transform_rows(const StageView& v, const uint8_t* raw_kb, uint8_t* a_kb, const float2* ss_kb,
                                               int chunk, int row_first, int nit) {
  uint64_t sc[4], sh[4];
  if (ACT) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 p = reinterpret_cast<const float4*>(ss_kb + chunk * 8)[i];
      sc[i] = pack2(p.x, p.y);
      sh[i] = pack2(p.z, p.w);
    }
  }
  const float* raw_c = reinterpret_cast<const float*>(raw_kb) + (chunk * 8) * (BOXW ? BOXW : v.box_w);
  uint8_t* a_hi = a_kb + chunk * (v.rows * 16);
  uint8_t* a_lo = a_hi + v.rows * 32;
  int it = 0;
  if (!DOWN) {  // (pooled rows already run two GELU batches per row)
#pragma unroll 1
    for (; it + 2 <= nit; it += 2) transform_row_group<DOWN, ACT, 2, BOXW>(v, raw_c, a_hi, a_lo, sc, sh, row_first + 32 * it);
  }
#pragma unroll 1
  for (; it < nit; ++it) transform_row_group<DOWN, ACT, 1, BOXW>(v, raw_c, a_hi, a_lo, sc, sh, row_first + 32 * it);
}

template <bool DOWN, int BOXW = 0>
__device__ __forceinline__ void transform_rows_n(const StageView& v, const uint8_t* raw_kb, uint8_t* a_kb, const float2* ss_kb,
                                                 int chunk, int row_first, int nit) {
  if (v.act) transform_rows<DOWN, true, BOXW>(v, raw_kb, a_kb, ss_kb, chunk, row_first, nit);
  else transform_rows<DOWN, false, BOXW>(v, raw_kb, a_kb, ss_kb, chunk, row_first, nit);
}

// nearest x2: one item = one SOURCE position of one K block -> two operand rows (GELU evaluated once)
__device__ __forceinline__ void transform_up2(const StageView& v, const uint8_t* raw_k, uint8_t* a_k, const float2* ss_k,
                                              int chunk, int ts) {
  float x[8];
  if (ts >= 0 && ts < v.t_src) {
    const float* raw = reinterpret_cast<const float*>(raw_k) + (chunk * 8) * v.box_w + (ts - v.x0);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = raw[e * v.box_w];
    if (v.act) affine_gelu8(x, reinterpret_cast<const float4*>(ss_k + chunk * 8));
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
  }
  uint8_t* a_hi = a_k + chunk * (v.rows * 16);
  uint8_t* a_lo = a_hi + v.rows * 32;
  const int r0 = 2 * ts - v.tcs;
  if (r0 >= 0 && r0 < v.n_rows) store_rows(x, a_hi, a_lo, r0, v.f16);
  if (r0 + 1 >= 0 && r0 + 1 < v.n_rows) store_rows(x, a_hi, a_lo, r0 + 1, v.f16);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Optional role profiler (debug flag 512): cycles per phase of block 0, read back with vqvs_debug_prof.
static __device__ unsigned long long g_prof[32];  // (one copy per translation unit; only the generic kernels' unit writes it)
#ifdef VQVS_PROF
static __device__ unsigned long long g_cta[160 * 4];  // per CTA: smid, first-tile-ready clock, end clock (globaltimer ns), tiles
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
#endif
#ifdef VQVS_PROF
#define PROF_ADD(slot, since) do { if (prof) { const long long now_ = clock64(); acc_[slot] += now_ - (since); (since) = now_; } } while (0)
#define PROF_DECL(cond) const bool prof = (cond); long long acc_[4] = {0, 0, 0, 0}; long long tprev = prof ? clock64() : 0
#define PROF_STORE(base) do { if (prof) for (int i_ = 0; i_ < 4; ++i_) g_prof[(base) + i_] = acc_[i_]; } while (0)
#else
#define PROF_ADD(slot, since) do { } while (0)
#define PROF_DECL(cond) do { } while (0)
#define PROF_STORE(base) do { } while (0)
#endif

struct Ring {  // running (slot, phase) of an mbarrier ring: no integer division on the critical path
  int idx, n;
  uint32_t ph;
  __device__ __forceinline__ Ring(int slots) : idx(0), n(slots), ph(0) {}
  __device__ __forceinline__ void next() {
    if (++idx == n) {
      idx = 0;
      ph ^= 1;
    }
  }
};

// LEAN = the common case (TMA staging + register-statistics epilogue) compiled WITHOUT the generic paths (direct
// global loads, per-tile butterfly epilogue): carrying unused code costs registers in the role loops (an unused
// 470-line experiment slowed this kernel by 10 %), so the hot instantiation contains only what it runs.
// KIND 0 = generic, 1 = LEAN, 2 = LEAN and PLAIN (no pooling / upsampling in the staged operands: 114 of the 130 convs
// of a UNet step), which also drops the resampling transforms.
template <int MT, int KIND>
__global__ void __launch_bounds__(THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_xa, const __grid_constant__ CUtensorMap tm_xb,
                 const __grid_constant__ CUtensorMap tm_sa, const __grid_constant__ CUtensorMap tm_sb, const VqvsConv d,
                 const Geo g, const VqvsGnFinalize fin) {
  // KIND 3 = PLAIN and additionally no 1x1-skip stages and resident weights (conv1 and identity-skip conv2 of every
  // 64-channel layer: the largest share of a step)
  constexpr bool LEAN = KIND != 0, PLAIN = KIND >= 2, SIMPLE = KIND == 3;
  extern __shared__ __align__(128) uint8_t smem[];
  // mbarriers: raw_full[8] raw_empty[8] b_full[8] b_empty[8] a_full[8] ab_empty[8] acc_full[2] acc_empty[2] w_full
  const uint32_t bar0 = smem_u32(smem);
#define RAW_FULL(i) (bar0 + 8u * (i))
#define RAW_EMPTY(i) (bar0 + 8u * (8 + (i)))
#define B_FULL(i) (bar0 + 8u * (16 + (i)))
#define B_EMPTY(i) (bar0 + 8u * (24 + (i)))
#define A_FULL(i) (bar0 + 8u * (32 + (i)))
#define AB_EMPTY(i) (bar0 + 8u * (40 + (i)))
#define ACC_FULL(i) (bar0 + 8u * (48 + (i)))
#define ACC_EMPTY(i) (bar0 + 8u * (50 + (i)))
#define W_FULL (bar0 + 8u * 52)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + 53 * 8);
  float2* s_ss = reinterpret_cast<float2*>(smem + g.off_ss);    // (scale, shift) of the current sample
  float* s_bias = reinterpret_cast<float*>(smem + g.off_bias);
  const int c_in = d.c_a + d.c_b;
#ifdef VQVS_PROF
  const int dbg_flags = LEAN ? (d.reserved_ & 512) : d.reserved_;  // profiling build: the role profiler runs in every kind
#else
  const int dbg_flags = LEAN ? 0 : d.reserved_;  // profiling / ablation switches exist only in the generic instantiation
#endif

  // warp index through a shuffle: tells the compiler it is warp-uniform, so the role branches are uniform and the
  // MMA / TMA warps can keep their loop state and descriptors in uniform registers (no R2UR per tcgen05.mma)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // tile schedule: every CTA owns a contiguous range of work items (sample-major, then N tile, then time), so
  // the epilogue can keep GroupNorm partial sums in registers across the tiles of a sample
  const int tiles_lo = g.tiles_total / (int)gridDim.x, tiles_rem = g.tiles_total % (int)gridDim.x;
  const int tile_first = (int)blockIdx.x * tiles_lo + min((int)blockIdx.x, tiles_rem);
  const int n_my_tiles = tiles_lo + ((int)blockIdx.x < tiles_rem ? 1 : 0);
  // (time tile, N tile, sample) of the CTA's first item, advanced incrementally: no division per tile
  const int tile0_tx = tile_first % g.tiles_t, tile0_nt = (tile_first / g.tiles_t) % g.n_tiles;
  const int tile0_n = tile_first / (g.tiles_t * g.n_tiles);
#define TILE_ITER_INIT() int it_tx = tile0_tx, it_nt = tile0_nt, it_n = tile0_n
#define TILE_COORDS(tile)     \
  const int nt = it_nt;       \
  const int n = it_n;         \
  const int t0 = it_tx * (TILE_M * MT);
#define TILE_ITER_NEXT() (++it_tx == g.tiles_t ? (it_tx = 0, (++it_nt == g.n_tiles ? (it_nt = 0, ++it_n) : 0)) : 0)

#ifdef VQVS_PROF
  const long long prof_t0 = clock64();
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_RAW_SLOTS; ++i) {
      mbar_init(RAW_FULL(i), 1);
      mbar_init(RAW_EMPTY(i), XFORM_WARPS);
    }
    for (int i = 0; i < MAX_B_SLOTS; ++i) {
      mbar_init(B_FULL(i), 1);
      mbar_init(B_EMPTY(i), 1);
    }
    for (int i = 0; i < MAX_AB_SLOTS; ++i) {
      mbar_init(A_FULL(i), XFORM_WARPS);
      mbar_init(AB_EMPTY(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(ACC_FULL(i), 1);
      mbar_init(ACC_EMPTY(i), EPI_WARPS * 32);
    }
    mbar_init(W_FULL, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_holder), g.tmem_cols);
  if (warp == TMA_RAW_WARP && lane == 0 && g.tma) {
    tma_prefetch_desc(&tm_xa);
    if (d.c_b) tma_prefetch_desc(&tm_xb);
    if (g.nkb_skip) {
      tma_prefetch_desc(&tm_sa);
      if (d.s_b) tma_prefetch_desc(&tm_sb);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int total_stages = g.main_stages + g.skip_stages;
  // Programmatic dependent launch (the 130 convs of a UNet step form one chain on one stream): the next conv may be
  // scheduled as soon as every CTA of this grid got here, so its CTAs take over SMs as ours retire and run their own
  // prologue (mbarrier init, TMEM allocation, resident weight load) under our tail; everything that reads what the
  // PREVIOUS grid wrote (activations, statistics, skip inputs) waits for that grid to complete.  Both instructions are
  // no-ops for a launch without the attribute.
  asm volatile("griddepcontrol.launch_dependents;");
  if (warp != TMA_W_WARP) asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef VQVS_PROF
  if ((dbg_flags & 512) && blockIdx.x == 0 && threadIdx.x == 0) g_prof[6] = clock64() - prof_t0;  // prologue + wait for the previous grid
  if ((dbg_flags & 512) && threadIdx.x == 0 && blockIdx.x < 160) {
    g_cta[blockIdx.x * 4 + 0] = smid();
    g_cta[blockIdx.x * 4 + 1] = gtime_ns();  // dependent data available
    g_cta[blockIdx.x * 4 + 3] = n_my_tiles;
  }
#endif

// Register budget: 640 threads x 96 registers (the launch bound) for every role.  Per-role budgets via setmaxnreg
// were tried (72/128, 56/96/112): the halo transform warp then runs spilling code on the critical path of every
// stage, and an overdrawn pool deadlocks in USETMAXREG.TRY_ALLOC; a uniform budget is both faster and simpler.
#define REG_DEC() do { } while (0)  // uniform budget: see above
#define REG_INC() do { } while (0)
  auto transform_role = [&](auto halo_c) {
    constexpr bool HALO = decltype(halo_c)::value;
    const int xtid = threadIdx.x - XFORM_WARP0 * 32, xwarp = warp - XFORM_WARP0;
    const int mwarp = xwarp - 1;  // 0..7 for the main-row warps
    // =========================== operand producers (transform warps) ===========================
    const Src main_src{d.xa, d.xb, d.c_a, d.c_b, d.t_in, d.t_out, d.resize};
    const Src skip_src{d.sa, d.sb, d.s_a, d.s_b, d.t_skip, d.t_out, d.skip_resize};
    Ring ab(g.ab_slots), rw(g.tma ? g.raw_slots : 1);
    int staged_n = -1;
#ifndef VQVS_PROF_XWARP
#define VQVS_PROF_XWARP 0  // which main transform warp the role profiler follows (0..7; warp 12 + n sits on scheduler n % 4)
#endif
    PROF_DECL((dbg_flags & 512) && blockIdx.x == 0 && xtid == 32 * (1 + VQVS_PROF_XWARP));
    // per-CTA constants of the two operand sources (main taps / 1x1 skip)
    StageView vm, vs;
    vm.rows = vs.rows = g.rows;
    vm.box_w = g.main_box_w;  vs.box_w = g.skip_box_w;
    vm.boxes = g.main_boxes;  vs.boxes = 1;
    vm.n_rows = g.rows;       vs.n_rows = TILE_M;
    vm.t_src = d.t_in;        vs.t_src = d.t_skip;
    vm.t_conv = vs.t_conv = d.t_out;
    vm.resize = d.resize;     vs.resize = d.skip_resize;
    vm.act = d.act && !(dbg_flags & 8);
    vs.act = false;
    vm.f16 = vs.f16 = g.prec == VQVS_PREC_F16;
    const int main_step = (TILE_M * g.main_origin_mul) / 2, skip_step = (TILE_M * g.skip_origin_mul) / 2;
    TILE_ITER_INIT();
    for (int k_local = 0; k_local < n_my_tiles; ++k_local, TILE_ITER_NEXT()) {
      TILE_COORDS(tile)
      (void)nt;
      if (d.act && n != staged_n) {  // per-sample GroupNorm/FiLM affine
        asm volatile("bar.sync 1, %0;" ::"n"(XFORM_WARPS * 32));
        if (fin.groups > 0) {
          // fused GroupNorm(+FiLM) finalize: the same fp64 formulas as gn_finalize_kernel, evaluated by the consumer
          // for the sample it is about to read (one separate launch per conv saved).  Two phases: one thread per GROUP
          // sums its channels' (sum, sumsq) (16-B loads, four in flight) and derives (mean, rstd) once; then every thread
          // turns two channels into (scale, shift).  (Per-channel re-reading of the whole group cost cg dependent L2 round
          // trips and one fp64 division + rsqrt per channel: ~6 us per sample change at 512 channels.)
          const int cg = c_in / fin.groups;
          double2* s_grp = reinterpret_cast<double2*>(smem + SMEM_GROUP_TABLE);
          const bool table = fin.groups <= 32;
          if (table) {
            if (xtid < fin.groups) {
              double s = 0.0, ss = 0.0;
              const int c0 = xtid * cg;
#pragma unroll 4
              for (int j = 0; j < cg; ++j) {
                const int cc = c0 + j;
                const double* st = cc < fin.c_a ? fin.stats_a + ((size_t)n * fin.c_a + cc) * 2
                                                : fin.stats_b + ((size_t)n * fin.c_b + (cc - fin.c_a)) * 2;
                const double2 v = __ldcg(reinterpret_cast<const double2*>(st));
                s += v.x;
                ss += v.y;
              }
              const double cnt = (double)cg * (double)fin.count;
              const double mean = s / cnt;
              double var = ss / cnt - mean * mean;
              var = var > 0.0 ? var : 0.0;
              s_grp[xtid] = make_double2(mean, rsqrt(var + 1e-5));
            }
            asm volatile("bar.sync 1, %0;" ::"n"(XFORM_WARPS * 32));
          }
          for (int i = xtid; i < c_in / 2; i += XFORM_WARPS * 32) {
            float sc2[2], sh2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 2 * i + e, grp = c / cg;
              double mean, rstd;
              if (table) {
                const double2 mr = s_grp[grp];
                mean = mr.x;
                rstd = mr.y;
              } else {
                double s = 0.0, ss = 0.0;
                for (int cc = grp * cg; cc < (grp + 1) * cg; ++cc) {
                  const double* st = cc < fin.c_a ? fin.stats_a + ((size_t)n * fin.c_a + cc) * 2
                                                  : fin.stats_b + ((size_t)n * fin.c_b + (cc - fin.c_a)) * 2;
                  s += __ldcg(st);
                  ss += __ldcg(st + 1);
                }
                const double cnt = (double)cg * (double)fin.count;
                mean = s / cnt;
                double var = ss / cnt - mean * mean;
                var = var > 0.0 ? var : 0.0;
                rstd = rsqrt(var + 1e-5);
              }
              double sc = rstd * (double)fin.gamma[c];
              double sh = (double)fin.beta[c] - mean * sc;
              if (fin.film) {
                const double fa = (double)fin.film[(size_t)n * fin.film_stride + c];
                const double fb = (double)fin.film[(size_t)n * fin.film_stride + c_in + c];
                sc = sc * (1.0 + fa);
                sh = sh * (1.0 + fa) + fb;
              }
              sc2[e] = (float)sc;
              sh2[e] = (float)sh;
            }
            reinterpret_cast<float4*>(s_ss)[i] = make_float4(sc2[0], sc2[1], sh2[0], sh2[1]);
          }
        } else {
          for (int i = xtid; i < c_in / 2; i += XFORM_WARPS * 32) {  // pair-major: {sc0, sc1, sh0, sh1}
            const float2 sc = *reinterpret_cast<const float2*>(d.scale + (size_t)n * c_in + 2 * i);
            const float2 sh = *reinterpret_cast<const float2*>(d.shift + (size_t)n * c_in + 2 * i);
            reinterpret_cast<float4*>(s_ss)[i] = make_float4(sc.x, sc.y, sh.x, sh.y);
          }
        }
        staged_n = n;
        asm volatile("bar.sync 1, %0;" ::"n"(XFORM_WARPS * 32));
      }
      const int x0m = (t0 / 2) * g.main_origin_mul + g.main_origin_off;  // t0 is a multiple of 128
      const int x0s = (t0 / 2) * g.skip_origin_mul;
      for (int st = 0; st < total_stages; ++st) {
        PROF_ADD(3, tprev);
        mbar_wait(AB_EMPTY(ab.idx), ab.ph ^ 1);
        PROF_ADD(0, tprev);
        const bool is_skip = !SIMPLE && st >= g.main_stages;
        const int kb0 = (is_skip ? st - g.main_stages : st) * g.kbs;
        const int nk = min(g.kbs, (is_skip ? g.nkb_skip : g.nkb_main) - kb0);
        uint8_t* a_slot0 = smem + g.off_ab + ab.idx * g.ab_slot_bytes;
        if (LEAN || g.tma) {
          mbar_wait(RAW_FULL(rw.idx), rw.ph);
          PROF_ADD(1, tprev);
          StageView v = is_skip ? vs : vm;
          v.x0 = is_skip ? x0s : x0m;
          v.tcs = is_skip ? t0 : t0 - g.pad;
          v.edge = v.tcs < 0 || v.tcs + MT * TILE_M + (v.n_rows - TILE_M) > v.t_conv;  // (covers every time tile of the item)
          const int step = is_skip ? skip_step : main_step;
          const float2* ss = s_ss + kb0 * KBLK;
          // Work split of a row-wise stage: the stage has 2*nk chunks of 8 channels; warps 0..7 each own ONE chunk
          // (q) and a 128*nk/... slice of the 128 main rows, so the chunk's (scale, shift) live in registers for
          // the whole stage; the 9th warp transforms the halo rows of every chunk.
          const int chunks = 2 * nk;                       // 2, 4 or 8 (stage sizes are 1, 2 or 4 K blocks)
          const int q = mwarp & (chunks - 1);
          const int csh = nk == 1 ? 1 : nk == 2 ? 2 : 3;                    // log2(chunks)
          const int row_first = ((mwarp >> csh) << (4 + csh)) + lane;      // 8/chunks warps share a chunk, 16*chunks rows each
          for (int j = 0; j < MT; ++j, v.x0 += step, v.tcs += TILE_M) {  // the time tiles of this item share the stage's weights
            const uint8_t* raw = smem + g.off_raw + rw.idx * g.raw_slot_bytes + j * g.kbs * g.raw_kb_bytes;
            uint8_t* a_slot = a_slot0 + j * g.kbs * g.a_kb_bytes;
            if (dbg_flags & 64) {
              // ablation: no staging work at all
            } else if (!PLAIN && v.resize == VQVS_RESIZE_UP2) {
              const int first = v.tcs >> 1;
              const int nsrc = ((v.tcs + v.n_rows - 1) >> 1) - first + 1;
              for (int i = xtid; i < nk * 2 * nsrc; i += XFORM_WARPS * 32) {
                const int qq = i / nsrc, jj = i - qq * nsrc;  // qq = (k block, chunk)
                transform_up2(v, raw + (qq >> 1) * g.raw_kb_bytes, a_slot + (qq >> 1) * g.a_kb_bytes, ss + (qq >> 1) * KBLK, qq & 1, first + jj);
              }
            } else {
              const bool down = !PLAIN && v.resize == VQVS_RESIZE_DOWN2;
              if constexpr (!HALO) {
                const uint8_t* raw_q = raw + (q >> 1) * g.raw_kb_bytes;
                uint8_t* a_q = a_slot + (q >> 1) * g.a_kb_bytes;
                const float2* ss_q = ss + (q >> 1) * KBLK;
                if (down) transform_rows_n<true>(v, raw_q, a_q, ss_q, q & 1, row_first, nk);
                else if (PLAIN && is_skip) transform_rows_n<false, PLAIN ? TILE_M : 0>(v, raw_q, a_q, ss_q, q & 1, row_first, nk);
                else transform_rows_n<false, PLAIN ? SIMPLE_BOXW : 0>(v, raw_q, a_q, ss_q, q & 1, row_first, nk);
              } else {
                const int n_extra = v.n_rows - TILE_M;
                for (int i = lane; i < chunks * n_extra; i += 32) {
                  const int qq = i / n_extra;  // (k block, chunk)
                  const int row = TILE_M + (i - qq * n_extra);
                  const int k = qq >> 1;
                  if (down) transform_rows_n<true>(v, raw + k * g.raw_kb_bytes, a_slot + k * g.a_kb_bytes, ss + k * KBLK, qq & 1, row, 1);
                  else if (PLAIN && is_skip) transform_rows_n<false, PLAIN ? TILE_M : 0>(v, raw + k * g.raw_kb_bytes, a_slot + k * g.a_kb_bytes, ss + k * KBLK, qq & 1, row, 1);
                  else transform_rows_n<false, PLAIN ? SIMPLE_BOXW : 0>(v, raw + k * g.raw_kb_bytes, a_slot + k * g.a_kb_bytes, ss + k * KBLK, qq & 1, row, 1);
                }
              }
            }
          }
        } else if constexpr (!LEAN) {
         for (int j = 0; j < MT; ++j) {
          uint8_t* a_slot = a_slot0 + j * g.kbs * g.a_kb_bytes;
          const int t0j = t0 + j * TILE_M;
          const Src& src = is_skip ? skip_src : main_src;
          const bool act = !is_skip && d.act;
          const int pad = is_skip ? 0 : g.pad;
          const int n_rows = TILE_M + 2 * pad;
          for (int i = xtid; i < nk * 2 * n_rows; i += XFORM_WARPS * 32) {
            const int q = i / n_rows, row = i - q * n_rows;
            const int c8 = kb0 * KBLK + q * 8;
            uint8_t* a_hi = a_slot + (q >> 1) * g.a_kb_bytes + (q & 1) * (g.rows * 16);
            produce_direct(src, n, c8, t0j - pad + row, act, reinterpret_cast<const float4*>(s_ss + c8), a_hi, a_hi + g.rows * 32, row,
                           g.prec == VQVS_PREC_F16);
          }
         }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(A_FULL(ab.idx));
          if (g.tma) mbar_arrive(RAW_EMPTY(rw.idx));
        }
        ab.next();
        rw.next();
        PROF_ADD(2, tprev);
      }
    }
    PROF_STORE(0);
  };  // transform_role
  if (warp > HALO_WARP) {
    transform_role(std::false_type{});
  } else if (warp >= EPI_WARP0 + EPI_WARPS) {
  REG_DEC();  // warpgroup 2 (setmaxnreg is warpgroup-collective)
  if (warp == HALO_WARP) {
    transform_role(std::true_type{});
  } else if (warp == TMA_RAW_WARP) {
    // =========================== TMA: raw activation boxes (warp-uniform loop, elected issue) ==========
    if (g.tma) {
      Ring rw(g.raw_slots);
      const uint32_t raw_base = smem_u32(smem + g.off_raw);
      const uint32_t box_bytes = KBLK * g.main_box_w * 4;
      TILE_ITER_INIT();
      for (int k_local = 0; k_local < n_my_tiles; ++k_local, TILE_ITER_NEXT()) {
        TILE_COORDS(tile)
        (void)nt;
        const int x0m = (t0 * g.main_origin_mul) / 2 + g.main_origin_off;
        const int x0s = (t0 * g.skip_origin_mul) / 2;
        for (int st = 0; st < total_stages; ++st) {
          const bool is_skip = !SIMPLE && st >= g.main_stages;
          const int kb0 = (is_skip ? st - g.main_stages : st) * g.kbs;
          const int nk = min(g.kbs, (is_skip ? g.nkb_skip : g.nkb_main) - kb0);
          mbar_wait(RAW_EMPTY(rw.idx), rw.ph ^ 1);
          if (elect_one()) {
            const uint32_t dst0 = raw_base + rw.idx * g.raw_slot_bytes;
            if (!is_skip) {
              mbar_expect_tx(RAW_FULL(rw.idx), MT * nk * g.main_boxes * box_bytes);
              for (int j = 0; j < MT; ++j) {
                const int xj = x0m + (j * TILE_M * g.main_origin_mul) / 2;
                for (int k = 0; k < nk; ++k) {
                  const int c16 = (kb0 + k) * KBLK;
                  const bool from_a = c16 < d.c_a;
                  const CUtensorMap* map = from_a ? &tm_xa : &tm_xb;
                  const int rowc = from_a ? n * d.c_a + c16 : n * d.c_b + (c16 - d.c_a);
                  const uint32_t dst = dst0 + (j * g.kbs + k) * g.raw_kb_bytes;
                  tma_box_2d(dst, map, xj, rowc, RAW_FULL(rw.idx));
                  if (g.main_boxes == 2) tma_box_2d(dst + box_bytes, map, xj + g.main_box_w, rowc, RAW_FULL(rw.idx));
                }
              }
            } else {
              mbar_expect_tx(RAW_FULL(rw.idx), MT * nk * KBLK * g.skip_box_w * 4);
              for (int j = 0; j < MT; ++j) {
                const int xj = x0s + (j * TILE_M * g.skip_origin_mul) / 2;
                for (int k = 0; k < nk; ++k) {
                  const int c16 = (kb0 + k) * KBLK;
                  const bool from_a = c16 < d.s_a;
                  const CUtensorMap* map = from_a ? &tm_sa : &tm_sb;
                  const int rowc = from_a ? n * d.s_a + c16 : n * d.s_b + (c16 - d.s_a);
                  tma_box_2d(dst0 + (j * g.kbs + k) * g.raw_kb_bytes, map, xj, rowc, RAW_FULL(rw.idx));
                }
              }
            }
          }
          __syncwarp();
          rw.next();
        }
      }
    }
  } else if (warp == TMA_W_WARP) {
    // =========================== TMA: weight image ===========================
    if (SIMPLE || g.w_resident) {  // loaded once, reused by every tile of this CTA
      if (elect_one()) {
        const uint8_t* wimg = reinterpret_cast<const uint8_t*>(d.w_packed);
        const uint32_t total = (uint32_t)g.per_tile_bytes;
        mbar_expect_tx(W_FULL, total);
        for (uint32_t o = 0; o < total; o += 32768) {
          const uint32_t nbytes = total - o < 32768 ? total - o : 32768;
          tma_bulk_g2s(smem_u32(smem + g.off_w + o), wimg + o, nbytes, W_FULL);
        }
      }
    } else {
      Ring br(g.b_slots);
      const uint32_t b_base = smem_u32(smem + g.off_b);
      const int nkb_all = g.nkb_main + g.nkb_skip;
      TILE_ITER_INIT();
      for (int k_local = 0; k_local < n_my_tiles; ++k_local, TILE_ITER_NEXT()) {
        const int nt = it_nt;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(d.w_packed) + (size_t)nt * g.per_tile_bytes;
        for (int kb = 0; kb < nkb_all; ++kb) {  // one K block of weights per slot; the image is laid out in pipeline order
          const uint32_t bytes = kb < g.nkb_main ? g.b_unit_main : g.b_unit_skip;
          mbar_wait(B_EMPTY(br.idx), br.ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(B_FULL(br.idx), bytes);
            tma_bulk_g2s(b_base + br.idx * g.b_slot_bytes, src, bytes, B_FULL(br.idx));
          }
          __syncwarp();
          src += bytes;
          br.next();
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer (warp-uniform loop, elected issue) ===========================
    const bool f16 = !SIMPLE && g.prec == VQVS_PREC_F16;
    const uint32_t idesc = make_idesc((g.stack ? 2 : 1) * g.n_tile, f16);
    // descriptor = constant fields + (address >> 4); the address field never carries into LBO
    const uint64_t a_const = make_desc(0, g.rows * 16, 128), b_const = make_desc(0, (g.stack ? 2 : 1) * g.n_tile * 16, 128);
    const uint32_t a_lo_c = (uint32_t)a_const, a_hi32 = (uint32_t)(a_const >> 32);
    const uint32_t b_lo_c = (uint32_t)b_const, b_hi32 = (uint32_t)(b_const >> 32);
    const uint32_t a_lo_off = (g.rows * 32) >> 4, b_lo_off = (g.n_tile * 32) >> 4, b_tap_off = (g.n_tile * (f16 ? 32 : 64)) >> 4;
    const uint32_t ab_base16 = smem_u32(smem + g.off_ab) >> 4, ab_slot16 = g.ab_slot_bytes >> 4;
    const uint32_t w_base16 = smem_u32(smem + g.off_w) >> 4;
    const uint32_t unit_main16 = g.b_unit_main >> 4, unit_skip16 = g.b_unit_skip >> 4;
    const uint32_t a_kb16 = g.a_kb_bytes >> 4;
    const bool streamed = !SIMPLE && !g.w_resident;
    const uint32_t b_base16 = smem_u32(smem + g.off_b) >> 4, b_slot16 = g.b_slot_bytes >> 4;
    if (!streamed) mbar_wait(W_FULL, 0);
    Ring ab(g.ab_slots), br(g.b_slots ? g.b_slots : 1);
    PROF_DECL((dbg_flags & 512) && blockIdx.x == 0 && lane == 0);
#ifdef VQVS_PROF
    long long prof_bwait = 0;
#endif
    for (int k_local = 0; k_local < n_my_tiles; ++k_local) {
      const int buf = g.nbuf == 2 ? (k_local & 1) : 0;  // (nbuf is 1 or 2: no integer division per tile)
      PROF_ADD(3, tprev);
      mbar_wait(ACC_EMPTY(buf), ((k_local >> (g.nbuf - 1)) & 1) ^ 1);  // epilogue drained this accumulator set
      tc_fence_after();
      PROF_ADD(2, tprev);
      const uint32_t d_tmem0 = tmem_base + buf * MT * g.acc_cols;
      uint32_t acc = 0;
      uint32_t w16 = w_base16;
      for (int st = 0; st < total_stages; ++st) {
        PROF_ADD(1, tprev);
        mbar_wait(A_FULL(ab.idx), ab.ph);
        tc_fence_after();
        PROF_ADD(0, tprev);
        const bool is_skip = !SIMPLE && st >= g.main_stages;
        const int kb0 = (is_skip ? st - g.main_stages : st) * g.kbs;
        const int nk = min(g.kbs, (is_skip ? g.nkb_skip : g.nkb_main) - kb0);
        const uint32_t a16 = ab_base16 + ab.idx * ab_slot16;
        const uint32_t unit16 = is_skip ? unit_skip16 : unit_main16;
        const int taps = SIMPLE ? 3 : is_skip ? 1 : d.ksize;
        const uint32_t tap_rows = is_skip ? 0u : (uint32_t)d.dilation;  // 16-B rows per tap shift
        // Descriptor low words advance by warp-uniform 32-bit adds computed by ALL lanes (uniform datapath); the MMAs of
        // one K block and one time tile sit behind ONE elected branch (mma_group_*).
        const bool issue = elect_one();
        const bool do_mma = issue && !(dbg_flags & 4);
        uint32_t a_k = a_lo_c + a16;
        const uint32_t a_step_j = g.kbs * a_kb16;
        const uint32_t tap2 = 2 * tap_rows;
#pragma unroll 1
        for (int k = 0; k < nk; ++k) {
          uint32_t b_k;
          if (streamed) {  // this K block's weights: one slot of the weight ring
#ifdef VQVS_PROF
            const long long bw0 = prof ? clock64() : 0;
#endif
            mbar_wait(B_FULL(br.idx), br.ph);
#ifdef VQVS_PROF
            if (prof) prof_bwait += clock64() - bw0;
#endif
            tc_fence_after();
            b_k = b_lo_c + b_base16 + br.idx * b_slot16;
          } else {
            b_k = b_lo_c + w16;
            w16 += unit16;
          }
          const uint32_t accf = acc | (uint32_t)(k != 0);  // 0 only for the first MMA of each accumulator of the item
          uint32_t a_j = a_k;
#pragma unroll
          for (int j = 0; j < MT; ++j) {  // the time tiles of the item share the K block's weights
            const uint32_t d_tmem = d_tmem0 + j * g.acc_cols;
            if (f16) {
              if (taps == 3) {
                if (do_mma)
                  mma_group_single3(d_tmem, a_hi32, b_hi32, idesc, accf, a_j, a_j + tap_rows, a_j + tap2, b_k, b_k + b_tap_off,
                                    b_k + 2 * b_tap_off);
              } else {
                if (do_mma) mma_group_single1(d_tmem, a_hi32, b_hi32, idesc, accf, a_j, b_k);
              }
            } else if (taps == 3) {
              if (g.stack) {
                if (do_mma)
                  mma_group_stack3(d_tmem, a_hi32, b_hi32, idesc, accf, a_j, a_j + a_lo_off, a_j + tap_rows, a_j + tap_rows + a_lo_off,
                                   a_j + tap2, a_j + tap2 + a_lo_off, b_k, b_k + b_tap_off, b_k + 2 * b_tap_off);
              } else {
                if (do_mma)
                  mma_group_split3(d_tmem, a_hi32, b_hi32, idesc, accf, a_j, a_j + a_lo_off, a_j + tap_rows, a_j + tap_rows + a_lo_off,
                                   a_j + tap2, a_j + tap2 + a_lo_off, b_k, b_k + b_lo_off, b_k + b_tap_off, b_k + b_tap_off + b_lo_off,
                                   b_k + 2 * b_tap_off, b_k + 2 * b_tap_off + b_lo_off);
              }
            } else {
              if (g.stack) {
                if (do_mma) mma_group_stack1(d_tmem, a_hi32, b_hi32, idesc, accf, a_j, a_j + a_lo_off, b_k);
              } else {
                if (do_mma) mma_group_split1(d_tmem, a_hi32, b_hi32, idesc, accf, a_j, a_j + a_lo_off, b_k, b_k + b_lo_off);
              }
            }
            a_j += a_step_j;
          }
          if (streamed) {
            if (issue) mma_commit(B_EMPTY(br.idx));  // the weight slot is free once these MMAs have read it
            __syncwarp();
            br.next();
          }
          a_k += a_kb16;
        }
        if (issue) {
          mma_commit(AB_EMPTY(ab.idx));
          if (st == total_stages - 1) mma_commit(ACC_FULL(buf));
        }
        __syncwarp();
        acc = 1;
        ab.next();
      }
    }
    PROF_STORE(8);
#ifdef VQVS_PROF
    if (prof) g_prof[4] = prof_bwait;  // weight-ring waits (also contained in mma.issue)
#endif
  }
  } else {
    // =========================== epilogue warps ===========================
    REG_INC();
    const int quarter = warp & 3;              // TMEM lanes [32*quarter, +32) belong to this warp
    const int half = (warp - EPI_WARP0) >> 2;  // the two warps of a quarter take alternate 32-column chunks
    const int etid = threadIdx.x - EPI_WARP0 * 32;
    int staged_nt = -1, stat_n = -1, stat_nt = 0;
    // running per-channel (sum, sumsq) of this warp's rows for up to 4 chunks, flushed when the sample changes
    double rs1[2] = {0, 0}, rs2[2] = {0, 0};
    PROF_DECL((dbg_flags & 512) && blockIdx.x == 0 && etid == 0);
    const int n_chunks32 = g.n_tile / 32;
    const bool tail16 = (g.n_tile & 31) != 0;
    const bool stats = d.stats_out && !(dbg_flags & 1);
    // ---- fast path: N tile of 64 or 128 channels, statistics kept in registers across the CTA's tiles ----
    // Each thread owns one row (time position) of the tile and NCH x 32 channel columns; it accumulates
    // (sum, sumsq) per channel PAIR in fp32 registers (2 instructions per element instead of the ~8 of a
    // per-tile transposing butterfly) and reduces across the warp's 32 rows only when the sample changes or
    // every STAT_FLUSH_TILES tiles (keeps the fp32 partial sums short), then one fp64 atomic per pair.
    // statistics granularity G granted by the caller (1 = per channel): NA = 32 / G accumulators per 32-column chunk
    // (eligibility and the accumulator plan are decided on the host: epilogue_plan())
    const int nch = g.epi_nch, na = g.epi_na;
    const bool fast = LEAN || g.epi_fast;
    auto run_fast = [&](auto nch_c, auto na_c, auto skipk_c, auto w16_c) {
      constexpr int NCH = decltype(nch_c)::value;
      constexpr int NA = decltype(na_c)::value;        // statistics accumulators per 32-column chunk (granularity 32 / NA)
      // W16: a 32-channel N tile (the level-0/1 layers of base_channels = 32 models: unet32, the guidance classifier).  Each
      // warp owns 16 columns (two 8-column pieces) and keeps PER-CHANNEL statistics (32 channels / 32 groups = 1 per group).
      constexpr bool W16 = decltype(w16_c)::value;
      constexpr int NSUB = W16 ? 2 : 4;                // 8-column pieces per chunk
      constexpr int GSH = W16 ? 0 : NA == 16 ? 1 : NA == 8 ? 2 : 3;  // log2 of the granularity
      constexpr int SKIPK = decltype(skipk_c)::value;  // 0: no identity skip, 1: identity (none / nearest x2), 2: identity, pooled
      constexpr int STAT_FLUSH_TILES = 32;
      float s1[NCH][NA], s2[NCH][NA];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < NA; ++i) s1[c][i] = s2[c][i] = 0.f;
      int since_flush = 0;
      auto flush = [&](int fn, int fnt) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          float arr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) arr[i] = 0.f;
#pragma unroll
          for (int i = 0; i < NA; ++i) {
            arr[i] = s1[c][i];
            arr[NA + i] = s2[c][i];
            s1[c][i] = s2[c][i] = 0.f;
          }
          const float r = column_sums32(arr, lane);  // lane l: l < NA -> sum of granule l, l < 2 NA -> sumsq of granule l - NA
          const int gi = lane < NA ? lane : lane - NA;
          const int co = fnt * g.n_tile + (W16 ? half * 16 : (half + EPI_SPLIT * c) * 32) + (gi << GSH);
          if (lane < 2 * NA && !(dbg_flags & 2))
            atomicAdd(d.stats_out + ((size_t)fn * d.c_out + co) * 2 + (lane < NA ? 0 : 1), (double)r);
        }
        since_flush = 0;
      };
      const int row = quarter * 32 + lane;
      const int skip_shift = d.skip_resize == VQVS_RESIZE_UP2 ? 1 : 0;
      const bool stack = W16 && g.stack;  // only 32-channel N tiles of the bf16x3 format are stacked (make_geo): compile-time false elsewhere
      TILE_ITER_INIT();
      for (int k_local = 0; k_local < n_my_tiles; ++k_local, TILE_ITER_NEXT()) {
        TILE_COORDS(tile)
        if (nt != staged_nt) {  // bias slice of this N tile
          asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32));
          for (int i = etid; i < g.n_tile; i += EPI_WARPS * 32) {
            const int co = nt * g.n_tile + i;
            float b = d.bias ? d.bias[co] : 0.f;
            if (d.skip_mode == VQVS_SKIP_CONV1X1 && d.b_skip) b += d.b_skip[co];
            s_bias[i] = b;
          }
          staged_nt = nt;
          asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32));
        }
        if (stats && (n != stat_n || nt != stat_nt || since_flush >= STAT_FLUSH_TILES)) {
          if (stat_n >= 0) flush(stat_n, stat_nt);
          stat_n = n;
          stat_nt = nt;
        }
        ++since_flush;
        if (SKIPK != 0) {
          // L2 prefetch of the NEXT item's identity-skip operands (same sample, next time tiles): lane l covers
          // channel l of each of this warp's chunks, the 32 (x2 when pooled) positions of the warp's TMEM quarter
          const int tn = t0 + MT * TILE_M + quarter * 32;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int co = nt * g.n_tile + (W16 ? half * 16 + (lane & 15) : (half + EPI_SPLIT * c) * 32 + lane);
            const float* sp = co < d.s_a ? d.sa + ((size_t)n * d.s_a + co) * d.t_skip
                                         : d.sb + ((size_t)n * d.s_b + (co - d.s_a)) * d.t_skip;
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              const int ts = SKIPK == 2 ? 2 * (tn + j * TILE_M) : (tn + j * TILE_M) >> skip_shift;
              if (ts < d.t_skip) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + ts));
                if (SKIPK == 2 && ts + 32 < d.t_skip) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + ts + 32));
              }
            }
          }
        }
        const int buf = g.nbuf == 2 ? (k_local & 1) : 0;  // (nbuf is 1 or 2: no integer division per tile)
        const uint32_t acc_par = (k_local >> (g.nbuf - 1)) & 1;
        bool waited = false;
#pragma unroll 1
        for (int j = 0; j < MT; ++j) {
          const uint32_t acc_addr = tmem_base + (buf * MT + j) * g.acc_cols + ((uint32_t)(quarter * 32) << 16);
          const int t = t0 + j * TILE_M + row;
          const bool t_ok = t < d.t_out;
          // Output and identity-skip pointers WALK the tile (8 channels per piece, one 64-bit multiply-add per access): a full
          // index computation per piece cost ~12 integer instructions of the ~66 a piece takes (ncu, profiles/r2_*).
          float* op = d.out + ((size_t)n * d.c_out + nt * g.n_tile + half * (W16 ? 16 : 32)) * d.t_out + t;
          const int ts_out = d.t_out, ts_skip = d.t_skip;
          const float* skp = nullptr;
          auto skip_chunk = [&](int c0_) {  // first piece of a 32-channel chunk (chunks never straddle the two concat sources)
            const int co_ = nt * g.n_tile + c0_;
            const float* sp = co_ < d.s_a ? d.sa + ((size_t)n * d.s_a + co_) * d.t_skip
                                          : d.sb + ((size_t)n * d.s_b + (co_ - d.s_a)) * d.t_skip;
            skp = sp + (SKIPK == 1 ? (t >> skip_shift) : 2 * t);
          };
          // identity-skip operands of the next 8-column piece (raw block input, resized on the fly)
          auto load_skip = [&](float* dst) {
            const float* p = skp;
            if (SKIPK == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                dst[i] = __ldg(p);
                p += ts_skip;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 pr = __ldg(reinterpret_cast<const float2*>(p));
                dst[i] = 0.5f * (pr.x + pr.y);
                p += ts_skip;
              }
            }
            skp = p;
          };
          float sk[8], skn[8];
          // the first piece's operands are requested BEFORE waiting for the accumulator, each later piece's while
          // the previous piece is being stored (the next item's lines were already prefetched into L2 above)
          if (SKIPK != 0 && t_ok) {
            skip_chunk(half * (W16 ? 16 : 32));
            load_skip(sk);
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
              const int c0 = (W16 ? half * 16 : (half + EPI_SPLIT * c) * 32) + sub * 8;  // first column of this 8-wide piece
              if (!waited) {
                PROF_ADD(1, tprev);
                mbar_wait(ACC_FULL(buf), acc_par);
                tc_fence_after();
                waited = true;
                PROF_ADD(0, tprev);
              }
              uint32_t vr[8], wr[8];
              tmem_ld8_nowait(acc_addr + c0, vr);
              if (stack) tmem_ld8_nowait(acc_addr + g.n_tile + c0, wr);
              constexpr int kLast = NSUB * NCH - 1;
              const bool has_next = c * NSUB + sub < kLast;
              if (SKIPK != 0 && t_ok && has_next) {
                if (sub == NSUB - 1) skip_chunk((half + EPI_SPLIT * (c + 1)) * 32);
                load_skip(skn);
              }
              tmem_ld_wait();
              if (j == MT - 1 && c == NCH - 1 && sub == NSUB - 1) {  // last TMEM read of the item: hand the accumulators back
                tc_fence_before();
                mbar_arrive(ACC_EMPTY(buf));
              }
              if (t_ok && !(dbg_flags & 16)) {
                const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0), b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
                const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float v[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {  // packed fp32x2 adds: accumulator (+ lo half) + bias (+ skip)
                  uint64_t o = add2(pack2(__uint_as_float(vr[2 * i]), __uint_as_float(vr[2 * i + 1])), pack2(bias[2 * i], bias[2 * i + 1]));
                  if (stack) o = add2(o, pack2(__uint_as_float(wr[2 * i]), __uint_as_float(wr[2 * i + 1])));
                  if (SKIPK != 0) o = add2(o, pack2(sk[2 * i], sk[2 * i + 1]));
                  unpack2(o, v[2 * i], v[2 * i + 1]);
                }
                {
                  float* p = op;
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    *p = v[i];
                    p += ts_out;
                  }
                }
                if (stats && !(dbg_flags & 128)) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    s1[c][(sub * 8 + i) >> GSH] += v[i];
                    s2[c][(sub * 8 + i) >> GSH] = fmaf(v[i], v[i], s2[c][(sub * 8 + i) >> GSH]);
                  }
                }
              }
              op += (sub == NSUB - 1 && !W16 ? 8 + 32 * (EPI_SPLIT - 1) : 8) * (ptrdiff_t)ts_out;  // next piece (next chunk of this warp after 4)
              if (SKIPK != 0 && has_next) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sk[i] = skn[i];
              }
            }
          }
        }
      }
      if (stats && stat_n >= 0) flush(stat_n, stat_nt);
    };
    if (fast) {
      const int skipk = d.skip_mode != VQVS_SKIP_IDENTITY ? 0 : d.skip_resize == VQVS_RESIZE_DOWN2 ? 2 : 1;
      auto dispatch_skip = [&](auto nch_c, auto na_c, auto w16_c) {
        if (skipk == 0) run_fast(nch_c, na_c, std::integral_constant<int, 0>{}, w16_c);
        else if (skipk == 1) run_fast(nch_c, na_c, std::integral_constant<int, 1>{}, w16_c);
        else run_fast(nch_c, na_c, std::integral_constant<int, 2>{}, w16_c);
      };
      using I1 = std::integral_constant<int, 1>;
      using I2 = std::integral_constant<int, 2>;
      using I4 = std::integral_constant<int, 4>;
      using I8 = std::integral_constant<int, 8>;
      using I16 = std::integral_constant<int, 16>;
      if (g.epi_w16) dispatch_skip(I1{}, I16{}, std::true_type{});
      else if (nch == 1) dispatch_skip(I1{}, I16{}, std::false_type{});
      else if (nch == 2 && na == 16) dispatch_skip(I2{}, I16{}, std::false_type{});
      else if (nch == 2) dispatch_skip(I2{}, I8{}, std::false_type{});
      else if (na == 8) dispatch_skip(I4{}, I8{}, std::false_type{});
      else dispatch_skip(I4{}, I4{}, std::false_type{});
    } else if constexpr (!LEAN) {
    auto flush_stats = [&](int fn, int fnt) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ch = half + EPI_SPLIT * i;
        if (ch < n_chunks32) {
          double* st = d.stats_out + ((size_t)fn * d.c_out + fnt * g.n_tile + ch * 32 + lane) * 2;
          atomicAdd(st, rs1[i]);
          atomicAdd(st + 1, rs2[i]);
          rs1[i] = rs2[i] = 0.0;
        }
      }
    };
    TILE_ITER_INIT();
    for (int k_local = 0; k_local < n_my_tiles; ++k_local, TILE_ITER_NEXT()) {
      TILE_COORDS(tile)
      if (nt != staged_nt) {  // bias slice of this N tile
        asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32));
        for (int i = etid; i < g.n_tile; i += EPI_WARPS * 32) {
          const int co = nt * g.n_tile + i;
          float b = d.bias ? d.bias[co] : 0.f;
          if (d.skip_mode == VQVS_SKIP_CONV1X1 && d.b_skip) b += d.b_skip[co];
          s_bias[i] = b;
        }
        staged_nt = nt;
        asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32));
      }
      if (stats && (n != stat_n || nt != stat_nt)) {
        if (stat_n >= 0) flush_stats(stat_n, stat_nt);
        stat_n = n;
        stat_nt = nt;
      }
      const int buf = g.nbuf == 2 ? (k_local & 1) : 0;  // (nbuf is 1 or 2: no integer division per tile)
      const uint32_t acc_par = (k_local >> (g.nbuf - 1)) & 1;
      const int row = quarter * 32 + lane;
      const bool skip_id = d.skip_mode == VQVS_SKIP_IDENTITY;
      bool released = false;  // this thread's ACC_EMPTY arrival (exactly one per item)
      bool waited = false;
      uint32_t acc_addr = 0;
      int t = 0;
      bool t_ok = false;
#pragma unroll 1
     for (int j = 0; j < MT; ++j) {  // the time tiles of this item
      acc_addr = tmem_base + (buf * MT + j) * g.acc_cols + ((uint32_t)(quarter * 32) << 16);
      t = t0 + j * TILE_M + row;
      t_ok = t < d.t_out;
#pragma unroll 1
      for (int ci = 0; ci < 8 / EPI_SPLIT; ++ci) {
        const int ch = half + EPI_SPLIT * ci;
        if (ch >= n_chunks32 || (dbg_flags & 16)) break;
        const int co0 = nt * g.n_tile + ch * 32;
        // identity-skip operands are fetched BEFORE waiting for the accumulator (latency overlaps the MMAs)
        float sk[32], sk_lo[32];
        if (skip_id && t_ok) {
          const float* sp = co0 < d.s_a ? d.sa + ((size_t)n * d.s_a + co0) * d.t_skip
                                        : d.sb + ((size_t)n * d.s_b + (co0 - d.s_a)) * d.t_skip;
          if (d.skip_resize == VQVS_RESIZE_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sk[j] = __ldg(sp + (size_t)j * d.t_skip + t);
          } else if (d.skip_resize == VQVS_RESIZE_UP2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sk[j] = __ldg(sp + (size_t)j * d.t_skip + (t >> 1));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 p = __ldg(reinterpret_cast<const float2*>(sp + (size_t)j * d.t_skip + 2 * t));
              sk[j] = 0.5f * (p.x + p.y);
            }
          }
        }
        if (!waited) {
          PROF_ADD(1, tprev);
          mbar_wait(ACC_FULL(buf), acc_par);
          tc_fence_after();
          PROF_ADD(0, tprev);
          waited = true;
        }
        float v[32];
        tmem_ld32(acc_addr + ch * 32, v);
        if (g.stack) {  // [W_hi ; W_lo] products sit in two column halves
          tmem_ld32(acc_addr + g.n_tile + ch * 32, sk_lo);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += sk_lo[j];
        }
        PROF_ADD(2, tprev);
        if (j == MT - 1 && ch + EPI_SPLIT >= n_chunks32 && !(tail16 && half == 0)) {  // last TMEM read of this warp: hand the accumulators back
          tc_fence_before();
          mbar_arrive(ACC_EMPTY(buf));
          released = true;
        }
        float* outp = d.out + ((size_t)n * d.c_out + co0) * d.t_out + t;
        const float* bias = s_bias + ch * 32;
        if (t_ok) {
          if (skip_id) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += sk[j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] += bias[j];
            *outp = v[j];
            outp += d.t_out;
          }
        } else {  // rows beyond the sequence end (last tile only): no store, no statistics
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        PROF_ADD(3, tprev);
        if (stats) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sk[j] = v[j] * v[j];
          const float s2 = column_sums32(sk, lane);
          const float s1 = column_sums32(v, lane);
          // lane l <-> channel co0 + l; static indices keep the accumulators in registers
          if (ci == 0) { rs1[0] += (double)s1; rs2[0] += (double)s2; }
          else if (ci == 1) { rs1[1] += (double)s1; rs2[1] += (double)s2; }
          else {  // wide N tiles (deep, short layers): straight to global
            double* st = d.stats_out + ((size_t)n * d.c_out + co0 + lane) * 2;
            atomicAdd(st, (double)s1);
            atomicAdd(st + 1, (double)s2);
          }
        }
      }
     }  // tiles of the item
      if (!released) {
        if (!waited) {
          PROF_ADD(1, tprev);
          mbar_wait(ACC_FULL(buf), acc_par);
          tc_fence_after();
          PROF_ADD(0, tprev);
        }
        float v[16];
        const int cbase = n_chunks32 * 32;
        const bool do_tail = tail16 && half == 0;
        if (do_tail) tmem_ld16(acc_addr + cbase, v);  // (stacking is never combined with a 16-column tail)
        tc_fence_before();
        mbar_arrive(ACC_EMPTY(buf));
        if (do_tail && !(dbg_flags & 16)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = nt * g.n_tile + cbase + j;
            float o = v[j] + s_bias[cbase + j];
            if (t_ok) {
              if (skip_id) {
                const float* sp = co < d.s_a ? d.sa + ((size_t)n * d.s_a + co) * d.t_skip
                                             : d.sb + ((size_t)n * d.s_b + (co - d.s_a)) * d.t_skip;
                if (d.skip_resize == VQVS_RESIZE_NONE) o += __ldg(sp + t);
                else if (d.skip_resize == VQVS_RESIZE_UP2) o += __ldg(sp + (t >> 1));
                else { const float2 p = __ldg(reinterpret_cast<const float2*>(sp + 2 * t)); o += 0.5f * (p.x + p.y); }
              }
              d.out[((size_t)n * d.c_out + co) * d.t_out + t] = o;
            } else {
              o = 0.f;
            }
            if (stats) {  // tiny configurations only: straight to global
              const float s1 = warp_sum(o), s2 = warp_sum(o * o);
              if (lane == 0) {
                double* st = d.stats_out + ((size_t)n * d.c_out + co) * 2;
                atomicAdd(st, (double)s1);
                atomicAdd(st + 1, (double)s2);
              }
            }
          }
        }
      }
    }
    if (stats && stat_n >= 0) flush_stats(stat_n, stat_nt);
    }  // generic path
    PROF_STORE(12);
    tc_fence_before();
  }
  __syncthreads();
#ifdef VQVS_PROF
  if ((dbg_flags & 512) && blockIdx.x == 0 && threadIdx.x == 0) g_prof[5] = clock64() - prof_t0;  // whole CTA
  if ((dbg_flags & 512) && threadIdx.x == 0 && blockIdx.x < 160) g_cta[blockIdx.x * 4 + 2] = gtime_ns();
#endif
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
#undef TILE_COORDS
#undef TILE_ITER_INIT
#undef TILE_ITER_NEXT
#undef RAW_FULL
#undef RAW_EMPTY
#undef B_FULL
#undef B_EMPTY
#undef A_FULL
#undef AB_EMPTY
#undef ACC_FULL
#undef ACC_EMPTY
#undef W_FULL
}

// ---------------------------------------------------------------------------
// Translation units.  This file is compiled once per kernel KIND with -DVQVS_KIND_TU=<kind> (each unit instantiates only
// that kind's kernels and exports one launcher) and once without (packer, self-test, C ABI): the instantiations
// dominate the build time and compile in parallel this way (__graft_entry__.build()).
// ---------------------------------------------------------------------------
#define VQVS_LAUNCHER_ARGS int mt, int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap* maps, const VqvsConv* d, \
                           const Geo* g, const VqvsGnFinalize* fin
cudaError_t launch_kind0(VQVS_LAUNCHER_ARGS);
cudaError_t launch_kind1(VQVS_LAUNCHER_ARGS);
cudaError_t launch_kind2(VQVS_LAUNCHER_ARGS);
cudaError_t launch_kind3(VQVS_LAUNCHER_ARGS);
cudaError_t read_prof(unsigned long long* host32);
cudaError_t read_prof1(unsigned long long* host32);
cudaError_t read_prof2(unsigned long long* host32);
cudaError_t read_prof3(unsigned long long* host32);
#ifdef VQVS_PROF
cudaError_t read_cta0(unsigned long long* h);
cudaError_t read_cta1(unsigned long long* h);
cudaError_t read_cta2(unsigned long long* h);
cudaError_t read_cta3(unsigned long long* h);
#endif

#ifdef VQVS_KIND_TU
template <int KIND>
static cudaError_t launch_kind_impl(VQVS_LAUNCHER_ARGS) {
  // per DEVICE: the opt-in shared-memory size is an attribute of the function in the current device's context
  static bool attr_done_dev[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  bool& attr_done = attr_done_dev[dev];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<1, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess && KIND != 3)
      e = cudaFuncSetAttribute(conv_umma_kernel<(KIND == 3 ? 1 : 2), KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = !getenv("VQVS_NO_PDL");
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (mt == 2 && KIND != 3)
    return cudaLaunchKernelEx(&cfg, conv_umma_kernel<(KIND == 3 ? 1 : 2), KIND>, maps[0], maps[1], maps[2], maps[3], *d, *g, *fin);
  return cudaLaunchKernelEx(&cfg, conv_umma_kernel<1, KIND>, maps[0], maps[1], maps[2], maps[3], *d, *g, *fin);
}
#if VQVS_KIND_TU == 0
cudaError_t launch_kind0(VQVS_LAUNCHER_ARGS) { return launch_kind_impl<0>(mt, grid, smem_bytes, stream, maps, d, g, fin); }
cudaError_t read_prof(unsigned long long* host32) { return cudaMemcpyFromSymbol(host32, g_prof, 32 * sizeof(unsigned long long)); }
#ifdef VQVS_PROF
cudaError_t read_cta0(unsigned long long* h) { return cudaMemcpyFromSymbol(h, g_cta, 640 * sizeof(unsigned long long)); }
#endif
#elif VQVS_KIND_TU == 1
cudaError_t launch_kind1(VQVS_LAUNCHER_ARGS) { return launch_kind_impl<1>(mt, grid, smem_bytes, stream, maps, d, g, fin); }
cudaError_t read_prof1(unsigned long long* host32) { return cudaMemcpyFromSymbol(host32, g_prof, 32 * sizeof(unsigned long long)); }
#ifdef VQVS_PROF
cudaError_t read_cta1(unsigned long long* h) { return cudaMemcpyFromSymbol(h, g_cta, 640 * sizeof(unsigned long long)); }
#endif
#elif VQVS_KIND_TU == 2
cudaError_t launch_kind2(VQVS_LAUNCHER_ARGS) { return launch_kind_impl<2>(mt, grid, smem_bytes, stream, maps, d, g, fin); }
cudaError_t read_prof2(unsigned long long* host32) { return cudaMemcpyFromSymbol(host32, g_prof, 32 * sizeof(unsigned long long)); }
#ifdef VQVS_PROF
cudaError_t read_cta2(unsigned long long* h) { return cudaMemcpyFromSymbol(h, g_cta, 640 * sizeof(unsigned long long)); }
#endif
#else
cudaError_t launch_kind3(VQVS_LAUNCHER_ARGS) { return launch_kind_impl<3>(mt, grid, smem_bytes, stream, maps, d, g, fin); }
cudaError_t read_prof3(unsigned long long* host32) { return cudaMemcpyFromSymbol(host32, g_prof, 32 * sizeof(unsigned long long)); }
#ifdef VQVS_PROF
cudaError_t read_cta3(unsigned long long* h) { return cudaMemcpyFromSymbol(h, g_cta, 640 * sizeof(unsigned long long)); }
#endif
#endif
}  // namespace umma
}  // namespace vqvs
#else  // main unit: packer, self-test, C ABI

// ---------------------------------------------------------------------------
// weight packer: fp32 [c_out, c_in, ksize] (+ [c_out, c_skip]) -> bf16 hi/lo smem image
// ---------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ w_skip, int c_out, int c_in,
                                    int ksize, int c_skip, Geo g, uint8_t* __restrict__ img) {
  const long long n_main = (long long)c_out * c_in * ksize;
  const long long total = n_main + (long long)c_out * c_skip;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float val;
    int co, ci, tap;
    bool skip = i >= n_main;
    if (!skip) {
      tap = (int)(i % ksize);
      ci = (int)((i / ksize) % c_in);
      co = (int)(i / ((long long)ksize * c_in));
      val = w[i];
    } else {
      const long long j = i - n_main;
      tap = 0;
      ci = (int)(j % c_skip);
      co = (int)(j / c_skip);
      val = w_skip[j];
    }
    const int nt = co / g.n_tile, row = co - nt * g.n_tile;
    const int kb = ci / KBLK, chunk = (ci % KBLK) / 8, e = ci % 8;
    long long off = (long long)nt * g.per_tile_bytes;
    const bool f16 = g.prec == VQVS_PREC_F16;
    if (!skip) off += (long long)kb * g.b_unit_main + (long long)tap * (g.n_tile * (f16 ? 32 : 64));
    else off += (long long)g.nkb_main * g.b_unit_main + (long long)kb * g.b_unit_skip;
    if (f16) {  // one fp16 image: [chunk][row][16 B]
      off += (long long)chunk * (g.n_tile * 16) + (long long)row * 16 + e * 2;
      *reinterpret_cast<__half*>(img + off) = __float2half_rn(val);
      continue;
    }
    // un-stacked: [hi: chunk][row][16 B] then [lo: ...]; stacked: [chunk][rows: hi 0..n_tile-1, lo n_tile..][16 B]
    off += (long long)chunk * ((g.stack ? 2 : 1) * g.n_tile * 16) + (long long)row * 16 + e * 2;
    const __nv_bfloat16 hi = __float2bfloat16_rn(val);
    const __nv_bfloat16 lo = __float2bfloat16_rn(val - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(img + off + (g.stack ? g.n_tile * 16 : g.n_tile * 32)) = lo;
  }
}

// ---------------------------------------------------------------------------
// self-test: D[128, n] = A[shift .. shift+128, :] * B^T through the production layout
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) selftest_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ dout, int n, int k, int row_shift, int variant) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int rows = TILE_M + row_shift;
  const int nkb = k / KBLK;
  const int a_kb = rows * 64, b_kb = n * 64;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + nkb * a_kb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int cols = 32;
  while (cols < n) cols *= 2;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&holder), cols);
  for (int i = threadIdx.x; i < nkb * 2 * rows; i += 128) {
    const int q = i / rows, r = i - q * rows;
    float v[8];
    for (int e = 0; e < 8; ++e) v[e] = a[(size_t)r * k + q * 8 + e];
    uint4 hi, lo;
    split8(v, &hi, &lo);
    uint8_t* p = a_s + (q >> 1) * a_kb + (q & 1) * (rows * 16);
    *reinterpret_cast<uint4*>(p + r * 16) = hi;
    *reinterpret_cast<uint4*>(p + rows * 32 + r * 16) = lo;
  }
  for (int i = threadIdx.x; i < nkb * 2 * n; i += 128) {
    const int q = i / n, r = i - q * n;
    float v[8];
    for (int e = 0; e < 8; ++e) v[e] = b[(size_t)r * k + q * 8 + e];
    uint4 hi, lo;
    split8(v, &hi, &lo);
    uint8_t* p = b_s + (q >> 1) * b_kb + (q & 1) * (n * 16);
    *reinterpret_cast<uint4*>(p + r * 16) = hi;
    *reinterpret_cast<uint4*>(p + n * 32 + r * 16) = lo;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = holder;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n);
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t a_hi = smem_u32(a_s + kb * a_kb) + row_shift * 16, a_lo = a_hi + rows * 32;
      const uint32_t b_hi = smem_u32(b_s + kb * b_kb), b_lo = b_hi + n * 32;
      uint32_t a_l = rows * 16, a_sb = 128, b_l = n * 16, b_sb = 128;
      if (variant == 1) { uint32_t t = a_l; a_l = a_sb; a_sb = t; t = b_l; b_l = b_sb; b_sb = t; }
      mma_bf16(tmem_base, make_desc(a_hi, a_l, a_sb), make_desc(b_hi, b_l, b_sb), idesc, acc);
      acc = 1;
      if (variant != 2) {  // variant 2: hi*hi only (plain bf16) to separate layout bugs from split bugs
        mma_bf16(tmem_base, make_desc(a_lo, a_l, a_sb), make_desc(b_hi, b_l, b_sb), idesc, 1);
        mma_bf16(tmem_base, make_desc(a_hi, a_l, a_sb), make_desc(b_lo, b_l, b_sb), idesc, 1);
      }
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) dout[(size_t)row * n + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, cols);
  }
}

}  // namespace umma
}  // namespace vqvs

// =============================================================================
// C ABI
// =============================================================================
using namespace vqvs;
using vqvs::umma::Geo;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// (compute capability, SM count) of the CURRENT device, cached per device ordinal (one process may drive several GPUs)
static int device_props(int* cc, int* sms) {
  static int cached_cc[64] = {}, cached_sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    (void)cudaGetLastError();
    return vqvs_device_info(cc, sms);
  }
  if (!cached_sms[dev]) {
    int c = 0, s = 0;
    const int rc = vqvs_device_info(&c, &s);
    if (rc != VQVS_OK) return rc;
    cached_cc[dev] = c;
    cached_sms[dev] = s;
  }
  *cc = cached_cc[dev];
  *sms = cached_sms[dev];
  return VQVS_OK;
}

// TMA needs 16-B aligned bases and row pitches (length % 4 == 0); otherwise the kernel reads directly.
static bool tma_eligible(const VqvsConv* d) {
  if (d->t_in % 4 || !aligned16(d->xa) || (d->c_b && !aligned16(d->xb))) return false;
  if (d->skip_mode == VQVS_SKIP_CONV1X1 && (d->t_skip % 4 || !aligned16(d->sa) || (d->s_b && !aligned16(d->sb)))) return false;
  return true;
}

static int umma_geo(const VqvsConv* d, Geo* g) {
  const int c_skip = d->skip_mode == VQVS_SKIP_CONV1X1 ? d->s_a + d->s_b : 0;
  if (d->c_a % umma::KBLK || d->c_b % umma::KBLK) return 0;
  if (c_skip && (d->s_a % umma::KBLK || d->s_b % umma::KBLK)) return 0;
  if (d->resize == VQVS_RESIZE_DOWN2 && (d->t_in & 1)) return 0;  // paired loads need even rows
  const bool tma = tma_eligible(d);
  const int prec = (d->reserved_ >> VQVS_CONV_PREC_SHIFT) & 3;
  if (prec != VQVS_PREC_BF16X3 && prec != VQVS_PREC_F16) return 0;
  auto geo = [&](int prefer_mt) {
    if (tma && umma::make_geo(d->c_a + d->c_b, d->c_out, d->ksize, d->dilation, c_skip, d->resize, d->skip_resize, true, g, prefer_mt, prec))
      return true;
    return umma::make_geo(d->c_a + d->c_b, d->c_out, d->ksize, d->dilation, c_skip, d->resize, d->skip_resize, false, g, prefer_mt, prec);
  };
  if (!geo(0)) return 0;
  if (g->mt == 2 && d->batch > 0 && d->t_out > 0) {
    // Two time tiles per work item halve the weight streaming, but (measured, tools/prof_roles.py with VQVS_FORCE_MT)
    // single tiles win when the epilogue also fetches an identity skip, and when the coarser items waste >= 10 % of
    // the last round of the persistent schedule.
    int sms = 0, cc = 0;
    if (device_props(&cc, &sms) != VQVS_OK || sms <= 0) sms = 148;
    const long long per = (long long)g->n_tiles * d->batch;
    const long long items2 = per * ceil_div(d->t_out, 2 * umma::TILE_M), items1 = per * ceil_div(d->t_out, umma::TILE_M);
    const long long rounds2 = 2 * ((items2 + sms - 1) / sms), rounds1 = (items1 + sms - 1) / sms;
    bool want1 = d->skip_mode == VQVS_SKIP_IDENTITY || 10 * rounds1 <= 9 * rounds2;
    // Re-measured per layer with VQVS_FORCE_MT after the prologue got cheaper (tools/op_profile.py, unet64 batch 64): every
    // bf16x3 layer with a 128-channel N tile is faster with single tiles (3-25 %: finer items, both accumulator sets in
    // flight), while the fp16 layers with a 1x1 skip conv (K up to 3*512 + 1024, the longest weight streams) keep two.
    static const bool old_rule = getenv("VQVS_MT_RULE_R1") != nullptr;  // (A/B aid: the rounds-based rule alone)
    if (!old_rule && prec == VQVS_PREC_BF16X3 && g->n_tile <= 128) want1 = true;
    if (!old_rule && prec == VQVS_PREC_F16 && d->skip_mode == VQVS_SKIP_CONV1X1 && items2 >= sms) want1 = false;  // (small batches keep the finer items)
    if (want1) {
      Geo g2 = *g;
      if (!geo(1)) *g = g2;  // keep the two-tile plan if a one-tile plan does not fit
    }
  }
  return 1;
}

extern "C" int vqvs_conv1d_umma_supported(const VqvsConv* d) {
  Geo g;
  return d && (d->ksize == 1 || d->ksize == 3) && d->dilation >= 1 && d->dilation <= 32 && umma_geo(d, &g);
}

extern "C" int64_t vqvs_packed_weight_bytes(int c_out, int c_in, int ksize, int c_skip, int prec) {
  Geo g;
  if ((prec != VQVS_PREC_BF16X3 && prec != VQVS_PREC_F16) || !umma::make_geo(c_in, c_out, ksize, 1, c_skip, 0, 0, false, &g, 0, prec))
    return -1;
  return (int64_t)g.n_tiles * g.per_tile_bytes;
}

extern "C" int vqvs_pack_conv_weights(const float* w, const float* w_skip, int c_out, int c_in, int ksize, int c_skip, int prec,
                                      void* packed, void* stream) {
  Geo g;
  VQVS_CHECK_ARG(w && packed && (c_skip == 0 || w_skip), "pack_conv_weights: null pointer");
  VQVS_CHECK_ARG(prec == VQVS_PREC_BF16X3 || prec == VQVS_PREC_F16, "pack_conv_weights: unknown operand format %d", prec);
  VQVS_CHECK_ARG(umma::make_geo(c_in, c_out, ksize, 1, c_skip, 0, 0, false, &g, 0, prec),
                 "pack_conv_weights: unsupported shape c_out=%d c_in=%d k=%d skip=%d", c_out, c_in, ksize, c_skip);
  umma::pack_weights_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(w, w_skip, c_out, c_in, ksize, c_skip, g, (uint8_t*)packed);
  VQVS_CHECK_LAUNCH("vqvs_pack_conv_weights");
  return VQVS_OK;
}

static int require_sm100() {
  int cc = 0, sms = 0;
  if (device_props(&cc, &sms) != VQVS_OK) return VQVS_ECUDA;
  if (cc / 10 != 10) {
    set_error("tcgen05 path needs an sm_100-class device, found sm_%d", cc);
    return VQVS_EARCH;
  }
  return VQVS_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// Encoded tensor maps, keyed by (base pointer, rows, length, box width): a plan launches the same convs over the same
// buffers every diffusion step, and cuTensorMapEncodeTiled costs ~1 us each -- up to four per conv, 130 convs per step,
// which is what bounds small batches.  (The only global mutable state of the library besides the error string.)
#include <mutex>
#include <unordered_map>
namespace {
struct MapKey {
  const void* base;
  int rows, t, box_w;
  bool operator==(const MapKey& o) const { return base == o.base && rows == o.rows && t == o.t && box_w == o.box_w; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= (size_t)k.t * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    return h ^ ((size_t)k.box_w << 48);
  }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
}  // namespace

static int encode_map_uncached(CUtensorMap* m, const float* base, int rows, int t, int box_w);
// [rows = batch*channels][t] fp32 tensor, boxes of 16 rows x box_w positions, zero fill outside.
static int encode_map(CUtensorMap* m, const float* base, int rows, int t, int box_w) {
  const MapKey key{base, rows, t, box_w};
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *m = it->second;
      return VQVS_OK;
    }
  }
  const int rc = encode_map_uncached(m, base, rows, t, box_w);
  if (rc == VQVS_OK) {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    if (g_map_cache.size() > 8192) g_map_cache.clear();  // (a descriptor only encodes address + geometry: stale entries are harmless)
    g_map_cache.emplace(key, *m);
  }
  return rc;
}
static int encode_map_uncached(CUtensorMap* m, const float* base, int rows, int t, int box_w) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("conv(umma): cuTensorMapEncodeTiled is not available from the driver");
    return VQVS_ECUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)t, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)t * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)umma::KBLK};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv(umma): cuTensorMapEncodeTiled failed with CUresult %d (rows=%d t=%d box=%d)", (int)r, rows, t, box_w);
    return VQVS_ECUDA;
  }
  return VQVS_OK;
}

static int g_last_kind = 0;  // (diagnostic: which translation unit's profile counters vqvs_debug_prof reads)

extern "C" int vqvs_conv1d_umma(const VqvsConv* d, void* stream) {
  VQVS_CHECK_ARG(d != nullptr, "conv(umma): null descriptor");
  int rc = require_sm100();
  if (rc) return rc;
  Geo g;
  VQVS_CHECK_ARG((d->ksize == 1 || d->ksize == 3) && d->dilation >= 1 && d->dilation <= 32 && umma_geo(d, &g),
                 "conv(umma): unsupported shape c_in=%d+%d c_out=%d k=%d dil=%d", d->c_a, d->c_b, d->c_out, d->ksize, d->dilation);
  const int expect = d->resize == VQVS_RESIZE_DOWN2 ? d->t_in / 2 : d->resize == VQVS_RESIZE_UP2 ? d->t_in * 2 : d->t_in;
  VQVS_CHECK_ARG(d->t_in > 0 && d->t_out == expect && d->batch > 0, "conv(umma): bad lengths");
  VQVS_CHECK_ARG(d->xa && (d->c_b == 0 || d->xb) && d->out && d->w_packed, "conv(umma): null pointer");
  VQVS_CHECK_ARG(!d->act || d->gn || (d->scale && d->shift), "conv(umma): act=1 needs scale/shift or a fused GroupNorm");
  VqvsGnFinalize fin;
  memset(&fin, 0, sizeof(fin));
  if (d->act && d->gn) {
    fin = *d->gn;
    VQVS_CHECK_ARG(fin.c_a == d->c_a && fin.c_b == d->c_b && fin.batch == d->batch, "conv(umma): fused GroupNorm shape mismatch");
    VQVS_CHECK_ARG(fin.groups > 0 && (fin.c_a + fin.c_b) % fin.groups == 0 && fin.count > 0, "conv(umma): fused GroupNorm bad groups");
    VQVS_CHECK_ARG(fin.stats_a && (fin.c_b == 0 || fin.stats_b) && fin.gamma && fin.beta, "conv(umma): fused GroupNorm null pointer");
  }
  if (d->skip_mode != VQVS_SKIP_NONE) {
    VQVS_CHECK_ARG(d->sa && d->s_a > 0 && (d->s_b == 0 || d->sb), "conv(umma): skip sources missing");
    const int sexp = d->skip_resize == VQVS_RESIZE_DOWN2 ? d->t_skip / 2 : d->skip_resize == VQVS_RESIZE_UP2 ? d->t_skip * 2 : d->t_skip;
    VQVS_CHECK_ARG(d->t_skip > 0 && sexp == d->t_out, "conv(umma): skip length mismatch");
    VQVS_CHECK_ARG(d->skip_resize != VQVS_RESIZE_DOWN2 || (d->t_skip & 1) == 0, "conv(umma): odd skip length with pooling");
    if (d->skip_mode == VQVS_SKIP_IDENTITY)
      VQVS_CHECK_ARG(d->s_a + d->s_b == d->c_out, "conv(umma): identity skip channel mismatch");
  }
  alignas(64) CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  if (g.tma) {
    if ((rc = encode_map(&maps[0], d->xa, d->batch * d->c_a, d->t_in, g.main_box_w))) return rc;
    if (d->c_b && (rc = encode_map(&maps[1], d->xb, d->batch * d->c_b, d->t_in, g.main_box_w))) return rc;
    if (g.nkb_skip) {
      if ((rc = encode_map(&maps[2], d->sa, d->batch * d->s_a, d->t_skip, g.skip_box_w))) return rc;
      if (d->s_b && (rc = encode_map(&maps[3], d->sb, d->batch * d->s_b, d->t_skip, g.skip_box_w))) return rc;
    }
  }
  int sm_count = 0, cc_unused = 0;
  if (device_props(&cc_unused, &sm_count) != VQVS_OK) return VQVS_ECUDA;
  g.tiles_t = ceil_div(d->t_out, umma::TILE_M * g.mt);  // work items along time (mt tiles each)
  g.tiles_total = g.tiles_t * g.n_tiles * d->batch;
  int grid = sm_count < g.tiles_total ? sm_count : g.tiles_total;
  g.tiles_per_cta = ceil_div(g.tiles_total, grid);  // round-robin schedule: every SM gets a CTA
  {  // epilogue plan (mirrors the role code): statistics granularity G granted by the caller -> accumulators per chunk
    const int n_chunks32 = g.n_tile / 32, nch = n_chunks32 / umma::EPI_SPLIT;
    const bool stats = d->stats_out != nullptr && !(d->reserved_ & 1);
    const int gran_log2 = !(d->reserved_ & VQVS_CONV_PAIR_STATS) ? 0 : ((d->reserved_ >> VQVS_CONV_STAT_GRAN_SHIFT) & 15) > 1
                                                                        ? ((d->reserved_ >> VQVS_CONV_STAT_GRAN_SHIFT) & 15) : 1;
    // as few accumulators as the granularity allows, at most 32 registers per kind in total
    const int na = !stats ? 4 : nch == 1 ? 16 : nch == 2 ? (gran_log2 >= 2 ? 8 : 16) : (gran_log2 >= 3 ? 4 : 8);
    const bool gran_ok = !stats || (gran_log2 >= 1 && (32 >> gran_log2) <= na);
    g.epi_w16 = 0;
    if (g.n_tile == 32 && g.n_tiles == 1) {  // 32 output channels: 16 columns per warp, per-channel statistics in 16 accumulators
      g.epi_nch = 1;
      g.epi_na = 16;
      g.epi_w16 = 1;
      g.epi_fast = !(d->reserved_ & 32);
    } else {
    g.epi_nch = nch;
    g.epi_na = na;
    g.epi_fast = gran_ok && !(g.n_tile & 31) && (nch == 1 || nch == 2 || nch == 4) && n_chunks32 == nch * umma::EPI_SPLIT &&
                 !(d->reserved_ & 32) &&
                 !(d->skip_mode == VQVS_SKIP_IDENTITY && d->s_b && (d->s_a & 31));  // skip chunks of 32 channels stay in one source
    }
  }
#ifdef VQVS_PROF
  const bool lean = g.tma && g.epi_fast && !(d->reserved_ & 0x1FF);  // (profiling build: bit 512 keeps the production kind)
#else
  const bool lean = g.tma && g.epi_fast && !(d->reserved_ & 0x3FF);  // any profiling / ablation bit selects the generic kernel
#endif
  // (PLAIN kinds compile the staging pitches in: 136 floats for the main taps at dilation 1 or 2, 128 for the 1x1 skip)
  const bool plain = d->resize == VQVS_RESIZE_NONE && (g.nkb_skip == 0 || d->skip_resize == VQVS_RESIZE_NONE) &&
                     g.main_box_w == umma::SIMPLE_BOXW && (g.nkb_skip == 0 || g.skip_box_w == umma::TILE_M);
  const bool simple = plain && g.nkb_skip == 0 && g.w_resident && g.mt == 1 && g.prec == VQVS_PREC_BF16X3 && g.n_tile <= 64 && d->ksize == 3;
  const int kind = !lean ? 0 : simple ? 3 : plain ? 2 : 1;
  g_last_kind = kind;
  cudaError_t le = kind == 3   ? umma::launch_kind3(g.mt, grid, g.smem_bytes, (cudaStream_t)stream, maps, d, &g, &fin)
                   : kind == 2 ? umma::launch_kind2(g.mt, grid, g.smem_bytes, (cudaStream_t)stream, maps, d, &g, &fin)
                   : kind == 1 ? umma::launch_kind1(g.mt, grid, g.smem_bytes, (cudaStream_t)stream, maps, d, &g, &fin)
                               : umma::launch_kind0(g.mt, grid, g.smem_bytes, (cudaStream_t)stream, maps, d, &g, &fin);
  if (le != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("vqvs_conv1d_umma: launch failed: %s", cudaGetErrorString(le));
    return VQVS_ECUDA;
  }
  VQVS_CHECK_LAUNCH("vqvs_conv1d_umma");
  return VQVS_OK;
}

extern "C" int vqvs_debug_geo(const VqvsConv* d, int* out16) {
  VQVS_CHECK_ARG(d && out16, "debug_geo: null pointer");
  Geo g;
  VQVS_CHECK_ARG((d->ksize == 1 || d->ksize == 3) && d->dilation >= 1 && d->dilation <= 32 && umma_geo(d, &g),
                 "debug_geo: shape not supported by the tcgen05 path");
  const int v[16] = {g.n_tiles, g.n_tile, g.stack, g.w_resident, g.kbs, g.mt, g.nbuf, g.ab_slots, g.ab_slot_bytes, g.raw_slots,
                     g.raw_slot_bytes, g.smem_bytes, g.tma, g.main_stages, g.skip_stages, g.b_slots};
  for (int i = 0; i < 16; ++i) out16[i] = v[i];
  return VQVS_OK;
}

extern "C" int vqvs_debug_prof(unsigned long long* host32) {
  cudaError_t e = g_last_kind == 3 ? umma::read_prof3(host32) : g_last_kind == 2 ? umma::read_prof2(host32)
                  : g_last_kind == 1 ? umma::read_prof1(host32) : umma::read_prof(host32);
  if (e != cudaSuccess) {
    set_error("vqvs_debug_prof: %s", cudaGetErrorString(e));
    return VQVS_ECUDA;
  }
  return VQVS_OK;
}

#ifdef VQVS_PROF
// profiling builds only: per-CTA (smid, start ns, end ns, tiles) of the last profiled launch (flag 512)
extern "C" int vqvs_debug_cta(unsigned long long* host640) {
  cudaError_t e = g_last_kind == 3 ? umma::read_cta3(host640) : g_last_kind == 2 ? umma::read_cta2(host640)
                  : g_last_kind == 1 ? umma::read_cta1(host640) : umma::read_cta0(host640);
  return e == cudaSuccess ? VQVS_OK : VQVS_ECUDA;
}
#endif

extern "C" int vqvs_umma_selftest(const float* a, const float* b, float* dout, int n, int k, int row_shift, int variant,
                                  void* stream) {
  int rc = require_sm100();
  if (rc) return rc;
  VQVS_CHECK_ARG(a && b && dout && n >= 16 && n <= 256 && n % 16 == 0 && k > 0 && k % 16 == 0 && row_shift >= 0,
                 "umma_selftest: bad arguments");
  const size_t smem = (size_t)(k / 16) * ((128 + row_shift) * 64 + n * 64);
  VQVS_CHECK_ARG(smem <= 200 * 1024, "umma_selftest: problem too large for one CTA");
  cudaError_t e = cudaFuncSetAttribute(umma::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("umma_selftest: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return VQVS_ECUDA;
  }
  umma::selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, dout, n, k, row_shift, variant);
  VQVS_CHECK_LAUNCH("vqvs_umma_selftest");
  return VQVS_OK;
}
#endif  // VQVS_KIND_TU
