// Dense projections of the hot path (in_proj / x_proj / out_proj / PatchMerging / PatchExpand / decoder linears,
// vmamba.py:679,195,725,616; MambaDecoder.py:17,39,82-83) as ONE hand-written sm_100a GEMM:
//
//     C[M,N] = A[M,K] · W[N,K]^T  (+ bias[N])  (+ residual[M,N] · rscale[N])      fp32 in / fp32 out
//
// 5th-generation tensor cores: `tcgen05.mma.cta_group::1.kind::tf32` issued by one elected thread, fp32 operands
// read as TF32 straight from shared memory (no conversion pass), fp32 accumulators in TMEM, operand tiles staged
// by TMA (128-byte swizzle) through a full/empty mbarrier ring, accumulators read back with `tcgen05.ld` and the
// bias / residual epilogue fused before the store.  Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM
// allocator + MMA issuer, warps 2-9 = epilogue (two per TMEM lane quarter `warp % 4`, alternating column chunks).
// Persistent CTAs loop over 128 x BN output tiles (BN <= 256, multiple of 16, a runtime value chosen per call) with
// two TMEM accumulator stages, so the loads and MMAs of the next tile overlap the epilogue of the current one.
// These GEMMs are HBM-bound at the Sigma shapes (K = 96..1536): what matters is one pass over A and one over C.
//
// X3 = true ("tf32x3", sigma_linear_tf32x3): fp32-grade products on the TF32 tensor pipe by error compensation,
//     A·W = A_hi·W_hi + A_lo·W_hi + A_hi·W_lo  (+ A_lo·W_lo ~ 2^-22, dropped),   x_hi = x with the low 13 mantissa bits cleared.
// CONV = true (sigma_conv3x3_tf32): the SAME kernel as an implicit-GEMM 3x3 convolution (pad 1, stride 1) over a channels-last
// (B, H, W, Cin) tensor: an M tile is an 8 x 16 pixel patch, the K loop runs over 9 taps x Cin blocks, and the A tile of tap
// (dy, dx) is ONE 4-D TMA box of the input shifted by (dy-1, dx-1) whose out-of-bounds fill is the zero padding (no im2col
// tensor); weights are (9, Cout, Cin); bias and an optional exact GELU are fused in the epilogue, whose per-warp store box is
// {32 channels, 16, 2 rows, 1}.  Replaces the ChannelAttentionBlock's two dense convolutions (vmamba.py:1749-1752).
//
// W_hi / W_lo come pre-split from the caller (weights: split once, sigma_split_tf32_fwd); the activations are split in
// shared memory by four extra warps between the TMA arrival and the MMA issue (hi written in place — exactly representable,
// so the tensor core's own TF32 conversion is the identity on it — lo into a second tile of the stage), and the issuer
// launches three MMAs per k-step.  The plain kernel's code is untouched by the template.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tma.cuh"

namespace sigma {

constexpr int GM_BM = 128;   // rows per CTA tile = UMMA_M
constexpr int GM_BK = 32;    // fp32 elements per k-block = one 128-byte swizzle row
constexpr int GM_UK = 8;     // UMMA_K for kind::tf32 (32 bytes)

struct alignas(64) GemmParams {
  CUtensorMap m_a, m_w, m_c, m_wlo;
  const float *bias, *residual, *rscale;
  float *C;
  long long ldr, ldc;
  int M, N, K, BN, stages, tmem_cols;
  int x3_keep_hi;   // X3: 1 = the splitter also rewrites the activations' hi part in place (needed only if the tensor core ROUNDS to TF32)
  int cH, cW, tiles_w, tiles_hw, kbc, act;   // CONV: image size, 8x16 patches per row / per image, Cin blocks per tap, activation
};

constexpr int CV_TH = 8, CV_TW = 16;   // pixel patch of one M tile (8 x 16 = GM_BM)

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_load_4d_g(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d_g(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows are 128 B apart, 8-row groups 1024 B apart (SBO), descriptor
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  Advancing along K inside the swizzle row = +32 B on the
// start address per UMMA_K.
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem) {
  const uint32_t addr = smem_u32(smem);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);          // start address, 16-byte units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

// instruction descriptor: D = fp32 (c_format 1), A = B = TF32 (format 2), both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

template <bool X3, bool CONV>
__global__ void __launch_bounds__(X3 ? 448 : 320) gemm_tf32_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int BN = p.BN, S = p.stages;
  const int a_bytes = GM_BM * GM_BK * 4, b_bytes = BN * GM_BK * 4;
  const int half_bytes = a_bytes + ((b_bytes + 1023) & ~1023);   // [A | W] ; X3: a second [A_lo | W_lo] follows
  const int stage_bytes = X3 ? 2 * half_bytes : half_bytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)S * stage_bytes);
  uint64_t *empty = full + S;
  uint64_t *acc_full = empty + S;      // [2]: accumulator stage complete (MMA -> epilogue)
  uint64_t *acc_empty = acc_full + 2;  // [2]: accumulator stage drained (epilogue -> MMA)
  uint64_t *split = acc_empty + 2;     // [S] (X3): the activations of the stage are split into hi / lo
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(split + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = CONV ? 9 * p.kbc : (p.K + GM_BK - 1) / GM_BK;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_m = CONV ? p.M : (p.M + GM_BM - 1) / GM_BM;      // CONV: p.M counts pixel patches
  const long long total = (long long)num_m * num_n;
  const int acc_cols = p.tmem_cols >> 1;   // columns per accumulator stage

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.m_a);
    tma_prefetch_desc(&p.m_w);
    tma_prefetch_desc(&p.m_c);
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&split[s], 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_mbar_init();
  }
  if (warp == 1) {  // TMEM allocation is a warp-wide operation; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Persistent CTA: tiles blockIdx.x, +gridDim.x, ... (n fastest, so consecutive tiles re-read the same A rows from
  // L2).  The three roles run their own loops over the same tile sequence and meet only at the mbarriers, so the
  // TMA loads of tile i+1 and the MMAs of tile i+1 overlap the epilogue of tile i (two TMEM accumulator stages).
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      long long it = 0;
      for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m0 = (int)(tile / num_n) * GM_BM, n0 = (int)(tile % num_n) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = (int)(it % S);
          mbar_wait_backoff(&empty[st], (uint32_t)(((it / S) & 1) ^ 1));
          unsigned char *sa = smem_raw + (size_t)st * stage_bytes;
          mbar_arrive_expect_tx(&full[st], (uint32_t)(a_bytes + (X3 ? 2 : 1) * b_bytes));
          if (CONV) {
            const int mt = (int)(tile / num_n);
            const int b = mt / p.tiles_hw, r = mt - b * p.tiles_hw, ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
            const int tap = kb / p.kbc, kc = kb - tap * p.kbc, dy = tap / 3, dx = tap - 3 * dy;
            // the tap's input patch: shifted by (dy-1, dx-1); rows / columns outside the image arrive as zeros = the padding
            tma_load_4d_g(sa, &p.m_a, &full[st], kc * GM_BK, tx * CV_TW + dx - 1, ty * CV_TH + dy - 1, b);
            tma_load_2d(sa + a_bytes, &p.m_w, &full[st], kc * GM_BK, tap * p.N + n0);
            if (X3) tma_load_2d(sa + half_bytes + a_bytes, &p.m_wlo, &full[st], kc * GM_BK, tap * p.N + n0);
          } else {
            tma_load_2d(sa, &p.m_a, &full[st], kb * GM_BK, m0);
            tma_load_2d(sa + a_bytes, &p.m_w, &full[st], kb * GM_BK, n0);
            if (X3) tma_load_2d(sa + half_bytes + a_bytes, &p.m_wlo, &full[st], kb * GM_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(GM_BM, BN);
      long long it = 0, tc = 0;
      for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, ++tc) {
        const int acc = (int)(tc & 1);
        mbar_wait(&acc_empty[acc], (uint32_t)(((tc >> 1) & 1) ^ 1));   // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = (int)(it % S);
          mbar_wait(X3 ? &split[st] : &full[st], (uint32_t)((it / S) & 1));
          tc_fence_after();
          const unsigned char *sa = smem_raw + (size_t)st * stage_bytes;
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + a_bytes);
          if (X3) {
            const uint64_t dal = umma_desc_sw128(sa + half_bytes), dbl = umma_desc_sw128(sa + half_bytes + a_bytes);
#pragma unroll
            for (int k = 0; k < GM_BK / GM_UK; ++k) {   // small terms first
              umma_tf32(tmem_d, dal + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
              umma_tf32(tmem_d, da + (uint64_t)(2 * k), dbl + (uint64_t)(2 * k), idesc, 1u);
              umma_tf32(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < GM_BK / GM_UK; ++k)   // +32 B along K inside the swizzle row = +2 in 16-byte units
              umma_tf32(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[st]);                   // frees the smem slot once these MMAs have read it
        }
        umma_commit(&acc_full[acc]);                 // accumulator of this tile complete
      }
    }
  } else if (X3 && warp >= 10) {
    // ===== activation splitters (X3): A tile -> (hi in place, lo in the stage's second half), elementwise at the same
    // swizzled offsets, then the async proxy (tensor core) may read both =====
    const int t = threadIdx.x - 320;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int st = (int)(it % S);
        mbar_wait(&full[st], (uint32_t)((it / S) & 1));
        uint4 *hi = reinterpret_cast<uint4 *>(smem_raw + (size_t)st * stage_bytes);
        float4 *lo = reinterpret_cast<float4 *>(smem_raw + (size_t)st * stage_bytes + half_bytes);
#pragma unroll
        for (int i = 0; i < GM_BM * GM_BK / 4 / 128; ++i) {
          const uint4 v = hi[t + 128 * i];
          const uint4 h = make_uint4(v.x & 0xFFFFE000u, v.y & 0xFFFFE000u, v.z & 0xFFFFE000u, v.w & 0xFFFFE000u);
          if (p.x3_keep_hi) hi[t + 128 * i] = h;
          lo[t + 128 * i] = make_float4(__uint_as_float(v.x) - __uint_as_float(h.x), __uint_as_float(v.y) - __uint_as_float(h.y),
                                        __uint_as_float(v.z) - __uint_as_float(h.z), __uint_as_float(v.w) - __uint_as_float(h.w));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split[st]);
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> (+bias, +residual·rscale) -> swizzled smem -> TMA store =====
    // Each warp owns 32 rows (its TMEM lane quarter) and moves them out in chunks of 32 columns: one thread = one
    // row, its 32 values go to a 128-byte shared row (16-byte chunks XOR-swizzled by row & 7, the layout the
    // SWIZZLE_128B tensor map expects), then ONE TMA store writes the 32 x 32 box as 128-byte rows.  TMA clips the
    // box at M and N, so ragged tiles need no masks.  One 4 KB staging buffer per warp (the tcgen05.ld of the next
    // chunk overlaps the previous store), which leaves room for a 4-deep operand ring at BN = 256.
    const int quarter = warp & 3;                    // TMEM lanes 32·quarter .. +31 are accessible to this warp
    const int chalf = (warp - 2) >> 2;               // two warps per lane quarter: even / odd 32-column chunks
    unsigned char *stage_c = smem_raw + (size_t)S * stage_bytes + 1024 + (size_t)(warp - 2) * 4096;
    long long tc = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, ++tc) {
      const int m0 = (int)(tile / num_n) * GM_BM, n0 = (int)(tile % num_n) * BN;
      const int acc = (int)(tc & 1);
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = CONV ? false : row < p.M;
      const float *rrow = (!CONV && p.residual && row_ok) ? p.residual + (long long)row * p.ldr : nullptr;
      // the residual row chunk is fetched one chunk ahead (the first one before the accumulator is even ready): a
      // thread-per-row read is 32 sectors per request and would otherwise sit between tcgen05.ld and the store
      float4 rpre[8];
      auto fetch_residual = [&](int n) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          rpre[q] = (rrow && n + 4 * q < p.N) ? __ldg(reinterpret_cast<const float4 *>(rrow + n + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (rrow) fetch_residual(n0 + 32 * chalf);
      mbar_wait(&acc_full[acc], (uint32_t)((tc >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols);
      for (int c0 = 32 * chalf; c0 < BN; c0 += 64) {
        const int n = n0 + c0;
        if (n >= p.N) break;                          // warp-uniform
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);          // warp-wide
        float4 rcur[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) rcur[q] = rpre[q];
        if (rrow && c0 + 64 < BN && n + 64 < p.N) fetch_residual(n + 64);
        float4 *dst = reinterpret_cast<float4 *>(stage_c + lane * 128);
        if (lane == 0) tma_store_wait_read<0>();     // the previous store of this warp has finished reading the buffer
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          const int nn = n + 4 * q;
          if (nn < p.N) {                             // N % 4 == 0, so a float4 is all-in or all-out
            if (p.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + nn));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (rrow) {
              const float4 r = rcur[q];
              if (p.rscale) {
                const float4 sc = __ldg(reinterpret_cast<const float4 *>(p.rscale + nn));
                o.x = fmaf(r.x, sc.x, o.x); o.y = fmaf(r.y, sc.y, o.y); o.z = fmaf(r.z, sc.z, o.z); o.w = fmaf(r.w, sc.w, o.w);
              } else {
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
              }
            }
          }
          if (CONV && p.act == 1) {   // nn.GELU() (exact, erf)
            o.x = 0.5f * o.x * (1.f + erff(o.x * 0.70710678118654752f)); o.y = 0.5f * o.y * (1.f + erff(o.y * 0.70710678118654752f));
            o.z = 0.5f * o.z * (1.f + erff(o.z * 0.70710678118654752f)); o.w = 0.5f * o.w * (1.f + erff(o.w * 0.70710678118654752f));
          }
          dst[q ^ (lane & 7)] = o;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (CONV) {   // this warp's 32 rows = patch rows 2·quarter, 2·quarter + 1 (16 pixels each)
            const int mt = (int)(tile / num_n);
            const int b = mt / p.tiles_hw, r = mt - b * p.tiles_hw, ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
            tma_store_4d_g(&p.m_c, stage_c, n, tx * CV_TW, ty * CV_TH + 2 * quarter, b);
          } else {
            tma_store_2d(&p.m_c, stage_c, n, m0 + quarter * 32);
          }
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);    // this warp no longer reads the accumulator stage
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host ----
typedef CUresult (*EncodeTiledFn2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void *get_tensor_map_encoder();  // ss2d_scan_host.cu

static int make_tmap_2d_sw128(CUtensorMap *map, const float *base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn2 fn = (EncodeTiledFn2)get_tensor_map_encoder();
  if (!fn) return SIGMA_ECUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t bdim[2] = {(cuuint32_t)GM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (gemm operand) failed (CUresult %d): rows=%lld cols=%lld ld=%lld box_rows=%d base=%p",
              (int)r, rows, cols, ld, box_rows, (const void *)base);
    return SIGMA_ECUDA;
  }
  return SIGMA_OK;
}

// BN is a multiple of 32 (the epilogue moves 32-column boxes); tiles that overhang N are clipped by TMA on both the
// W load (zero fill) and the C store.
// With many row tiles (the throughput regime) this is the widest divisor of N (the A tile is re-read once per column tile);
// with few (small batch: M of a few hundred rows) narrower tiles put more CTAs to work: cost = waves over the 148 persistent
// CTAs x (bn + 128), the shared-memory bytes an MMA k-step reads for a 128 x bn tile.
static int pick_bn(int N, long long m_tiles = 1 << 30) {
  const int full = N <= 256 ? ((N + 31) / 32) * 32 : 256;
  int best_bn = full;
  long long best = -1;
  for (int bn = full; bn >= 64; bn -= 32) {
    const long long tiles = m_tiles * ((N + bn - 1) / bn);    // a last column tile that overhangs N is clipped by TMA, but costs a full tile
    const long long cost = ((tiles + 147) / 148) * (bn + 128);
    if (best < 0 || cost < best || (cost == best && N % bn == 0 && N % best_bn != 0)) { best = cost; best_bn = bn; }
  }
  return best_bn;
}

int gemm_pick_bn_hook(int N, long long m_tiles) { return pick_bn(N, m_tiles); }   // api.cu test hook

static int make_tmap_c(CUtensorMap *map, const float *base, long long rows, long long cols, long long ld) {
  EncodeTiledFn2 fn = (EncodeTiledFn2)get_tensor_map_encoder();
  if (!fn) return SIGMA_ECUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t bdim[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (gemm output) failed (CUresult %d): rows=%lld cols=%lld ld=%lld base=%p", (int)r, rows,
              cols, ld, (const void *)base);
    return SIGMA_ECUDA;
  }
  return SIGMA_OK;
}

// x -> (hi = x with the low 13 mantissa bits cleared, lo = x - hi): the operand split of the tf32x3 GEMM, for the weights
__global__ void split_tf32_kernel(const float *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

int split_tf32_launch(const float *x, float *hi, float *lo, long long n, cudaStream_t stream) {
  if (n == 0) return SIGMA_OK;
  split_tf32_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(x, hi, lo, n);
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}

// The activation splitter does not rewrite the hi part by default: kind::tf32 TRUNCATES its fp32 operands on B200 (the x3 parity
// tests pass at 4e-6 either way, round 2 call 8), so x and x_hi give the same MMA.  SIGMA_X3_KEEP_HI=1 restores the rewrite
// (needed if a device rounded to nearest instead).
static int x3_keep_hi() {
  static const int v = [] { const char *e = getenv("SIGMA_X3_KEEP_HI"); return e ? atoi(e) : 0; }();
  return v;
}

// W_lo == nullptr: plain TF32 (one MMA per k-step); else tf32x3 with W = W_hi
int gemm_tf32_launch(const float *A, long long lda, const float *W, const float *W_lo, const float *bias, const float *residual,
                     long long ldr, const float *rscale, float *C, long long ldc, long long M, int N, int K, cudaStream_t stream) {
  if (M == 0) return SIGMA_OK;
  const bool x3 = W_lo != nullptr;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.x3_keep_hi = x3_keep_hi();
  p.bias = bias; p.residual = residual; p.rscale = rscale; p.C = C; p.ldr = ldr; p.ldc = ldc;
  p.M = (int)M; p.N = N; p.K = K;
  p.BN = pick_bn(N, (M + GM_BM - 1) / GM_BM);
  if (const char *e = getenv("SIGMA_GEMM_BN_RULE")) { if (e[0] == 'o') p.BN = pick_bn(N); }   // "old": ignore the row-tile count
  p.tmem_cols = 2 * (p.BN <= 16 ? 16 : p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256);   // two accumulator stages
  if (p.tmem_cols < 32) p.tmem_cols = 32;
  int rc;
  if ((rc = make_tmap_2d_sw128(&p.m_a, A, M, K, lda, GM_BM))) return rc;
  if ((rc = make_tmap_2d_sw128(&p.m_w, W, N, K, K, p.BN))) return rc;
  if (x3 && (rc = make_tmap_2d_sw128(&p.m_wlo, W_lo, N, K, K, p.BN))) return rc;
  if ((rc = make_tmap_c(&p.m_c, C, M, N, ldc))) return rc;
  const int a_bytes = GM_BM * GM_BK * 4, b_bytes = p.BN * GM_BK * 4;
  const int stage_bytes = (a_bytes + ((b_bytes + 1023) & ~1023)) * (x3 ? 2 : 1);
  p.stages = std::max(2, std::min(8, (192 * 1024) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 /*barriers*/ + 8 * 4096 /*epilogue staging*/;
  const void *kern = x3 ? (const void *)gemm_tf32_kernel<true, false> : (const void *)gemm_tf32_kernel<false, false>;
  SIGMA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)((N + p.BN - 1) / p.BN) * ((M + GM_BM - 1) / GM_BM);
  const int ctas_per_sm = std::max(1, std::min(std::min(2, 512 / p.tmem_cols), (int)((220 * 1024) / (smem + 1024))));
  const unsigned grid = (unsigned)std::min<long long>(total, 148LL * ctas_per_sm);   // persistent CTAs
  if (x3) gemm_tf32_kernel<true, false><<<grid, 448, smem, stream>>>(p);
  else gemm_tf32_kernel<false, false><<<grid, 320, smem, stream>>>(p);
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}

int make_tmap_generic(CUtensorMap *map, CUtensorMapDataType dtype, int rank, const void *base, const uint64_t *dims,
                      const uint64_t *strides_bytes, const uint32_t *box, CUtensorMapSwizzle swz, CUtensorMapL2promotion promo);   // scan_op_tma.cu

// 3x3 convolution, pad 1, stride 1, channels-last: y (B,H,W,Cout) = conv(x (B,H,W,Cin), W9 (9, Cout, Cin)) + bias, optional GELU.
// W9_lo == nullptr: plain TF32; else tf32x3 with W9 = W9_hi.
int conv3x3_tf32_launch(const float *x, const float *W9, const float *W9_lo, const float *bias, int act, float *y, int B, int H, int W,
                        int Cin, int Cout, cudaStream_t stream) {
  if (B == 0) return SIGMA_OK;
  const bool x3 = W9_lo != nullptr;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.x3_keep_hi = x3_keep_hi();
  p.bias = bias; p.C = y;
  p.N = Cout; p.K = Cin;
  p.cH = H; p.cW = W; p.act = act;
  p.tiles_w = (W + CV_TW - 1) / CV_TW;
  p.tiles_hw = p.tiles_w * ((H + CV_TH - 1) / CV_TH);
  p.M = B * p.tiles_hw;
  p.kbc = (Cin + GM_BK - 1) / GM_BK;
  p.BN = pick_bn(Cout);
  p.tmem_cols = 2 * (p.BN <= 16 ? 16 : p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256);
  if (p.tmem_cols < 32) p.tmem_cols = 32;
  int rc;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {(uint32_t)GM_BK, (uint32_t)CV_TW, (uint32_t)CV_TH, 1};
    if ((rc = make_tmap_generic(&p.m_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B))) return rc;
    uint64_t dimc[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strc[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    uint32_t boxc[4] = {32, (uint32_t)CV_TW, 2, 1};
    if ((rc = make_tmap_generic(&p.m_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, dimc, strc, boxc, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE))) return rc;
  }
  if ((rc = make_tmap_2d_sw128(&p.m_w, W9, 9LL * Cout, Cin, Cin, p.BN))) return rc;
  if (x3 && (rc = make_tmap_2d_sw128(&p.m_wlo, W9_lo, 9LL * Cout, Cin, Cin, p.BN))) return rc;
  const int a_bytes = GM_BM * GM_BK * 4, b_bytes = p.BN * GM_BK * 4;
  const int stage_bytes = (a_bytes + ((b_bytes + 1023) & ~1023)) * (x3 ? 2 : 1);
  p.stages = std::max(2, std::min(8, (192 * 1024) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 8 * 4096;
  const void *kern = x3 ? (const void *)gemm_tf32_kernel<true, true> : (const void *)gemm_tf32_kernel<false, true>;
  SIGMA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)((Cout + p.BN - 1) / p.BN) * p.M;
  const int ctas_per_sm = std::max(1, std::min(std::min(2, 512 / p.tmem_cols), (int)((220 * 1024) / (smem + 1024))));
  const unsigned grid = (unsigned)std::min<long long>(total, 148LL * ctas_per_sm);
  if (x3) gemm_tf32_kernel<true, true><<<grid, 448, smem, stream>>>(p);
  else gemm_tf32_kernel<false, true><<<grid, 320, smem, stream>>>(p);
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}

}  // namespace sigma
