// a4+a5 — fused multi-direction SS2D scan, channels-last (see include/sigma_b200.h: sigma_ss2d_scan_fwd).
//
// One launch does, for every direction k of an SS2D block:
//   CrossScan (index math: each direction is a walk over the SAME channels-last tensor, expressed as a
//   4-D TMA tensor map — row-major = tiles along L, column-major = tiles along H at fixed w, reversed =
//   the same tiles walked backwards) -> dt_proj (R-term dot product, W_dt row in registers) -> +bias ->
//   softplus -> selective scan (state in registers, one MUFU.EX2 per element) -> D skip -> store at the
//   POSITION the value belongs to (so CrossMerge's un-flip / un-transpose disappear).
//
// Mapping: LPC lanes per channel (template; 1, 2 or 4), each holding N/LPC states in registers; the lanes of a
// channel are adjacent, a warp covers 32/LPC consecutive channels, B/C/dt_r are broadcast shared reads.
// LPC = 1 has no shuffles and no redundant work: per (channel, position) N x (FMUL, MUFU.EX2, FMUL, FFMA, FFMA)
// + R FFMA (dt_proj) + softplus (measured 150 thread-instr per (channel, position) at N=16, i.e. 4.7 issue
// clk vs 4.5 MUFU clk per sub-partition) but exposes only batch x K x D/32 warps; LPC = 2 / 4 trade a few
// shuffles for 2x / 4x the warps when that product cannot fill 148 SMs (profiles/r01_scan_*.txt).
// CTA = (channel tile DT, direction k [x L-segment], image b) = DT/32 consumer warps + one TMA producer warp.
// Tiles of LT scan positions are staged HBM -> shared by TMA (cp.async.bulk.tensor) through an NST-deep ring
// guarded by full/empty mbarriers, so consumer warps never wait for each other (no CTA-wide barrier in the
// loop); y goes straight from registers to HBM (a warp writes 32 consecutive channels of one position = one
// 128-byte row).
// Per group of 4 positions the delta' of the NEXT group is computed while the recurrence of the current one
// runs (software pipelining: the only serial dependency is the fma h = a·h + b).
#pragma once
#include "scan_core.cuh"
#include "tma.cuh"

namespace sigma {

template <int N>
struct Ss2dCfg {
  static constexpr int LT = N >= 16 ? 16 : 32;   // scan positions per tile
  static constexpr int NST = N >= 16 ? 4 : 3;    // TMA ring depth
};

struct alignas(64) Ss2dParams {
  CUtensorMap m_xc[4], m_dbl[4];
  const float *dtw, *dtb, *A, *Ds;
  float *y, *carry;
  int D, N, R, Cp, kind, batch, ndir;
  long long Lseq;
  int I[4], O[4], rev[4];
  long long istride[4], ostride[4];   // y element strides of the inner / outer walk index
  int nsplit, tiles_per_split;
};

__host__ __device__ inline size_t ss2d_smem_bytes(int LT, int DT, int NST, int Cp, bool cross) {
  const size_t stage = (size_t)LT * DT + (size_t)LT * Cp * (cross ? 2 : 1);
  return NST * stage * sizeof(float) + 128 /*barriers*/;
}

template <int SPT, int RP>
struct Ss2dThread {
  float h[SPT], a2[SPT], W[RP];
  float bias, Dv, sumdl;
  int ch, q, lane;   // channel within the CTA tile, lane within the channel's group, lane within the warp
  bool ok;
};

// delta' and u for the 4 positions of group j (tile rows 4j..4j+3).  The LPC lanes of a channel split the
// four dot-product + softplus evaluations between them and exchange the results by shuffle.
template <int N, int LPC, int RP>
__device__ __forceinline__ void group_prologue(const Ss2dThread<N / LPC, RP> &t, const float *sXC, const float *sDB, int DT,
                                               int j, float (&dl)[4], float (&u)[4]) {
  constexpr int Cp = 2 * N + RP;  // x_dbl row length: [B | C | dt_r padded to RP] (sigma_ss2d_padded_cp)
  constexpr int PPL = 4 / LPC;    // positions evaluated by this lane
  float own[PPL];
#pragma unroll
  for (int e = 0; e < PPL; ++e) {
    const float *row = sDB + (4 * j + t.q * PPL + e) * Cp + 2 * N;
    float acc0 = t.bias, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;   // 4 accumulators: short dependency chains
#pragma unroll
    for (int c = 0; c < RP / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(row + 4 * c);   // broadcast read
      acc0 = fmaf(t.W[4 * c + 0], v.x, acc0);
      acc1 = fmaf(t.W[4 * c + 1], v.y, acc1);
      acc2 = fmaf(t.W[4 * c + 2], v.z, acc2);
      acc3 = fmaf(t.W[4 * c + 3], v.w, acc3);
    }
    own[e] = softplus20((acc0 + acc1) + (acc2 + acc3));
  }
  if (LPC == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dl[i] = own[i % PPL];
  } else {
    const int base = t.lane & ~(LPC - 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) dl[i] = __shfl_sync(0xffffffffu, own[i % PPL], base + i / PPL);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) u[e] = sXC[(4 * j + e) * DT + t.ch];
}

// recurrence over `cnt` (<= 4) positions of group j, in walk order (REV: descending tile rows)
template <int N, int LPC, int RP, bool WITH_Y, bool REV, bool FULL>
__device__ __forceinline__ void group_body(Ss2dThread<N / LPC, RP> &t, const float *sDB, const float *sDC, float *yrow,
                                           long long ystride, int j, const float (&dl)[4], const float (&u)[4],
                                           int cnt) {
  constexpr int Cp = 2 * N + RP;
  constexpr int SPT = N / LPC;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = REV ? 3 - ii : ii;
    if (FULL || i < cnt) {
      float Bs[SPT], Cs[SPT];
      const float *rb = sDB + (4 * j + i) * Cp + t.q * SPT;
      const float *rc = sDC + (4 * j + i) * Cp + N + t.q * SPT;
#pragma unroll
      for (int s4 = 0; s4 < SPT / 4; ++s4) {
        const float4 bv = *reinterpret_cast<const float4 *>(rb + 4 * s4);   // broadcast reads
        Bs[4 * s4] = bv.x; Bs[4 * s4 + 1] = bv.y; Bs[4 * s4 + 2] = bv.z; Bs[4 * s4 + 3] = bv.w;
        if (WITH_Y) {
          const float4 cv = *reinterpret_cast<const float4 *>(rc + 4 * s4);
          Cs[4 * s4] = cv.x; Cs[4 * s4 + 1] = cv.y; Cs[4 * s4 + 2] = cv.z; Cs[4 * s4 + 3] = cv.w;
        } else {
          Cs[4 * s4] = Cs[4 * s4 + 1] = Cs[4 * s4 + 2] = Cs[4 * s4 + 3] = 0.f;
        }
      }
      float y = 0.f;
      scan_step<SPT, WITH_Y>(t.h, t.a2, dl[i], u[i], Bs, Cs, y);
      if (WITH_Y) {
        y = channel_reduce<LPC>(y);
        if (t.ok && t.q == 0) yrow[(long long)(4 * j + i) * ystride] = fmaf(t.Dv, u[i], y);
      } else {
        t.sumdl += dl[i];
      }
    }
  }
}

// Full group with the exponentials issued ONE POSITION AHEAD of their use: a_cur holds exp2(delta'·A) of the
// position about to be consumed, and while its fma chains run the SPT MUFU.EX2 of the following position are
// already in flight.  An in-order warp then never waits on a MUFU it has just issued (a single warp per
// sub-partition with 16 independent ex2 in flight reaches 15/16 of the MUFU peak: scripts/mufu_bench.cu).
template <int N, int LPC, int RP, bool WITH_Y, bool REV>
__device__ __forceinline__ void group_body_pipe(Ss2dThread<N / LPC, RP> &t, const float *sDB, const float *sDC, float *yrow,
                                                long long ystride, int j, const float (&dl)[4], const float (&u)[4],
                                                float dl_after, float (&a_cur)[N / LPC]) {
  constexpr int Cp = 2 * N + RP;
  constexpr int SPT = N / LPC;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = REV ? 3 - ii : ii;
    const float dn = ii < 3 ? dl[REV ? i - 1 : i + 1] : dl_after;
    float a_nxt[SPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s) a_nxt[s] = ex2(dn * t.a2[s]);
    const float *rb = sDB + (4 * j + i) * Cp + t.q * SPT;
    const float *rc = sDC + (4 * j + i) * Cp + N + t.q * SPT;
    const float dlu = dl[i] * u[i];
    // all state updates first (SPT independent fma), then the C·h dot product on 4 accumulators: no fma waits
    // on a result produced less than ~4 instructions earlier (an in-order warp stalls on every such pair)
#pragma unroll
    for (int s4 = 0; s4 < SPT / 4; ++s4) {
      const float4 bv = *reinterpret_cast<const float4 *>(rb + 4 * s4);
      t.h[4 * s4 + 0] = fmaf(a_cur[4 * s4 + 0], t.h[4 * s4 + 0], dlu * bv.x);
      t.h[4 * s4 + 1] = fmaf(a_cur[4 * s4 + 1], t.h[4 * s4 + 1], dlu * bv.y);
      t.h[4 * s4 + 2] = fmaf(a_cur[4 * s4 + 2], t.h[4 * s4 + 2], dlu * bv.z);
      t.h[4 * s4 + 3] = fmaf(a_cur[4 * s4 + 3], t.h[4 * s4 + 3], dlu * bv.w);
    }
    float y = 0.f;
    if (WITH_Y) {
      float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
#pragma unroll
      for (int s4 = 0; s4 < SPT / 4; ++s4) {
        const float4 cv = *reinterpret_cast<const float4 *>(rc + 4 * s4);
        y0 = fmaf(t.h[4 * s4 + 0], cv.x, y0);
        y1 = fmaf(t.h[4 * s4 + 1], cv.y, y1);
        y2 = fmaf(t.h[4 * s4 + 2], cv.z, y2);
        y3 = fmaf(t.h[4 * s4 + 3], cv.w, y3);
      }
      y = (y0 + y1) + (y2 + y3);
    }
    if (WITH_Y) {
      y = channel_reduce<LPC>(y);
      if (t.ok && t.q == 0) yrow[(long long)(4 * j + i) * ystride] = fmaf(t.Dv, u[i], y);
    } else {
      t.sumdl += dl[i];
    }
#pragma unroll
    for (int s = 0; s < SPT; ++s) a_cur[s] = a_nxt[s];
  }
}

template <int N, int LPC, int RP, bool WITH_Y, bool REV>
__device__ __forceinline__ void scan_tile(Ss2dThread<N / LPC, RP> &t, const float *sXC, const float *sDB, const float *sDC,
                                          float *yrow, long long ystride, int DT, int npos) {
  const int nfull = npos >> 2, rem = npos & 3;
  float dl[4], u[4];
  if (REV && rem) {  // the ragged group comes first when walking backwards
    group_prologue<N, LPC, RP>(t, sXC, sDB, DT, nfull, dl, u);
    group_body<N, LPC, RP, WITH_Y, REV, false>(t, sDB, sDC, yrow, ystride, nfull, dl, u, rem);
  }
  if (nfull > 0) {
    int j = REV ? nfull - 1 : 0;
    group_prologue<N, LPC, RP>(t, sXC, sDB, DT, j, dl, u);
    float a_cur[N / LPC];
#pragma unroll
    for (int s = 0; s < N / LPC; ++s) a_cur[s] = ex2(dl[REV ? 3 : 0] * t.a2[s]);
#pragma unroll 1
    for (int g = 0; g < nfull; ++g) {
      // next group's delta'/u first (clamped index: the last iteration recomputes a valid group, unused),
      // so its loads / dot products / softplus overlap this group's exponentials and fma chains
      const int jn = REV ? max(j - 1, 0) : min(j + 1, nfull - 1);
      float dln[4], un[4];
      group_prologue<N, LPC, RP>(t, sXC, sDB, DT, jn, dln, un);
      group_body_pipe<N, LPC, RP, WITH_Y, REV>(t, sDB, sDC, yrow, ystride, j, dl, u, dln[REV ? 3 : 0], a_cur);
#pragma unroll
      for (int i = 0; i < 4; ++i) { dl[i] = dln[i]; u[i] = un[i]; }
      j = REV ? j - 1 : j + 1;
    }
  }
  if (!REV && rem) {
    group_prologue<N, LPC, RP>(t, sXC, sDB, DT, nfull, dl, u);
    group_body<N, LPC, RP, WITH_Y, REV, false>(t, sDB, sDC, yrow, ystride, nfull, dl, u, rem);
  }
}

template <int N, int LPC, int RP, int MODE>
__global__ void __launch_bounds__(288) ss2d_scan_kernel(const __grid_constant__ Ss2dParams p) {
  constexpr int LT = Ss2dCfg<N>::LT, NST = Ss2dCfg<N>::NST;
  constexpr bool WITH_Y = MODE != MODE_SUMMARY;

  extern __shared__ __align__(1024) unsigned char smem_raw[];  // TMA destinations need 128-byte alignment
  float *stages = reinterpret_cast<float *>(smem_raw);
  constexpr int Cp = 2 * N + RP;  // == p.Cp (checked on the host)
  const bool cross = p.kind == SIGMA_DIRS_CROSS;

  const int tid = threadIdx.x;
  constexpr int SPT = N / LPC;               // states per thread
  const int NTC = blockDim.x - 32;           // consumer threads; the last warp is the TMA producer
  const int DT = NTC / LPC;                  // channels per CTA
  const int nwarps_c = NTC >> 5;
  const bool is_producer = tid >= NTC;
  const int xc_fl = LT * DT, dbl_fl = LT * Cp;
  const int stage_fl = xc_fl + dbl_fl * (cross ? 2 : 1);
  uint64_t *full = reinterpret_cast<uint64_t *>(stages + NST * stage_fl);
  uint64_t *empty = full + NST;

  Ss2dThread<SPT, RP> t;
  t.lane = tid & 31;
  t.q = tid % LPC;
  t.ch = tid / LPC;
  const int d0 = blockIdx.x * DT;
  const int d = d0 + t.ch;
  t.ok = !is_producer && d < p.D;
  const int k = cross ? 0 : blockIdx.y / p.nsplit;
  const int split = cross ? blockIdx.y : blockIdx.y - k * p.nsplit;
  const int b = blockIdx.z;
  const int half = p.batch >> 1;        // CROSS: images [0,half) are modality 0 (rgb), [half,batch) modality 1
  const int kw = cross ? (b >= half ? 1 : 0) : k;                 // which weight set (direction / modality)
  const int bC = cross ? (b >= half ? b - half : b + half) : b;   // C of the OTHER modality (vmamba.py:1530,1536)
  const int I = p.I[k], O = p.O[k];
  const bool rev = p.rev[k] != 0;
  const int TPO = (I + LT - 1) / LT, ntiles = O * TPO;
  const int t0 = split * p.tiles_per_split, t1 = min(ntiles, t0 + p.tiles_per_split);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);           // one arrive (the producer's expect_tx) + the TMA bytes
      mbar_init(&empty[s], nwarps_c);   // one arrive per consumer warp
    }
    fence_mbar_init();
  }
  __syncthreads();

  auto tile_coord = [&](int tau, int &o, int &i0) {
    const int tm = rev ? ntiles - 1 - tau : tau;
    o = tm / TPO;
    i0 = (tm - o * TPO) * LT;
  };

  if (is_producer) {
    // ===== TMA producer warp: one elected lane refills a ring slot as soon as every consumer warp released it =====
    if (tid == NTC) {
      tma_prefetch_desc(&p.m_xc[k]);
      tma_prefetch_desc(&p.m_dbl[k]);
      const uint32_t tx_bytes = (uint32_t)(stage_fl * sizeof(float));
      for (int tau = t0; tau < t1; ++tau) {
        const int it = tau - t0, st = it % NST;
        mbar_wait(&empty[st], (uint32_t)(((it / NST) & 1) ^ 1));   // fresh barrier: parity 1 passes immediately
        float *dst = stages + st * stage_fl;
        int o, i0;
        tile_coord(tau, o, i0);
        mbar_arrive_expect_tx(&full[st], tx_bytes);
        tma_load_4d(dst, &p.m_xc[k], &full[st], d0, i0, o, b);
        tma_load_4d(dst + xc_fl, &p.m_dbl[k], &full[st], 0, i0, o, b);
        if (cross) tma_load_4d(dst + xc_fl + dbl_fl, &p.m_dbl[k], &full[st], 0, i0, o, bC);
      }
    }
    return;
  }

  // ===== consumer warps: one channel per thread =====
  const long long wd = (long long)kw * p.D + (t.ok ? d : 0);
#pragma unroll
  for (int s = 0; s < SPT; ++s) {
    t.a2[s] = t.ok ? p.A[wd * N + t.q * SPT + s] * kLog2e : 0.f;
    t.h[s] = 0.f;
  }
#pragma unroll
  for (int r = 0; r < RP; ++r) t.W[r] = (t.ok && r < p.R) ? p.dtw[wd * p.R + r] : 0.f;
  t.bias = t.ok ? p.dtb[wd] : 0.f;
  t.Dv = t.ok ? p.Ds[wd] : 0.f;
  t.sumdl = 0.f;
  float *carry_row = nullptr;
  if (MODE != MODE_SERIAL) {
    carry_row = p.carry + ((((long long)b * p.ndir + k) * p.D + (t.ok ? d : 0)) * p.nsplit + split) * 2 * N;
    if (MODE == MODE_APPLY && t.ok) {
#pragma unroll
      for (int s = 0; s < SPT; ++s) t.h[s] = carry_row[N + t.q * SPT + s];
    }
  }
  float *ybase = p.y + (((long long)k * p.batch + b) * p.Lseq) * p.D + (t.ok ? d : 0);
  const long long istride = p.istride[k], ostride = p.ostride[k];

  for (int tau = t0; tau < t1; ++tau) {
    const int it = tau - t0;
    const int st = it % NST;
    mbar_wait(&full[st], (uint32_t)((it / NST) & 1));

    const float *sXC = stages + st * stage_fl;
    const float *sDB = sXC + xc_fl;
    const float *sDC = cross ? sDB + dbl_fl : sDB;
    int o, i0;
    tile_coord(tau, o, i0);
    const int npos = min(LT, I - i0);
    float *yrow = ybase + (long long)o * ostride + (long long)i0 * istride;

    if (rev) scan_tile<N, LPC, RP, WITH_Y, true>(t, sXC, sDB, sDC, yrow, istride, DT, npos);
    else     scan_tile<N, LPC, RP, WITH_Y, false>(t, sXC, sDB, sDC, yrow, istride, DT, npos);

    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[st]);   // this warp is done with ring slot st
  }

  if (MODE == MODE_SUMMARY && t.ok) {
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
      carry_row[t.q * SPT + s] = ex2(t.a2[s] * t.sumdl);
      carry_row[N + t.q * SPT + s] = t.h[s];
    }
  }
}

// host-side launcher for one (N, LPC, RP) instantiation; defined per RP in ss2d_scan_rp*.cu.
// `nthreads` = consumer threads per CTA (LPC per channel); the launcher adds the producer warp.
template <int N, int LPC, int RP>
int ss2d_launch(const Ss2dParams &p, int nthreads, cudaStream_t stream);

}  // namespace sigma
