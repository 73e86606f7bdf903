// a4+a5 — fused multi-direction SS2D scan, channels-last (see include/sigma_b200.h: sigma_ss2d_scan_fwd).
//
// One launch does, for every direction k of an SS2D block:
//   CrossScan (index math: each direction is a walk over the SAME channels-last tensor, expressed as a
//   4-D TMA tensor map — row-major = tiles along L, column-major = tiles along H at fixed w, reversed =
//   the same tiles walked backwards) -> dt_proj (R-term dot product, W_dt row in registers) -> +bias ->
//   softplus -> selective scan (state in registers, one MUFU.EX2 per element) -> D skip -> store at the
//   POSITION the value belongs to (so CrossMerge's un-flip / un-transpose disappear).
//
// Mapping: one thread owns CPT (1 or 2) channels with all N states of each in registers; a warp covers 32
// consecutive channels (and the 32 that lie DT/2 further for CPT = 2), so every global / shared access of a warp
// is a 128-byte row and B / C / dt_r are broadcast shared reads shared by the thread's channels.  No shuffles.
// The recurrence runs on packed fp32x2 instructions (FFMA2 / FMUL2 over state pairs): per (channel, position)
// and state pair FMUL2 + 2 MUFU.EX2 + FMUL2 + FFMA2 + FFMA2, i.e. 3 issue slots per element instead of 5.
// History (profiles/r01_scan_variants.txt): 4 lanes/channel was issue-bound; 1 thread/channel with scalar
// fp32 left MUFU, the LDS return path (2N floats of B/C per channel-position) and issue each ~50 % busy;
// CPT = 2 halves the B/C traffic per channel and FFMA2 halves the fp32 issue slots.
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
  int ablate;  // timing experiments only (SIGMA_SCAN_ABLATE): 1 = no y store, 2 = no per-group prologue, 4 = no TMA reload
};

__host__ __device__ inline size_t ss2d_smem_bytes(int LT, int DT, int NST, int Cp, bool cross) {
  const size_t stage = (size_t)LT * DT + (size_t)LT * Cp * (cross ? 2 : 1);
  return NST * stage * sizeof(float) + 128 /*barriers*/;
}

template <int N, int CPT, int RP>
struct Ss2dThread {
  float h[CPT][N], a2[CPT][N], W[CPT][RP];
  float bias[CPT], Dv[CPT], sumdl[CPT];
  int ch;          // first channel of this thread inside the CTA tile; the c-th is ch + c*DT/CPT
  int ablate;
  bool ok[CPT];
};

// delta' and u of both channels for the 4 positions of group j (tile rows 4j..4j+3).  dt_r is read once per
// position (broadcast LDS.128) and used for all CPT channels; the dot products run on FFMA2 pairs.
template <int N, int CPT, int RP>
__device__ __forceinline__ void group_prologue(const Ss2dThread<N, CPT, RP> &t, const float *sXC, const float *sDB, int DT,
                                               int j, float (&dl)[CPT][4], float (&u)[CPT][4]) {
  constexpr int Cp = 2 * N + RP;  // x_dbl row length: [B | C | dt_r padded to RP] (sigma_ss2d_padded_cp)
  const int cstride = DT / CPT;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float *row = sDB + (4 * j + e) * Cp + 2 * N;
    f2 acc[CPT][2];
#pragma unroll
    for (int c = 0; c < CPT; ++c) { acc[c][0] = f2{t.bias[c], 0.f}; acc[c][1] = f2{0.f, 0.f}; }
#pragma unroll
    for (int q4 = 0; q4 < RP / 4; ++q4) {
      const float4 v = *reinterpret_cast<const float4 *>(row + 4 * q4);   // broadcast read, shared by CPT channels
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        acc[c][0] = fma2(f2{t.W[c][4 * q4 + 0], t.W[c][4 * q4 + 1]}, f2{v.x, v.y}, acc[c][0]);
        acc[c][1] = fma2(f2{t.W[c][4 * q4 + 2], t.W[c][4 * q4 + 3]}, f2{v.z, v.w}, acc[c][1]);
      }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      dl[c][e] = softplus20((acc[c][0].x + acc[c][0].y) + (acc[c][1].x + acc[c][1].y));
      u[c][e] = sXC[(4 * j + e) * DT + t.ch + c * cstride];
    }
  }
}

// recurrence over `cnt` (<= 4) positions of group j, in walk order (REV: descending tile rows).
// Per position B and C are read ONCE (2·N/4 broadcast LDS.128) and reused by the CPT channels of the thread;
// per channel and state pair: FMUL2 (exp arguments), 2 x MUFU.EX2, FMUL2 (delta·u·B), FFMA2 (h), FFMA2 (C·h).
template <int N, int CPT, int RP, bool WITH_Y, bool REV, bool FULL>
__device__ __forceinline__ void group_body(Ss2dThread<N, CPT, RP> &t, const float *sDB, const float *sDC, float *yrow,
                                           long long ystride, int ycstride, int j, const float (&dl)[CPT][4],
                                           const float (&u)[CPT][4], int cnt) {
  constexpr int Cp = 2 * N + RP;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = REV ? 3 - ii : ii;
    if (FULL || i < cnt) {
      const float *rb = sDB + (4 * j + i) * Cp;
      const float *rc = sDC + (4 * j + i) * Cp + N;
      f2 yacc[CPT][2];
#pragma unroll
      for (int c = 0; c < CPT; ++c) yacc[c][0] = yacc[c][1] = f2{0.f, 0.f};
#pragma unroll
      for (int s4 = 0; s4 < N / 4; ++s4) {
        const float4 bv = *reinterpret_cast<const float4 *>(rb + 4 * s4);   // broadcast reads
        float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (WITH_Y) cv = *reinterpret_cast<const float4 *>(rc + 4 * s4);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const float d = dl[c][i], du = dl[c][i] * u[c][i];
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {                                   // state pair (4·s4 + 2·hp, +1)
            const int s = 4 * s4 + 2 * hp;
            const f2 arg = mul2(f2{d, d}, f2{t.a2[c][s], t.a2[c][s + 1]});
            const f2 a = f2{ex2(arg.x), ex2(arg.y)};
            const f2 bb = mul2(f2{du, du}, hp == 0 ? f2{bv.x, bv.y} : f2{bv.z, bv.w});
            const f2 hn = fma2(a, f2{t.h[c][s], t.h[c][s + 1]}, bb);
            t.h[c][s] = hn.x; t.h[c][s + 1] = hn.y;
            if (WITH_Y) yacc[c][hp] = fma2(hn, hp == 0 ? f2{cv.x, cv.y} : f2{cv.z, cv.w}, yacc[c][hp]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (WITH_Y) {
          const float y = (yacc[c][0].x + yacc[c][0].y) + (yacc[c][1].x + yacc[c][1].y);
          if (t.ok[c] && !(t.ablate & 1)) yrow[(long long)(4 * j + i) * ystride + c * ycstride] = fmaf(t.Dv[c], u[c][i], y);
          if (t.ablate & 1) t.sumdl[c] += y;
        } else {
          t.sumdl[c] += dl[c][i];
        }
      }
    }
  }
}

template <int N, int CPT, int RP, bool WITH_Y, bool REV>
__device__ __forceinline__ void scan_tile(Ss2dThread<N, CPT, RP> &t, const float *sXC, const float *sDB, const float *sDC,
                                          float *yrow, long long ystride, int DT, int npos) {
  const int nfull = npos >> 2, rem = npos & 3;
  const int ycs = DT / CPT;
  float dl[CPT][4], u[CPT][4];
  if (REV && rem) {  // the ragged group comes first when walking backwards
    group_prologue<N, CPT, RP>(t, sXC, sDB, DT, nfull, dl, u);
    group_body<N, CPT, RP, WITH_Y, REV, false>(t, sDB, sDC, yrow, ystride, ycs, nfull, dl, u, rem);
  }
  if (nfull > 0) {
    int j = REV ? nfull - 1 : 0;
    group_prologue<N, CPT, RP>(t, sXC, sDB, DT, j, dl, u);
#pragma unroll 1
    for (int g = 0; g < nfull; ++g) {
      // next group's delta'/u first (clamped index: the last iteration recomputes a valid group, unused),
      // so its loads / dot products / softplus overlap this group's exponentials and fma chains
      const int jn = REV ? max(j - 1, 0) : min(j + 1, nfull - 1);
      float dln[CPT][4], un[CPT][4];
      if (!(t.ablate & 2)) group_prologue<N, CPT, RP>(t, sXC, sDB, DT, jn, dln, un);
      else {
#pragma unroll
        for (int c = 0; c < CPT; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) { dln[c][i] = dl[c][i] * 1.0001f; un[c][i] = u[c][i]; }
      }
      group_body<N, CPT, RP, WITH_Y, REV, true>(t, sDB, sDC, yrow, ystride, ycs, j, dl, u, 4);
#pragma unroll
      for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) { dl[c][i] = dln[c][i]; u[c][i] = un[c][i]; }
      j = REV ? j - 1 : j + 1;
    }
  }
  if (!REV && rem) {
    group_prologue<N, CPT, RP>(t, sXC, sDB, DT, nfull, dl, u);
    group_body<N, CPT, RP, WITH_Y, REV, false>(t, sDB, sDC, yrow, ystride, ycs, nfull, dl, u, rem);
  }
}

template <int N, int CPT, int RP, int MODE>
__global__ void __launch_bounds__(160, (N >= 16 ? 3 : 4)) ss2d_scan_kernel(const __grid_constant__ Ss2dParams p) {
  constexpr int LT = Ss2dCfg<N>::LT, NST = Ss2dCfg<N>::NST;
  constexpr bool WITH_Y = MODE != MODE_SUMMARY;

  extern __shared__ __align__(1024) unsigned char smem_raw[];  // TMA destinations need 128-byte alignment
  float *stages = reinterpret_cast<float *>(smem_raw);
  constexpr int Cp = 2 * N + RP;  // == p.Cp (checked on the host)
  const bool cross = p.kind == SIGMA_DIRS_CROSS;

  const int tid = threadIdx.x;
  const int NTC = blockDim.x - 32;           // consumer threads; the last warp is the TMA producer
  const int DT = NTC * CPT;                  // channels per CTA: thread t owns channels t and t + NTC (CPT = 2)
  const int nwarps_c = NTC >> 5;
  const bool is_producer = tid >= NTC;
  const int xc_fl = LT * DT, dbl_fl = LT * Cp;
  const int stage_fl = xc_fl + dbl_fl * (cross ? 2 : 1);
  uint64_t *full = reinterpret_cast<uint64_t *>(stages + NST * stage_fl);
  uint64_t *empty = full + NST;

  const int d0 = blockIdx.x * DT;
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
        if ((p.ablate & 4) && it >= NST) break;
        mbar_wait_backoff(&empty[st], (uint32_t)(((it / NST) & 1) ^ 1));   // fresh barrier: parity 1 passes immediately
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

  // ===== consumer warps: CPT channels per thread, all N states of each in registers =====
  Ss2dThread<N, CPT, RP> t;
  t.ch = tid;
  t.ablate = p.ablate;
  float *carry_row[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int d = d0 + tid + c * NTC;
    t.ok[c] = d < p.D;
    const long long wd = (long long)kw * p.D + (t.ok[c] ? d : 0);
#pragma unroll
    for (int s = 0; s < N; ++s) {
      t.a2[c][s] = t.ok[c] ? p.A[wd * N + s] * kLog2e : 0.f;
      t.h[c][s] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) t.W[c][r] = (t.ok[c] && r < p.R) ? p.dtw[wd * p.R + r] : 0.f;
    t.bias[c] = t.ok[c] ? p.dtb[wd] : 0.f;
    t.Dv[c] = t.ok[c] ? p.Ds[wd] : 0.f;
    t.sumdl[c] = 0.f;
    carry_row[c] = nullptr;
    if (MODE != MODE_SERIAL) {
      carry_row[c] = p.carry + ((((long long)b * p.ndir + k) * p.D + (t.ok[c] ? d : 0)) * p.nsplit + split) * 2 * N;
      if (MODE == MODE_APPLY && t.ok[c]) {
#pragma unroll
        for (int s = 0; s < N; ++s) t.h[c][s] = carry_row[c][N + s];
      }
    }
  }
  float *ybase = p.y + (((long long)k * p.batch + b) * p.Lseq) * p.D + min(d0 + tid, p.D - 1);
  const long long istride = p.istride[k], ostride = p.ostride[k];

  for (int tau = t0; tau < t1; ++tau) {
    const int it = tau - t0;
    const int st = it % NST;
    if (!((p.ablate & 4) && it >= NST)) mbar_wait(&full[st], (uint32_t)((it / NST) & 1));

    const float *sXC = stages + st * stage_fl;
    const float *sDB = sXC + xc_fl;
    const float *sDC = cross ? sDB + dbl_fl : sDB;
    int o, i0;
    tile_coord(tau, o, i0);
    const int npos = min(LT, I - i0);
    float *yrow = ybase + (long long)o * ostride + (long long)i0 * istride;

    if (rev) scan_tile<N, CPT, RP, WITH_Y, true>(t, sXC, sDB, sDC, yrow, istride, DT, npos);
    else     scan_tile<N, CPT, RP, WITH_Y, false>(t, sXC, sDB, sDC, yrow, istride, DT, npos);

    __syncwarp();
    if ((tid & 31) == 0 && !(p.ablate & 4)) mbar_arrive(&empty[st]);   // this warp is done with ring slot st
  }

  if (MODE == MODE_SUMMARY || (p.ablate & 1)) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (t.ok[c] && carry_row[c] != nullptr) {
#pragma unroll
        for (int s = 0; s < N; ++s) {
          carry_row[c][s] = ex2(t.a2[c][s] * t.sumdl[c]);
          carry_row[c][N + s] = t.h[c][s];
        }
      }
    }
  }
}

// host-side launcher for one (N, CPT, RP) instantiation; defined per RP in ss2d_scan_rp*.cu.
// `nthreads` = consumer threads per CTA (each owning CPT channels); the launcher adds the producer warp.
template <int N, int CPT, int RP>
int ss2d_launch(const Ss2dParams &p, int nthreads, cudaStream_t stream);

}  // namespace sigma
