// a4+a5 — fused multi-direction SS2D scan, channels-last (see include/sigma_b200.h: sigma_ss2d_scan_fwd).
//
// One launch does, for every direction k of an SS2D block:
//   CrossScan (index math: each direction is a walk over the SAME channels-last tensor, expressed as a
//   4-D TMA tensor map — row-major = tiles along L, column-major = tiles along H at fixed w, reversed =
//   the same tiles walked backwards) -> dt_proj (R-term dot product, W_dt row in registers) -> +bias ->
//   softplus -> selective scan (state in registers, one MUFU.EX2 per element) -> D skip -> store at the
//   POSITION the value belongs to (so CrossMerge's un-flip / un-transpose disappear).
//
// CTA = (channel tile DT, direction k [x L-segment], image b).  Tiles of LT scan positions are staged
// HBM -> shared by TMA (cp.async.bulk.tensor, mbarrier complete_tx) through an NST-deep ring; the y tile
// goes back shared -> HBM by TMA store.  Per group of 4 positions the delta' of the NEXT group is computed
// while the recurrence of the current one runs (software pipelining: the only serial dependency is the
// fma h = a·h + b), and all 4·SPT exponentials of a group are independent of h.
#pragma once
#include "scan_core.cuh"
#include "tma.cuh"

namespace sigma {

template <int LPC>
struct Ss2dCfg {
  static constexpr int NW = LPC >= 4 ? 8 : (LPC == 2 ? 8 : 4);  // warps per CTA
  static constexpr int LT = LPC >= 2 ? 32 : 16;                 // scan positions per tile
  static constexpr int NST = 3;                                 // TMA ring depth
  static constexpr int CPW = 32 / LPC;
  static constexpr int DT = NW * CPW;                           // channels per CTA: 64 / 128 / 128
  static constexpr int NTHREADS = NW * 32;
};

struct alignas(64) Ss2dParams {
  CUtensorMap m_xc[4], m_dbl[4], m_y[4];
  const float *dtw, *dtb, *A, *Ds;
  float *carry;
  int D, N, R, Cp, kind, batch;
  int I[4], O[4], rev[4];
  int nsplit, tiles_per_split;
};

__host__ __device__ inline size_t ss2d_smem_bytes(int LT, int DT, int NST, int Cp, bool cross) {
  const size_t stage = (size_t)LT * DT + (size_t)LT * Cp * (cross ? 2 : 1);
  return (NST * stage + 2 * (size_t)LT * DT) * sizeof(float) + 64 /*barriers*/ + 128 /*alignment slack*/;
}

template <int SPT, int LPC, int RP>
struct Ss2dThread {
  float h[SPT], a2[SPT], W[RP];
  float bias, Dv, sumdl;
  int lane, q, ch;
};

// delta' for the 4 positions of group j (tile rows 4j..4j+3) + their u values.
template <int SPT, int LPC, int RP, int DT>
__device__ __forceinline__ void group_prologue(const Ss2dThread<SPT, LPC, RP> &t, const float *sXC, const float *sDB,
                                               int Cp, int j, float (&dl)[4], float (&u)[4]) {
  constexpr int N = SPT * LPC;
  constexpr int PPL = LPC >= 4 ? 1 : 4 / LPC;  // positions whose delta this lane evaluates
  const int first = LPC >= 4 ? (t.q & 3) : t.q * PPL;
  float own[PPL];
#pragma unroll
  for (int e = 0; e < PPL; ++e) {
    const float *row = sDB + (4 * j + first + e) * Cp + 2 * N;
    float acc = t.bias;
#pragma unroll
    for (int c = 0; c < RP / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4 *>(row + 4 * c);
      acc = fmaf(t.W[4 * c + 0], v.x, acc);
      acc = fmaf(t.W[4 * c + 1], v.y, acc);
      acc = fmaf(t.W[4 * c + 2], v.z, acc);
      acc = fmaf(t.W[4 * c + 3], v.w, acc);
    }
    own[e] = softplus20(acc);
  }
  if (LPC == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dl[i] = own[i % PPL];
  } else {
    const int base = t.lane & ~(LPC - 1);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dl[i] = __shfl_sync(0xffffffffu, own[i % PPL], base + (LPC >= 4 ? i : i / PPL));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = sXC[(4 * j + i) * DT + t.ch];
}

// recurrence over `cnt` (<= 4) positions of group j, in walk order (REV: descending tile rows).
template <int SPT, int LPC, int RP, int DT, bool WITH_Y, bool REV, bool FULL>
__device__ __forceinline__ void group_body(Ss2dThread<SPT, LPC, RP> &t, const float *sDB, const float *sDC, float *sY,
                                           int Cp, int j, const float (&dl)[4], const float (&u)[4], int cnt) {
  constexpr int N = SPT * LPC;
  float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = REV ? 3 - ii : ii;
    if (FULL || i < cnt) {
      float Bs[SPT], Cs[SPT];
      const float *rb = sDB + (4 * j + i) * Cp + t.q * SPT;
      const float *rc = sDC + (4 * j + i) * Cp + N + t.q * SPT;
#pragma unroll
      for (int s4 = 0; s4 < SPT / 4; ++s4) {
        const float4 bv = *reinterpret_cast<const float4 *>(rb + 4 * s4);
        Bs[4 * s4] = bv.x; Bs[4 * s4 + 1] = bv.y; Bs[4 * s4 + 2] = bv.z; Bs[4 * s4 + 3] = bv.w;
        if (WITH_Y) {
          const float4 cv = *reinterpret_cast<const float4 *>(rc + 4 * s4);
          Cs[4 * s4] = cv.x; Cs[4 * s4 + 1] = cv.y; Cs[4 * s4 + 2] = cv.z; Cs[4 * s4 + 3] = cv.w;
        } else {
          Cs[4 * s4] = Cs[4 * s4 + 1] = Cs[4 * s4 + 2] = Cs[4 * s4 + 3] = 0.f;
        }
      }
      scan_step<SPT, WITH_Y>(t.h, t.a2, dl[i], u[i], Bs, Cs, y[i]);
      if (!WITH_Y) t.sumdl += dl[i];
    }
  }
  if (WITH_Y) {
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = channel_reduce<LPC>(y[i]);
    if (t.q == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (FULL || i < cnt) sY[(4 * j + i) * DT + t.ch] = fmaf(t.Dv, u[i], y[i]);
    }
  }
}

template <int SPT, int LPC, int RP, int DT, bool WITH_Y, bool REV>
__device__ __forceinline__ void scan_tile(Ss2dThread<SPT, LPC, RP> &t, const float *sXC, const float *sDB,
                                          const float *sDC, float *sY, int Cp, int npos) {
  const int nfull = npos >> 2, rem = npos & 3;
  float dl[4], u[4];
  if (REV && rem) {  // the ragged group comes first when walking backwards
    group_prologue<SPT, LPC, RP, DT>(t, sXC, sDB, Cp, nfull, dl, u);
    group_body<SPT, LPC, RP, DT, WITH_Y, REV, false>(t, sDB, sDC, sY, Cp, nfull, dl, u, rem);
  }
  if (nfull > 0) {
    int j = REV ? nfull - 1 : 0;
    group_prologue<SPT, LPC, RP, DT>(t, sXC, sDB, Cp, j, dl, u);
#pragma unroll 1
    for (int g = 0; g < nfull; ++g) {
      // next group's delta'/u first (clamped index: the last iteration recomputes a valid group, unused),
      // so its loads / softplus / shuffles overlap this group's exponentials and fma chain
      const int jn = REV ? max(j - 1, 0) : min(j + 1, nfull - 1);
      float dln[4], un[4];
      group_prologue<SPT, LPC, RP, DT>(t, sXC, sDB, Cp, jn, dln, un);
      group_body<SPT, LPC, RP, DT, WITH_Y, REV, true>(t, sDB, sDC, sY, Cp, j, dl, u, 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) { dl[i] = dln[i]; u[i] = un[i]; }
      j = REV ? j - 1 : j + 1;
    }
  }
  if (!REV && rem) {
    group_prologue<SPT, LPC, RP, DT>(t, sXC, sDB, Cp, nfull, dl, u);
    group_body<SPT, LPC, RP, DT, WITH_Y, REV, false>(t, sDB, sDC, sY, Cp, nfull, dl, u, rem);
  }
}

template <int SPT, int LPC, int RP, int MODE>
__global__ void __launch_bounds__(Ss2dCfg<LPC>::NTHREADS) ss2d_scan_kernel(const __grid_constant__ Ss2dParams p) {
  using Cfg = Ss2dCfg<LPC>;
  constexpr int DT = Cfg::DT, LT = Cfg::LT, NST = Cfg::NST, CPW = Cfg::CPW;
  constexpr int N = SPT * LPC;
  constexpr bool WITH_Y = MODE != MODE_SUMMARY;

  extern __shared__ __align__(1024) unsigned char smem_raw[];  // TMA destinations need 128-byte alignment
  float *stages = reinterpret_cast<float *>(smem_raw);
  const int Cp = p.Cp;
  const bool cross = p.kind == SIGMA_DIRS_CROSS;
  const int xc_fl = LT * DT, dbl_fl = LT * Cp;
  const int stage_fl = xc_fl + dbl_fl * (cross ? 2 : 1);
  float *sYbase = stages + NST * stage_fl;
  uint64_t *full = reinterpret_cast<uint64_t *>(sYbase + 2 * xc_fl);

  const int tid = threadIdx.x;
  Ss2dThread<SPT, LPC, RP> t;
  t.lane = tid & 31;
  t.q = t.lane % LPC;
  t.ch = (tid >> 5) * CPW + t.lane / LPC;
  const int d0 = blockIdx.x * DT;
  const int d = d0 + t.ch;
  const bool ch_ok = d < p.D;
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

  const long long wd = (long long)kw * p.D + (ch_ok ? d : 0);
#pragma unroll
  for (int s = 0; s < SPT; ++s) {
    t.a2[s] = ch_ok ? p.A[wd * N + t.q * SPT + s] * kLog2e : 0.f;
    t.h[s] = 0.f;
  }
#pragma unroll
  for (int r = 0; r < RP; ++r) t.W[r] = (ch_ok && r < p.R) ? p.dtw[wd * p.R + r] : 0.f;
  t.bias = ch_ok ? p.dtb[wd] : 0.f;
  t.Dv = ch_ok ? p.Ds[wd] : 0.f;
  t.sumdl = 0.f;
  float *carry_row = nullptr;
  if (MODE != MODE_SERIAL) {
    const int ndir = cross ? 1 : (int)(gridDim.y / p.nsplit);
    carry_row = p.carry + ((((long long)b * ndir + k) * p.D + (ch_ok ? d : 0)) * p.nsplit + split) * 2 * N;
    if (MODE == MODE_APPLY && ch_ok) {
#pragma unroll
      for (int s = 0; s < SPT; ++s) t.h[s] = carry_row[N + t.q * SPT + s];
    }
  }

  if (tid == 0) {
    tma_prefetch_desc(&p.m_xc[k]);
    tma_prefetch_desc(&p.m_dbl[k]);
    if (WITH_Y) tma_prefetch_desc(&p.m_y[k]);
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const uint32_t tx_bytes = (uint32_t)(stage_fl * sizeof(float));
  auto tile_coord = [&](int tau, int &o, int &i0) {
    const int tm = rev ? ntiles - 1 - tau : tau;
    o = tm / TPO;
    i0 = (tm - o * TPO) * LT;
  };
  auto issue = [&](int tau) {
    const int st = (tau - t0) % NST;
    float *dst = stages + st * stage_fl;
    int o, i0;
    tile_coord(tau, o, i0);
    mbar_arrive_expect_tx(&full[st], tx_bytes);
    tma_load_4d(dst, &p.m_xc[k], &full[st], d0, i0, o, b);
    tma_load_4d(dst + xc_fl, &p.m_dbl[k], &full[st], 0, i0, o, b);
    if (cross) tma_load_4d(dst + xc_fl + dbl_fl, &p.m_dbl[k], &full[st], 0, i0, o, bC);
  };

  if (tid == 0)
    for (int tau = t0; tau < min(t1, t0 + NST - 1); ++tau) issue(tau);

  for (int tau = t0; tau < t1; ++tau) {
    const int it = tau - t0;
    const int st = it % NST;
    if (tid == 0 && tau + NST - 1 < t1) issue(tau + NST - 1);  // slot freed by the barrier that ended iteration tau-1
    mbar_wait(&full[st], (uint32_t)((it / NST) & 1));

    const float *sXC = stages + st * stage_fl;
    const float *sDB = sXC + xc_fl;
    const float *sDC = cross ? sDB + dbl_fl : sDB;
    float *sY = sYbase + (it & 1) * xc_fl;
    int o, i0;
    tile_coord(tau, o, i0);
    const int npos = min(LT, I - i0);

    if (rev) scan_tile<SPT, LPC, RP, DT, WITH_Y, true>(t, sXC, sDB, sDC, sY, Cp, npos);
    else     scan_tile<SPT, LPC, RP, DT, WITH_Y, false>(t, sXC, sDB, sDC, sY, Cp, npos);

    if (WITH_Y) {
      fence_proxy_async();                      // my y-tile writes -> visible to the TMA store
      if (tid == 0) tma_store_wait_read<0>();   // store of tile tau-1 has finished reading the other y buffer
    }
    __syncthreads();                            // everyone done with input stage st and with this y tile
    if (WITH_Y && tid == 0) {
      tma_store_4d(&p.m_y[k], sY, d0, i0, o, b);
      tma_store_commit();
    }
  }
  if (WITH_Y && tid == 0) tma_store_wait_all<0>();

  if (MODE == MODE_SUMMARY && ch_ok) {
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
      carry_row[t.q * SPT + s] = ex2(t.a2[s] * t.sumdl);
      carry_row[N + t.q * SPT + s] = t.h[s];
    }
  }
}

// host-side launcher for one (N, RP) instantiation; defined per RP in ss2d_scan_rp*.cu
template <int SPT, int LPC, int RP>
int ss2d_launch(const Ss2dParams &p, int ndir, cudaStream_t stream);

}  // namespace sigma
