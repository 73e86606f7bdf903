// f1 — backward of the FUSED multi-direction SS2D scan, channels-last (see include/sigma_b200.h: sigma_ss2d_scan_bwd).
//
// Replaces, for training, the autograd of CrossScan + dt_proj einsum + SelectiveScan + CrossMerge
// (vmamba.py:80-121, 195-215 and selective_scan_bwd_kernel.cuh:68-274) without ever materialising CrossScan's (B,4,D,L)
// copy: every direction is the same 4-D TMA walk over the channels-last tensors that the forward uses.  Three sweeps:
//   1. state sweep  (ss2d_state_kernel, walk order): delta' = softplus(dt_r·W_dt + bias) -> `delta` slabs (K,B,L,D), and the
//      state h at the start of every 16-position tile -> `hs`; L-segments by MODE_SUMMARY -> scan_combine_kernel -> MODE_APPLY;
//   2. reverse summaries (ss2d_bwd_kernel<MODE_SUMMARY>, only with L-segments) + scan_combine_rev_kernel: the dh entering
//      every segment (L-parallel reverse sweep);
//   3. main sweep (ss2d_bwd_kernel, tiles walked BACKWARDS): per tile recompute h at every position from the tile's start
//      state into shared memory, then the reverse recurrence dh_l = a_{l+1}·dh_{l+1} + dy_l·C_l producing
//        du      -> TMA REDUCE-ADD (cp.reduce.async.bulk.tensor .add.f32) straight into dxc (B,L,D): the four directions'
//                   contributions meet in L2, no (B,4,D,L) gradient tensor and no CrossScan backward;
//        ddelta  -> `ddelta` slabs (K,B,L,D) (pre-softplus; the caller turns them into d dt_r and dW_dt with two GEMMs);
//        dB, dC  -> summed over the warp's channels by the transposing shuffle reduction of the op-level backward, then
//                   coalesced red.global.add into dxdbl (B,L,K,Cp);
//        dA, dDs, d dt_bias -> per-thread accumulators, one atomic per channel at the end.
// Thread mapping as scan_op_bwd_tma.cu: d_state 16 -> 2 lanes per channel (8 states each), d_state 4 -> 1; a CTA = 64
// channels of one (direction, image [, L-segment]); tiles of 16 positions through a TMA ring (full mbarrier + last-arriver
// refill).  Kinds SIGMA_DIRS_CROSS4 and SIGMA_DIRS_SEQ2 (SS2D and ConMB); d_state in {4, 16}; D % 64 == 0.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "scan_core.cuh"
#include "tma.cuh"

namespace sigma {

int make_tmap_f32_4d(CUtensorMap *map, const void *base, const uint64_t dims[4], const uint64_t strides_bytes[3], const uint32_t box[4]);
int pick_segments(long long ctas_base, int ntiles, long long slots, double pass_factor, int max_split);   // scan_op_tma.cu
cudaError_t prep_kernel_once(const void *fn);                                                            // scan_op_tma.cu
__global__ void scan_combine_kernel(float *carry, long long nrows, int nsplit, int NP);                  // scan_op.cu
__global__ void scan_combine_rev_kernel(float *carry, long long nrows, int nsplit, int NP);              // scan_op_bwd_tma.cu

constexpr int FB_LT = 16;   // positions per tile
constexpr int FB_DT = 64;   // channels per CTA

struct alignas(64) Ss2dBwdParams {
  CUtensorMap m_xc[4], m_dy[4], m_dl[4], m_dbl[4];    // loads: boxes {64 ch, 16 pos} / {Cp, 16 pos}
  CUtensorMap m_dxc[4], m_dd[4];                      // per-warp outputs: boxes {CPW ch, 16 pos}
  const float *dtw, *dtb, *A, *Ds, *hs_in;
  float *hs, *dxdbl, *dA, *dDs, *ddtb, *carry;
  int D, N, R, Cp, K, batch;
  long long Lseq;
  int I[4], O[4], rev[4];
  long long psi[4], pso[4];     // position = o·pso + i·psi
  int nsplit, tiles_per_split, max_tiles, nst;
};

template <int N> struct FbCfg {
  static constexpr int LPC = N >= 16 ? 2 : 1;
  static constexpr int NS = N / LPC;
  static constexpr int CPW = 32 / LPC;
};

__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// see scan_op_bwd_tma.cu
template <int NV, int OFF>
__device__ __forceinline__ float fb_transpose_reduce(float (&v)[NV], int lane, int &which) {
  if constexpr (OFF == 0) {
    return v[0];
  } else if constexpr (NV > 1) {
    const bool up = (lane & OFF) != 0;
    float w[NV / 2];
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) {
      const float send = up ? v[j] : v[j + NV / 2];
      const float keep = up ? v[j + NV / 2] : v[j];
      w[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    which = which * 2 + (up ? 1 : 0);
    return fb_transpose_reduce<NV / 2, OFF / 2>(w, lane, which);
  } else {
    float w[1] = {v[0] + __shfl_xor_sync(0xffffffffu, v[0], OFF)};
    return fb_transpose_reduce<1, OFF / 2>(w, lane, which);
  }
}

// everything the three kernels share: CTA coordinates, the direction's tile geometry, the ring
struct FbWalk {
  int k, split, b, d0, t0, t1, I, TPO, ntiles;
  bool rev;
  __device__ __forceinline__ void tile(int tau, int &o, int &i0, int &npos) const {
    const int tm = rev ? ntiles - 1 - tau : tau;
    o = tm / TPO;
    i0 = (tm - o * TPO) * FB_LT;
    npos = min(FB_LT, I - i0);
  }
};

__device__ __forceinline__ FbWalk fb_walk(const Ss2dBwdParams &p) {
  FbWalk w;
  w.d0 = blockIdx.x * FB_DT;
  w.k = blockIdx.y / p.nsplit;
  w.split = blockIdx.y - w.k * p.nsplit;
  w.b = blockIdx.z;
  w.I = p.I[w.k];
  w.rev = p.rev[w.k] != 0;
  w.TPO = (w.I + FB_LT - 1) / FB_LT;
  w.ntiles = p.O[w.k] * w.TPO;
  w.t0 = w.split * p.tiles_per_split;
  w.t1 = min(w.ntiles, w.t0 + p.tiles_per_split);
  return w;
}

// ---------------------------------------------------------------------------------------------------------------------
// 1. state sweep: delta' slabs + tile-start states (walk order)
// ---------------------------------------------------------------------------------------------------------------------
template <int N, int MODE>
__global__ void __launch_bounds__(128, 3) ss2d_state_kernel(const __grid_constant__ Ss2dBwdParams p) {
  constexpr int LPC = FbCfg<N>::LPC, NS = FbCfg<N>::NS, CPW = FbCfg<N>::CPW, NT = FB_DT * LPC;
  // delta' does not depend on the state: the pass that sees a position FIRST (MODE_SERIAL, or MODE_SUMMARY when the walk is
  // cut into L-segments) computes it and stores the slab tile; MODE_APPLY reads it back instead of repeating the dot product
  constexpr bool COMPUTE = MODE != MODE_APPLY, STATES = MODE != MODE_SUMMARY;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *smem = reinterpret_cast<float *>(smem_raw);
  const int NST = p.nst, Cp = p.Cp, R = p.R;
  const int xc_fl = FB_LT * FB_DT, dbl_fl = FB_LT * Cp, stage_fl = xc_fl + dbl_fl + (COMPUTE ? 0 : xc_fl);   // xc | dbl | [delta']
  float *stage_all = smem + NST * stage_fl;                // per-warp delta staging [16][CPW] (TMA store source: keep it aligned)
  float *sW = stage_all + (NT / 32) * FB_LT * CPW;         // W_dt rows of this CTA's channels, pitch R + 1
  uint64_t *full = reinterpret_cast<uint64_t *>(sW + ((FB_DT * (R + 1) + 1) & ~1));
  uint32_t *done = reinterpret_cast<uint32_t *>(full + NST);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  const int half = lane / CPW, cl = lane - half * CPW, c = warp * CPW + cl, n0 = half * NS;
  const FbWalk w = fb_walk(p);
  const int d = w.d0 + c;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
    fence_mbar_init();
  }
  if (COMPUTE) {
    for (int i = tid; i < FB_DT * R; i += NT) {
      const int cc = i / R, r = i - cc * R;
      sW[cc * (R + 1) + r] = p.dtw[((long long)w.k * p.D + w.d0 + cc) * R + r];
    }
  }
  __syncthreads();
  if (w.t0 >= w.t1) return;
  const uint32_t tx = (uint32_t)(stage_fl * sizeof(float));
  auto request_tile = [&](int tau, int st) {
    int o, i0, npos;
    w.tile(tau, o, i0, npos);
    float *dst = smem + st * stage_fl;
    mbar_arrive_expect_tx(&full[st], tx);
    tma_load_4d(dst, &p.m_xc[w.k], &full[st], w.d0, i0, o, w.b);
    tma_load_4d(dst + xc_fl, &p.m_dbl[w.k], &full[st], 0, i0, o, w.b);
    if (!COMPUTE) tma_load_4d(dst + xc_fl + dbl_fl, &p.m_dl[w.k], &full[st], w.d0, i0, o, w.k * p.batch + w.b);
  };
  if (tid == 0) for (int tau = w.t0; tau < min(w.t1, w.t0 + NST); ++tau) request_tile(tau, tau - w.t0);

  float h[NS], a2[NS];
  const long long wd = (long long)w.k * p.D + d;
#pragma unroll
  for (int s = 0; s < NS; ++s) { a2[s] = p.A[wd * N + n0 + s] * kLog2e; h[s] = 0.f; }
  const float bias = p.dtb[wd];
  float sumdl = 0.f;
  float *carry_row = p.carry + ((((long long)w.b * p.K + w.k) * p.D + d) * p.nsplit + w.split) * 2 * N;
  if (MODE == MODE_APPLY) {
#pragma unroll
    for (int s = 0; s < NS; ++s) h[s] = carry_row[N + n0 + s];
  }
  const float *wrow = sW + c * (R + 1);
  float *stg = stage_all + warp * FB_LT * CPW;

  int st = 0, ph = 0;
  for (int tau = w.t0; tau < w.t1; ++tau) {
    int o, i0, npos;
    w.tile(tau, o, i0, npos);
    mbar_spin(&full[st], (uint32_t)ph);
    const float *sXC = smem + st * stage_fl, *sDB = sXC + xc_fl, *sDL = sDB + dbl_fl;
    if (STATES) {   // state at the start of the tile (walk order)
      float4 *hp = reinterpret_cast<float4 *>(p.hs + (((((long long)w.k * p.batch + w.b) * p.max_tiles + tau) * p.D + d) * N + n0));
#pragma unroll
      for (int q = 0; q < NS / 4; ++q) hp[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    }
#pragma unroll 1
    for (int s = 0; s < npos; ++s) {
      const int r = w.rev ? npos - 1 - s : s;
      const float *row = sDB + r * Cp;
      float dl;
      if (COMPUTE) {
        dl = 0.f;
        if (half == 0) {           // one lane per channel evaluates dt_proj + softplus; its partner lane (d_state 16) receives it
          float acc = bias;
          for (int q = 0; q < R; ++q) acc = fmaf(wrow[q], row[2 * N + q], acc);
          dl = softplus20(acc);
          stg[r * CPW + cl] = dl;
        }
        if (LPC == 2) dl = __shfl_sync(0xffffffffu, dl, cl);
      } else {
        dl = sDL[r * FB_DT + c];
      }
      const float du = dl * sXC[r * FB_DT + c];
#pragma unroll
      for (int q = 0; q < NS / 4; ++q) {
        const float4 bv = *reinterpret_cast<const float4 *>(row + n0 + 4 * q);
        h[4 * q] = fmaf(ex2(dl * a2[4 * q]), h[4 * q], du * bv.x);
        h[4 * q + 1] = fmaf(ex2(dl * a2[4 * q + 1]), h[4 * q + 1], du * bv.y);
        h[4 * q + 2] = fmaf(ex2(dl * a2[4 * q + 2]), h[4 * q + 2], du * bv.z);
        h[4 * q + 3] = fmaf(ex2(dl * a2[4 * q + 3]), h[4 * q + 3], du * bv.w);
      }
      sumdl += dl;
    }
    if (COMPUTE) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&p.m_dd[w.k], stg, w.d0 + warp * CPW, i0, o, w.k * p.batch + w.b);   // m_dd doubles as the delta-slab map here
        tma_store_commit();
        tma_store_wait_read<0>();
      }
    }
    __syncwarp();
    if (lane == 0 && tau + NST < w.t1) {
      const uint32_t old = smem_inc_acq_rel(&done[st]);
      if ((old + 1) % (uint32_t)nwarps == 0) request_tile(tau + NST, st);
    }
    if (++st == NST) { st = 0; ph ^= 1; }
  }
  if (COMPUTE && lane == 0) tma_store_wait_all<0>();
  if (MODE == MODE_SUMMARY) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      carry_row[n0 + s] = ex2(a2[s] * sumdl);
      carry_row[N + n0 + s] = h[s];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2 + 3. reverse summaries (MODE_SUMMARY) and the main backward sweep (MODE_SERIAL / MODE_APPLY)
// ---------------------------------------------------------------------------------------------------------------------
template <int N, int MODE>
__global__ void __launch_bounds__(128, 2) ss2d_bwd_kernel(const __grid_constant__ Ss2dBwdParams p) {
  constexpr int LPC = FbCfg<N>::LPC, NS = FbCfg<N>::NS, CPW = FbCfg<N>::CPW, NT = FB_DT * LPC;
  constexpr bool MAIN = MODE != MODE_SUMMARY;
  constexpr float kLn2 = 0.6931471805599453f;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *smem = reinterpret_cast<float *>(smem_raw);
  const int NST = p.nst, Cp = p.Cp;
  const int xc_fl = FB_LT * FB_DT, dbl_fl = FB_LT * Cp;
  const int stage_fl = (MAIN ? 3 : 2) * xc_fl + dbl_fl;            // [xc] dy dl dbl
  float *stage_all = smem + NST * stage_fl;                        // per-warp staging: du [16][CPW], ddelta [16][CPW]
  float4 *sH = reinterpret_cast<float4 *>(stage_all + (MAIN ? (NT / 32) * 2 * FB_LT * CPW : 0));   // [16][NS/4][NT]
  uint64_t *full = reinterpret_cast<uint64_t *>(sH + (MAIN ? FB_LT * (NS / 4) * NT : 0));
  uint32_t *done = reinterpret_cast<uint32_t *>(full + NST);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  const int half = lane / CPW, cl = lane - half * CPW, c = warp * CPW + cl, n0 = half * NS;
  const FbWalk w = fb_walk(p);
  const int d = w.d0 + c;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
    fence_mbar_init();
  }
  __syncthreads();
  if (w.t0 >= w.t1) {
    if (MODE == MODE_SUMMARY) {   // empty trailing segment: identity summary
      float *cr = p.carry + ((((long long)w.b * p.K + w.k) * p.D + d) * p.nsplit + w.split) * 2 * N;
#pragma unroll
      for (int s = 0; s < NS; ++s) { cr[n0 + s] = 1.f; cr[N + n0 + s] = 0.f; }
    }
    return;
  }
  const int ntl = w.t1 - w.t0;
  const uint32_t tx = (uint32_t)(stage_fl * sizeof(float));
  // ring order kk = 0.. walks tiles t1-1 down to t0
  auto request_tile = [&](int kk, int st) {
    int o, i0, npos;
    w.tile(w.t1 - 1 - kk, o, i0, npos);
    float *dst = smem + st * stage_fl;
    mbar_arrive_expect_tx(&full[st], tx);
    if (MAIN) {
      tma_load_4d(dst, &p.m_xc[w.k], &full[st], w.d0, i0, o, w.b);
      dst += xc_fl;
    }
    tma_load_4d(dst, &p.m_dy[w.k], &full[st], w.d0, i0, o, w.b);
    tma_load_4d(dst + xc_fl, &p.m_dl[w.k], &full[st], w.d0, i0, o, w.k * p.batch + w.b);
    tma_load_4d(dst + 2 * xc_fl, &p.m_dbl[w.k], &full[st], 0, i0, o, w.b);
  };
  if (tid == 0) for (int kk = 0; kk < min(ntl, NST); ++kk) request_tile(kk, kk);

  float a2[NS], dh[NS], dAacc[NS];
  const long long wd = (long long)w.k * p.D + d;
#pragma unroll
  for (int s = 0; s < NS; ++s) { a2[s] = p.A[wd * N + n0 + s] * kLog2e; dh[s] = 0.f; dAacc[s] = 0.f; }
  float *carry_row = p.carry + ((((long long)w.b * p.K + w.k) * p.D + d) * p.nsplit + w.split) * 2 * N;
  if (MODE == MODE_APPLY) {
#pragma unroll
    for (int s = 0; s < NS; ++s) dh[s] = carry_row[N + n0 + s];
  }
  const float Dv = MAIN ? p.Ds[wd] : 0.f;
  float dDacc = 0.f, dbacc = 0.f, sumdl = 0.f;
  float *sdu = stage_all + warp * 2 * FB_LT * CPW, *sdd = sdu + FB_LT * CPW;
  float4 *sHt = sH + tid;
  float *dxrow0 = p.dxdbl + (long long)w.b * p.Lseq * p.K * Cp + (long long)w.k * Cp;   // + pos·K·Cp

  int st = 0, ph = 0;
  for (int kk = 0; kk < ntl; ++kk) {
    const int tau = w.t1 - 1 - kk;
    int o, i0, npos;
    w.tile(tau, o, i0, npos);
    mbar_spin(&full[st], (uint32_t)ph);
    const float *base = smem + st * stage_fl;
    const float *sXC = base, *sDY = base + (MAIN ? xc_fl : 0), *sDL = sDY + xc_fl, *sDB = sDL + xc_fl;

    if (MAIN) {
      // ---- forward inside the tile from its start state (walk order), h after every step -> shared memory ----
      float h[NS];
      const float4 *hp = reinterpret_cast<const float4 *>(p.hs_in + (((((long long)w.k * p.batch + w.b) * p.max_tiles + tau) * p.D + d) * N + n0));
#pragma unroll
      for (int q = 0; q < NS / 4; ++q) { const float4 v = hp[q]; h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w; }
#pragma unroll 1
      for (int s = 0; s < npos; ++s) {
        const int r = w.rev ? npos - 1 - s : s;
        const float dl = sDL[r * FB_DT + c];
        const float du = dl * sXC[r * FB_DT + c];
        const float *row = sDB + r * Cp + n0;
#pragma unroll
        for (int q = 0; q < NS / 4; ++q) {
          const float4 bv = *reinterpret_cast<const float4 *>(row + 4 * q);
          h[4 * q] = fmaf(ex2(dl * a2[4 * q]), h[4 * q], du * bv.x);
          h[4 * q + 1] = fmaf(ex2(dl * a2[4 * q + 1]), h[4 * q + 1], du * bv.y);
          h[4 * q + 2] = fmaf(ex2(dl * a2[4 * q + 2]), h[4 * q + 2], du * bv.z);
          h[4 * q + 3] = fmaf(ex2(dl * a2[4 * q + 3]), h[4 * q + 3], du * bv.w);
          sHt[(s * (NS / 4) + q) * NT] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        }
      }
    }

    // ---- reverse recurrence over the tile's steps ----
    float dAt[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) dAt[s] = 0.f;
#pragma unroll 1
    for (int s = npos - 1; s >= 0; --s) {
      const int r = w.rev ? npos - 1 - s : s;
      const float dl = sDL[r * FB_DT + c], dy = sDY[r * FB_DT + c];
      const float *row = sDB + r * Cp + n0;
      if (!MAIN) {
#pragma unroll
        for (int q = 0; q < NS / 4; ++q) {
          const float4 cv = *reinterpret_cast<const float4 *>(row + N + 4 * q);
          dh[4 * q] = fmaf(dy, cv.x, dh[4 * q]) * ex2(dl * a2[4 * q]);
          dh[4 * q + 1] = fmaf(dy, cv.y, dh[4 * q + 1]) * ex2(dl * a2[4 * q + 1]);
          dh[4 * q + 2] = fmaf(dy, cv.z, dh[4 * q + 2]) * ex2(dl * a2[4 * q + 2]);
          dh[4 * q + 3] = fmaf(dy, cv.w, dh[4 * q + 3]) * ex2(dl * a2[4 * q + 3]);
        }
        sumdl += dl;
        continue;
      }
      const float u = sXC[r * FB_DT + c];
      const float dlu = dl * u;
      float cB[NS], cC[NS];
      float s1 = 0.f, s2 = 0.f;     // Σ dh·B and Σ t·a2 over this lane's states
#pragma unroll
      for (int q = 0; q < NS / 4; ++q) {
        const float4 bv = *reinterpret_cast<const float4 *>(row + 4 * q);
        const float4 cv = *reinterpret_cast<const float4 *>(row + N + 4 * q);
        const float4 hv = sHt[(s * (NS / 4) + q) * NT];
        const float Bq[4] = {bv.x, bv.y, bv.z, bv.w}, Cq[4] = {cv.x, cv.y, cv.z, cv.w}, Hq[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int sI = 4 * q + e;
          const float a = ex2(dl * a2[sI]);
          const float dhn = fmaf(dy, Cq[e], dh[sI]);          // gradient reaching h at this step
          cC[sI] = dy * Hq[e];                                 // dC term
          const float t = dhn * fmaf(-dlu, Bq[e], Hq[e]);      // dh · a·h_prev,  a·h_prev = h - delta·u·B
          s1 = fmaf(dhn, Bq[e], s1);
          s2 = fmaf(t, a2[sI], s2);
          dAt[sI] = fmaf(t, dl, dAt[sI]);
          cB[sI] = dhn * dlu;                                  // dB term
          dh[sI] = dhn * a;
        }
      }
      // dB / dC: sum over the CPW channels of this warp that share the lane's state set, one coalesced red per row
      int wb = 0, wc = 0;
      const float rB = fb_transpose_reduce<NS, CPW / 2>(cB, lane, wb);
      const float rC = fb_transpose_reduce<NS, CPW / 2>(cC, lane, wc);
      constexpr int DUP = CPW / NS;
      if ((cl & (DUP - 1)) == 0) {
        float *dst = dxrow0 + ((long long)o * p.pso[w.k] + (long long)(i0 + r) * p.psi[w.k]) * p.K * Cp;
        atomicAdd(dst + n0 + wb, rB);
        atomicAdd(dst + N + n0 + wb, rC);
      }
      if (LPC == 2) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
      }
      float ddl = fmaf(u, s1, s2 * kLn2);
      const float duv = fmaf(dy, Dv, dl * s1);
      dDacc = fmaf(dy, u, dDacc);
      ddl *= 1.f - ex2(-dl * kLog2e);                          // softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x))
      dbacc += ddl;
      if (half == 0) { sdu[r * CPW + cl] = duv; sdd[r * CPW + cl] = ddl; }
    }
    if (MAIN) {
#pragma unroll
      for (int s = 0; s < NS; ++s) dAacc[s] += dAt[s];
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_4d(&p.m_dxc[w.k], sdu, w.d0 + warp * CPW, i0, o, w.b);
        tma_store_4d(&p.m_dd[w.k], sdd, w.d0 + warp * CPW, i0, o, w.k * p.batch + w.b);
        tma_store_commit();
        tma_store_wait_read<0>();
      }
    }
    __syncwarp();
    if (lane == 0 && kk + NST < ntl) {
      const uint32_t old = smem_inc_acq_rel(&done[st]);
      if ((old + 1) % (uint32_t)nwarps == 0) request_tile(kk + NST, st);
    }
    if (++st == NST) { st = 0; ph ^= 1; }
  }
  if (MAIN) {
    if (lane == 0) tma_store_wait_all<0>();
#pragma unroll
    for (int s = 0; s < NS; ++s) atomicAdd(&p.dA[wd * N + n0 + s], dAacc[s]);
    if (half == 0) {
      atomicAdd(&p.dDs[wd], dDacc);
      atomicAdd(&p.ddtb[wd], dbacc);
    }
  } else {
#pragma unroll
    for (int s = 0; s < NS; ++s) { carry_row[n0 + s] = ex2(a2[s] * sumdl); carry_row[N + n0 + s] = dh[s]; }
  }
}

// ---- host ----
constexpr int kFbMaxSplit = 64;

static size_t fb_al(size_t v) { return (v + 255) & ~(size_t)255; }

int ss2d_save_tiles(int kind, int H, int W);
static int fb_max_tiles(int kind, int H, int W) { return ss2d_save_tiles(kind, H, W); }
int ss2d_save_tiles(int kind, int H, int W) {
  const long long L = (long long)H * W;
  if (kind == SIGMA_DIRS_SEQ2) return (int)((2 * L + FB_LT - 1) / FB_LT);
  return (int)std::max<long long>((L + FB_LT - 1) / FB_LT, (long long)W * ((H + FB_LT - 1) / FB_LT));
}

size_t ss2d_scan_hs_bytes(int kind, int batch, int H, int W, int D, int N) {
  const int K = kind == SIGMA_DIRS_CROSS4 ? 4 : 2;
  return (size_t)K * batch * fb_max_tiles(kind, H, W) * D * N * sizeof(float);
}

// workspace = [hs (K, batch, max_tiles, D, N)] [forward carries] [reverse carries]
size_t ss2d_scan_bwd_workspace_bytes(int kind, int batch, int H, int W, int D, int N) {
  const int K = kind == SIGMA_DIRS_CROSS4 ? 4 : 2;
  const size_t carry = (size_t)batch * K * D * kFbMaxSplit * 2 * N * sizeof(float);
  return fb_al((size_t)K * batch * fb_max_tiles(kind, H, W) * D * N * sizeof(float)) + 2 * fb_al(carry);
}

// delta / ddelta: (K, batch, Lseq, D) slabs stored at the position a value belongs to; dxc (batch, Lseq, D) and dxdbl
// (batch, Lseq, K, Cp) are ACCUMULATED INTO after being zeroed here; dA (K·D, N), dDs (K·D), ddtb (K, D) overwritten.
int ss2d_scan_bwd(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                  const float *dy, float *delta, float *dxc, float *ddelta, float *dxdbl, float *dA, float *dDs, float *ddtb, int batch,
                  int H, int W, int D, int N, int R, int Cp, void *ws, size_t ws_bytes, int force_split, cudaStream_t stream,
                  const float *hs_saved) {
  if (ws == nullptr || ws_bytes < ss2d_scan_bwd_workspace_bytes(kind, batch, H, W, D, N)) {
    set_error("sigma_ss2d_scan_bwd: workspace too small (%zu < %zu)", ws_bytes, ss2d_scan_bwd_workspace_bytes(kind, batch, H, W, D, N));
    return SIGMA_EWORKSPACE;
  }
  Ss2dBwdParams p;
  memset(&p, 0, sizeof(p));
  const int K = kind == SIGMA_DIRS_CROSS4 ? 4 : 2;
  const long long Lseq = kind == SIGMA_DIRS_SEQ2 ? 2LL * H * W : (long long)H * W;
  p.dtw = dtw; p.dtb = dtb; p.A = A; p.Ds = Ds;
  p.dxdbl = dxdbl; p.dA = dA; p.dDs = dDs; p.ddtb = ddtb;
  p.D = D; p.N = N; p.R = R; p.Cp = Cp; p.K = K; p.batch = batch; p.Lseq = Lseq;
  p.max_tiles = fb_max_tiles(kind, H, W);
  const size_t hs_b = fb_al((size_t)K * batch * p.max_tiles * D * N * sizeof(float));
  const size_t carry_b = fb_al((size_t)batch * K * D * kFbMaxSplit * 2 * N * sizeof(float));
  p.hs = (float *)ws;
  p.hs_in = hs_saved ? hs_saved : p.hs;   // hs_saved: the training forward already wrote delta' and the block-start states
  float *fcarry = (float *)((char *)ws + hs_b), *rcarry = (float *)((char *)ws + hs_b + carry_b);
  const int CPW = N >= 16 ? 16 : 32;
  int rc, max_tiles = 0;
  for (int k = 0; k < K; ++k) {
    const bool colmajor = kind == SIGMA_DIRS_CROSS4 && (k & 1);
    p.rev[k] = kind == SIGMA_DIRS_CROSS4 ? (k >= 2) : (k == 1);
    uint64_t dims[4], str[3];
    uint32_t box[4] = {(uint32_t)FB_DT, (uint32_t)FB_LT, 1, 1}, boxw[4] = {(uint32_t)CPW, (uint32_t)FB_LT, 1, 1};
    if (!colmajor) {
      p.I[k] = (int)Lseq; p.O[k] = 1; p.psi[k] = 1; p.pso[k] = 0;
      dims[0] = D; dims[1] = Lseq; dims[2] = 1; dims[3] = batch;
      str[0] = (uint64_t)D * 4; str[1] = (uint64_t)Lseq * D * 4; str[2] = (uint64_t)Lseq * D * 4;
    } else {   // inner index h at fixed w: position h·W + w
      p.I[k] = H; p.O[k] = W; p.psi[k] = W; p.pso[k] = 1;
      dims[0] = D; dims[1] = H; dims[2] = W; dims[3] = batch;
      str[0] = (uint64_t)W * D * 4; str[1] = (uint64_t)D * 4; str[2] = (uint64_t)Lseq * D * 4;
    }
    if ((rc = make_tmap_f32_4d(&p.m_xc[k], xc, dims, str, box))) return rc;
    if ((rc = make_tmap_f32_4d(&p.m_dy[k], dy, dims, str, box))) return rc;
    if ((rc = make_tmap_f32_4d(&p.m_dxc[k], dxc, dims, str, boxw))) return rc;
    dims[3] = (uint64_t)K * batch;   // slabs: image index k·batch + b
    if ((rc = make_tmap_f32_4d(&p.m_dl[k], delta, dims, str, box))) return rc;
    dims[3] = batch;
    uint32_t boxd[4] = {(uint32_t)Cp, (uint32_t)FB_LT, 1, 1};
    dims[0] = Cp;
    const uint64_t pos = (uint64_t)K * Cp * 4;
    if (!colmajor) { str[0] = pos; str[1] = Lseq * pos; str[2] = Lseq * pos; }
    else { str[0] = W * pos; str[1] = pos; str[2] = Lseq * pos; }
    if ((rc = make_tmap_f32_4d(&p.m_dbl[k], xdbl + (long long)k * Cp, dims, str, boxd))) return rc;
    max_tiles = std::max(max_tiles, p.O[k] * ((p.I[k] + FB_LT - 1) / FB_LT));
  }
  // L-segments: all directions share tiles_per_split (directions with fewer tiles get empty trailing segments)
  const int lpc = N >= 16 ? 2 : 1;
  int nsplit = pick_segments((long long)(D / FB_DT) * K * batch, max_tiles, 148LL * (lpc == 2 ? 2 : 4), 1.3, kFbMaxSplit);
  if (force_split > 0) nsplit = std::min(force_split, kFbMaxSplit);
  nsplit = std::max(1, std::min(nsplit, max_tiles));
  p.tiles_per_split = (max_tiles + nsplit - 1) / nsplit;
  p.nsplit = (max_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.nst = 2;

  SIGMA_CHECK_CUDA(cudaMemsetAsync(dxc, 0, (size_t)batch * Lseq * D * sizeof(float), stream));
  SIGMA_CHECK_CUDA(cudaMemsetAsync(dxdbl, 0, (size_t)batch * Lseq * K * Cp * sizeof(float), stream));
  SIGMA_CHECK_CUDA(cudaMemsetAsync(dA, 0, (size_t)K * D * N * sizeof(float), stream));
  SIGMA_CHECK_CUDA(cudaMemsetAsync(dDs, 0, (size_t)K * D * sizeof(float), stream));
  SIGMA_CHECK_CUDA(cudaMemsetAsync(ddtb, 0, (size_t)K * D * sizeof(float), stream));

  // the state sweep stores delta' through m_dd (per-warp boxes over the delta slabs); the main sweep re-points it at ddelta
  auto make_dd = [&](float *slab) -> int {
    for (int k = 0; k < K; ++k) {
      const bool colmajor = kind == SIGMA_DIRS_CROSS4 && (k & 1);
      uint64_t dims[4], str[3];
      uint32_t boxw[4] = {(uint32_t)CPW, (uint32_t)FB_LT, 1, 1};
      if (!colmajor) {
        dims[0] = D; dims[1] = Lseq; dims[2] = 1; dims[3] = (uint64_t)K * batch;
        str[0] = (uint64_t)D * 4; str[1] = (uint64_t)Lseq * D * 4; str[2] = (uint64_t)Lseq * D * 4;
      } else {
        dims[0] = D; dims[1] = H; dims[2] = W; dims[3] = (uint64_t)K * batch;
        str[0] = (uint64_t)W * D * 4; str[1] = (uint64_t)D * 4; str[2] = (uint64_t)Lseq * D * 4;
      }
      int r = make_tmap_f32_4d(&p.m_dd[k], slab, dims, str, boxw);
      if (r) return r;
    }
    return SIGMA_OK;
  };
  if ((rc = make_dd(delta))) return rc;
  Ss2dBwdParams ps = p;
  if ((rc = make_dd(ddelta))) return rc;
  Ss2dBwdParams pm = p;
  // ps holds m_dd -> delta (state sweep), pm holds m_dd -> ddelta (main sweep)
  auto go = [&](auto tag) -> int {
    constexpr int NN = decltype(tag)::value;
    constexpr int LPC = FbCfg<NN>::LPC, NS = FbCfg<NN>::NS, CPWc = FbCfg<NN>::CPW, NT = FB_DT * LPC;
    dim3 grid(D / FB_DT, K * pm.nsplit, batch), block(NT);
    const long long nrows = (long long)batch * K * D, tot = nrows * NN;
    const size_t st_smem = ((size_t)pm.nst * (2 * FB_LT * FB_DT + FB_LT * Cp) + FB_DT * (R + 1) + 2 + (NT / 32) * FB_LT * CPWc) * sizeof(float) + 256;
    const size_t sm_smem = ((size_t)pm.nst * (2 * FB_LT * FB_DT + FB_LT * Cp)) * sizeof(float) + 256;
    const size_t mn_smem = ((size_t)pm.nst * (3 * FB_LT * FB_DT + FB_LT * Cp) + (NT / 32) * 2 * FB_LT * CPWc + (size_t)FB_LT * NS * NT) * sizeof(float) + 256;
    auto run = [&](auto kern, const Ss2dBwdParams &pp, size_t smem) -> int {
      SIGMA_CHECK_CUDA(prep_kernel_once((const void *)kern));
      kern<<<grid, block, smem, stream>>>(pp);
      SIGMA_CHECK_LAUNCH();
      return SIGMA_OK;
    };
    int r;
    ps.carry = fcarry;
    if (hs_saved != nullptr) {
      // nothing to recompute
    } else if (pm.nsplit == 1) {
      if ((r = run(ss2d_state_kernel<NN, MODE_SERIAL>, ps, st_smem))) return r;
    } else {
      if ((r = run(ss2d_state_kernel<NN, MODE_SUMMARY>, ps, st_smem))) return r;
      scan_combine_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(fcarry, nrows, pm.nsplit, NN);
      SIGMA_CHECK_LAUNCH();
      if ((r = run(ss2d_state_kernel<NN, MODE_APPLY>, ps, st_smem))) return r;
    }
    pm.carry = rcarry;
    if (pm.nsplit == 1) return run(ss2d_bwd_kernel<NN, MODE_SERIAL>, pm, mn_smem);
    if ((r = run(ss2d_bwd_kernel<NN, MODE_SUMMARY>, pm, sm_smem))) return r;
    scan_combine_rev_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(rcarry, nrows, pm.nsplit, NN);
    SIGMA_CHECK_LAUNCH();
    return run(ss2d_bwd_kernel<NN, MODE_APPLY>, pm, mn_smem);
  };
  if (N == 16) return go(std::integral_constant<int, 16>{});
  return go(std::integral_constant<int, 4>{});
}

}  // namespace sigma
