// Host side of the fused SS2D scan: tensor maps, L-segment planning, dispatch (sigma_ss2d_scan_fwd).
#include <stdlib.h>

#include <algorithm>

#include "ss2d_scan.cuh"

namespace sigma {

// ---- tensor map creation through the runtime-resolved driver entry point ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = (EncodeTiledFn)ptr;
  return fn;
}

void *get_tensor_map_encoder() { return (void *)get_encode_fn(); }

int make_tmap_f32_4d(CUtensorMap *map, const void *base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4]) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return SIGMA_ECUDA;
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bdim[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) "
              "box=(%u,%u,%u,%u) base=%p",
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
              (unsigned long long)strides_bytes[2], box[0], box[1], box[2], box[3], base);
    return SIGMA_ECUDA;
  }
  return SIGMA_OK;
}

static int pad_rp(int R) {
  const int opts[] = {4, 8, 12, 16, 24, 32, 48, 64};
  for (int o : opts)
    if (R <= o) return o;
  return -1;
}

template <int N, int CPT>
static int dispatch_rp(const Ss2dParams &p, int nthreads, int ctas, cudaStream_t s) {
  switch (pad_rp(p.R)) {
    case 4: return ss2d_launch<N, CPT, 4>(p, nthreads, ctas, s);
    case 8: return ss2d_launch<N, CPT, 8>(p, nthreads, ctas, s);
    case 12: return ss2d_launch<N, CPT, 12>(p, nthreads, ctas, s);
    case 16: return ss2d_launch<N, CPT, 16>(p, nthreads, ctas, s);
    case 24: return ss2d_launch<N, CPT, 24>(p, nthreads, ctas, s);
    case 32: return ss2d_launch<N, CPT, 32>(p, nthreads, ctas, s);
    case 48: return ss2d_launch<N, CPT, 48>(p, nthreads, ctas, s);
    case 64: return ss2d_launch<N, CPT, 64>(p, nthreads, ctas, s);
  }
  set_error("sigma_ss2d_scan_fwd: dt_rank %d > 64 unsupported", p.R);
  return SIGMA_EUNSUPPORTED;
}

static int kind_dirs(int kind) { return kind == SIGMA_DIRS_CROSS4 ? 4 : (kind == SIGMA_DIRS_SEQ2 ? 2 : 1); }
static int lt_for(int N) { return N >= 16 ? Ss2dCfg<16>::LT : Ss2dCfg<4>::LT; }

// warps per CTA: a count <= maxw whose channel tile (32·cpt channels per warp) divides D, tried in the order 4, 2, 3, 1 —
// 3-warp CTAs (D = 192: 2 x 96 channels) measured 10 % slower than 2-warp ones (3 x 64) on the four-direction scans at the
// same 12 resident warps per SM, every other shape is fastest with 4 (profiles/r02_scan_warps_sweep.txt); ragged D falls
// back to enough warps to cover it with a partially filled last CTA (TMA zero-fills, y goes to the sink)
static int pick_warps(int D, int cpt, int maxw) {
  const int cpw = 32 * cpt;
  const int order[4] = {4, 2, 3, 1};
  for (int w : order)
    if (w <= maxw && D % (cpw * w) == 0) return w;
  return std::max(1, std::min(maxw, (D + cpw - 1) / cpw));
}

constexpr int kMaxSplit = 32;

int ss2d_save_tiles(int kind, int H, int W);   // ss2d_scan_bwd.cu: 16-position blocks of the longest walk

// Which register budget to run (`CTAS` of ss2d_scan_kernel), from B200 measurements of the Sigma shapes at B = 74
// (profiles/r02_scan_ctas_sweep.txt): the scans are MUFU / MIO-limited, so 16 or 20 resident warps instead of 12 change
// little — d_state 16 gains 3 % at dt_rank 24 (stage 2, nine blocks) and loses 5-12 % at dt_rank 6 / 48; d_state 4 is
// flat or slower.  Only that one case leaves the default.
static int ss2d_pick_ctas(int N, int rp) {
  return (N == 16 && rp == 24) ? 4 : 3;
}

// Number of L-segments for a grid of `ctas` CTAs of `nw` warps walking `ntiles` tiles each.  Model: CTAs spread evenly over the
// 148 SMs, the kernel lasts as long as its busiest SM; a warp's time per tile grows with the warps sharing its sub-partition
// (MUFU / issue are shared: per-warp cost 1 : 1.5 : 2.25 for 1 : 2 : 3 warps); a segmented walk runs two passes (summary +
// apply) and pays a fixed ring fill / parameter load per CTA (~2 tiles).  The three constants are a least-squares fit to the
// B200 sweep of every Sigma call shape x {1, 2, 4, 8} images x 10 forced counts (profiles/r02_ss2d_split_sweep.txt): the
// model's choices total 15.1 ms there, the per-row optimum 14.9 ms, the round-1 rule 19.9 ms.  Candidates within 3 % keep
// the smaller count.
static int ss2d_pick_segments(long long ctas, int nw, int ntiles, int N) {
  (void)N;                                                    // d_state 4 and 16 fit the same constants
  const double pf = 1.95;
  const int cap_warps = 12;                                   // resident warps per SM (168 registers)
  double best = 1e300;
  int best_n = 1;
  for (int n = 1; n <= std::min(kMaxSplit, ntiles); ++n) {
    const int tps = (ntiles + n - 1) / n, ne = (ntiles + tps - 1) / tps;
    if (ne != n) continue;
    const long long per_sm = (ctas * ne + 147) / 148;         // CTAs on the busiest SM
    const long long warps_sm = per_sm * nw;
    const long long rounds = (warps_sm + cap_warps - 1) / cap_warps;
    const long long w_smsp = (std::min<long long>(warps_sm, cap_warps) + 3) / 4;   // warps sharing a sub-partition
    const double share = std::max(1.0, 0.75 * (double)w_smsp);
    const double cost = (double)rounds * (tps + 2) * share * (ne > 1 ? pf : 1.0);
    if (cost < best * 0.97) { best = cost; best_n = ne; }
  }
  return best_n;
}

int ss2d_pick_segments_hook(long long ctas, int nw, int ntiles, int N) { return ss2d_pick_segments(ctas, nw, ntiles, N); }   // api.cu test hook

size_t ss2d_scan_workspace_bytes(int kind, int batch, int D, int N) {
  return (size_t)batch * kind_dirs(kind) * D * kMaxSplit * 2 * N * sizeof(float);
}

int ss2d_scan_fwd(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A,
                  const float *Ds, float *y, int batch, int H, int W, int D, int N, int R, int Cp, void *ws,
                  size_t ws_bytes, int force_split, cudaStream_t stream, float *dsave, float *hsave) {
  Ss2dParams p;
  memset(&p, 0, sizeof(p));
  p.dtw = dtw; p.dtb = dtb; p.A = A; p.Ds = Ds; p.y = y; p.carry = (float *)ws;
  p.D = D; p.N = N; p.R = R; p.Cp = Cp; p.kind = kind; p.batch = batch;
  p.dsave = dsave; p.hsave = hsave; p.save_tiles = hsave ? ss2d_save_tiles(kind, H, W) : 0;
  const int ndir = kind_dirs(kind);
  p.ndir = ndir;
  if (const char *e = getenv("SIGMA_SCAN_ABLATE")) p.ablate = atoi(e);
  const int K = kind == SIGMA_DIRS_CROSS ? 1 : ndir;        // x_dbl rows per position
  const long long Lseq = kind == SIGMA_DIRS_SEQ2 ? 2LL * H * W : (long long)H * W;
  p.Lseq = Lseq;
  const int LT = lt_for(N);
  const int cpt = 1;   // channels per thread; CPT = 2 (shared B/C reads) lost at every Sigma shape (profiles/r01_scan_variants.txt)
  int maxw = 4;  // warps per CTA (Ss2dCfg::MAXW; the kernels' register budget assumes 128 threads)
  if (const char *e = getenv("SIGMA_SCAN_WARPS")) maxw = std::max(1, std::min(4, atoi(e)));
  const int NW = pick_warps(D, cpt, maxw), DT = 32 * cpt * NW;
  int rc;
  int max_tiles = 0;
  for (int k = 0; k < ndir; ++k) {
    const bool colmajor = kind == SIGMA_DIRS_CROSS4 && (k & 1);
    p.rev[k] = (kind == SIGMA_DIRS_CROSS4) ? (k >= 2) : (kind == SIGMA_DIRS_SEQ2 ? (k == 1) : 0);
    uint64_t dims[4], str[3];
    uint32_t box[4] = {(uint32_t)DT, (uint32_t)LT, 1, 1};
    // channels-last activations (batch, Lseq, D)
    if (!colmajor) {
      p.I[k] = (int)Lseq; p.O[k] = 1;
      p.istride[k] = D; p.ostride[k] = 0;
      dims[0] = D; dims[1] = Lseq; dims[2] = 1; dims[3] = batch;
      str[0] = (uint64_t)D * 4; str[1] = (uint64_t)Lseq * D * 4; str[2] = (uint64_t)Lseq * D * 4;
    } else {  // walk h (inner) at fixed w (outer): l1 = w·H + h  <->  position h·W + w   (vmamba.py:87)
      p.I[k] = H; p.O[k] = W;
      p.istride[k] = (long long)W * D; p.ostride[k] = D;
      dims[0] = D; dims[1] = H; dims[2] = W; dims[3] = batch;
      str[0] = (uint64_t)W * D * 4; str[1] = (uint64_t)D * 4; str[2] = (uint64_t)Lseq * D * 4;
    }
    if ((rc = make_tmap_f32_4d(&p.m_xc[k], xc, dims, str, box))) return rc;
    // x_dbl (batch, Lseq, K, Cp): direction k's row starts at column k·Cp
    uint32_t boxd[4] = {(uint32_t)Cp, (uint32_t)LT, 1, 1};
    dims[0] = Cp;
    const uint64_t pos = (uint64_t)K * Cp * 4;
    if (!colmajor) { str[0] = pos; str[1] = Lseq * pos; str[2] = Lseq * pos; }
    else { str[0] = W * pos; str[1] = pos; str[2] = Lseq * pos; }
    if ((rc = make_tmap_f32_4d(&p.m_dbl[k], xdbl + (long long)(kind == SIGMA_DIRS_CROSS ? 0 : k) * Cp, dims, str, boxd)))
      return rc;
    max_tiles = std::max(max_tiles, p.O[k] * ((p.I[k] + LT - 1) / LT));
  }
  // L-segments (MODE_SUMMARY -> combine -> MODE_APPLY): a second pass over the data, so only when the unsplit grid leaves SM
  // sub-partitions without a warp.  ss2d_pick_segments models the busiest SM (the old rule, "fill 592 warp slots", ignored
  // wave quantisation: 312 CTAs of 2 warps put 6 warps on some SMs where 288 put 4 on every SM).  SIGMA_SCAN_SPLIT_RULE=old
  // restores the round-1 rule for comparison.
  const long long ctas = (long long)((D + DT - 1) / DT) * ndir * batch;
  const long long warps = ctas * NW;
  const long long full = 148LL * 4;
  int nsplit = 1;
  const char *rule = getenv("SIGMA_SCAN_SPLIT_RULE");
  if (rule && rule[0] == 'o') {
    if (warps < full) {
      const long long want = N >= 16 ? full : 2 * full;
      nsplit = (int)std::min<long long>((want + warps - 1) / warps, kMaxSplit);
    }
  } else if (warps < 3 * full) {
    nsplit = ss2d_pick_segments(ctas, NW, max_tiles, N);
  }
  if (force_split > 0) nsplit = std::min(force_split, kMaxSplit);
  if (ws == nullptr || ws_bytes < ss2d_scan_workspace_bytes(kind, batch, D, N)) {
    if (force_split > 1) { set_error("sigma_ss2d_scan_fwd: workspace too small for %d segments", force_split); return SIGMA_EWORKSPACE; }
    nsplit = 1;
  }
  nsplit = std::max(1, std::min(nsplit, max_tiles));
  // all directions share tiles_per_split; directions with fewer tiles simply get empty trailing segments
  p.tiles_per_split = (max_tiles + nsplit - 1) / nsplit;
  p.nsplit = nsplit;

  // register budget (ss2d_scan.cuh): which `__launch_bounds__(128, CTAS)` build runs.  SIGMA_SCAN_CTAS overrides.
  int rbud = ss2d_pick_ctas(N, pad_rp(R));
  if (const char *e = getenv("SIGMA_SCAN_CTAS")) rbud = std::max(3, std::min(5, atoi(e)));
  if (N != 16) rbud = 3;
  rbud = std::min(rbud, 4);
  {
    // TMA ring depth: as many stages as fit without lowering the register-limited occupancy (227 KB per SM, 1 KB
    // reserved per CTA).
    // Deep rings matter: a tile is requested when the LAST warp releases its slot and needed by the FIRST warp
    // nst-1 tiles later; with 3-4 stages the warps spun on the full barrier ~80 times per tile (ncu, round 1).
    const size_t stage = ((size_t)LT * DT + (size_t)LT * Cp * (kind == SIGMA_DIRS_CROSS ? 2 : 1)) * sizeof(float);
    // resident CTAs per SM by registers (e.g. 168 per thread under __launch_bounds__(128, 3): 3 / 4 / 6 / 12 for 4 / 3 / 2 / 1 warps)
    const int ctas_sm = std::max(rbud, std::min(16, 65536 / (32 * NW * ss2d_reg_cap(rbud))));
    const size_t budget = (227 * 1024) / ctas_sm - 1024 - 128;
    p.nst = (int)std::max<size_t>(2, std::min<size_t>(Ss2dCfg<16>::MAX_NST, budget / stage));
    if (const char *e = getenv("SIGMA_SCAN_NST")) p.nst = std::max(2, std::min(Ss2dCfg<16>::MAX_NST, atoi(e)));
  }
  const int nthreads = 32 * NW;
  switch (N) {
    case 4: return dispatch_rp<4, 1>(p, nthreads, rbud, stream);
    case 8: return dispatch_rp<8, 1>(p, nthreads, rbud, stream);
    case 16: return dispatch_rp<16, 1>(p, nthreads, rbud, stream);
  }
  set_error("sigma_ss2d_scan_fwd: d_state=%d unsupported by the fused kernel (4, 8, 16)", N);
  return SIGMA_EUNSUPPORTED;
}

}  // namespace sigma
