"""CPU: the two launch heuristics of the latency regime, held to the recorded B200 sweeps through host-only test hooks.
* `ss2d_pick_segments` (L-segment count of the fused scan, a cost model of the busiest SM): replayed on every row of
  profiles/r02_ss2d_split_sweep.txt (15 call shapes x {1,2,4,8} images x 10 forced counts) its choices must stay within 6 % of the
  per-row optimum in total and well ahead of the round-1 rule.
* `pick_bn` (GEMM tile width): the widest divisor in the throughput regime, narrower tiles when few row tiles exist."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPLITS = [1, 2, 3, 4, 6, 8, 12, 16, 24, 32]
# name -> (streams per image, H, W, D, d_state, directions, sequence factor)
SHAPES = {"enc0": (2, 120, 160, 192, 16, 4, 1), "enc1": (2, 60, 80, 384, 16, 4, 1), "enc2": (2, 30, 40, 768, 16, 4, 1),
          "enc3": (2, 15, 20, 1536, 16, 4, 1), "dec0": (1, 120, 160, 192, 4, 4, 1), "dec1": (1, 60, 80, 384, 4, 4, 1),
          "dec2": (1, 30, 40, 768, 4, 4, 1), "conmb0": (1, 120, 160, 192, 4, 2, 2), "cromb0": (2, 120, 160, 192, 4, 1, 1),
          "conmb1": (1, 60, 80, 384, 4, 2, 2), "cromb1": (2, 60, 80, 384, 4, 1, 1), "conmb2": (1, 30, 40, 768, 4, 2, 2),
          "cromb2": (2, 30, 40, 768, 4, 1, 1), "conmb3": (1, 15, 20, 1536, 4, 2, 2), "cromb3": (2, 15, 20, 1536, 4, 1, 1)}


@pytest.fixture(scope="module")
def L():
    from sigma_b200 import _lib
    return _lib.lib()


def _rows():
    rows = []
    for ln in open(os.path.join(ROOT, "profiles", "r02_ss2d_split_sweep.txt")):
        p = ln.split()
        if len(p) == 15 and p[1].isdigit() and p[0] in SHAPES:
            rows.append((p[0], int(p[1]), [float(v) for v in p[2:12]], float(p[12]), float(p[14])))
    return rows


def test_segment_model_tracks_the_recorded_sweep(L):
    rows = _rows()
    assert len(rows) == 60
    tot_model = tot_best = tot_old = 0.0
    for name, images, times, t_old, t_best in rows:
        spi, H, W, D, N, K, seq = SHAPES[name]
        LT = 16 if N >= 16 else 32
        nw = next(w for w in (4, 2, 3, 1) if D % (32 * w) == 0)          # pick_warps (ss2d_scan_host.cu)
        ctas = (D // (32 * nw)) * K * images * spi
        ntiles = max(-(-(H * W * seq) // LT), W * (-(-H // LT)) if K == 4 else 0)
        n = L.sigma_test_pick_segments(ctas, nw, ntiles, N)
        assert 1 <= n <= 32
        lo = max(s for s in SPLITS if s <= n)
        hi = min(s for s in SPLITS if s >= n)
        tot_model += max(times[SPLITS.index(lo)], times[SPLITS.index(hi)])          # conservative between measured counts
        tot_best += t_best
        tot_old += t_old
    assert tot_model <= 1.06 * tot_best, (tot_model, tot_best)
    assert tot_model <= 0.85 * tot_old, (tot_model, tot_old)


def test_segment_model_leaves_full_grids_alone(L):
    # B = 74 shapes: every sub-partition has its warps already; a second pass can only lose
    assert L.sigma_test_pick_segments(3 * 4 * 148, 2, 1280, 16) == 1       # enc0
    assert L.sigma_test_pick_segments(6 * 4 * 148, 4, 80, 16) == 1         # enc2
    assert L.sigma_test_pick_segments(2 * 4 * 74, 2, 640, 4) == 1          # dec0
    assert L.sigma_test_pick_segments(1, 4, 1, 16) == 1                    # one tile cannot be cut


@pytest.mark.parametrize("N,want", [(96, 96), (160, 160), (176, 192), (192, 192), (384, 192), (768, 256), (1536, 256), (3072, 256), (9, 32)])
def test_gemm_tile_width_throughput_regime(L, N, want):
    assert L.sigma_test_pick_bn(N, 1 << 30) == want


@pytest.mark.parametrize("N", [96, 192, 384, 768, 1536, 3072])
@pytest.mark.parametrize("m_tiles", [1, 2, 5, 19, 75, 300, 22200])
def test_gemm_tile_width_is_a_valid_tile(L, N, m_tiles):
    bn = L.sigma_test_pick_bn(N, m_tiles)
    assert bn % 32 == 0 and 32 <= bn <= 256
    wide = L.sigma_test_pick_bn(N, 1 << 30)
    assert bn <= wide                                   # never wider than the throughput choice
    if m_tiles * -(-N // wide) >= 148 * 4:
        assert bn == wide                               # enough tiles for every persistent CTA: nothing to gain from narrow tiles
    if m_tiles <= 5 and N >= 768:
        assert bn <= 128                                # a handful of row tiles: spread the columns over more CTAs
