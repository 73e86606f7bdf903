// EXPERIMENTAL: d_state-16 scan with NPOLY of the 8 state pairs per position taking their exponentials from the
// FMA pipe (ex2_poly, ss2d_scan.cuh).  Selected only by SIGMA_SCAN_POLY=<NPOLY>; never on the default path.
#define SIGMA_RP 24
#define SIGMA_SCAN_POLY_TU
#include "ss2d_scan_inst.inc"
namespace sigma {
template int ss2d_launch<16, 1, 24, 1>(const Ss2dParams &, int, cudaStream_t);
template int ss2d_launch<16, 1, 24, 2>(const Ss2dParams &, int, cudaStream_t);
template int ss2d_launch<16, 1, 24, 3>(const Ss2dParams &, int, cudaStream_t);
}
