#define SIGMA_RP 64
#include "ss2d_scan_inst.inc"
