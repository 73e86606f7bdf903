#define SIGMA_RP 8
#include "ss2d_scan_inst.inc"
