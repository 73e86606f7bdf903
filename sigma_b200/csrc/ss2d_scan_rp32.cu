#define SIGMA_RP 32
#include "ss2d_scan_inst.inc"
