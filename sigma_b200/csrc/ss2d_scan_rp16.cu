#define SIGMA_RP 16
#include "ss2d_scan_inst.inc"
