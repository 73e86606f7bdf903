#define SIGMA_RP 48
#include "ss2d_scan_inst.inc"
