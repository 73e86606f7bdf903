#define SIGMA_RP 4
#include "ss2d_scan_inst.inc"
