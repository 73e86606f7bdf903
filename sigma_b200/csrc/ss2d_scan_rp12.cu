#define SIGMA_RP 12
#include "ss2d_scan_inst.inc"
