#define SIGMA_RP 24
#include "ss2d_scan_inst.inc"
