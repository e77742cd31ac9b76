// dense FK + Jacobian kernels for float paths of 23..32 joints (see nbx_fkjac_wide.inc)
#define NBX_WIDE_LO 23
#define NBX_WIDE_HI 32
#define NBX_WIDE_FN nbx_launch_wide_b_f32
#include "nbx_fkjac_wide.inc"
