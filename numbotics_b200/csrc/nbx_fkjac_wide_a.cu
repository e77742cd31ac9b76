// dense FK + Jacobian kernels for float paths of 9..22 joints (see nbx_fkjac_wide.inc)
#define NBX_WIDE_LO 9
#define NBX_WIDE_HI 22
#define NBX_WIDE_FN nbx_launch_wide_a_f32
#include "nbx_fkjac_wide.inc"
