// G2 instantiation of the bucket MSM.
#define B2Z_CURVE G2
#include "msm_impl.inc"
