// G1 instantiation of the bucket MSM (+ the curve-independent host helpers).
#define B2Z_CURVE G1
#define B2Z_MSM_COMMON 1
#include "msm_impl.inc"
