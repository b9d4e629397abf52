// the two-warp isothermal family (33..64 x-nodes) with rxn_MHC compiled in next to rxn_BV (custom_functions.jl:233-298)
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_WIDE 1
#define PLB_MHC 1
#define PLB_NS widemhc
#include "plb_variant.cuh"
