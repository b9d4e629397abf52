// the two-warp thermal family (33..64 x-nodes) with rxn_MHC compiled in next to rxn_BV
#define PLB_TH 1
#define PLB_SEI 0
#define PLB_WIDE 1
#define PLB_MHC 1
#define PLB_NS wthmhc
#include "plb_variant.cuh"
