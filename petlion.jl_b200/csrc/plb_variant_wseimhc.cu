// the two-warp SEI family (33..64 x-nodes) with rxn_MHC compiled in next to rxn_BV
#define PLB_TH 0
#define PLB_SEI 1
#define PLB_WIDE 1
#define PLB_MHC 1
#define PLB_NS wseimhc
#include "plb_variant.cuh"
