// temperature = true with aging = :SEI (up to 32 x-nodes) with rxn_MHC compiled in next to rxn_BV
#define PLB_TH 1
#define PLB_SEI 1
#define PLB_WIDE 0
#define PLB_MHC 1
#define PLB_NS thseimhc
#include "plb_variant.cuh"
