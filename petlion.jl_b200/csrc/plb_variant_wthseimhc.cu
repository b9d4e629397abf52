// temperature = true with aging = :SEI on 33..64 x-nodes with rxn_MHC compiled in next to rxn_BV
#define PLB_TH 1
#define PLB_SEI 1
#define PLB_WIDE 1
#define PLB_MHC 1
#define PLB_NS wthseimhc
#include "plb_variant.cuh"
