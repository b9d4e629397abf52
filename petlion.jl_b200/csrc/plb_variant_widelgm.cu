// NMC_LGM50 / LiC6_LGM50 (params.jl:514-849), isothermal, on 33..64 x-nodes: the two-warp family instantiated for that chemistry only
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_WIDE 1
#define PLB_ONLY_CHEM CHEM_LGM
#define PLB_NS widelgm
#include "plb_variant.cuh"
