// NMC_LGM50 / LiC6_LGM50 (Chen et al. 2020; params.jl:514-849), temperature = true: the thermal family instantiated for that
// chemistry only (its own OCVs and electrolyte laws, laws_generated.cuh); selected by plb_create for cathode = NMC_LGM50
#define PLB_TH 1
#define PLB_SEI 0
#define PLB_ONLY_CHEM CHEM_LGM
#define PLB_NS thlgm
#include "plb_variant.cuh"
