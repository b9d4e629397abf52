// the isothermal family once more, with rxn_MHC (custom_functions.jl:233-298) compiled in next to rxn_BV: only models that ask
// for it are sent here, so the default family keeps its code, registers and occupancy
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_MHC 1
#define PLB_NS isomhc
#include "plb_variant.cuh"
