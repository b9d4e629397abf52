// the isothermal family once more, with the concentration-rate inputs dc_s_* / dc_e_* (input_methods.jl:190-245) compiled in:
// only runs that ask for them are sent here, so the headline family (plb_variant_iso.cu) keeps its code and registers
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_DC 1
#define PLB_NS isodc
#include "plb_variant.cuh"
