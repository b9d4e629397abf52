// the isothermal family compiled for N_r_p = N_r_n = 12 radial nodes per particle (params.jl:134-136): its own stencil,
// eigen-basis, lane registers, workspace stride and recipe tables (laws_generated.cuh, namespace nr12); selected by plb_create
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_NR 12
#define PLB_NS iso12
#include "plb_variant.cuh"
