// temperature = true together with aging = :SEI (grids of up to 32 x-nodes): N = 372 for N = (10,10,10), N_a = N_z = 10
#define PLB_TH 1
#define PLB_SEI 1
#define PLB_NS thsei
#include "plb_variant.cuh"
