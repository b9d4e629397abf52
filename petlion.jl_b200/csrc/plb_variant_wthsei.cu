// temperature = true together with aging = :SEI on grids with 33..64 x-nodes (two warps per system): N = 722 for N = (20,20,20)
#define PLB_TH 1
#define PLB_SEI 1
#define PLB_WIDE 1
#define PLB_NS wthsei
#include "plb_variant.cuh"
