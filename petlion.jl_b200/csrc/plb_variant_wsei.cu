// aging = :SEI on grids with 33..64 x-nodes (BASELINE configs[4]: N = (20,20,20), 642 DAEs): two warps per system
#define PLB_TH 0
#define PLB_SEI 1
#define PLB_WIDE 1
#define PLB_NS wsei
#include "plb_variant.cuh"
