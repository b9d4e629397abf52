// isothermal model family on grids with 33..64 x-nodes (e.g. N = (20,20,20): 601 DAEs): two warps per system
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_WIDE 1
#define PLB_NS wide
#include "plb_variant.cuh"
