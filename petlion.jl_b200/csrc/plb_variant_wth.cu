// thermal model family on grids with 33..64 x-nodes (two warps per system): e.g. N = 681 for N = (20,20,20), N_a = N_z = 10
#define PLB_TH 1
#define PLB_SEI 0
#define PLB_WIDE 1
#define PLB_NS wth
#include "plb_variant.cuh"
