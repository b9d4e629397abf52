// thermal model family (temperature = true): N = 351 for N = (10,10,10), N_a = N_z = 10, N_r = 10
#define PLB_TH 1
#define PLB_SEI 0
#define PLB_NS th
#include "plb_variant.cuh"
