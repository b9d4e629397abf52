// isothermal model family (temperature = false): N = 301 for N = (10,10,10), N_r = 10
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_NS iso
#include "plb_variant.cuh"
