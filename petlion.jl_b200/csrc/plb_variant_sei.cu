// aging = :SEI model family (isothermal): N = 322 for N = (10,10,10), N_r = 10
#define PLB_TH 0
#define PLB_SEI 1
#define PLB_NS sei
#include "plb_variant.cuh"
