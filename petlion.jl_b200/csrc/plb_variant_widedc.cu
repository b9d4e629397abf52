// the isothermal two-warps-per-system family with the concentration-rate inputs dc_s_* / dc_e_* compiled in
#define PLB_TH 0
#define PLB_SEI 0
#define PLB_WIDE 1
#define PLB_DC 1
#define PLB_NS widedc
#include "plb_variant.cuh"
