// the thermal family with Fickian_method = :spectral (Chebyshev collocation in the particles, residuals.jl:181-235; "BETA" in
// the reference): a dense constant particle block and a j coupling on every radial row (laws_generated.cuh, namespace sp10)
#define PLB_TH 1
#define PLB_SEI 0
#define PLB_SPECTRAL 1
#define PLB_NS thsp
#include "plb_variant.cuh"
