// float64 / reference-exact mode of every entry point.  MUST be compiled with -fmad=false: the
// expression order in kernels.cuh is then exactly the arithmetic of the reference's python.
#define EMEI_REAL double
#define EMEI_FN(name) name##_f64
#include "impl.inc"
