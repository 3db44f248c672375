// Bit-exact build of the hot-path kernels: compiled with -fmad=false so that every float32 operation is
// individually rounded (IEEE div/sqrt, no flush-to-zero), like the reference's .NET arithmetic.
#define FT_NS ft_strict
#include "kernels_impl.cuh"
