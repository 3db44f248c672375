// Contracted build of the hot-path kernels (-fmad=true): FT_FLAG_FAST_MATH.  Not bit-identical to the
// reference; parity for this build is tolerance-based.
#define FT_NS ft_fast
#define FT_WITH_CONE 1
#include "kernels_impl.cuh"
