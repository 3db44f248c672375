// Bit-exact build of the hot-path kernels (like kernels_strict.cu) that also evaluates the extension cone primitive
// (DESIGN.md 4.8); launched only for scenes that contain a cone.
#define FT_NS ft_strict_cone
#define FT_WITH_CONE 1
#include "kernels_impl.cuh"
