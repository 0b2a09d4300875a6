// tdb200_fast_inst_lm_sym2.cu -- the packed-int16 Log-MAP decoder kernels reading received QPSK symbols (float): the soft
// demapper is fused into the load stage (device code in tdb200_fast_kernel.cuh).
#include "tdb200_fast_kernel.cuh"

namespace tdb200 {
typedef void (*fast_kernel_fn)(FastArgs);
fast_kernel_fn fast_pick_lm_sym2(const FastGeom &g) { return pick_kernel_lm_t<kLlrSymQpskF32>(g); }
}  // namespace tdb200
