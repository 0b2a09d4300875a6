// tdb200_fast_inst_s8.cu -- the packed-int16 decoder kernels for TDB200_LLR_S8 channel LLRs
// (all geometry variants; device code in tdb200_fast_kernel.cuh).
#include "tdb200_fast_kernel.cuh"

namespace tdb200 {
typedef void (*fast_kernel_fn)(FastArgs);
fast_kernel_fn fast_pick_s8(const FastGeom &g) { return pick_kernel_t<TDB200_LLR_S8, false>(g); }
}  // namespace tdb200
