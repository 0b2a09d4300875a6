// tdb200_fast_inst_f64.cu -- the packed-int16 decoder kernels for TDB200_LLR_F64 channel LLRs
// (all geometry variants; device code in tdb200_fast_kernel.cuh).
#include "tdb200_fast_kernel.cuh"

namespace tdb200 {
typedef void (*fast_kernel_fn)(FastArgs);
fast_kernel_fn fast_pick_f64(const FastGeom &g) { return pick_kernel_t<TDB200_LLR_F64, false>(g); }
}  // namespace tdb200
