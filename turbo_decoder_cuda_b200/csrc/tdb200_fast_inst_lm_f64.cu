// tdb200_fast_inst_lm_f64.cu -- the packed-int16 Log-MAP decoder kernels (TDB200_ALGO_LOGMAP_S16) for
// TDB200_LLR_F64 channel LLRs (device code in tdb200_fast_kernel.cuh).
#include "tdb200_fast_kernel.cuh"

namespace tdb200 {
typedef void (*fast_kernel_fn)(FastArgs);
fast_kernel_fn fast_pick_lm_f64(const FastGeom &g) { return pick_kernel_lm_t<TDB200_LLR_F64>(g); }
}  // namespace tdb200
