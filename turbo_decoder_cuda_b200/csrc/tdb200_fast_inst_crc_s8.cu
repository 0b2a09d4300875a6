// tdb200_fast_inst_crc_s8.cu -- the packed-int16 decoder kernels with the CRC stopping rule (early_term 2 / 3)
// for TDB200_LLR_S8 channel LLRs: separate instantiations, so the default kernels stay free of that path.
#include "tdb200_fast_kernel.cuh"

namespace tdb200 {
typedef void (*fast_kernel_fn)(FastArgs);
fast_kernel_fn fast_pick_crc_s8(const FastGeom &g) { return pick_kernel_t<TDB200_LLR_S8, true>(g); }
}  // namespace tdb200
