// tdb200_fast_inst_crc_f16.cu -- the packed-int16 decoder kernels with the CRC stopping rule (early_term 2 / 3)
// for TDB200_LLR_F16 channel LLRs: separate instantiations, so the default kernels stay free of that path.
#include "tdb200_fast_kernel.cuh"

namespace tdb200 {
typedef void (*fast_kernel_fn)(FastArgs);
fast_kernel_fn fast_pick_crc_f16(const FastGeom &g) { return pick_kernel_t<TDB200_LLR_F16, true>(g); }
}  // namespace tdb200
