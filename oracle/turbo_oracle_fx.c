/* placeholder: fixed-point model is written together with the CUDA kernel */
#include "turbo_oracle.h"
int tdo_fx_decode(const float *llr_in, const int *pi, const tdo_fx_params *p,
                  int *bits_out, int *le_out, int *overflow)
{ (void)llr_in; (void)pi; (void)p; (void)bits_out; (void)le_out; (void)overflow; return -1; }
