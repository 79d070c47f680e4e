// rc_reduce_ext_b16.cu -- reductions, vecdot and allclose of b16 (body: rc_reduce_extx_body.cuh)
#define RC_EXTX_KIND 1
#include "rc_reduce_extx_body.cuh"
