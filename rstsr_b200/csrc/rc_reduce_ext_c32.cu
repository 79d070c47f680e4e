// rc_reduce_ext_c32.cu -- reductions, vecdot and allclose of c32 (body: rc_reduce_extx_body.cuh)
#define RC_EXTX_KIND 2
#include "rc_reduce_extx_body.cuh"
