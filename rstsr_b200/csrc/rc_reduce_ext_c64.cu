// rc_reduce_ext_c64.cu -- reductions, vecdot and allclose of c64 (body: rc_reduce_extx_body.cuh)
#define RC_EXTX_KIND 3
#include "rc_reduce_extx_body.cuh"
