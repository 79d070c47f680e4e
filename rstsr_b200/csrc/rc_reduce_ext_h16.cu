// rc_reduce_ext_h16.cu -- reductions, vecdot and allclose of h16 (body: rc_reduce_extx_body.cuh)
#define RC_EXTX_KIND 0
#include "rc_reduce_extx_body.cuh"
