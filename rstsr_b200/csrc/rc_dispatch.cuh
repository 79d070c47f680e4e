// rc_dispatch.cuh -- dtype switch helpers for the dispatch translation units.
#pragma once
#include "rc_elementwise.cuh"
#include "rc_functors.cuh"
#include "rc_ops.hpp"

namespace rc {

[[noreturn]] inline void unsupported(const char *what, rc_dtype t) {
    raise(RC_ERR_UNIMPLEMENTED, std::string(what) + " is not implemented for dtype " + dtype_name(t));
}

#define RC_CASE(DT, CT, FUNCTOR) \
    case DT: ew_launch<FUNCTOR<CT>>(dev, c, args); return;

#define RC_SWITCH_INT(FUNCTOR)          \
    RC_CASE(RC_I8, int8_t, FUNCTOR)     \
    RC_CASE(RC_I16, int16_t, FUNCTOR)   \
    RC_CASE(RC_I32, int32_t, FUNCTOR)   \
    RC_CASE(RC_I64, int64_t, FUNCTOR)   \
    RC_CASE(RC_U8, uint8_t, FUNCTOR)    \
    RC_CASE(RC_U16, uint16_t, FUNCTOR)  \
    RC_CASE(RC_U32, uint32_t, FUNCTOR)  \
    RC_CASE(RC_U64, uint64_t, FUNCTOR)

#define RC_SWITCH_SIGNED_INT(FUNCTOR)   \
    RC_CASE(RC_I8, int8_t, FUNCTOR)     \
    RC_CASE(RC_I16, int16_t, FUNCTOR)   \
    RC_CASE(RC_I32, int32_t, FUNCTOR)   \
    RC_CASE(RC_I64, int64_t, FUNCTOR)

#define RC_SWITCH_FLOAT(FUNCTOR)        \
    RC_CASE(RC_F32, float, FUNCTOR)     \
    RC_CASE(RC_F64, double, FUNCTOR)

#define RC_SWITCH_NUM(FUNCTOR) RC_SWITCH_INT(FUNCTOR) RC_SWITCH_FLOAT(FUNCTOR)

}  // namespace rc
