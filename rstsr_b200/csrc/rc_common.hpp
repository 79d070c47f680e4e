// rc_common.hpp -- shared host-side definitions of librstsr_cuda.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rstsr_cuda.h"

namespace rc {

// ---- error transport: C++ exception inside, rc_status + thread-local message at the boundary ----
struct Error : std::runtime_error {
    rc_status code;
    Error(rc_status c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

void set_last_error(const std::string &msg);

[[noreturn]] inline void raise(rc_status code, const std::string &msg) { throw Error(code, msg); }

#define RC_CHECK(cond, code, msg)                     \
    do {                                              \
        if (!(cond)) ::rc::raise((code), (msg));      \
    } while (0)

// Wraps the body of an extern "C" entry point.
template <class F>
inline int guard(F &&f) noexcept {
    try {
        f();
        return RC_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc &) {
        set_last_error("host allocation failed");
        return RC_ERR_MEMORY;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return RC_ERR_RUNTIME;
    } catch (...) {
        set_last_error("unknown error");
        return RC_ERR_RUNTIME;
    }
}

// ---- Layout (host mirror of rc_layout with std::vector storage) ----
struct Layout {
    std::vector<int64_t> shape;
    std::vector<int64_t> stride;
    int64_t offset = 0;

    int ndim() const { return (int)shape.size(); }
    int64_t size() const {
        int64_t s = 1;
        for (auto d : shape) s *= d;
        return s;
    }
};

Layout from_c(const rc_layout *l);
void to_c(const Layout &l, rc_layout *out);

inline size_t dtype_size(rc_dtype t) {
    switch (t) {
        case RC_BOOL: case RC_I8: case RC_U8: return 1;
        case RC_I16: case RC_U16: return 2;
        case RC_I32: case RC_U32: case RC_F32: return 4;
        case RC_I64: case RC_U64: case RC_F64: case RC_C32: return 8;
        case RC_F16: case RC_BF16: return 2;
        case RC_C64: return 16;
    }
    raise(RC_ERR_INVALID_VALUE, "unknown dtype");
}
inline bool dtype_is_float(rc_dtype t) { return t == RC_F32 || t == RC_F64; }
inline bool dtype_is_half(rc_dtype t) { return t == RC_F16 || t == RC_BF16; }
inline bool dtype_is_complex(rc_dtype t) { return t == RC_C32 || t == RC_C64; }
inline bool dtype_is_extended(rc_dtype t) { return dtype_is_half(t) || dtype_is_complex(t); }
inline bool dtype_is_signed_int(rc_dtype t) { return t == RC_I8 || t == RC_I16 || t == RC_I32 || t == RC_I64; }
inline bool dtype_is_unsigned_int(rc_dtype t) { return t == RC_U8 || t == RC_U16 || t == RC_U32 || t == RC_U64; }
inline bool dtype_is_int(rc_dtype t) { return dtype_is_signed_int(t) || dtype_is_unsigned_int(t); }
const char *dtype_name(rc_dtype t);

}  // namespace rc
