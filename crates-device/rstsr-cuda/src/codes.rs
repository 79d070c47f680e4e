//! Values of the C enumerations of include/rstsr_cuda.h (they cross the boundary as `c_int`).
use core::ffi::c_int;

// rc_dtype
pub const RC_BOOL: c_int = 0;
pub const RC_I8: c_int = 1;
pub const RC_I16: c_int = 2;
pub const RC_I32: c_int = 3;
pub const RC_I64: c_int = 4;
pub const RC_U8: c_int = 5;
pub const RC_U16: c_int = 6;
pub const RC_U32: c_int = 7;
pub const RC_U64: c_int = 8;
pub const RC_F32: c_int = 9;
pub const RC_F64: c_int = 10;
pub const RC_F16: c_int = 11;
pub const RC_BF16: c_int = 12;
pub const RC_C32: c_int = 13;
pub const RC_C64: c_int = 14;

// rc_order / rc_iter_order
pub const RC_ROW_MAJOR: c_int = 0;
pub const RC_COL_MAJOR: c_int = 1;

// rc_binop
pub const RC_ADD: c_int = 0;
pub const RC_SUB: c_int = 1;
pub const RC_MUL: c_int = 2;
pub const RC_DIV: c_int = 3;
pub const RC_REM: c_int = 4;
pub const RC_BITOR: c_int = 5;
pub const RC_BITAND: c_int = 6;
pub const RC_BITXOR: c_int = 7;
pub const RC_SHL: c_int = 8;
pub const RC_SHR: c_int = 9;
pub const RC_MAXIMUM: c_int = 10;
pub const RC_MINIMUM: c_int = 11;
pub const RC_FLOOR_DIVIDE: c_int = 12;
pub const RC_POW: c_int = 13;
pub const RC_ATAN2: c_int = 14;
pub const RC_COPYSIGN: c_int = 15;
pub const RC_HYPOT: c_int = 16;
pub const RC_LOGADDEXP: c_int = 17;
pub const RC_NEXTAFTER: c_int = 18;
pub const RC_EQ: c_int = 32;
pub const RC_NE: c_int = 33;
pub const RC_LT: c_int = 34;
pub const RC_LE: c_int = 35;
pub const RC_GT: c_int = 36;
pub const RC_GE: c_int = 37;

// rc_unop
pub const RC_NEG: c_int = 0;
pub const RC_NOT: c_int = 1;
pub const RC_ABS: c_int = 2;
pub const RC_SQUARE: c_int = 3;
pub const RC_SIGN: c_int = 4;
pub const RC_SQRT: c_int = 5;
pub const RC_EXP: c_int = 6;
pub const RC_EXPM1: c_int = 7;
pub const RC_LOG: c_int = 8;
pub const RC_LOG2: c_int = 9;
pub const RC_LOG10: c_int = 10;
pub const RC_SIN: c_int = 11;
pub const RC_COS: c_int = 12;
pub const RC_TAN: c_int = 13;
pub const RC_ASIN: c_int = 14;
pub const RC_ACOS: c_int = 15;
pub const RC_ATAN: c_int = 16;
pub const RC_SINH: c_int = 17;
pub const RC_COSH: c_int = 18;
pub const RC_TANH: c_int = 19;
pub const RC_ASINH: c_int = 20;
pub const RC_ACOSH: c_int = 21;
pub const RC_ATANH: c_int = 22;
pub const RC_FLOOR: c_int = 23;
pub const RC_CEIL: c_int = 24;
pub const RC_ROUND: c_int = 25;
pub const RC_TRUNC: c_int = 26;
pub const RC_RECIPROCAL: c_int = 27;
pub const RC_CONJ: c_int = 28;
pub const RC_REAL: c_int = 29;
pub const RC_IMAG: c_int = 30;
pub const RC_ISNAN: c_int = 48;
pub const RC_ISINF: c_int = 49;
pub const RC_ISFINITE: c_int = 50;
pub const RC_SIGNBIT: c_int = 51;

// rc_redop
pub const RC_SUM: c_int = 0;
pub const RC_PROD: c_int = 1;
pub const RC_MAX: c_int = 2;
pub const RC_MIN: c_int = 3;
pub const RC_MEAN: c_int = 4;
pub const RC_VAR: c_int = 5;
pub const RC_STD: c_int = 6;
pub const RC_L2_NORM: c_int = 7;
pub const RC_ARGMIN: c_int = 8;
pub const RC_ARGMAX: c_int = 9;
pub const RC_ALL: c_int = 10;
pub const RC_ANY: c_int = 11;
pub const RC_COUNT_NONZERO: c_int = 12;

// rc_uplo / rc_symm
pub const RC_UPLO_U: c_int = 0;
pub const RC_UPLO_L: c_int = 1;
pub const RC_SYMM_SY: c_int = 0;
pub const RC_SYMM_HE: c_int = 1;
pub const RC_SYMM_AY: c_int = 2;
pub const RC_SYMM_AH: c_int = 3;
pub const RC_SYMM_N: c_int = 4;
