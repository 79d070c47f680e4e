//! Element types a GPU device can hold: POD numerics (the reference is generic over any `T: Clone`).
use crate::codes::*;
use core::ffi::c_int;
use core::mem::MaybeUninit;

/// A type librstsr_cuda.so knows: `CODE` is its `rc_dtype`.
///
/// # Safety
/// Implementors must be plain-old-data with the size and bit layout the C side assumes for `CODE`.
pub unsafe trait CudaDType: Copy + Send + Sync + 'static {
    const CODE: c_int;
}

macro_rules! impl_dtype {
    ($($T:ty => $code:expr),* $(,)?) => { $( unsafe impl CudaDType for $T { const CODE: c_int = $code; } )* };
}

impl_dtype!(
    bool => RC_BOOL, i8 => RC_I8, i16 => RC_I16, i32 => RC_I32, i64 => RC_I64,
    u8 => RC_U8, u16 => RC_U16, u32 => RC_U32, u64 => RC_U64, f32 => RC_F32, f64 => RC_F64,
);

#[cfg(target_pointer_width = "64")]
impl_dtype!(isize => RC_I64, usize => RC_U64);

// half and complex element types (round 2): same bit layout as half::{f16, bf16} and num::Complex<f32 | f64>
impl_dtype!(num::complex::Complex<f32> => RC_C32, num::complex::Complex<f64> => RC_C64);
#[cfg(feature = "half")]
impl_dtype!(half::f16 => RC_F16, half::bf16 => RC_BF16);

/// `MaybeUninit<T>` storage is the same bytes as `T` storage (uninit_impl / assume_init_impl).
unsafe impl<T: CudaDType> CudaDType for MaybeUninit<T> {
    const CODE: c_int = T::CODE;
}
