//! Unary functions `a = f(b)` / `a = f(a)`: counterpart of
//! rstsr-core/src/feature_rayon/auto_impl/op_binary_common.rs:10-239.
//!
//! The reference computes the math functions on `b.into_float()` (`DTypeIntoFloatAPI`, integers -> f64) and writes
//! `T::FloatType`.  On the device a float input is ONE fused kernel; an integer input is cast to f64 first
//! (`rc_assign`, element-exact) and then takes the same kernel -- two launches, identical values.
use crate::prelude_dev::*;
use rstsr_dtype_traits::DTypeIntoFloatAPI;

/// `a = f(into_float(b))` for any element type.
fn unary_into_float<T, D>(
    dev: &DeviceCuda,
    code: c_int,
    a: &mut CudaRaw<MaybeUninit<<T as DTypeIntoFloatAPI>::FloatType>>,
    la: &Layout<D>,
    b: &CudaRaw<T>,
    lb: &Layout<D>,
) -> Result<()>
where
    T: CudaDType + DTypeIntoFloatAPI,
    T::FloatType: CudaDType,
    D: DimAPI,
{
    let tf = <T::FloatType as CudaDType>::CODE;
    if T::CODE == tf {
        return check(unsafe { ffi::rc_unary_muta_refb(dev.raw(), code, tf, a.ptr, &cl(la), b.ptr, &cl(lb)) });
    }
    // integer input: cast into the output buffer (same layout as the result), then apply f in place
    check(unsafe { ffi::rc_assign(dev.raw(), tf, a.ptr, &cl(la), T::CODE, b.ptr, &cl(lb)) })?;
    check(unsafe { ffi::rc_unary_muta(dev.raw(), code, tf, a.ptr, &cl(la)) })
}

#[duplicate_item(
     OpAPI             CODE           ;
    [OpAcosAPI      ] [RC_ACOS      ];
    [OpAcoshAPI     ] [RC_ACOSH     ];
    [OpAsinAPI      ] [RC_ASIN      ];
    [OpAsinhAPI     ] [RC_ASINH     ];
    [OpAtanAPI      ] [RC_ATAN      ];
    [OpAtanhAPI     ] [RC_ATANH     ];
    [OpCeilAPI      ] [RC_CEIL      ];
    [OpConjAPI      ] [RC_CONJ      ];
    [OpCosAPI       ] [RC_COS       ];
    [OpCoshAPI      ] [RC_COSH      ];
    [OpExpAPI       ] [RC_EXP       ];
    [OpExpm1API     ] [RC_EXPM1     ];
    [OpFloorAPI     ] [RC_FLOOR     ];
    [OpInvAPI       ] [RC_RECIPROCAL];
    [OpLogAPI       ] [RC_LOG       ];
    [OpLog2API      ] [RC_LOG2      ];
    [OpLog10API     ] [RC_LOG10     ];
    [OpReciprocalAPI] [RC_RECIPROCAL];
    [OpRoundAPI     ] [RC_ROUND     ];
    [OpSinAPI       ] [RC_SIN       ];
    [OpSinhAPI      ] [RC_SINH      ];
    [OpSqrtAPI      ] [RC_SQRT      ];
    [OpTanAPI       ] [RC_TAN       ];
    [OpTanhAPI      ] [RC_TANH      ];
    [OpTruncAPI     ] [RC_TRUNC     ];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + DTypeIntoFloatAPI,
    T::FloatType: CudaDType,
    D: DimAPI,
{
    type TOut = T::FloatType;

    fn op_muta_refb(&self, a: &mut CudaRaw<MaybeUninit<Self::TOut>>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        unary_into_float::<T, D>(self, CODE, a, la, b, lb)
    }

    fn op_muta(&self, a: &mut CudaRaw<MaybeUninit<Self::TOut>>, la: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta(self.raw(), CODE, <Self::TOut as CudaDType>::CODE, a.ptr, &cl(la)) })
    }
}

// same-type ops: square (b * b), sign (ext_sign)
#[duplicate_item(
     OpAPI         CODE       ;
    [OpSquareAPI] [RC_SQUARE];
    [OpSignAPI  ] [RC_SIGN  ];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + num::Num,
    D: DimAPI,
{
    type TOut = T;

    fn op_muta_refb(&self, a: &mut CudaRaw<MaybeUninit<T>>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta_refb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb)) })
    }

    fn op_muta(&self, a: &mut CudaRaw<MaybeUninit<T>>, la: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta(self.raw(), CODE, T::CODE, a.ptr, &cl(la)) })
    }
}

// real element types: abs keeps the type (ExtNum::AbsOut = T), real is the identity, imag is zero
#[duplicate_item(
     OpAPI       CODE     ;
    [OpAbsAPI ] [RC_ABS ];
    [OpRealAPI] [RC_REAL];
    [OpImagAPI] [RC_IMAG];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + num::Num,
    D: DimAPI,
{
    type TOut = T;

    fn op_muta_refb(&self, a: &mut CudaRaw<MaybeUninit<T>>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta_refb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb)) })
    }

    fn op_muta(&self, a: &mut CudaRaw<MaybeUninit<T>>, la: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta(self.raw(), CODE, T::CODE, a.ptr, &cl(la)) })
    }
}

// boolean output (:94-117); `op_muta` is unreachable in the reference as well (the output type differs)
#[duplicate_item(
     OpAPI           CODE          ;
    [OpSignBitAPI ] [RC_SIGNBIT ];
    [OpIsFiniteAPI] [RC_ISFINITE];
    [OpIsInfAPI   ] [RC_ISINF   ];
    [OpIsNanAPI   ] [RC_ISNAN   ];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + num::Float,
    D: DimAPI,
{
    type TOut = bool;

    fn op_muta_refb(&self, a: &mut CudaRaw<MaybeUninit<bool>>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta_refb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb)) })
    }

    fn op_muta(&self, _a: &mut CudaRaw<MaybeUninit<bool>>, _la: &Layout<D>) -> Result<()> {
        let type_b = core::any::type_name::<T>();
        unreachable!("{:?} is not supported in this function.", type_b);
    }
}
