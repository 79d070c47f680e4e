//! Binary functions with promotion: counterpart of
//! rstsr-core/src/feature_rayon/auto_impl/op_ternary_common.rs:6-189.  `TA` and `TB` may differ: the library promotes
//! (`rc_op_mutc_*_ex`, the table of rstsr-dtype-traits/src/promotion.rs) and runs the single fused kernel when no
//! operand needs a cast.
use crate::prelude_dev::*;
use num::pow::Pow;
use rstsr_dtype_traits::{DTypeIntoFloatAPI, DTypePromoteAPI};

macro_rules! ternary_ex_body {
    ($CODE: expr, $TOut: ty) => {
        fn op_mutc_refa_refb(
            &self,
            c: &mut CudaRaw<MaybeUninit<$TOut>>,
            lc: &Layout<D>,
            a: &CudaRaw<TA>,
            la: &Layout<D>,
            b: &CudaRaw<TB>,
            lb: &Layout<D>,
        ) -> Result<()> {
            check(unsafe {
                ffi::rc_op_mutc_refa_refb_ex(
                    self.raw(), $CODE, <$TOut as CudaDType>::CODE, c.ptr, &cl(lc), TA::CODE, a.ptr, &cl(la), TB::CODE, b.ptr, &cl(lb),
                )
            })
        }

        fn op_mutc_refa_numb(&self, c: &mut CudaRaw<MaybeUninit<$TOut>>, lc: &Layout<D>, a: &CudaRaw<TA>, la: &Layout<D>, b: TB) -> Result<()> {
            let b_host = &b as *const TB as *const c_void;
            check(unsafe {
                ffi::rc_op_mutc_refa_numb_ex(self.raw(), $CODE, <$TOut as CudaDType>::CODE, c.ptr, &cl(lc), TA::CODE, a.ptr, &cl(la), TB::CODE, b_host)
            })
        }

        fn op_mutc_numa_refb(&self, c: &mut CudaRaw<MaybeUninit<$TOut>>, lc: &Layout<D>, a: TA, b: &CudaRaw<TB>, lb: &Layout<D>) -> Result<()> {
            let a_host = &a as *const TA as *const c_void;
            check(unsafe {
                ffi::rc_op_mutc_numa_refb_ex(self.raw(), $CODE, <$TOut as CudaDType>::CODE, c.ptr, &cl(lc), TA::CODE, a_host, TB::CODE, b.ptr, &cl(lb))
            })
        }
    };
}

// output with special promotion (:6-73): promote_pair, into_float, f
#[duplicate_item(
     OpAPI               CODE           ;
    [OpATan2API       ] [RC_ATAN2     ];
    [OpCopySignAPI    ] [RC_COPYSIGN  ];
    [OpHypotAPI       ] [RC_HYPOT     ];
    [OpNextAfterAPI   ] [RC_NEXTAFTER ];
    [OpLogAddExpAPI   ] [RC_LOGADDEXP ];
)]
impl<TA, TB, D> OpAPI<TA, TB, D> for DeviceCudaAutoImpl
where
    TA: CudaDType + DTypePromoteAPI<TB, Res: DTypeIntoFloatAPI<FloatType: CudaDType>>,
    TB: CudaDType,
    D: DimAPI,
{
    type TOut = <TA::Res as DTypeIntoFloatAPI>::FloatType;

    ternary_ex_body!(CODE, <TA::Res as DTypeIntoFloatAPI>::FloatType);
}

// general promotion (:75-139): TOut = Res
#[duplicate_item(
     OpAPI               CODE              ;
    [OpMaximumAPI     ] [RC_MAXIMUM      ];
    [OpMinimumAPI     ] [RC_MINIMUM      ];
    [OpFloorDivideAPI ] [RC_FLOOR_DIVIDE ];
)]
impl<TA, TB, D> OpAPI<TA, TB, D> for DeviceCudaAutoImpl
where
    TA: CudaDType + DTypePromoteAPI<TB, Res: CudaDType>,
    TB: CudaDType,
    D: DimAPI,
{
    type TOut = TA::Res;

    ternary_ex_body!(CODE, TA::Res);
}

// comparisons (:75-139): compared after promotion, TOut = bool
#[duplicate_item(
     OpAPI               CODE    ;
    [OpEqualAPI       ] [RC_EQ ];
    [OpNotEqualAPI    ] [RC_NE ];
    [OpGreaterAPI     ] [RC_GT ];
    [OpGreaterEqualAPI] [RC_GE ];
    [OpLessAPI        ] [RC_LT ];
    [OpLessEqualAPI   ] [RC_LE ];
)]
impl<TA, TB, D> OpAPI<TA, TB, D> for DeviceCudaAutoImpl
where
    TA: CudaDType + DTypePromoteAPI<TB, Res: CudaDType>,
    TB: CudaDType,
    D: DimAPI,
{
    type TOut = bool;

    ternary_ex_body!(CODE, bool);
}

// pow (:141-189): `TA: Pow<TB>`, TOut = TA::Output.  num implements Pow for float ^ same float (powf),
// float ^ i8/u8/i16/u16/i32 (powi) and integer ^ u8/u16/u32/usize (wrapping power) -- always with Output = TA;
// the library rejects any other pair with UnImplemented.
impl<TA, TB, D> OpPowAPI<TA, TB, D> for DeviceCudaAutoImpl
where
    TA: CudaDType + Pow<TB, Output = TA>,
    TB: CudaDType,
    D: DimAPI,
{
    type TOut = TA;

    ternary_ex_body!(RC_POW, TA);
}

// OpIsCloseAPI (operators/ops/op_ternary_common.rs:59-102), TE = f64; both operands of one type on the device
impl<T, D> OpIsCloseAPI<T, T, D, f64> for DeviceCudaAutoImpl
where
    T: CudaDType + DTypePromoteAPI<T>,
    D: DimAPI,
{
    fn op_mutc_refa_refb(
        &self,
        c: &mut CudaRaw<MaybeUninit<bool>>,
        lc: &Layout<D>,
        a: &CudaRaw<T>,
        la: &Layout<D>,
        b: &CudaRaw<T>,
        lb: &Layout<D>,
        isclose_args: &IsCloseArgs<f64>,
    ) -> Result<()> {
        let IsCloseArgs { rtol, atol, equal_nan } = isclose_args;
        check(unsafe {
            ffi::rc_isclose(self.raw(), T::CODE, c.ptr, &cl(lc), a.ptr, &cl(la), b.ptr, &cl(lb), *rtol, *atol, *equal_nan as c_int)
        })
    }

    fn op_mutc_refa_numb(
        &self,
        c: &mut CudaRaw<MaybeUninit<bool>>,
        lc: &Layout<D>,
        a: &CudaRaw<T>,
        la: &Layout<D>,
        b: T,
        isclose_args: &IsCloseArgs<f64>,
    ) -> Result<()> {
        // the scalar becomes a one-element device buffer broadcast over `la`'s shape (stride 0 everywhere)
        let sb = DeviceCreationAnyAPI::<T>::from_cpu_vec(self, &[b])?;
        let lb = unsafe { Layout::new_unchecked(la.shape().clone(), la.new_stride(), 0) };
        self.op_mutc_refa_refb(c, lc, a, la, sb.raw(), &lb, isclose_args)
    }

    fn op_mutc_numa_refb(
        &self,
        c: &mut CudaRaw<MaybeUninit<bool>>,
        lc: &Layout<D>,
        a: T,
        b: &CudaRaw<T>,
        lb: &Layout<D>,
        isclose_args: &IsCloseArgs<f64>,
    ) -> Result<()> {
        let sa = DeviceCreationAnyAPI::<T>::from_cpu_vec(self, &[a])?;
        let la = unsafe { Layout::new_unchecked(lb.shape().clone(), lb.new_stride(), 0) };
        self.op_mutc_refa_refb(c, lc, sa.raw(), &la, b, lb, isclose_args)
    }
}
