//! `a op= b`, consuming forms, neg / not: counterpart of
//! rstsr-core/src/feature_rayon/auto_impl/op_binary_arithmetic.rs:4-113.
use crate::prelude_dev::*;

// a = a op b  (Op*AssignAPI :4-32)
#[duplicate_item(
     OpAPI               Op             CODE       ;
    [OpAddAssignAPI   ] [AddAssign   ] [RC_ADD   ];
    [OpSubAssignAPI   ] [SubAssign   ] [RC_SUB   ];
    [OpMulAssignAPI   ] [MulAssign   ] [RC_MUL   ];
    [OpDivAssignAPI   ] [DivAssign   ] [RC_DIV   ];
    [OpRemAssignAPI   ] [RemAssign   ] [RC_REM   ];
    [OpBitOrAssignAPI ] [BitOrAssign ] [RC_BITOR ];
    [OpBitAndAssignAPI] [BitAndAssign] [RC_BITAND];
    [OpBitXorAssignAPI] [BitXorAssign] [RC_BITXOR];
    [OpShlAssignAPI   ] [ShlAssign   ] [RC_SHL   ];
    [OpShrAssignAPI   ] [ShrAssign   ] [RC_SHR   ];
)]
impl<T, D> OpAPI<T, T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Op<T>,
    D: DimAPI,
{
    fn op_muta_refb(&self, a: &mut CudaRaw<T>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_op_muta_refb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb), 0) })
    }

    fn op_muta_numb(&self, a: &mut CudaRaw<T>, la: &Layout<D>, b: T) -> Result<()> {
        check(unsafe { ffi::rc_op_muta_numb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), &b as *const T as *const c_void, 0) })
    }
}

// a = a op b with `a` the consumed (owned) left operand (OpLConsume*API :34-62)
#[duplicate_item(
     OpAPI                 Op       CODE       ;
    [OpLConsumeAddAPI   ] [Add   ] [RC_ADD   ];
    [OpLConsumeSubAPI   ] [Sub   ] [RC_SUB   ];
    [OpLConsumeMulAPI   ] [Mul   ] [RC_MUL   ];
    [OpLConsumeDivAPI   ] [Div   ] [RC_DIV   ];
    [OpLConsumeRemAPI   ] [Rem   ] [RC_REM   ];
    [OpLConsumeBitOrAPI ] [BitOr ] [RC_BITOR ];
    [OpLConsumeBitAndAPI] [BitAnd] [RC_BITAND];
    [OpLConsumeBitXorAPI] [BitXor] [RC_BITXOR];
    [OpLConsumeShlAPI   ] [Shl   ] [RC_SHL   ];
    [OpLConsumeShrAPI   ] [Shr   ] [RC_SHR   ];
)]
impl<T, D> OpAPI<T, T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Op<T, Output = T>,
    D: DimAPI,
{
    fn op_muta_refb(&self, a: &mut CudaRaw<T>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_op_muta_refb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb), 0) })
    }

    fn op_muta_numb(&self, a: &mut CudaRaw<T>, la: &Layout<D>, b: T) -> Result<()> {
        check(unsafe { ffi::rc_op_muta_numb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), &b as *const T as *const c_void, 0) })
    }
}

// b = a op b with `b` the consumed (owned) RIGHT operand (OpRConsume*API :64-92): `reverse` = 1 on the C side
#[duplicate_item(
     OpAPI                 Op       CODE       ;
    [OpRConsumeAddAPI   ] [Add   ] [RC_ADD   ];
    [OpRConsumeSubAPI   ] [Sub   ] [RC_SUB   ];
    [OpRConsumeMulAPI   ] [Mul   ] [RC_MUL   ];
    [OpRConsumeDivAPI   ] [Div   ] [RC_DIV   ];
    [OpRConsumeRemAPI   ] [Rem   ] [RC_REM   ];
    [OpRConsumeBitOrAPI ] [BitOr ] [RC_BITOR ];
    [OpRConsumeBitAndAPI] [BitAnd] [RC_BITAND];
    [OpRConsumeBitXorAPI] [BitXor] [RC_BITXOR];
    [OpRConsumeShlAPI   ] [Shl   ] [RC_SHL   ];
    [OpRConsumeShrAPI   ] [Shr   ] [RC_SHR   ];
)]
impl<T, D> OpAPI<T, T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Op<T, Output = T>,
    D: DimAPI,
{
    fn op_muta_refb(&self, b: &mut CudaRaw<T>, lb: &Layout<D>, a: &CudaRaw<T>, la: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_op_muta_refb(self.raw(), CODE, T::CODE, b.ptr, &cl(lb), a.ptr, &cl(la), 1) })
    }

    fn op_muta_numb(&self, b: &mut CudaRaw<T>, lb: &Layout<D>, a: T) -> Result<()> {
        check(unsafe { ffi::rc_op_muta_numb(self.raw(), CODE, T::CODE, b.ptr, &cl(lb), &a as *const T as *const c_void, 1) })
    }
}

// a = op b / a = op a  (OpNegAPI, OpNotAPI :94-113)
#[duplicate_item(
     OpAPI      Op    CODE    ;
    [OpNegAPI] [Neg] [RC_NEG];
    [OpNotAPI] [Not] [RC_NOT];
)]
impl<T, D> OpAPI<T, T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Op<Output = T>,
    D: DimAPI,
{
    fn op_muta_refb(&self, a: &mut CudaRaw<MaybeUninit<T>>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta_refb(self.raw(), CODE, T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb)) })
    }

    fn op_muta(&self, a: &mut CudaRaw<T>, la: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_unary_muta(self.raw(), CODE, T::CODE, a.ptr, &cl(la)) })
    }
}
