//! `c = a op b`: counterpart of rstsr-core/src/feature_rayon/auto_impl/op_ternary_arithmetic.rs:3-56.
//! The reference bounds `TA: Op<TB, Output = TC>`; the primitive impls of `core::ops` pair a type with itself, so the
//! device implements the traits for `TA = TB = TC = T` (integer arithmetic wraps, as in the reference's release CI).
use crate::prelude_dev::*;

#[duplicate_item(
     OpAPI         Op       CODE       ;
    [OpAddAPI   ] [Add   ] [RC_ADD   ];
    [OpSubAPI   ] [Sub   ] [RC_SUB   ];
    [OpMulAPI   ] [Mul   ] [RC_MUL   ];
    [OpDivAPI   ] [Div   ] [RC_DIV   ];
    [OpRemAPI   ] [Rem   ] [RC_REM   ];
    [OpBitOrAPI ] [BitOr ] [RC_BITOR ];
    [OpBitAndAPI] [BitAnd] [RC_BITAND];
    [OpBitXorAPI] [BitXor] [RC_BITXOR];
    [OpShlAPI   ] [Shl   ] [RC_SHL   ];
    [OpShrAPI   ] [Shr   ] [RC_SHR   ];
)]
impl<T, D> OpAPI<T, T, T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Op<T, Output = T>,
    D: DimAPI,
{
    fn op_mutc_refa_refb(
        &self,
        c: &mut CudaRaw<MaybeUninit<T>>,
        lc: &Layout<D>,
        a: &CudaRaw<T>,
        la: &Layout<D>,
        b: &CudaRaw<T>,
        lb: &Layout<D>,
    ) -> Result<()> {
        check(unsafe { ffi::rc_op_mutc_refa_refb(self.raw(), CODE, T::CODE, c.ptr, &cl(lc), a.ptr, &cl(la), b.ptr, &cl(lb)) })
    }

    fn op_mutc_refa_numb(&self, c: &mut CudaRaw<MaybeUninit<T>>, lc: &Layout<D>, a: &CudaRaw<T>, la: &Layout<D>, b: T) -> Result<()> {
        let b_host = &b as *const T as *const c_void;
        check(unsafe { ffi::rc_op_mutc_refa_numb(self.raw(), CODE, T::CODE, c.ptr, &cl(lc), a.ptr, &cl(la), b_host) })
    }

    fn op_mutc_numa_refb(&self, c: &mut CudaRaw<MaybeUninit<T>>, lc: &Layout<D>, a: T, b: &CudaRaw<T>, lb: &Layout<D>) -> Result<()> {
        let a_host = &a as *const T as *const c_void;
        check(unsafe { ffi::rc_op_mutc_numa_refb(self.raw(), CODE, T::CODE, c.ptr, &cl(lc), a_host, b.ptr, &cl(lb)) })
    }
}
