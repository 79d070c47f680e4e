//! `pack_tri` / `unpack_tri`: counterpart of rstsr-core/src/feature_rayon/auto_impl/op_tri.rs
//! (traits: rstsr-core/src/operators/ops/op_tri.rs:3-30).
use crate::prelude_dev::*;

fn uplo_code(uplo: FlagUpLo) -> c_int {
    match uplo {
        FlagUpLo::U => RC_UPLO_U,
        FlagUpLo::L => RC_UPLO_L,
    }
}

fn symm_code(symm: FlagSymm) -> c_int {
    match symm {
        FlagSymm::Sy => RC_SYMM_SY,
        FlagSymm::He => RC_SYMM_HE,
        FlagSymm::Ay => RC_SYMM_AY,
        FlagSymm::Ah => RC_SYMM_AH,
        FlagSymm::N => RC_SYMM_N,
    }
}

impl<T> OpPackTriAPI<T> for DeviceCudaAutoImpl
where
    T: CudaDType,
{
    fn pack_tri(&self, a: &mut CudaRaw<MaybeUninit<T>>, la: &Layout<IxD>, b: &CudaRaw<T>, lb: &Layout<IxD>, uplo: FlagUpLo) -> Result<()> {
        check(unsafe { ffi::rc_pack_tri(self.raw(), T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb), uplo_code(uplo)) })
    }
}

#[duplicate_item(T; [f32]; [f64])]
impl OpUnpackTriAPI<T> for DeviceCudaAutoImpl {
    fn unpack_tri(
        &self,
        a: &mut CudaRaw<MaybeUninit<T>>,
        la: &Layout<IxD>,
        b: &CudaRaw<T>,
        lb: &Layout<IxD>,
        uplo: FlagUpLo,
        symm: FlagSymm,
    ) -> Result<()> {
        check(unsafe { ffi::rc_unpack_tri(self.raw(), T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb), uplo_code(uplo), symm_code(symm)) })
    }
}
