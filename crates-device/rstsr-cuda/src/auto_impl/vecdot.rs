//! `vecdot`: counterpart of rstsr-core/src/feature_rayon/auto_impl/vecdot.rs (trait: operators/linalg.rs:3-21);
//! `c` is caller-allocated (tensor/linalg/vecdot.rs:236-241).  One pass over both inputs, fixed-order accumulation.
use crate::prelude_dev::*;

impl<T, DA, DB, DC> DeviceVecdotAPI<T, T, T, DA, DB, DC> for DeviceCudaAutoImpl
where
    T: CudaDType + Mul<Output = T> + num::Zero,
    DA: DimAPI,
    DB: DimAPI,
    DC: DimAPI,
{
    fn vecdot(
        &self,
        c: &mut CudaRaw<MaybeUninit<T>>,
        lc: &Layout<DC>,
        a: &CudaRaw<T>,
        la: &Layout<DA>,
        b: &CudaRaw<T>,
        lb: &Layout<DB>,
        axes_a: &[isize],
        axes_b: &[isize],
    ) -> Result<()> {
        rstsr_assert_eq!(axes_a.len(), axes_b.len(), InvalidValue, "axes_a and axes_b should have the same length")?;
        let xa: Vec<i64> = axes_a.iter().map(|&x| x as i64).collect();
        let xb: Vec<i64> = axes_b.iter().map(|&x| x as i64).collect();
        check(unsafe {
            ffi::rc_vecdot(self.raw(), T::CODE, c.ptr, &cl(lc), a.ptr, &cl(la), b.ptr, &cl(lb), xa.as_ptr(), xb.as_ptr(), xa.len() as c_int)
        })
    }
}
