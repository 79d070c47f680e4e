//! Assignment: counterpart of rstsr-core/src/feature_rayon/auto_impl/assignment.rs:3-49.
use crate::prelude_dev::*;

impl<TC, TA, DC, DA> OpAssignArbitaryAPI<TC, DC, DA, TA> for DeviceCudaAutoImpl
where
    TC: CudaDType,
    TA: CudaDType + DTypeCastAPI<TC>,
    DC: DimAPI,
    DA: DimAPI,
{
    /// k-th element of `lc` <- k-th element of `la`, k counted in the device's default order
    /// (cpu_serial/assignment.rs:39-67).  The order goes along explicitly: clones share one C handle.
    fn assign_arbitary(&self, c: &mut CudaRaw<TC>, lc: &Layout<DC>, a: &CudaRaw<TA>, la: &Layout<DA>) -> Result<()> {
        check(unsafe {
            ffi::rc_assign_arbitary_order(self.raw(), self.order(), TC::CODE, c.ptr, &cl(lc), TA::CODE, a.ptr, &cl(la))
        })
    }

    fn assign_arbitary_uninit(
        &self,
        c: &mut CudaRaw<MaybeUninit<TC>>,
        lc: &Layout<DC>,
        a: &CudaRaw<TA>,
        la: &Layout<DA>,
    ) -> Result<()> {
        check(unsafe {
            ffi::rc_assign_arbitary_order(self.raw(), self.order(), TC::CODE, c.ptr, &cl(lc), TA::CODE, a.ptr, &cl(la))
        })
    }
}

impl<TC, TA, D> OpAssignAPI<TC, D, TA> for DeviceCudaAutoImpl
where
    TC: CudaDType,
    TA: CudaDType + DTypeCastAPI<TC>,
    D: DimAPI,
{
    fn assign(&self, c: &mut CudaRaw<TC>, lc: &Layout<D>, a: &CudaRaw<TA>, la: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_assign(self.raw(), TC::CODE, c.ptr, &cl(lc), TA::CODE, a.ptr, &cl(la)) })
    }

    fn assign_uninit(&self, c: &mut CudaRaw<MaybeUninit<TC>>, lc: &Layout<D>, a: &CudaRaw<TA>, la: &Layout<D>) -> Result<()> {
        check(unsafe { ffi::rc_assign(self.raw(), TC::CODE, c.ptr, &cl(lc), TA::CODE, a.ptr, &cl(la)) })
    }

    fn fill(&self, c: &mut CudaRaw<TC>, lc: &Layout<D>, fill: TA) -> Result<()> {
        check(unsafe { ffi::rc_fill(self.raw(), TC::CODE, c.ptr, &cl(lc), TA::CODE, &fill as *const TA as *const c_void) })
    }
}
