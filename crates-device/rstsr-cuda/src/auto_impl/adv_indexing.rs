//! `index_select`: counterpart of rstsr-core/src/feature_rayon/auto_impl/adv_indexing.rs.
use crate::prelude_dev::*;

impl<T, D> DeviceIndexSelectAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType,
    D: DimAPI + DimSmallerOneAPI,
    D::SmallerOne: DimAPI,
{
    /// `indices` stay a host slice (the trait takes `&[usize]`); the library uploads them.
    fn index_select(
        &self,
        c: &mut CudaRaw<MaybeUninit<T>>,
        lc: &Layout<D>,
        a: &CudaRaw<T>,
        la: &Layout<D>,
        axis: usize,
        indices: &[usize],
    ) -> Result<()> {
        let idx: Vec<i64> = indices.iter().map(|&i| i as i64).collect();
        check(unsafe {
            ffi::rc_index_select(self.raw(), T::CODE, c.ptr, &cl(lc), a.ptr, &cl(la), axis as c_int, idx.as_ptr(), idx.len() as i64)
        })
    }
}
