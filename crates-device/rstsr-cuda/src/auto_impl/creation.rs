//! Creation: counterpart of rstsr-core/src/feature_rayon/auto_impl/creation.rs:5-119.
use crate::prelude_dev::*;
use num::Num;

impl<T> DeviceCreationAnyAPI<T> for DeviceCudaAutoImpl
where
    T: CudaDType,
{
    unsafe fn empty_impl(&self, len: usize) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>> {
        Ok(Storage::new(CudaRaw::<T>::alloc(self, len)?.into(), self.clone()))
    }

    fn full_impl(&self, len: usize, fill: T) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>>
    where
        T: Clone,
    {
        let raw = CudaRaw::<T>::alloc(self, len)?;
        let l = cl(&[len].c());
        check(unsafe { ffi::rc_fill(self.raw(), T::CODE, raw.ptr, &l, T::CODE, &fill as *const T as *const c_void) })?;
        Ok(Storage::new(raw.into(), self.clone()))
    }

    fn outof_cpu_vec(&self, vec: Vec<T>) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>> {
        self.from_cpu_vec(&vec)
    }

    fn from_cpu_vec(&self, vec: &[T]) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>>
    where
        T: Clone,
    {
        let raw = CudaRaw::<T>::alloc(self, vec.len())?;
        check(unsafe { ffi::rc_memcpy_h2d(self.raw(), raw.ptr, vec.as_ptr() as *const c_void, raw.nbytes()) })?;
        self.synchronize()?; // pageable source: the copy must have left `vec` before the caller may drop it
        Ok(Storage::new(raw.into(), self.clone()))
    }

    fn uninit_impl(&self, len: usize) -> Result<Storage<DataOwned<CudaRaw<MaybeUninit<T>>>, MaybeUninit<T>, Self>> {
        Ok(Storage::new(CudaRaw::<MaybeUninit<T>>::alloc(self, len)?.into(), self.clone()))
    }

    unsafe fn assume_init_impl(
        storage: Storage<DataOwned<CudaRaw<MaybeUninit<T>>>, MaybeUninit<T>, Self>,
    ) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>> {
        let (data, device) = storage.into_raw_parts();
        Ok(Storage::new(data.into_raw().assume_init().into(), device))
    }
}

impl<T> DeviceCreationNumAPI<T> for DeviceCudaAutoImpl
where
    T: CudaDType + Num,
{
    fn zeros_impl(&self, len: usize) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>> {
        let raw = CudaRaw::<T>::alloc(self, len)?;
        check(unsafe { ffi::rc_memset(self.raw(), raw.ptr, 0, raw.nbytes()) })?;
        Ok(Storage::new(raw.into(), self.clone()))
    }

    fn ones_impl(&self, len: usize) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>> {
        self.full_impl(len, T::one())
    }
}

impl<T> DeviceCreationArangeAPI<T> for DeviceCudaAutoImpl
where
    T: CudaDType + PartialOrd + Num,
{
    fn arange_impl(&self, start: T, end: T, step: T) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>> {
        rstsr_assert!(step != T::zero(), InvalidValue)?;
        let (mut ptr, mut n) = (core::ptr::null_mut(), 0i64);
        check(unsafe {
            ffi::rc_arange(
                self.raw(),
                T::CODE,
                &start as *const T as *const c_void,
                &end as *const T as *const c_void,
                &step as *const T as *const c_void,
                &mut ptr,
                &mut n,
            )
        })?;
        Ok(Storage::new(unsafe { CudaRaw::<T>::from_raw(ptr, n as usize, self) }.into(), self.clone()))
    }
}

#[duplicate_item(T; [f32]; [f64])]
impl DeviceCreationComplexFloatAPI<T> for DeviceCudaAutoImpl {
    fn linspace_impl(&self, start: T, end: T, n: usize, endpoint: bool) -> Result<Storage<DataOwned<CudaRaw<T>>, T, Self>> {
        let mut ptr = core::ptr::null_mut();
        check(unsafe {
            ffi::rc_linspace(
                self.raw(),
                T::CODE,
                &start as *const T as *const c_void,
                &end as *const T as *const c_void,
                n as i64,
                endpoint as c_int,
                &mut ptr,
            )
        })?;
        Ok(Storage::new(unsafe { CudaRaw::<T>::from_raw(ptr, n, self) }.into(), self.clone()))
    }
}

impl<T> DeviceCreationTriAPI<T> for DeviceCudaAutoImpl
where
    T: CudaDType + Num,
{
    fn tril_impl<D>(&self, raw: &mut CudaRaw<T>, layout: &Layout<D>, k: isize) -> Result<()>
    where
        D: DimAPI,
    {
        check(unsafe { ffi::rc_tril(self.raw(), T::CODE, raw.ptr, &cl(layout), k as i64) })
    }

    fn triu_impl<D>(&self, raw: &mut CudaRaw<T>, layout: &Layout<D>, k: isize) -> Result<()>
    where
        D: DimAPI,
    {
        check(unsafe { ffi::rc_triu(self.raw(), T::CODE, raw.ptr, &cl(layout), k as i64) })
    }
}
