//! `DeviceChangeAPI` (rstsr-core/src/storage/conversion.rs:3-21) for DeviceCuda, after the pattern of
//! `crates-device/rstsr-openblas/src/conversion.rs:3-48`.  CPU <-> CPU there is a zero-copy re-tag; CPU <-> CUDA and
//! CUDA <-> CUDA are real copies, so `change_device` always returns owned storage.
use crate::prelude_dev::*;

/// host `Vec<T>` storage (any CPU device) -> DeviceCuda: one H2D copy of the whole raw buffer; the layout is kept.
macro_rules! impl_cpu_to_cuda {
    ($DevCpu: ty) => {
        impl<'a, R, T, D> DeviceChangeAPI<'a, DeviceCuda, R, T, D> for $DevCpu
        where
            T: CudaDType + 'a,
            D: DimAPI,
            R: DataCloneAPI<Data = Vec<T>>,
        {
            type Repr = DataOwned<CudaRaw<T>>;
            type ReprTo = DataOwned<CudaRaw<T>>;

            fn change_device(tensor: TensorAny<R, T, $DevCpu, D>, device: &DeviceCuda) -> Result<TensorAny<Self::Repr, T, DeviceCuda, D>> {
                let (storage, layout) = tensor.into_raw_parts();
                let host: &Vec<T> = storage.raw();
                let raw = CudaRaw::<T>::alloc(device, host.len())?;
                check(unsafe { ffi::rc_memcpy_h2d(device.raw(), raw.ptr, host.as_ptr() as *const c_void, raw.nbytes()) })?;
                device.synchronize()?; // `host` may be freed as soon as we return
                Ok(TensorAny::new(Storage::new(raw.into(), device.clone()), layout))
            }

            fn into_device(tensor: TensorAny<R, T, $DevCpu, D>, device: &DeviceCuda) -> Result<TensorAny<DataOwned<CudaRaw<T>>, T, DeviceCuda, D>> {
                DeviceChangeAPI::change_device(tensor, device)
            }

            fn to_device(tensor: &'a TensorAny<R, T, $DevCpu, D>, device: &DeviceCuda) -> Result<TensorAny<Self::ReprTo, T, DeviceCuda, D>> {
                DeviceChangeAPI::change_device(tensor.view(), device)
            }
        }
    };
}

/// DeviceCuda -> host `Vec<T>` storage: one D2H copy (synchronises).
macro_rules! impl_cuda_to_cpu {
    ($DevCpu: ty) => {
        impl<'a, R, T, D> DeviceChangeAPI<'a, $DevCpu, R, T, D> for DeviceCuda
        where
            T: CudaDType + 'a,
            D: DimAPI,
            R: DataCloneAPI<Data = CudaRaw<T>>,
        {
            type Repr = DataOwned<Vec<T>>;
            type ReprTo = DataOwned<Vec<T>>;

            fn change_device(tensor: TensorAny<R, T, DeviceCuda, D>, device: &$DevCpu) -> Result<TensorAny<Self::Repr, T, $DevCpu, D>> {
                let (storage, layout) = tensor.into_raw_parts();
                let host = <DeviceCuda as DeviceStorageAPI<T>>::to_cpu_vec(&storage)?;
                Ok(TensorAny::new(Storage::new(host.into(), device.clone()), layout))
            }

            fn into_device(tensor: TensorAny<R, T, DeviceCuda, D>, device: &$DevCpu) -> Result<TensorAny<DataOwned<Vec<T>>, T, $DevCpu, D>> {
                DeviceChangeAPI::change_device(tensor, device)
            }

            fn to_device(tensor: &'a TensorAny<R, T, DeviceCuda, D>, device: &$DevCpu) -> Result<TensorAny<Self::ReprTo, T, $DevCpu, D>> {
                DeviceChangeAPI::change_device(tensor.view(), device)
            }
        }
    };
}

impl_cpu_to_cuda!(DeviceCpuSerial);
impl_cuda_to_cpu!(DeviceCpuSerial);
#[cfg(feature = "rayon")]
impl_cpu_to_cuda!(DeviceCpuRayon);
#[cfg(feature = "rayon")]
impl_cuda_to_cpu!(DeviceCpuRayon);
#[cfg(feature = "faer")]
impl_cpu_to_cuda!(DeviceFaer);
#[cfg(feature = "faer")]
impl_cuda_to_cpu!(DeviceFaer);

/// DeviceCuda -> DeviceCuda: another stream of the same GPU, or another GPU (NVLink peer copy); stream-ordered on both.
impl<'a, R, T, D> DeviceChangeAPI<'a, DeviceCuda, R, T, D> for DeviceCuda
where
    T: CudaDType + 'a,
    D: DimAPI,
    R: DataCloneAPI<Data = CudaRaw<T>>,
{
    type Repr = DataOwned<CudaRaw<T>>;
    type ReprTo = DataOwned<CudaRaw<T>>;

    fn change_device(tensor: TensorAny<R, T, DeviceCuda, D>, device: &DeviceCuda) -> Result<TensorAny<Self::Repr, T, DeviceCuda, D>> {
        let (storage, layout) = tensor.into_raw_parts();
        let src: &CudaRaw<T> = storage.raw();
        let dst = CudaRaw::<T>::alloc(device, src.len())?;
        check(unsafe { ffi::rc_memcpy_peer(device.raw(), dst.ptr, src.dev.raw(), src.ptr, src.nbytes()) })?;
        Ok(TensorAny::new(Storage::new(dst.into(), device.clone()), layout))
    }

    fn into_device(tensor: TensorAny<R, T, DeviceCuda, D>, device: &DeviceCuda) -> Result<TensorAny<DataOwned<CudaRaw<T>>, T, DeviceCuda, D>> {
        DeviceChangeAPI::change_device(tensor, device)
    }

    fn to_device(tensor: &'a TensorAny<R, T, DeviceCuda, D>, device: &DeviceCuda) -> Result<TensorAny<Self::ReprTo, T, DeviceCuda, D>> {
        DeviceChangeAPI::change_device(tensor.view(), device)
    }
}

#[cfg(test)]
mod test {
    use super::*;

    #[test]
    fn test_device_conversion_cpu_serial() {
        // mirrors crates-device/rstsr-openblas/src/conversion.rs:55-66
        let device_serial = DeviceCpuSerial::default();
        let device = DeviceCuda::default();
        let a = linspace((1.0, 5.0, 5, &device));
        let b = a.to_device(&device_serial);
        println!("{b:?}");
        let a = linspace((1.0, 5.0, 5, &device_serial));
        let a_view = a.view();
        let b = a_view.to_device(&device);
        println!("{b:?}");
    }
}
