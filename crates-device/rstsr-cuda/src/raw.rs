//! `CudaRaw<T>`: the `Raw` of `DeviceRawAPI<T>` -- what `Vec<T>` is to the CPU devices
//! (rstsr-core/src/device_faer/device.rs:62-64).  Drop frees (stream-ordered), Clone is a device-to-device copy.
use crate::device::{check, DeviceCuda};
use crate::ffi;
use core::ffi::c_void;
use core::marker::PhantomData;
use core::mem::MaybeUninit;

pub struct CudaRaw<T> {
    pub(crate) ptr: *mut c_void,
    pub(crate) len: usize,
    pub(crate) dev: DeviceCuda,
    pub(crate) _t: PhantomData<T>,
}

// device memory is only touched through stream-ordered calls on `dev`
unsafe impl<T: Send> Send for CudaRaw<T> {}
unsafe impl<T: Sync> Sync for CudaRaw<T> {}

impl<T> CudaRaw<T> {
    /// `len` elements of uninitialised device memory (rc_malloc; the pool keeps freed blocks cached).
    pub fn alloc(dev: &DeviceCuda, len: usize) -> rstsr_core::prelude_dev::Result<Self> {
        let mut ptr = core::ptr::null_mut();
        check(unsafe { ffi::rc_malloc(dev.raw(), len * core::mem::size_of::<T>(), &mut ptr) })?;
        Ok(Self { ptr, len, dev: dev.clone(), _t: PhantomData })
    }

    /// Adopt memory the library allocated for us (`rc_reduce_axes`, `rc_arange`, ...).
    ///
    /// # Safety
    /// `ptr` must come from this device's allocator and hold `len` elements of `T`.
    pub unsafe fn from_raw(ptr: *mut c_void, len: usize, dev: &DeviceCuda) -> Self {
        Self { ptr, len, dev: dev.clone(), _t: PhantomData }
    }

    pub fn len(&self) -> usize {
        self.len
    }

    pub fn is_empty(&self) -> bool {
        self.len == 0
    }

    pub fn as_ptr(&self) -> *const c_void {
        self.ptr
    }

    pub fn as_mut_ptr(&mut self) -> *mut c_void {
        self.ptr
    }

    pub fn device(&self) -> &DeviceCuda {
        &self.dev
    }

    pub(crate) fn nbytes(&self) -> usize {
        self.len * core::mem::size_of::<T>()
    }
}

impl<T> CudaRaw<MaybeUninit<T>> {
    /// `Vec<MaybeUninit<T>>` -> `Vec<T>` of the CPU devices (auto_impl/creation.rs:41-53): same buffer, new type.
    ///
    /// # Safety
    /// Every element must have been written.
    pub unsafe fn assume_init(self) -> CudaRaw<T> {
        let me = core::mem::ManuallyDrop::new(self);
        CudaRaw { ptr: me.ptr, len: me.len, dev: me.dev.clone(), _t: PhantomData }
    }
}

impl<T> Clone for CudaRaw<T> {
    fn clone(&self) -> Self {
        let out = CudaRaw::<T>::alloc(&self.dev, self.len).expect("device allocation failed in CudaRaw::clone");
        check(unsafe { ffi::rc_memcpy_d2d(self.dev.raw(), out.ptr, self.ptr, self.nbytes()) })
            .expect("device copy failed in CudaRaw::clone");
        out
    }
}

impl<T> Drop for CudaRaw<T> {
    fn drop(&mut self) {
        if !self.ptr.is_null() {
            unsafe { ffi::rc_free(self.dev.raw(), self.ptr) };
        }
    }
}

impl<T> core::fmt::Debug for CudaRaw<T> {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "CudaRaw<{}>[{}] @ {:?} on cuda:{}", core::any::type_name::<T>(), self.len, self.ptr, self.dev.ordinal())
    }
}
