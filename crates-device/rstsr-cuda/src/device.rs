//! `DeviceCuda`: {ordinal, default order, stream} behind a shared handle -- the counterpart of
//! `crates-device/rstsr-openblas/src/device.rs:5-133` (and of `DeviceFaer`, rstsr-core/src/device_faer/device.rs:5-140).
use crate::prelude_dev::*;
use core::ffi::CStr;
use std::sync::Arc;

/// Owner of the C handle; destroyed (stream synchronised, workspaces freed) when the last clone goes away.
pub(crate) struct Handle {
    ptr: *mut ffi::rc_device,
}
unsafe impl Send for Handle {}
unsafe impl Sync for Handle {} // every entry point of the library is re-entrant on one handle

impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { ffi::rc_device_destroy(self.ptr) };
    }
}

/// A CUDA device.  Clones share the handle (and therefore the stream and the reduction workspace).
///
/// `default_order` is kept on the Rust side, like `DeviceCpuRayon::default_order` (feature_rayon/device.rs:41-51):
/// `set_default_order` on one clone must not change another clone, so ops that depend on the order pass it
/// explicitly (`rc_assign_arbitary_order`) instead of mutating the shared C handle.
#[derive(Clone)]
pub struct DeviceCuda {
    handle: Arc<Handle>,
    ordinal: i32,
    default_order: FlagOrder,
}

impl core::fmt::Debug for DeviceCuda {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "DeviceCuda {{ ordinal: {}, default_order: {:?} }}", self.ordinal, self.default_order)
    }
}

/// `rc_status` -> `RSTSRError` (rstsr-common/src/error.rs:12-49).
pub fn check(status: c_int) -> Result<()> {
    if status == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::rc_last_error()) }.to_string_lossy().into_owned();
    match status {
        1 => rstsr_raise!(ValueOutOfRange, "{msg}"),
        2 => rstsr_raise!(InvalidValue, "{msg}"),
        3 => rstsr_raise!(InvalidLayout, "{msg}"),
        4 => rstsr_raise!(RuntimeError, "{msg}"),
        5 => rstsr_raise!(DeviceMismatch, "{msg}"),
        6 => rstsr_raise!(UnImplemented, "{msg}"),
        7 => rstsr_raise!(MemoryError, "{msg}"),
        9 => rstsr_raise!(IndexError, "{msg}"),
        _ => rstsr_raise!(DeviceError, "{msg}"),
    }
}

/// `Layout<D>` -> `rc_layout` (element strides and offset; at most 16 axes).
pub fn cl<D: DimAPI>(l: &Layout<D>) -> ffi::rc_layout {
    let (shape, stride) = (l.shape(), l.stride());
    let nd = shape.as_ref().len();
    assert!(nd <= ffi::RC_MAX_NDIM, "DeviceCuda supports at most {} axes", ffi::RC_MAX_NDIM);
    let mut out = ffi::rc_layout { ndim: nd as i32, shape: [0; 16], stride: [0; 16], offset: l.offset() as i64 };
    for i in 0..nd {
        out.shape[i] = shape.as_ref()[i] as i64;
        out.stride[i] = stride.as_ref()[i] as i64;
    }
    out
}

/// `rc_layout` -> `Layout<IxD>` (outputs whose layout the callee chose: `*_axes` reductions).
pub fn layout_from_c(l: &ffi::rc_layout) -> Layout<IxD> {
    let nd = l.ndim as usize;
    let shape: Vec<usize> = l.shape[..nd].iter().map(|&d| d as usize).collect();
    let stride: Vec<isize> = l.stride[..nd].iter().map(|&s| s as isize).collect();
    unsafe { Layout::new_unchecked(shape, stride, l.offset as usize) }
}

pub(crate) fn order_code(order: FlagOrder) -> c_int {
    match order {
        FlagOrder::C => RC_ROW_MAJOR,
        FlagOrder::F => RC_COL_MAJOR,
    }
}

impl DeviceCuda {
    /// GPU `ordinal` with a stream of its own (cf. `DeviceBLAS::new(num_threads)`).
    pub fn new(ordinal: usize) -> Result<Self> {
        let mut ptr = core::ptr::null_mut();
        check(unsafe { ffi::rc_device_create(ordinal as c_int, RC_ROW_MAJOR, &mut ptr) })?;
        Ok(Self { handle: Arc::new(Handle { ptr }), ordinal: ordinal as i32, default_order: FlagOrder::default() })
    }

    /// Ops are enqueued on a stream the caller owns (`cudaStream_t`).
    ///
    /// # Safety
    /// `stream` must be a valid stream of GPU `ordinal` that outlives the device.
    pub unsafe fn with_stream(ordinal: usize, stream: *mut c_void) -> Result<Self> {
        let mut ptr = core::ptr::null_mut();
        check(ffi::rc_device_create_on_stream(ordinal as c_int, RC_ROW_MAJOR, stream, &mut ptr))?;
        Ok(Self { handle: Arc::new(Handle { ptr }), ordinal: ordinal as i32, default_order: FlagOrder::default() })
    }

    pub fn device_count() -> Result<usize> {
        let mut n: c_int = 0;
        check(unsafe { ffi::rc_device_count(&mut n) })?;
        Ok(n as usize)
    }

    #[inline]
    pub fn ordinal(&self) -> i32 {
        self.ordinal
    }

    #[inline]
    pub(crate) fn raw(&self) -> *mut ffi::rc_device {
        self.handle.ptr
    }

    #[inline]
    pub(crate) fn order(&self) -> c_int {
        order_code(self.default_order)
    }

    /// Block until everything enqueued on this device's stream has finished.
    pub fn synchronize(&self) -> Result<()> {
        check(unsafe { ffi::rc_device_synchronize(self.raw()) })
    }

    /// Work enqueued on `self` from now on runs after what `other` has enqueued so far (same GPU).
    pub fn wait(&self, other: &DeviceCuda) -> Result<()> {
        check(unsafe { ffi::rc_device_wait(self.raw(), other.raw()) })
    }

    /// NUMA node of the GPU's PCIe root (for `rc_host_alloc_on_node` staging buffers); `None` if unknown.
    pub fn numa_node(&self) -> Option<usize> {
        let mut node: c_int = -1;
        let st = unsafe { ffi::rc_device_numa_node(self.raw(), &mut node) };
        (st == 0 && node >= 0).then_some(node as usize)
    }
}

impl Default for DeviceCuda {
    /// GPU 0.  Panics without a CUDA device: there is no CPU fallback.
    fn default() -> Self {
        DeviceCuda::new(0).expect("DeviceCuda::default(): no usable CUDA device (there is no CPU fallback)")
    }
}

impl DeviceBaseAPI for DeviceCuda {
    /// Same GPU, same stream (= same handle) and same default order: two devices with different streams are not
    /// ordered against each other, so tensors must move between them through `to_device` (cf. DeviceFaer:
    /// equal pool size & order, device_faer/device.rs:46-51).
    fn same_device(&self, other: &Self) -> bool {
        let mut same: c_int = 0;
        let ok = unsafe { ffi::rc_device_same_device(self.raw(), other.raw(), &mut same) } == 0;
        ok && same != 0 && self.default_order == other.default_order
    }

    fn default_order(&self) -> FlagOrder {
        self.default_order
    }

    fn set_default_order(&mut self, order: FlagOrder) {
        self.default_order = order;
    }
}

impl<T> DeviceRawAPI<T> for DeviceCuda {
    type Raw = CudaRaw<T>;
}

impl<T> DeviceStorageAPI<T> for DeviceCuda {
    fn len<R>(storage: &Storage<R, T, Self>) -> usize
    where
        R: DataAPI<Data = Self::Raw>,
    {
        storage.raw().len()
    }

    fn to_cpu_vec<R>(storage: &Storage<R, T, Self>) -> Result<Vec<T>>
    where
        Self::Raw: Clone,
        R: DataAPI<Data = Self::Raw>,
    {
        let raw = storage.raw();
        let mut out: Vec<T> = Vec::with_capacity(raw.len());
        // rc_memcpy_d2h synchronises the stream: the vector is complete on return
        check(unsafe { ffi::rc_memcpy_d2h(raw.dev.raw(), out.as_mut_ptr() as *mut c_void, raw.ptr, raw.nbytes()) })?;
        unsafe { out.set_len(raw.len()) };
        Ok(out)
    }

    fn into_cpu_vec<R>(storage: Storage<R, T, Self>) -> Result<Vec<T>>
    where
        Self::Raw: Clone,
        R: DataCloneAPI<Data = Self::Raw>,
    {
        Self::to_cpu_vec(&storage)
    }

    #[inline]
    fn get_index<R>(storage: &Storage<R, T, Self>, index: usize) -> T
    where
        T: Clone,
        R: DataAPI<Data = Self::Raw>,
    {
        let raw = storage.raw();
        assert!(index < raw.len(), "index {index} out of bounds for storage of {} elements", raw.len());
        let mut out = MaybeUninit::<T>::uninit();
        let src = unsafe { (raw.ptr as *const u8).add(index * core::mem::size_of::<T>()) } as *const c_void;
        check(unsafe { ffi::rc_memcpy_d2h(raw.dev.raw(), out.as_mut_ptr() as *mut c_void, src, core::mem::size_of::<T>()) })
            .expect("device read failed in get_index");
        unsafe { out.assume_init() }
    }

    /// Not honourable on a GPU device: there is no host pointer into device memory.
    fn get_index_ptr<R>(_storage: &Storage<R, T, Self>, _index: usize) -> *const T
    where
        R: DataAPI<Data = Self::Raw>,
    {
        panic!("DeviceCuda: get_index_ptr is not available (device memory has no host address); use get_index / to_cpu_vec")
    }

    fn get_index_mut_ptr<R>(_storage: &mut Storage<R, T, Self>, _index: usize) -> *mut T
    where
        R: DataMutAPI<Data = Self::Raw>,
    {
        panic!("DeviceCuda: get_index_mut_ptr is not available (device memory has no host address); use set_index")
    }

    #[inline]
    fn set_index<R>(storage: &mut Storage<R, T, Self>, index: usize, value: T)
    where
        R: DataMutAPI<Data = Self::Raw>,
    {
        let raw = storage.raw_mut();
        assert!(index < raw.len(), "index {index} out of bounds for storage of {} elements", raw.len());
        let dst = unsafe { (raw.ptr as *mut u8).add(index * core::mem::size_of::<T>()) } as *mut c_void;
        check(unsafe { ffi::rc_memcpy_h2d(raw.dev.raw(), dst, &value as *const T as *const c_void, core::mem::size_of::<T>()) })
            .and_then(|_| raw.dev.synchronize()) // `value` is dropped when we return
            .expect("device write failed in set_index");
    }
}

impl<T> DeviceAPI<T> for DeviceCuda {}

// umbrella bounds (rstsr-core/src/operators/combined_trait.rs:6-55), as device_faer/device.rs:127-140
impl<T, D> DeviceNumAPI<T, D> for DeviceCuda
where
    T: CudaDType + num::Num,
    D: DimAPI,
{
}

impl<D> DeviceComplexFloatAPI<f32, D> for DeviceCuda where D: DimAPI {}
impl<D> DeviceComplexFloatAPI<f64, D> for DeviceCuda where D: DimAPI {}
// Complex<f32> / Complex<f64>: the library covers storage, copies, casts, + - * /, neg, conj / abs / real / imag / square,
// the listed math functions and sum / prod / mean (include/rstsr_cuda.h, RC_C32 / RC_C64); the umbrella bound asks for the
// full unary set (asin .. atanh, floor-like ops are real-only), so it is NOT claimed for complex types yet.
