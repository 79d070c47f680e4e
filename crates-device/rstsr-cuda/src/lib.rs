#![allow(clippy::needless_return)]
#![allow(non_camel_case_types)]
#![doc = include_str!("../readme.md")]
//! File-for-file counterpart of `crates-device/rstsr-openblas/src/lib.rs:1-29` (minus BLAS / LAPACK).

pub mod auto_impl;
pub mod codes;
#[cfg(feature = "comm")]
pub mod comm;
pub mod conversion;
pub mod device;
pub mod dtype;
pub mod ffi;
pub mod prelude_dev;
pub mod raw;

pub use device::DeviceCuda;
pub use dtype::CudaDType;
pub use raw::CudaRaw;

// the auto_impl files are written against this alias, like the reference's `DeviceRayonAutoImpl`
pub(crate) use DeviceCuda as DeviceCudaAutoImpl;
