pub(crate) use crate::codes::*;
pub(crate) use crate::device::{check, cl, layout_from_c};
pub(crate) use crate::dtype::CudaDType;
pub(crate) use crate::ffi;
pub(crate) use crate::raw::CudaRaw;
pub(crate) use crate::DeviceCuda;
pub(crate) use crate::DeviceCudaAutoImpl;
pub(crate) use core::ffi::{c_int, c_void};
pub use rstsr_core::prelude_dev::*;
