//! One file per file of `rstsr-core/src/feature_rayon/auto_impl/` (which DeviceFaer and every BLAS device share
//! verbatim through symlinks, crates-device/rstsr-openblas/src/rayon_auto_impl/): the trait method marshals its
//! arguments and makes exactly ONE call into librstsr_cuda.so.
pub mod adv_indexing;
pub mod assignment;
pub mod creation;
pub mod op_binary_arithmetic;
pub mod op_binary_common;
pub mod op_ternary_arithmetic;
pub mod op_ternary_common;
pub mod op_tri;
pub mod op_with_func;
pub mod reduction;
pub mod vecdot;
