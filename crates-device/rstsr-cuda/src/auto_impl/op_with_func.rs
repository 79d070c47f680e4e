//! The closure traits `Op_Mut*_API<.., F>` (rstsr-core/src/operators/ops/op_with_func.rs:5-107; CPU impls in
//! feature_rayon/auto_impl/op_with_func.rs:5-117) take a Rust closure per element.  A closure cannot run on the
//! device, so `DeviceCuda` does NOT implement them: `map*` (tensor/map_elementwise.rs) and the element iterators do not
//! compile against it, by design -- bring the tensor to a CPU device first (`to_device(&DeviceFaer::default())`).
//!
//! What callers usually want from `map` is covered by the named ops of this device (the unary math functions, the
//! binary functions with promotion, `isclose`); anything else is a CPU job or a new kernel.
