// cf. crates-device/rstsr-openblas/tests/tests_core_row.rs:1-14: the reference's device-agnostic test body
// (rstsr-core/tests/core_func, symlinked) re-run on DeviceCuda with a row-major default order.
mod core_func; // symlink -> ../../../rstsr-core/tests/core_func
mod test_utils; // symlink -> ../../../rstsr-core/tests/test_utils

pub use rstsr::prelude::*;
pub use std::sync::LazyLock;
pub use test_utils::TestCfg;

pub use rstsr_cuda::DeviceCuda as DeviceType;

pub static TESTCFG: LazyLock<TestCfg<DeviceType>> = LazyLock::new(|| {
    let mut device = DeviceType::default(); // GPU 0, its own stream
    device.set_default_order(RowMajor);
    // bodies that need a trait DeviceCuda cannot honour (closure ops, element iterators, matmul) are skipped at run time
    TestCfg::init(device, vec!["test_matmul", "test_matrix_transpose", "test_map", "test_iter"], None)
});
