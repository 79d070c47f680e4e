// the same body with a column-major default order (broadcasting aligns at the left, F-contiguous outputs)
mod core_func; // symlink -> ../../../rstsr-core/tests/core_func
mod test_utils; // symlink -> ../../../rstsr-core/tests/test_utils

pub use rstsr::prelude::*;
pub use std::sync::LazyLock;
pub use test_utils::TestCfg;

pub use rstsr_cuda::DeviceCuda as DeviceType;

pub static TESTCFG: LazyLock<TestCfg<DeviceType>> = LazyLock::new(|| {
    let mut device = DeviceType::default();
    device.set_default_order(ColMajor);
    TestCfg::init(device, vec!["test_matmul", "test_matrix_transpose", "test_map", "test_iter"], None)
});
