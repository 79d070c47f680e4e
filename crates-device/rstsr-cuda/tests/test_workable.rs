// cf. crates-device/rstsr-openblas/tests/test_workable.rs: the smallest end-to-end use of the device.
use rstsr::prelude::*;
use rstsr_cuda::DeviceCuda;

#[test]
fn test_broadcast_add_and_axis_sum() {
    let device = DeviceCuda::default();
    let a = rt::arange((12.0, &device)).into_shape([3, 4]);
    let b = rt::arange((4.0, &device));
    let c = &a + &b; // (3, 4) + (4,): one fused kernel
    let s = c.sum_axes(0);
    assert_eq!(s.to_vec(), vec![12.0, 18.0, 24.0, 30.0]);
    assert_eq!(c.sum_all(), 84.0);
}

#[test]
fn test_permuted_to_contig_round_trip() {
    let device = DeviceCuda::default();
    let a = rt::arange((24.0, &device)).into_shape([2, 3, 4]);
    let t = a.transpose([2, 0, 1]).to_contig(RowMajor); // strided copy through the shared-memory tile kernel
    let host = t.to_device(&DeviceCpuSerial::default());
    assert_eq!(host.shape(), &[4, 2, 3]);
    assert_eq!(host[[1, 0, 2]], 9.0);
}
