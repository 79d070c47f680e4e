//! `extern "C"` block of librstsr_cuda.so -- GENERATED from include/rstsr_cuda.h by scripts/gen_rust_ffi.py
//! (do not edit; tests/test_cabi.py keeps it in lock-step with the header and with the ctypes stub rstsr_b200/_ffi.py).
//! Enumerations cross the boundary as `c_int`; their Rust-side values live in `crate::codes`.
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_int, c_void};

pub const RC_MAX_NDIM: usize = 16;
pub const RC_COMM_ID_BYTES: usize = 128;

/// `Layout<IxD>` as the C side sees it (rstsr-common/src/layout/layoutbase.rs:15-23): element strides, element offset.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rc_layout {
    pub ndim: i32,
    pub shape: [i64; RC_MAX_NDIM],
    pub stride: [i64; RC_MAX_NDIM],
    pub offset: i64,
}

/// Opaque device handle: {ordinal, default order, stream, workspaces}.
#[repr(C)]
pub struct rc_device {
    _private: [u8; 0],
}

/// Opaque communicator: NCCL comm + NVLink peer window.
#[repr(C)]
pub struct rc_comm {
    _private: [u8; 0],
}

#[link(name = "rstsr_cuda")]
extern "C" {
    pub fn rc_last_error() -> *const c_char;
    pub fn rc_version() -> *const c_char;
    pub fn rc_device_count(count: *mut c_int) -> c_int;
    pub fn rc_device_create(ordinal: c_int, default_order: c_int, out: *mut *mut rc_device) -> c_int;
    pub fn rc_device_create_on_stream(
        ordinal: c_int,
        default_order: c_int,
        cuda_stream: *mut c_void,
        out: *mut *mut rc_device,
    ) -> c_int;
    pub fn rc_device_destroy(dev: *mut rc_device) -> c_int;
    pub fn rc_device_default_order(dev: *const rc_device, out: *mut c_int) -> c_int;
    pub fn rc_device_set_default_order(dev: *mut rc_device, order: c_int) -> c_int;
    pub fn rc_device_same_device(a: *const rc_device, b: *const rc_device, same: *mut c_int) -> c_int;
    pub fn rc_device_numa_node(dev: *const rc_device, node: *mut c_int) -> c_int;
    pub fn rc_device_ordinal(dev: *const rc_device, ordinal: *mut c_int) -> c_int;
    pub fn rc_device_stream(dev: *const rc_device, cuda_stream: *mut *mut c_void) -> c_int;
    pub fn rc_device_synchronize(dev: *mut rc_device) -> c_int;
    pub fn rc_device_launch_count(dev: *const rc_device, count: *mut u64) -> c_int;
    pub fn rc_malloc(dev: *mut rc_device, nbytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn rc_free(dev: *mut rc_device, ptr: *mut c_void) -> c_int;
    pub fn rc_memcpy_h2d(dev: *mut rc_device, dst_dev: *mut c_void, src_host: *const c_void, nbytes: usize) -> c_int;
    pub fn rc_memcpy_d2h(dev: *mut rc_device, dst_host: *mut c_void, src_dev: *const c_void, nbytes: usize) -> c_int;
    pub fn rc_memcpy_d2d(dev: *mut rc_device, dst_dev: *mut c_void, src_dev: *const c_void, nbytes: usize) -> c_int;
    pub fn rc_memset(dev: *mut rc_device, dst_dev: *mut c_void, byte: c_int, nbytes: usize) -> c_int;
    pub fn rc_memcpy_h2d_async(
        dev: *mut rc_device,
        dst_dev: *mut c_void,
        src_host: *const c_void,
        nbytes: usize,
    ) -> c_int;
    pub fn rc_memcpy_d2h_async(
        dev: *mut rc_device,
        dst_host: *mut c_void,
        src_dev: *const c_void,
        nbytes: usize,
    ) -> c_int;
    pub fn rc_memcpy2d_d2h_async(
        dev: *mut rc_device,
        dst_host: *mut c_void,
        dst_pitch: usize,
        src_dev: *const c_void,
        src_pitch: usize,
        width_bytes: usize,
        height: usize,
    ) -> c_int;
    pub fn rc_memcpy2d_h2d_async(
        dev: *mut rc_device,
        dst_dev: *mut c_void,
        dst_pitch: usize,
        src_host: *const c_void,
        src_pitch: usize,
        width_bytes: usize,
        height: usize,
    ) -> c_int;
    pub fn rc_device_wait(waiter: *mut rc_device, signaler: *mut rc_device) -> c_int;
    pub fn rc_memcpy_peer(
        dst_dev: *mut rc_device,
        dst_ptr: *mut c_void,
        src_dev: *mut rc_device,
        src_ptr: *const c_void,
        nbytes: usize,
    ) -> c_int;
    pub fn rc_get_index(
        dev: *mut rc_device,
        dtype: c_int,
        a: *const c_void,
        index: i64,
        host_out: *mut c_void,
    ) -> c_int;
    pub fn rc_set_index(
        dev: *mut rc_device,
        dtype: c_int,
        a: *mut c_void,
        index: i64,
        host_value: *const c_void,
    ) -> c_int;
    pub fn rc_host_alloc(nbytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn rc_host_alloc_on_node(nbytes: usize, node: c_int, out: *mut *mut c_void, bound_out: *mut c_int) -> c_int;
    pub fn rc_host_free(ptr: *mut c_void) -> c_int;
    pub fn rc_dtype_size(dtype: c_int) -> usize;
    pub fn rc_layout_check(l: *const rc_layout) -> c_int;
    pub fn rc_layout_bounds_index(l: *const rc_layout, min_out: *mut i64, max_out: *mut i64) -> c_int;
    pub fn rc_layout_c_contig(l: *const rc_layout, out: *mut c_int) -> c_int;
    pub fn rc_layout_f_contig(l: *const rc_layout, out: *mut c_int) -> c_int;
    pub fn rc_layout_new_contig(
        shape: *const i64,
        ndim: c_int,
        order: c_int,
        offset: i64,
        out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_layout_broadcast(
        la: *const rc_layout,
        lb: *const rc_layout,
        order: c_int,
        la_out: *mut rc_layout,
        lb_out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_layout_for_binary_op(
        la: *const rc_layout,
        lb: *const rc_layout,
        order: c_int,
        lc_out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_layout_for_array_copy(
        la: *const rc_layout,
        iter_order: c_int,
        default_order: c_int,
        lc_out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_layout_for_reduce(
        la: *const rc_layout,
        axes: *const i64,
        naxes: c_int,
        lo_out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_layout_reshapeable(
        la: *const rc_layout,
        shape: *const i64,
        ndim: c_int,
        order: c_int,
        viewable: *mut c_int,
        out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_layout_equal(a: *const rc_layout, b: *const rc_layout, equal: *mut c_int) -> c_int;
    pub fn rc_assign(
        dev: *mut rc_device,
        tc: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        ta: c_int,
        a: *const c_void,
        la: *const rc_layout,
    ) -> c_int;
    pub fn rc_assign_arbitary(
        dev: *mut rc_device,
        tc: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        ta: c_int,
        a: *const c_void,
        la: *const rc_layout,
    ) -> c_int;
    pub fn rc_assign_arbitary_order(
        dev: *mut rc_device,
        order: c_int,
        tc: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        ta: c_int,
        a: *const c_void,
        la: *const rc_layout,
    ) -> c_int;
    pub fn rc_fill(
        dev: *mut rc_device,
        tc: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        tf: c_int,
        fill: *const c_void,
    ) -> c_int;
    pub fn rc_arange(
        dev: *mut rc_device,
        dtype: c_int,
        start: *const c_void,
        end: *const c_void,
        step: *const c_void,
        out_dev: *mut *mut c_void,
        n_out: *mut i64,
    ) -> c_int;
    pub fn rc_linspace(
        dev: *mut rc_device,
        dtype: c_int,
        start: *const c_void,
        end: *const c_void,
        n: i64,
        endpoint: c_int,
        out_dev: *mut *mut c_void,
    ) -> c_int;
    pub fn rc_tril(dev: *mut rc_device, dtype: c_int, a: *mut c_void, la: *const rc_layout, k: i64) -> c_int;
    pub fn rc_triu(dev: *mut rc_device, dtype: c_int, a: *mut c_void, la: *const rc_layout, k: i64) -> c_int;
    pub fn rc_op_mutc_refa_refb(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        a: *const c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
    ) -> c_int;
    pub fn rc_op_mutc_refa_numb(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        a: *const c_void,
        la: *const rc_layout,
        b_host_scalar: *const c_void,
    ) -> c_int;
    pub fn rc_op_mutc_numa_refb(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        a_host_scalar: *const c_void,
        b: *const c_void,
        lb: *const rc_layout,
    ) -> c_int;
    pub fn rc_op_muta_refb(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *mut c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
        reverse: c_int,
    ) -> c_int;
    pub fn rc_op_muta_numb(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *mut c_void,
        la: *const rc_layout,
        b_host_scalar: *const c_void,
        reverse: c_int,
    ) -> c_int;
    pub fn rc_unary_muta_refb(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *mut c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
    ) -> c_int;
    pub fn rc_unary_muta(dev: *mut rc_device, op: c_int, dtype: c_int, a: *mut c_void, la: *const rc_layout) -> c_int;
    pub fn rc_redop_out_dtype(op: c_int, dtype: c_int, out: *mut c_int) -> c_int;
    pub fn rc_binop_out_dtype(op: c_int, dtype: c_int, out: *mut c_int) -> c_int;
    pub fn rc_dtype_promote(ta: c_int, tb: c_int, out: *mut c_int) -> c_int;
    pub fn rc_binop_out_dtype_ex(op: c_int, ta: c_int, tb: c_int, out: *mut c_int) -> c_int;
    pub fn rc_op_mutc_refa_refb_ex(
        dev: *mut rc_device,
        op: c_int,
        tc: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        ta: c_int,
        a: *const c_void,
        la: *const rc_layout,
        tb: c_int,
        b: *const c_void,
        lb: *const rc_layout,
    ) -> c_int;
    pub fn rc_op_mutc_refa_numb_ex(
        dev: *mut rc_device,
        op: c_int,
        tc: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        ta: c_int,
        a: *const c_void,
        la: *const rc_layout,
        tb: c_int,
        b_host_scalar: *const c_void,
    ) -> c_int;
    pub fn rc_op_mutc_numa_refb_ex(
        dev: *mut rc_device,
        op: c_int,
        tc: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        ta: c_int,
        a_host_scalar: *const c_void,
        tb: c_int,
        b: *const c_void,
        lb: *const rc_layout,
    ) -> c_int;
    pub fn rc_isclose(
        dev: *mut rc_device,
        dtype: c_int,
        c_bool: *mut c_void,
        lc: *const rc_layout,
        a: *const c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
        rtol: f64,
        atol: f64,
        equal_nan: c_int,
    ) -> c_int;
    pub fn rc_unop_out_dtype(op: c_int, dtype: c_int, out: *mut c_int) -> c_int;
    pub fn rc_reduce_all(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        host_out: *mut c_void,
    ) -> c_int;
    pub fn rc_reduce_all_device(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        dev_out: *mut c_void,
    ) -> c_int;
    pub fn rc_reduce_axes(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        axes: *const i64,
        naxes: c_int,
        out_dev: *mut *mut c_void,
        lo_out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_reduce_axes_into(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        axes: *const i64,
        naxes: c_int,
        out_dev: *mut c_void,
        lo: *const rc_layout,
    ) -> c_int;
    pub fn rc_reduce_unraveled_arg_all(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        index_out: *mut i64,
    ) -> c_int;
    pub fn rc_reduce_unraveled_arg_axes(
        dev: *mut rc_device,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        axes: *const i64,
        naxes: c_int,
        out_dev: *mut *mut c_void,
        lo_out: *mut rc_layout,
    ) -> c_int;
    pub fn rc_vecdot(
        dev: *mut rc_device,
        dtype: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        a: *const c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
        axes_a: *const i64,
        axes_b: *const i64,
        naxes: c_int,
    ) -> c_int;
    pub fn rc_allclose_all(
        dev: *mut rc_device,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
        rtol: f64,
        atol: f64,
        equal_nan: c_int,
        result: *mut c_int,
    ) -> c_int;
    pub fn rc_index_select(
        dev: *mut rc_device,
        dtype: c_int,
        c: *mut c_void,
        lc: *const rc_layout,
        a: *const c_void,
        la: *const rc_layout,
        axis: c_int,
        indices: *const i64,
        n_indices: i64,
    ) -> c_int;
    pub fn rc_pack_tri(
        dev: *mut rc_device,
        dtype: c_int,
        a: *mut c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
        uplo: c_int,
    ) -> c_int;
    pub fn rc_unpack_tri(
        dev: *mut rc_device,
        dtype: c_int,
        a: *mut c_void,
        la: *const rc_layout,
        b: *const c_void,
        lb: *const rc_layout,
        uplo: c_int,
        symm: c_int,
    ) -> c_int;
    pub fn rc_comm_get_unique_id(id: *mut u8) -> c_int;
    pub fn rc_comm_init_rank(
        dev: *mut rc_device,
        nranks: c_int,
        rank: c_int,
        id: *const u8,
        out: *mut *mut rc_comm,
    ) -> c_int;
    pub fn rc_comm_destroy(comm: *mut rc_comm) -> c_int;
    pub fn rc_comm_info(comm: *const rc_comm, nranks: *mut c_int, rank: *mut c_int, peer_window: *mut c_int) -> c_int;
    pub fn rc_comm_set_peer_window(comm: *mut rc_comm, enable: c_int) -> c_int;
    pub fn rc_comm_all_reduce(
        comm: *mut rc_comm,
        op: c_int,
        dtype: c_int,
        buf_dev: *mut c_void,
        count: usize,
    ) -> c_int;
    pub fn rc_reduce_all_sharded(
        dev: *mut rc_device,
        comm: *mut rc_comm,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        n_global: i64,
        host_out: *mut c_void,
    ) -> c_int;
    pub fn rc_reduce_axes_sharded(
        dev: *mut rc_device,
        comm: *mut rc_comm,
        op: c_int,
        dtype: c_int,
        a: *const c_void,
        la: *const rc_layout,
        axes: *const i64,
        naxes: c_int,
        n_reduced_global: i64,
        out_dev: *mut c_void,
        lo: *const rc_layout,
    ) -> c_int;
}
