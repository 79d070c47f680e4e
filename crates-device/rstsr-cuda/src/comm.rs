//! Multi-GPU: one `DeviceCuda` (process or thread) per GPU; shards are independent except for reductions whose
//! sharded axis is reduced (the one exchange step of this path).  Results up to 256 KiB are combined by ONE kernel
//! over the NVLink peer window (fold of the local partial states + stores into every rank's window + rank-ordered
//! fold: bitwise identical on every rank and run to run); larger ones by ncclAllReduce.
use crate::prelude_dev::*;

pub struct Comm {
    ptr: *mut ffi::rc_comm,
    dev: DeviceCuda,
}
unsafe impl Send for Comm {}

impl Comm {
    /// 128 bytes to broadcast from rank 0 to every rank (ncclUniqueId).
    pub fn unique_id() -> Result<[u8; ffi::RC_COMM_ID_BYTES]> {
        let mut id = [0u8; ffi::RC_COMM_ID_BYTES];
        check(unsafe { ffi::rc_comm_get_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }

    /// Collective over all ranks.
    pub fn new(dev: &DeviceCuda, nranks: usize, rank: usize, id: &[u8; ffi::RC_COMM_ID_BYTES]) -> Result<Self> {
        let mut ptr = core::ptr::null_mut();
        check(unsafe { ffi::rc_comm_init_rank(dev.raw(), nranks as c_int, rank as c_int, id.as_ptr(), &mut ptr) })?;
        Ok(Self { ptr, dev: dev.clone() })
    }

    /// (nranks, rank, peer window in use)
    pub fn info(&self) -> Result<(usize, usize, bool)> {
        let (mut n, mut r, mut p) = (0, 0, 0);
        check(unsafe { ffi::rc_comm_info(self.ptr, &mut n, &mut r, &mut p) })?;
        Ok((n as usize, r as usize, p != 0))
    }

    /// `*_all` of a tensor split over the ranks: `la` is this rank's shard, `n_global` the global element count
    /// (mean = combined sum / n_global).  `op`: RC_SUM / RC_PROD / RC_MAX / RC_MIN / RC_MEAN.
    pub fn reduce_all_sharded<T: CudaDType, D: DimAPI>(&self, op: c_int, a: &CudaRaw<T>, la: &Layout<D>, n_global: usize) -> Result<T> {
        let mut out = MaybeUninit::<T>::uninit();
        check(unsafe {
            ffi::rc_reduce_all_sharded(self.dev.raw(), self.ptr, op, T::CODE, a.ptr, &cl(la), n_global as i64, out.as_mut_ptr() as *mut c_void)
        })?;
        Ok(unsafe { out.assume_init() })
    }

    /// `*_axes` where the sharded axis is reduced: every rank ends with the complete (out, lo).
    #[allow(clippy::too_many_arguments)]
    pub fn reduce_axes_sharded<T: CudaDType, D: DimAPI>(
        &self,
        op: c_int,
        a: &CudaRaw<T>,
        la: &Layout<D>,
        axes: &[isize],
        n_reduced_global: usize,
        out: &mut CudaRaw<T>,
        lo: &Layout<IxD>,
    ) -> Result<()> {
        let ax: Vec<i64> = axes.iter().map(|&x| x as i64).collect();
        check(unsafe {
            ffi::rc_reduce_axes_sharded(
                self.dev.raw(), self.ptr, op, T::CODE, a.ptr, &cl(la), ax.as_ptr(), ax.len() as c_int, n_reduced_global as i64, out.ptr, &cl(lo),
            )
        })
    }

    /// In-place all-reduce of a dense device buffer.
    pub fn all_reduce<T: CudaDType>(&self, op: c_int, buf: &mut CudaRaw<T>) -> Result<()> {
        check(unsafe { ffi::rc_comm_all_reduce(self.ptr, op, T::CODE, buf.ptr, buf.len()) })
    }
}

impl Drop for Comm {
    fn drop(&mut self) {
        unsafe { ffi::rc_comm_destroy(self.ptr) };
    }
}
