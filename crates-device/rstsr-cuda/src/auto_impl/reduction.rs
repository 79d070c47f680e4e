//! Reductions: counterpart of rstsr-core/src/feature_rayon/auto_impl/reduction.rs:7-763.
//! `*_all` returns a host scalar (the kernel writes it into a mapped host slot; one stream sync, no D2H copy);
//! `*_axes`: the CALLEE allocates the output and chooses its layout (operators/reduction.rs:26-32).
//! Results are run-to-run deterministic (fixed-order two-pass grid reduction), which rayon's fold is not.
use crate::prelude_dev::*;
use num::{Float, FromPrimitive, One, Zero};
use rstsr_dtype_traits::ExtReal;

fn reduce_all<T: CudaDType, TO: CudaDType, D: DimAPI>(dev: &DeviceCuda, op: c_int, a: &CudaRaw<T>, la: &Layout<D>) -> Result<TO> {
    let mut out = MaybeUninit::<TO>::uninit();
    check(unsafe { ffi::rc_reduce_all(dev.raw(), op, T::CODE, a.ptr, &cl(la), out.as_mut_ptr() as *mut c_void) })?;
    Ok(unsafe { out.assume_init() })
}

#[allow(clippy::type_complexity)]
fn reduce_axes<T: CudaDType, TO: CudaDType, D: DimAPI>(
    dev: &DeviceCuda,
    op: c_int,
    a: &CudaRaw<T>,
    la: &Layout<D>,
    axes: &[isize],
) -> Result<(Storage<DataOwned<CudaRaw<TO>>, TO, DeviceCuda>, Layout<IxD>)> {
    let ax: Vec<i64> = axes.iter().map(|&x| x as i64).collect();
    let mut ptr = core::ptr::null_mut();
    let mut lo = cl(&[0usize].c());
    check(unsafe { ffi::rc_reduce_axes(dev.raw(), op, T::CODE, a.ptr, &cl(la), ax.as_ptr(), ax.len() as c_int, &mut ptr, &mut lo) })?;
    let layout = layout_from_c(&lo);
    let raw = unsafe { CudaRaw::<TO>::from_raw(ptr, layout.size().max(1), dev) };
    Ok((Storage::new(raw.into(), dev.clone()), layout))
}

// sum / prod / max / min: TOut = T (:7-165)
#[duplicate_item(
     OpAPI       func_all   func_axes   CODE      Bound                         ;
    [OpSumAPI ] [sum_all ] [sum_axes ] [RC_SUM ] [Zero + Add<Output = T>       ];
    [OpProdAPI] [prod_all] [prod_axes] [RC_PROD] [One + Mul<Output = T>        ];
    [OpMaxAPI ] [max_all ] [max_axes ] [RC_MAX ] [ExtReal                      ];
    [OpMinAPI ] [min_all ] [min_axes ] [RC_MIN ] [ExtReal                      ];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Bound,
    D: DimAPI,
{
    type TOut = T;

    /// max / min of a zero-size tensor: InvalidValue, as auto_impl/reduction.rs:51-53,95-97 (raised by the library).
    fn func_all(&self, a: &CudaRaw<T>, la: &Layout<D>) -> Result<T> {
        reduce_all::<T, T, D>(self, CODE, a, la)
    }

    fn func_axes(&self, a: &CudaRaw<T>, la: &Layout<D>, axes: &[isize]) -> Result<(Storage<DataOwned<CudaRaw<T>>, T, Self>, Layout<IxD>)> {
        reduce_axes::<T, T, D>(self, CODE, a, la, axes)
    }
}

// mean / var / std / l2_norm on real floats: TOut = T (= T::Real) (:167-354)
#[duplicate_item(
     OpAPI         func_all      func_axes      CODE         ;
    [OpMeanAPI  ] [mean_all   ] [mean_axes   ] [RC_MEAN   ];
    [OpVarAPI   ] [var_all    ] [var_axes    ] [RC_VAR    ];
    [OpStdAPI   ] [std_all    ] [std_axes    ] [RC_STD    ];
    [OpL2NormAPI] [l2_norm_all] [l2_norm_axes] [RC_L2_NORM];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Float + FromPrimitive,
    D: DimAPI,
{
    type TOut = T;

    fn func_all(&self, a: &CudaRaw<T>, la: &Layout<D>) -> Result<T> {
        reduce_all::<T, T, D>(self, CODE, a, la)
    }

    fn func_axes(&self, a: &CudaRaw<T>, la: &Layout<D>, axes: &[isize]) -> Result<(Storage<DataOwned<CudaRaw<T>>, T, Self>, Layout<IxD>)> {
        reduce_axes::<T, T, D>(self, CODE, a, la, axes)
    }
}

// argmin / argmax / count_nonzero: TOut = usize (u64 on the device) (:356-464, 536-570)
#[duplicate_item(
     OpAPI               func_all            func_axes            CODE               Bound              ;
    [OpArgMinAPI      ] [argmin_all       ] [argmin_axes       ] [RC_ARGMIN       ] [PartialOrd       ];
    [OpArgMaxAPI      ] [argmax_all       ] [argmax_axes       ] [RC_ARGMAX       ] [PartialOrd       ];
    [OpCountNonZeroAPI] [count_nonzero_all] [count_nonzero_axes] [RC_COUNT_NONZERO] [PartialEq + Zero ];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + Bound,
    D: DimAPI,
{
    type TOut = usize;

    fn func_all(&self, a: &CudaRaw<T>, la: &Layout<D>) -> Result<usize> {
        reduce_all::<T, usize, D>(self, CODE, a, la)
    }

    fn func_axes(&self, a: &CudaRaw<T>, la: &Layout<D>, axes: &[isize]) -> Result<(Storage<DataOwned<CudaRaw<usize>>, usize, Self>, Layout<IxD>)> {
        reduce_axes::<T, usize, D>(self, CODE, a, la, axes)
    }
}

// all / any over bool (:466-534)
#[duplicate_item(
     OpAPI      func_all  func_axes  CODE    ;
    [OpAllAPI] [all_all] [all_axes] [RC_ALL];
    [OpAnyAPI] [any_all] [any_axes] [RC_ANY];
)]
impl<D> OpAPI<bool, D> for DeviceCudaAutoImpl
where
    D: DimAPI,
{
    type TOut = bool;

    fn func_all(&self, a: &CudaRaw<bool>, la: &Layout<D>) -> Result<bool> {
        reduce_all::<bool, bool, D>(self, CODE, a, la)
    }

    fn func_axes(&self, a: &CudaRaw<bool>, la: &Layout<D>, axes: &[isize]) -> Result<(Storage<DataOwned<CudaRaw<bool>>, bool, Self>, Layout<IxD>)> {
        reduce_axes::<bool, bool, D>(self, CODE, a, la, axes)
    }
}

// sum of a bool tensor = count of `true` (:678-715)
impl<D> OpSumBoolAPI<D> for DeviceCudaAutoImpl
where
    D: DimAPI,
{
    fn sum_all(&self, a: &CudaRaw<bool>, la: &Layout<D>) -> Result<usize> {
        reduce_all::<bool, usize, D>(self, RC_COUNT_NONZERO, a, la)
    }

    fn sum_axes(&self, a: &CudaRaw<bool>, la: &Layout<D>, axes: &[isize]) -> Result<(Storage<DataOwned<CudaRaw<usize>>, usize, Self>, Layout<IxD>)> {
        reduce_axes::<bool, usize, D>(self, RC_COUNT_NONZERO, a, la, axes)
    }
}

// unraveled argmin / argmax (:572-676).  `_all`: the device returns the row-major flat index, unravelled here.
// `_axes` would need `IxD = Vec<usize>` ELEMENTS in device memory -- not a POD element type: UnImplemented.
#[duplicate_item(
     OpAPI                  func_all               func_axes               CODE        ;
    [OpUnraveledArgMinAPI] [unraveled_argmin_all] [unraveled_argmin_axes] [RC_ARGMIN];
    [OpUnraveledArgMaxAPI] [unraveled_argmax_all] [unraveled_argmax_axes] [RC_ARGMAX];
)]
impl<T, D> OpAPI<T, D> for DeviceCudaAutoImpl
where
    T: CudaDType + PartialOrd,
    D: DimAPI,
{
    fn func_all(&self, a: &CudaRaw<T>, la: &Layout<D>) -> Result<D> {
        let flat = reduce_all::<T, usize, D>(self, CODE, a, la)?;
        Ok(unsafe { la.shape().unravel_index_c(flat) })
    }

    fn func_axes(&self, _a: &CudaRaw<T>, _la: &Layout<D>, _axes: &[isize]) -> Result<(Storage<DataOwned<CudaRaw<IxD>>, IxD, Self>, Layout<IxD>)> {
        rstsr_raise!(UnImplemented, "DeviceCuda: unraveled arg* over axes yields Vec<usize> elements; use argmin_axes / argmax_axes and unravel on the host")
    }
}

// allclose (:717-763), TE = f64
impl<T, D> OpAllCloseAPI<T, T, f64, D> for DeviceCudaAutoImpl
where
    T: CudaDType,
    D: DimAPI,
{
    fn allclose_all(&self, a: &CudaRaw<T>, la: &Layout<D>, b: &CudaRaw<T>, lb: &Layout<D>, isclose_args: &IsCloseArgs<f64>) -> Result<bool> {
        let IsCloseArgs { rtol, atol, equal_nan } = isclose_args;
        let mut r: c_int = 0;
        check(unsafe { ffi::rc_allclose_all(self.raw(), T::CODE, a.ptr, &cl(la), b.ptr, &cl(lb), *rtol, *atol, *equal_nan as c_int, &mut r) })?;
        Ok(r != 0)
    }

    fn allclose_axes(
        &self,
        _a: &CudaRaw<T>,
        _la: &Layout<D>,
        _b: &CudaRaw<T>,
        _lb: &Layout<D>,
        _axes: &[isize],
        _isclose_args: &IsCloseArgs<f64>,
    ) -> Result<(Storage<DataOwned<CudaRaw<bool>>, bool, Self>, Layout<IxD>)> {
        unimplemented!("allclose_axes is unimplemented in the reference as well (auto_impl/reduction.rs:752-762)")
    }
}
