//! batched.rs — UNCOMPILED HERE (no Rust toolchain in the build image).
//!
//! The reference's `special` surface on the batched B200 library, same signatures, ordering and scaling flags:
//!   * `i0`   — drop-in for /root/reference/src/special/mod.rs:8-13
//!   * `kv`   — drop-in for /root/reference/src/special/kv.rs:11-25 (NaN sanitising, `unwrap()` panic on Err)
//!   * `kve`  — /root/reference/src/special/kv.rs:39-41 (`kv(v, z) * z.exp()`, post-multiplication)
//!   * `sinc` — /root/reference/src/special/trig.rs:4-13
//! plus the widened batched entry points the north star names (`jv_array`, `yv_array`, `iv_array`, `kv_array`,
//! `hankel1_array`, `hankel2_array` and their `*e_array` scaled forms, `airy_array`, `gamma_array`, `lgamma_array`,
//! `jn_sequence`): order first, then z (kv.rs:11); unscaled `xv` / scaled `xve` as separate names (kv / kve);
//! owned ndarray in, `Result<owned ndarray, i32>` out where the first Err in index order wins (mod.rs:8,11-12).
use super::ffi::*;
use ndarray::{Array, Array1, Array2, Dimension};
use num::{Complex, Float, NumCast, Zero};
use std::os::raw::c_int;

/// GPUs a host batch is sliced over (contiguous slices, one host thread + two streams per device, no collective).
fn exec() -> SpbExec {
    let g = std::env::var("SCIPORT_B200_DEVICES").ok().and_then(|s| s.parse::<c_int>().ok()).unwrap_or(1);
    SpbExec::host(g)
}

fn lib_ok(rc: c_int) {
    assert_eq!(rc, 0, "libsciport_b200: {}", last_error());
}

/// `.collect::<Result<_, i32>>()` of mod.rs:11-12 without per-element codes: the kernels keep the smallest failing
/// (index, ierr) of the call in one device word.
fn first_err(which: c_int) -> Result<(), i32> {
    let (mut idx, mut code) = (-1i64, 0i32);
    lib_ok(unsafe { spb_last_first_error(which, &mut idx, &mut code) });
    if idx >= 0 { Err(code) } else { Ok(()) }
}

/// One family over a batch: `v` is one broadcast order (len 1) or one order per point.
fn bessel_array(func: c_int, kode: c_int, v: &[f64], z: Array1<Complex<f64>>) -> Result<Array1<Complex<f64>>, i32> {
    let n = z.len();
    assert!(v.len() == 1 || v.len() == n, "v must be a scalar or match z");
    let z = z.as_standard_layout();
    let mut out = Array1::<Complex<f64>>::zeros(n);
    let stride = if v.len() == 1 { 0 } else { 1 };
    lib_ok(unsafe {
        spb_bessel(func, kode, v.as_ptr(), stride, z.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut(),
                   std::ptr::null_mut(), n as i64, 1, &exec())
    });
    first_err(0).map(|_| out)
}

/// Drop-in for `special::i0` (mod.rs:8-13): I_0 of a real array through the complex routine (`v.into()`).
pub fn i0(x: Array1<f64>) -> Result<Array1<Complex<f64>>, i32> {
    let n = x.len();
    let x = x.as_standard_layout();
    let mut out = Array1::<Complex<f64>>::zeros(n);
    let order = 0.0f64; // the literal of mod.rs:10
    lib_ok(unsafe {
        spb_bessel_real(SPB_FN_I, SPB_KODE_UNSCALED, &order, 0, x.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut(),
                        std::ptr::null_mut(), n as i64, 1, &exec())
    });
    first_err(0).map(|_| out)
}

/// Drop-in for `special::kv` (kv.rs:11-25).  Negative orders are evaluated inside the library (K_-v = K_v), so
/// `kv(-3.5, 1000)` (kv.rs:53-54) returns 0 like the reference instead of an Err.
pub fn kv(mut v: f64, mut z: Complex<f64>) -> Complex<f64> {
    if z.is_nan() {
        z = Complex::zero(); // kv.rs:12-14
    }
    if v.is_nan() {
        v = 0.0; // kv.rs:16-18
    }
    let mut out = Complex::<f64>::zero();
    let mut ierr = 0i32;
    lib_ok(unsafe {
        spb_bessel(SPB_FN_K, SPB_KODE_UNSCALED, &v, 0, &z, &mut out, &mut ierr, std::ptr::null_mut(), 1, 1, &exec())
    });
    let res: Result<Complex<f64>, i32> = if ierr == 0 { Ok(out) } else { Err(ierr) };
    if res.is_err() {
        println!("{v} {z}"); // kv.rs:20-22
    }
    res.unwrap() // kv.rs:24
}

/// `special::kve` (kv.rs:39-41)
pub fn kve(v: f64, z: Complex<f64>) -> Complex<f64> {
    kv(v, z) * z.exp()
}

/// `special::sinc` (trig.rs:4-13): computed in f64 on the device, cast back to T.
pub fn sinc<T: Float, D: Dimension>(x: Array<T, D>) -> Array<T, D> {
    let shape = x.raw_dim();
    let xf: Vec<f64> = x.iter().map(|a| <f64 as NumCast>::from(*a).unwrap()).collect();
    let mut out = vec![0.0f64; xf.len()];
    lib_ok(unsafe { spb_sinc(xf.as_ptr(), out.as_mut_ptr(), xf.len() as i64, &exec()) });
    Array::from_shape_vec(shape, out.into_iter().map(|a| T::from(a).unwrap()).collect()).unwrap()
}

macro_rules! family {
    ($plain:ident, $scaled:ident, $func:expr) => {
        /// unscaled, per-element (or one broadcast) order; first Err in index order wins
        pub fn $plain(v: &[f64], z: Array1<Complex<f64>>) -> Result<Array1<Complex<f64>>, i32> {
            bessel_array($func, SPB_KODE_UNSCALED, v, z)
        }
        /// exponentially scaled (KODE = 2)
        pub fn $scaled(v: &[f64], z: Array1<Complex<f64>>) -> Result<Array1<Complex<f64>>, i32> {
            bessel_array($func, SPB_KODE_SCALED, v, z)
        }
    };
}
family!(jv_array, jve_array, SPB_FN_J);
family!(yv_array, yve_array, SPB_FN_Y);
family!(iv_array, ive_array, SPB_FN_I);
family!(kv_array, kve_array, SPB_FN_K);
family!(hankel1_array, hankel1e_array, SPB_FN_H1);
family!(hankel2_array, hankel2e_array, SPB_FN_H2);

/// J_v and Y_v in one pass (shared classification and I pass).
pub fn jvyv_array(v: &[f64], z: Array1<Complex<f64>>)
                  -> Result<(Array1<Complex<f64>>, Array1<Complex<f64>>), i32> {
    let n = z.len();
    let z = z.as_standard_layout();
    let (mut j, mut y) = (Array1::<Complex<f64>>::zeros(n), Array1::<Complex<f64>>::zeros(n));
    let stride = if v.len() == 1 { 0 } else { 1 };
    let nul = std::ptr::null_mut::<i32>();
    lib_ok(unsafe {
        spb_bessel_jy(SPB_KODE_UNSCALED, v.as_ptr(), stride, z.as_ptr(), j.as_mut_ptr(), y.as_mut_ptr(), nul, nul, nul,
                      nul, n as i64, &exec())
    });
    first_err(0)?;
    first_err(1)?;
    Ok((j, y))
}

/// Order sequence J_{v0+k}(z_i), k < n_orders, as an (n_orders, n) array (order-major: BASELINE config 5).
pub fn jn_sequence(n_orders: usize, v0: f64, z: Array1<Complex<f64>>) -> Result<Array2<Complex<f64>>, i32> {
    let n = z.len();
    let z = z.as_standard_layout();
    let mut out = Array2::<Complex<f64>>::zeros((n_orders, n));
    lib_ok(unsafe {
        spb_bessel(SPB_FN_J, SPB_KODE_UNSCALED, &v0, 0, z.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut(),
                   std::ptr::null_mut(), n as i64, n_orders as c_int, &exec())
    });
    first_err(0).map(|_| out)
}

/// Airy Ai / Ai' / Bi / Bi' (`which` = SPB_AIRY_*), unscaled.
pub fn airy_array(which: c_int, z: Array1<Complex<f64>>) -> Result<Array1<Complex<f64>>, i32> {
    let n = z.len();
    let z = z.as_standard_layout();
    let mut out = Array1::<Complex<f64>>::zeros(n);
    lib_ok(unsafe {
        spb_airy(which, SPB_KODE_UNSCALED, z.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut(),
                 n as i64, &exec())
    });
    first_err(0).map(|_| out)
}

pub fn gamma_array(x: Array1<f64>) -> Array1<f64> {
    let x = x.as_standard_layout();
    let mut out = Array1::<f64>::zeros(x.len());
    lib_ok(unsafe { spb_gamma(x.as_ptr(), out.as_mut_ptr(), x.len() as i64, &exec()) });
    out
}

pub fn lgamma_array(x: Array1<f64>) -> Array1<f64> {
    let x = x.as_standard_layout();
    let mut out = Array1::<f64>::zeros(x.len());
    lib_ok(unsafe { spb_lgamma(x.as_ptr(), out.as_mut_ptr(), x.len() as i64, &exec()) });
    out
}
