//! mod.rs — UNCOMPILED HERE.  /root/reference/src/special/mod.rs:1-13 with the `b200` switch: with the feature on,
//! the same public names resolve to the batched library; with it off the file is the reference's own.
#[cfg(feature = "b200")]
mod batched;
#[cfg(feature = "b200")]
mod ffi;
#[cfg(feature = "b200")]
pub use batched::*;

#[cfg(not(feature = "b200"))]
mod kv;
#[cfg(not(feature = "b200"))]
mod trig;
#[cfg(not(feature = "b200"))]
pub use kv::*;
#[cfg(not(feature = "b200"))]
pub use trig::*;

#[cfg(not(feature = "b200"))]
pub fn i0(x: ndarray::Array1<f64>) -> Result<ndarray::Array1<num::Complex<f64>>, i32> {
    // the reference's scalar path (mod.rs:8-13)
    x.mapv(|v| complex_bessel_rs::bessel_i::bessel_i(0.0, v.into())).into_iter().collect()
}
