//! ffi.rs — UNCOMPILED HERE (no Rust toolchain in the build image).
//!
//! `extern "C"` view of include/sciport_b200.h.  `num::Complex<f64>` is `#[repr(C)] { re, im }`, i.e. `spb_c128`
//! (= CUDA `double2`), so arrays of it cross the boundary without conversion.
#![allow(non_camel_case_types, dead_code)]
use num::Complex;
use std::os::raw::{c_char, c_int, c_void};

/// `spb_exec` (sciport_b200.h): where the buffers live and which GPUs to use.
#[repr(C)]
pub struct SpbExec {
    pub n_devices: c_int,        // 0 | 1: one device; > 1: contiguous slice per device (host pointers)
    pub devices: *const c_int,   // device ordinals, null = 0..n_devices-1 (or the current device)
    pub device_ptrs: c_int,      // 0: host memory (pinned or pageable); 1: device memory on devices[0]
    pub stream: *mut c_void,     // cudaStream_t for device_ptrs = 1
    pub flags: c_int,            // SPB_SEM_* | SPB_RAW_NEG_ORDERS | ...
}

impl SpbExec {
    /// host buffers, `n_devices` GPUs (0 or 1 = one)
    pub fn host(n_devices: c_int) -> Self {
        SpbExec { n_devices, devices: std::ptr::null(), device_ptrs: 0, stream: std::ptr::null_mut(), flags: 0 }
    }
}

// enum spb_fn
pub const SPB_FN_J: c_int = 0;
pub const SPB_FN_Y: c_int = 1;
pub const SPB_FN_I: c_int = 2; // replaces complex_bessel_rs::bessel_i::bessel_i (mod.rs:10)
pub const SPB_FN_K: c_int = 3; // replaces complex_bessel_rs::bessel_k::bessel_k (kv.rs:19)
pub const SPB_FN_H1: c_int = 4;
pub const SPB_FN_H2: c_int = 5;
// enum spb_kode
pub const SPB_KODE_UNSCALED: c_int = 1;
pub const SPB_KODE_SCALED: c_int = 2;
// enum spb_airy_which
pub const SPB_AIRY_AI: c_int = 0;
pub const SPB_AIRY_AIP: c_int = 1;
pub const SPB_AIRY_BI: c_int = 2;
pub const SPB_AIRY_BIP: c_int = 3;
// enum spb_flags
pub const SPB_SEM_SCIPY: c_int = 1;
pub const SPB_RAW_NEG_ORDERS: c_int = 2;

extern "C" {
    pub fn spb_bessel(func: c_int, kode: c_int, v: *const f64, v_stride: i64, z: *const Complex<f64>,
                      out: *mut Complex<f64>, ierr: *mut i32, nz: *mut i32, n: i64, n_orders: c_int,
                      ex: *const SpbExec) -> c_int;
    pub fn spb_bessel_real(func: c_int, kode: c_int, v: *const f64, v_stride: i64, x: *const f64,
                           out: *mut Complex<f64>, ierr: *mut i32, nz: *mut i32, n: i64, n_orders: c_int,
                           ex: *const SpbExec) -> c_int;
    pub fn spb_bessel_jy(kode: c_int, v: *const f64, v_stride: i64, z: *const Complex<f64>,
                         out_j: *mut Complex<f64>, out_y: *mut Complex<f64>, ierr_j: *mut i32, ierr_y: *mut i32,
                         nz_j: *mut i32, nz_y: *mut i32, n: i64, ex: *const SpbExec) -> c_int;
    pub fn spb_airy(which: c_int, kode: c_int, z: *const Complex<f64>, out: *mut Complex<f64>, ierr: *mut i32,
                    nz: *mut i32, n: i64, ex: *const SpbExec) -> c_int;
    pub fn spb_gamma(x: *const f64, out: *mut f64, n: i64, ex: *const SpbExec) -> c_int;
    pub fn spb_lgamma(x: *const f64, out: *mut f64, n: i64, ex: *const SpbExec) -> c_int;
    pub fn spb_cgamma(z: *const Complex<f64>, out: *mut Complex<f64>, n: i64, ex: *const SpbExec) -> c_int;
    pub fn spb_clgamma(z: *const Complex<f64>, out: *mut Complex<f64>, n: i64, ex: *const SpbExec) -> c_int;
    pub fn spb_sinc(x: *const f64, out: *mut f64, n: i64, ex: *const SpbExec) -> c_int;
    pub fn spb_first_error(ierr: *const i32, n: i64, idx: *mut i64, code: *mut i32, ex: *const SpbExec) -> c_int;
    pub fn spb_last_first_error(which: c_int, idx: *mut i64, code: *mut i32) -> c_int;
    pub fn spb_kaiser(m: i64, beta: f64, sym: c_int, out: *mut f64, ex: *const SpbExec) -> c_int;
    pub fn spb_lanczos(m: i64, sym: c_int, out: *mut f64, ex: *const SpbExec) -> c_int;
    pub fn spb_window(kind: c_int, m: i64, sym: c_int, params: *const f64, n_params: c_int, out: *mut f64,
                      ex: *const SpbExec) -> c_int;
    pub fn spb_device_count() -> c_int;
    pub fn spb_last_error() -> *const c_char;
}

/// Text of the last library error of this thread.
pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(spb_last_error()).to_string_lossy().into_owned() }
}
