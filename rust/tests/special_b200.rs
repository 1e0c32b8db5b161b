//! special_b200.rs — UNCOMPILED HERE.  The reference's own i0 test (/root/reference/tests/special.rs:23-46: 500
//! rounds of up to 100 random points in [0, 1) against `scipy.special.i0` through an embedded CPython, abs/rel 1e-14)
//! pointed at the batched path, plus the kv call of /root/reference/src/special/kv.rs:50-58.
#![cfg(feature = "b200")]
use ndarray::Array1;
use num::complex::Complex64;
use numpy::{PyArray1, PyArrayMethods, ToPyArray};
use pyo3::prelude::*;
use rand::Rng;
use sciport_rs::special::{i0, kv, kve};

#[test]
fn i0_matches_scipy() {
    let mut rng = rand::thread_rng();
    for _ in 0..500 {
        let len = rng.gen_range(1..100);
        let x: Vec<f64> = (0..len).map(|_| rng.gen_range(0.0..1.0)).collect();
        let got = i0(Array1::from_vec(x.clone())).unwrap();
        let want: Vec<f64> = Python::with_gil(|py| {
            let sc = py.import_bound("scipy.special").unwrap();
            let r = sc.call_method1("i0", (x.to_pyarray_bound(py),)).unwrap();
            r.downcast::<PyArray1<f64>>().unwrap().to_vec().unwrap()
        });
        for (g, w) in got.iter().zip(want.iter()) {
            assert!((g.re - w).abs() <= 1e-14 * w.abs().max(1.0) && g.im == 0.0, "{g} vs {w}");
        }
    }
}

#[test]
fn kv_negative_order_and_kve() {
    // kv.rs:52-57
    let c1 = Complex64::new(1000.0, 0.0);
    assert_eq!(kv(-3.5, c1), Complex64::new(0.0, 0.0));
    let _ = kve(0.0, 1.0.into());
    let res3 = kve(3.5, 1.0 / Complex64::new(1.2, 0.3));
    assert!(res3.is_finite());
}
