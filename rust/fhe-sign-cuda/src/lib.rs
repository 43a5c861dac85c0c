//! fhe-sign-cuda: `extern "C"` binding of libfhe_sign_cuda.so plus FheUint look-alikes whose operators
//! forward to the GPU, so that the reference's `src/biguint.rs` keeps its source apart from its `use` lines:
//!
//! ```text
//! - use tfhe::prelude::*;
//! - use tfhe::{FheUint32, FheUint64, ClientKey};
//! + use fhe_sign_cuda::prelude::*;
//! + use fhe_sign_cuda::{FheUint32, FheUint64, ClientKey};
//! ```
//!
//! Written blind (no Rust toolchain in the build image); every signature mirrors include/fhe_sign_cuda.h.
#![allow(non_camel_case_types)]
use std::cell::RefCell;
use std::ffi::CStr;
use std::ops::{Add, BitAnd, Div, Mul, Shr};
use std::os::raw::{c_char, c_void};
use std::rc::Rc;

pub mod sys {
    use super::*;
    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct fsc_params {
        pub lwe_dim: u32, pub glwe_dim: u32, pub poly_size: u32, pub pbs_base_log: u32, pub pbs_level: u32,
        pub ks_base_log: u32, pub ks_level: u32, pub message_modulus: u32, pub carry_modulus: u32, pub acc_bits: u32,
    }
    pub enum fsc_ctx {}
    pub enum fsc_radix {}
    extern "C" {
        pub fn fsc_ctx_create(p: *const fsc_params, device: i32, stream: usize, out: *mut *mut fsc_ctx) -> i32;
        pub fn fsc_ctx_destroy(ctx: *mut fsc_ctx) -> i32;
        pub fn fsc_last_error(ctx: *const fsc_ctx) -> *const c_char;
        pub fn fsc_keys_upload(ctx: *mut fsc_ctx, bsk: *const u64, bsk_words: usize, ksk: *const u64, ksk_words: usize) -> i32;
        pub fn fsc_radix_from_lwe(ctx: *mut fsc_ctx, blocks: *const u64, n_blocks: usize, out: *mut *mut fsc_radix) -> i32;
        pub fn fsc_radix_to_lwe(ctx: *mut fsc_ctx, r: *mut fsc_radix, blocks: *mut u64) -> i32;
        pub fn fsc_radix_trivial(ctx: *mut fsc_ctx, v: *const u8, n_bytes: usize, n_blocks: usize, out: *mut *mut fsc_radix) -> i32;
        pub fn fsc_radix_clone(ctx: *mut fsc_ctx, a: *const fsc_radix, out: *mut *mut fsc_radix) -> i32;
        pub fn fsc_radix_free(ctx: *mut fsc_ctx, a: *mut fsc_radix) -> i32;
        pub fn fsc_radix_binary(ctx: *mut fsc_ctx, op: u32, a: *const fsc_radix, b: *const fsc_radix, out: *mut *mut fsc_radix) -> i32;
        pub fn fsc_radix_scalar(ctx: *mut fsc_ctx, op: u32, a: *const fsc_radix, s: *const u8, n_bytes: usize, out: *mut *mut fsc_radix) -> i32;
        pub fn fsc_radix_cast(ctx: *mut fsc_ctx, a: *const fsc_radix, n_blocks: usize, out: *mut *mut fsc_radix) -> i32;
        // fused schedule: a * b without wrapping, and a * b + addend with one carry propagation
        pub fn fsc_radix_mul_wide(ctx: *mut fsc_ctx, a: *const fsc_radix, b: *const fsc_radix, out_blocks: usize, out: *mut *mut fsc_radix) -> i32;
        pub fn fsc_radix_mul_add_wide(ctx: *mut fsc_ctx, a: *const fsc_radix, b: *const fsc_radix, addend: *const fsc_radix, out_blocks: usize, out: *mut *mut fsc_radix) -> i32;
        // on-disk formats (csrc/keyfile.cpp): expanded server key written by the signer's side, read by the GPU host
        pub fn fsc_server_keys_load(path: *const c_char, params: *mut fsc_params, bsk: *mut *mut u64, bsk_words: *mut usize,
                                    ksk: *mut *mut u64, ksk_words: *mut usize) -> i32;
        pub fn fsc_buffer_free(buffer: *mut u64) -> i32;
    }
    pub const OP_ADD: u32 = 0; pub const OP_MUL: u32 = 2; pub const OP_MIN: u32 = 3; pub const OP_SHR: u32 = 5;
    pub const OP_AND: u32 = 7; pub const OP_DIV: u32 = 12;
}

/// GPU server key: what `set_server_key` installs (replaces tfhe::ServerKey, src/biguint.rs:278).
pub struct GpuServerKey { ctx: *mut sys::fsc_ctx }
impl Drop for GpuServerKey { fn drop(&mut self) { unsafe { sys::fsc_ctx_destroy(self.ctx); } } }

thread_local! { static SERVER: RefCell<Option<Rc<GpuServerKey>>> = RefCell::new(None); }

/// Mirrors tfhe::set_server_key: thread-local, used implicitly by every operator below.
pub fn set_server_key(key: GpuServerKey) { SERVER.with(|s| *s.borrow_mut() = Some(Rc::new(key))); }

fn server() -> Rc<GpuServerKey> { SERVER.with(|s| s.borrow().clone()).expect("set_server_key was not called on this thread") }

fn check(key: &GpuServerKey, rc: i32) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(sys::fsc_last_error(key.ctx)) }.to_string_lossy().into_owned();
        panic!("fhe-sign-cuda error {}: {}", rc, msg);      // the reference unwrap()s / panics at the same places
    }
}

impl GpuServerKey {
    /// Server key from an FSCFILE1 container written by `fsc_server_keys_save` (no secrets inside).
    pub fn from_file(path: &str, device: i32) -> Result<Self, String> {
        let c = std::ffi::CString::new(path).map_err(|e| e.to_string())?;
        let mut p = sys::fsc_params { lwe_dim: 0, glwe_dim: 0, poly_size: 0, pbs_base_log: 0, pbs_level: 0, ks_base_log: 0,
                                      ks_level: 0, message_modulus: 0, carry_modulus: 0, acc_bits: 0 };
        let (mut bsk, mut ksk) = (std::ptr::null_mut(), std::ptr::null_mut());
        let (mut nb, mut nk) = (0usize, 0usize);
        let rc = unsafe { sys::fsc_server_keys_load(c.as_ptr(), &mut p, &mut bsk, &mut nb, &mut ksk, &mut nk) };
        if rc != 0 { return Err(format!("fsc_server_keys_load failed with status {}", rc)); }
        p.acc_bits = 32;
        let key = unsafe { Self::new(p, device, std::slice::from_raw_parts(bsk, nb), std::slice::from_raw_parts(ksk, nk)) };
        unsafe { sys::fsc_buffer_free(bsk); }
        key
    }

    /// `bsk`: standard-domain bootstrapping key, `ksk`: keyswitching key, as exported by the client-side keygen.
    pub fn new(params: sys::fsc_params, device: i32, bsk: &[u64], ksk: &[u64]) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::fsc_ctx_create(&params, device, 0, &mut ctx) };
        if rc != 0 { return Err(unsafe { CStr::from_ptr(sys::fsc_last_error(std::ptr::null())) }.to_string_lossy().into_owned()); }
        let key = GpuServerKey { ctx };
        let rc = unsafe { sys::fsc_keys_upload(ctx, bsk.as_ptr(), bsk.len(), ksk.as_ptr(), ksk.len()) };
        if rc != 0 { return Err(unsafe { CStr::from_ptr(sys::fsc_last_error(ctx)) }.to_string_lossy().into_owned()); }
        Ok(key)
    }
}

/// N-block radix ciphertext living on the GPU; FheUint8/32/64 = 4/16/32 blocks.
pub struct FheUint<const BLOCKS: usize> { h: *mut sys::fsc_radix, key: Rc<GpuServerKey> }
pub type FheUint8 = FheUint<4>;
pub type FheUint32 = FheUint<16>;
pub type FheUint64 = FheUint<32>;

impl<const B: usize> Drop for FheUint<B> { fn drop(&mut self) { unsafe { sys::fsc_radix_free(self.key.ctx, self.h); } } }
impl<const B: usize> Clone for FheUint<B> {
    fn clone(&self) -> Self {
        let mut h = std::ptr::null_mut();
        check(&self.key, unsafe { sys::fsc_radix_clone(self.key.ctx, self.h, &mut h) });
        FheUint { h, key: self.key.clone() }
    }
}
impl<const B: usize> FheUint<B> {
    fn binary(&self, op: u32, rhs: &Self) -> Self {
        let mut h = std::ptr::null_mut();
        check(&self.key, unsafe { sys::fsc_radix_binary(self.key.ctx, op, self.h, rhs.h, &mut h) });
        FheUint { h, key: self.key.clone() }
    }
    fn scalar(&self, op: u32, s: u64) -> Self {
        let bytes = s.to_le_bytes();
        let mut h = std::ptr::null_mut();
        check(&self.key, unsafe { sys::fsc_radix_scalar(self.key.ctx, op, self.h, bytes.as_ptr(), 8, &mut h) });
        FheUint { h, key: self.key.clone() }
    }
    /// FheUintM::cast_from(FheUintN) (src/biguint.rs:110,116,135-137)
    pub fn cast_from<const A: usize>(x: FheUint<A>) -> Self {
        let mut h = std::ptr::null_mut();
        check(&x.key, unsafe { sys::fsc_radix_cast(x.key.ctx, x.h, B, &mut h) });
        FheUint { h, key: x.key.clone() }
    }
    pub fn min(&self, rhs: &Self) -> Self { self.binary(sys::OP_MIN, rhs) }
    /// blocks encrypted by the client (ClientKey::encrypt below) are handed to the device here
    pub fn from_blocks(blocks: &[u64]) -> Self {
        let key = server();
        let mut h = std::ptr::null_mut();
        check(&key, unsafe { sys::fsc_radix_from_lwe(key.ctx, blocks.as_ptr(), B, &mut h) });
        FheUint { h, key }
    }
    pub fn to_blocks(&self) -> Vec<u64> {
        let mut out = vec![0u64; B * 2049];
        check(&self.key, unsafe { sys::fsc_radix_to_lwe(self.key.ctx, self.h, out.as_mut_ptr()) });
        out
    }
}
impl<const B: usize> Add for FheUint<B> { type Output = Self; fn add(self, r: Self) -> Self { self.binary(sys::OP_ADD, &r) } }
impl<const B: usize> Mul for FheUint<B> { type Output = Self; fn mul(self, r: Self) -> Self { self.binary(sys::OP_MUL, &r) } }
impl<const B: usize> Shr<u64> for &FheUint<B> { type Output = FheUint<B>; fn shr(self, r: u64) -> FheUint<B> { self.scalar(sys::OP_SHR, r) } }
impl<const B: usize> Shr<&FheUint<B>> for &FheUint<B> { type Output = FheUint<B>; fn shr(self, r: &FheUint<B>) -> FheUint<B> { self.binary(sys::OP_SHR, r) } }
impl<const B: usize> BitAnd<u64> for &FheUint<B> { type Output = FheUint<B>; fn bitand(self, r: u64) -> FheUint<B> { self.scalar(sys::OP_AND, r) } }
impl<const B: usize> Div<u64> for &FheUint<B> { type Output = FheUint<B>; fn div(self, r: u64) -> FheUint<B> { self.scalar(sys::OP_DIV, r) } }

pub mod prelude { pub use super::{set_server_key, FheUint, FheUint32, FheUint64, FheUint8}; }
#[allow(dead_code)] fn _unused(_: *mut c_void) {}
