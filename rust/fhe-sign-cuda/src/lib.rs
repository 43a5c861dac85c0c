//! fhe-sign-cuda: the slice of the `tfhe` 0.10 high-level API that fhe-sign uses, over libfhe_sign_cuda.so (B200).
//!
//! Every name the reference imports from `tfhe` exists here with the same shape, so the crate can stand in for it through
//! a Cargo rename and **no source line of the reference changes**:
//!
//! ```toml
//! # Cargo.toml of key-protocol (was: tfhe = { version = "*", features = [...] })
//! tfhe = { package = "fhe-sign-cuda", path = "../fhe-sign-cuda" }
//! ```
//!
//! | reference use (file:line) | here |
//! |---|---|
//! | `ConfigBuilder::default().build()` (src/biguint.rs:276) | [`ConfigBuilder`], [`Config`] = the 2_2 parameter set |
//! | `tfhe::generate_keys(config)` (:277) | [`generate_keys`] -> ([`ClientKey`], [`ServerKey`]) via `fsc_client_keygen` (OS entropy) |
//! | `tfhe::set_server_key(server_key)` (:278) | [`set_server_key`]: `fsc_ctx_create` + `fsc_keys_upload`, thread-local |
//! | `FheUint32::try_encrypt(d, client_key)` (:26,207), `-> Result<_, tfhe::Error>` (:17) | [`prelude::FheTryEncrypt`], [`Error`] |
//! | `digit.decrypt(client_key)` (:70) | [`prelude::FheDecrypt`] |
//! | `FheUint64::cast_from(x)` (:110,135), `x.cast_into()` (src/perf_test.rs:40) | [`prelude::CastFrom`], [`prelude::CastInto`] |
//! | `a + b`, `a * b`, `&a + &b`, `&a * &b` (:138,223,248; src/perf_test.rs:28,32) | `Add`, `Mul` by value and by reference |
//! | `&x >> 32u64`, `&x & 0xFFFFFFFFu64` (:141,143), `x & 1_u8` (src/perf_test.rs:48) | `Shr<u64>`, `BitAnd<u64>`, `BitAnd<u8>` |
//! | `&a >> &b` (src/perf_test.rs:36), `a.min(&c)` (:44), `&a / 5` (:54) | `Shr<&FheUint>`, [`prelude::FheOrd`], `Div<u32>` |
//! | `encrypted * 456u32`, `encrypted + 456u32` (src/schnorr.rs:588,604) | `Mul<u32>`, `Add<u32>` |
//! | `FheBool` (imported, unused: src/perf_test.rs:5) | [`FheBool`] |
//!
//! Written blind: the build image has no Rust toolchain, so this file has never been compiled; every `extern "C"` signature
//! mirrors include/fhe_sign_cuda.h and the Python ctypes binding (fhe_sign_b200/{capi,radix,client}.py), which IS exercised
//! by the test-suite, follows the same call sequences.  Run `cargo check` before relying on it.
#![allow(non_camel_case_types)]
use std::cell::RefCell;
use std::ffi::CStr;
use std::fmt;
use std::ops::{Add, BitAnd, Div, Mul, Shr};
use std::os::raw::c_char;
use std::rc::Rc;

pub mod sys {
    use std::os::raw::c_char;
    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct fsc_params {
        pub lwe_dim: u32, pub glwe_dim: u32, pub poly_size: u32, pub pbs_base_log: u32, pub pbs_level: u32,
        pub ks_base_log: u32, pub ks_level: u32, pub message_modulus: u32, pub carry_modulus: u32, pub acc_bits: u32,
    }
    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct fsc_noise_params {
        pub noise_kind: u32, pub lwe_tuniform_bound: u32, pub glwe_tuniform_bound: u32, pub reserved: u32,
        pub lwe_noise_std: f64, pub glwe_noise_std: f64,
    }
    pub enum fsc_ctx {}
    pub enum fsc_radix {}
    pub enum fsc_client {}
    pub const FSC_PEER_HANDLE_BYTES: usize = 128;
    extern "C" {
        pub fn fsc_ctx_create(p: *const fsc_params, device: i32, stream: usize, out: *mut *mut fsc_ctx) -> i32;
        pub fn fsc_ctx_destroy(ctx: *mut fsc_ctx) -> i32;
        pub fn fsc_last_error(ctx: *const fsc_ctx) -> *const c_char;
        pub fn fsc_sync(ctx: *mut fsc_ctx) -> i32;
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
        pub fn fsc_radix_slice(ctx: *mut fsc_ctx, a: *const fsc_radix, first: usize, n_blocks: usize, out: *mut *mut fsc_radix) -> i32;
        pub fn fsc_radix_concat(ctx: *mut fsc_ctx, parts: *const *const fsc_radix, n_parts: usize, out: *mut *mut fsc_radix) -> i32;
        // multi-GPU level sharding owned by the library: peer-mapped block pools, exchange fused into the blind rotation
        pub fn fsc_peer_pool_export(ctx: *mut fsc_ctx, capacity_blocks: usize, handle_out: *mut u8) -> i32;
        pub fn fsc_peer_pool_connect(ctx: *mut fsc_ctx, rank: i32, world: i32, min_width: usize, handles: *const u8) -> i32;
        pub fn fsc_peer_pool_disconnect(ctx: *mut fsc_ctx) -> i32;
        // client side (host CPU)
        pub fn fsc_client_keygen(p: *const fsc_params, noise: *const fsc_noise_params, out: *mut *mut fsc_client) -> i32;
        pub fn fsc_client_keygen_seeded(p: *const fsc_params, noise: *const fsc_noise_params, seed: u64, out: *mut *mut fsc_client) -> i32;
        pub fn fsc_client_free(c: *mut fsc_client) -> i32;
        pub fn fsc_client_last_error(c: *const fsc_client) -> *const c_char;
        pub fn fsc_client_server_keys(c: *const fsc_client, bsk: *mut *const u64, bsk_words: *mut usize, ksk: *mut *const u64, ksk_words: *mut usize) -> i32;
        pub fn fsc_client_encrypt_blocks(c: *mut fsc_client, values: *const u8, n_blocks: usize, out_blocks: *mut u64) -> i32;
        pub fn fsc_client_decrypt_blocks(c: *mut fsc_client, blocks: *const u64, n_blocks: usize, values: *mut u8, noise: *mut i64) -> i32;
        pub fn fsc_client_save(c: *const fsc_client, path: *const c_char) -> i32;
        pub fn fsc_client_load(path: *const c_char, out: *mut *mut fsc_client) -> i32;
        // on-disk formats (csrc/keyfile.cpp): expanded server key written by the signer's side, read by the GPU host
        pub fn fsc_server_keys_save(c: *const fsc_client, path: *const c_char) -> i32;
        pub fn fsc_server_keys_load(path: *const c_char, params: *mut fsc_params, bsk: *mut *mut u64, bsk_words: *mut usize,
                                    ksk: *mut *mut u64, ksk_words: *mut usize) -> i32;
        pub fn fsc_buffer_free(buffer: *mut u64) -> i32;
    }
    pub const OP_ADD: u32 = 0; pub const OP_SUB: u32 = 1; pub const OP_MUL: u32 = 2; pub const OP_MIN: u32 = 3; pub const OP_MAX: u32 = 4;
    pub const OP_SHR: u32 = 5; pub const OP_SHL: u32 = 6; pub const OP_AND: u32 = 7; pub const OP_DIV: u32 = 12; pub const OP_REM: u32 = 13;
}

// ---- errors: `tfhe::Error` (src/biguint.rs:17, src/schnorr.rs:154,235) ---------------------------------------------------
#[derive(Debug, Clone)]
pub struct Error { pub status: i32, pub message: String }
impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "fhe-sign-cuda error {}: {}", self.status, self.message) }
}
impl std::error::Error for Error {}
fn cstr(p: *const c_char) -> String { if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() } }

// ---- configuration: `ConfigBuilder::default().build()` (src/biguint.rs:276) ----------------------------------------------
/// PARAM_MESSAGE_2_CARRY_2_KS_PBS as plain data (Gaussian flavour n = 834; `tuniform()` selects n = 887).
#[derive(Clone, Copy)]
pub struct Config { pub params: sys::fsc_params, pub noise: sys::fsc_noise_params, pub device: i32 }
#[derive(Clone, Copy)]
pub struct ConfigBuilder { cfg: Config }
impl Default for ConfigBuilder {
    fn default() -> Self {
        let params = sys::fsc_params { lwe_dim: 834, glwe_dim: 1, poly_size: 2048, pbs_base_log: 23, pbs_level: 1, ks_base_log: 3,
                                       ks_level: 5, message_modulus: 4, carry_modulus: 4, acc_bits: 32 };
        let noise = sys::fsc_noise_params { noise_kind: 0, lwe_tuniform_bound: 0, glwe_tuniform_bound: 0, reserved: 0,
                                            lwe_noise_std: 3.5539902359442825e-06, glwe_noise_std: 2.845267479601915e-15 };
        ConfigBuilder { cfg: Config { params, noise, device: 0 } }
    }
}
impl ConfigBuilder {
    pub fn tuniform(mut self) -> Self {
        self.cfg.params.lwe_dim = 887; self.cfg.params.pbs_base_log = 22;
        self.cfg.noise = sys::fsc_noise_params { noise_kind: 1, lwe_tuniform_bound: 46, glwe_tuniform_bound: 17, reserved: 0,
                                                 lwe_noise_std: 0.0, glwe_noise_std: 0.0 };
        self
    }
    /// the reference's accumulator width (64) instead of the engine's default (32)
    pub fn accumulator_bits(mut self, bits: u32) -> Self { self.cfg.params.acc_bits = bits; self }
    pub fn device(mut self, device: i32) -> Self { self.cfg.device = device; self }
    pub fn build(self) -> Config { self.cfg }
}

// ---- keys -----------------------------------------------------------------------------------------------------------------
struct ClientInner { h: *mut sys::fsc_client, cfg: Config }
impl Drop for ClientInner { fn drop(&mut self) { unsafe { sys::fsc_client_free(self.h); } } }
/// `tfhe::ClientKey`: secret keys on the host CPU; `Clone` is a shared handle (src/biguint.rs:52 clones it per value).
#[derive(Clone)]
pub struct ClientKey { inner: Rc<ClientInner> }
/// `tfhe::ServerKey` as `generate_keys` returns it: the expanded key material, still on the host.  `set_server_key` puts it on the GPU.
pub struct ServerKey { cfg: Config, bsk: Vec<u64>, ksk: Vec<u64> }

/// `tfhe::generate_keys(config)` (src/biguint.rs:277): OS entropy, like the reference (`seeder_unix`, Cargo.toml:9).
pub fn generate_keys(config: Config) -> (ClientKey, ServerKey) {
    let mut h = std::ptr::null_mut();
    let rc = unsafe { sys::fsc_client_keygen(&config.params, &config.noise, &mut h) };
    if rc != 0 { panic!("key generation failed ({}): {}", rc, cstr(unsafe { sys::fsc_client_last_error(std::ptr::null()) })); }
    let ck = ClientKey { inner: Rc::new(ClientInner { h, cfg: config }) };
    let sk = ck.server_key();
    (ck, sk)
}
impl ClientKey {
    fn server_key(&self) -> ServerKey {
        let (mut b, mut k) = (std::ptr::null(), std::ptr::null());
        let (mut nb, mut nk) = (0usize, 0usize);
        unsafe {
            sys::fsc_client_server_keys(self.inner.h, &mut b, &mut nb, &mut k, &mut nk);
            ServerKey { cfg: self.inner.cfg, bsk: std::slice::from_raw_parts(b, nb).to_vec(), ksk: std::slice::from_raw_parts(k, nk).to_vec() }
        }
    }
    fn encrypt(&self, value: u64, n_blocks: usize) -> Result<Vec<u64>, Error> {
        let digits: Vec<u8> = (0..n_blocks).map(|i| ((value >> (2 * i).min(62)) & 3) as u8 * ((2 * i < 64) as u8)).collect();
        let words = self.inner.cfg.params.glwe_dim as usize * self.inner.cfg.params.poly_size as usize + 1;
        let mut out = vec![0u64; n_blocks * words];
        let rc = unsafe { sys::fsc_client_encrypt_blocks(self.inner.h, digits.as_ptr(), n_blocks, out.as_mut_ptr()) };
        if rc != 0 { return Err(Error { status: rc, message: cstr(unsafe { sys::fsc_client_last_error(self.inner.h) }) }); }
        Ok(out)
    }
    fn decrypt(&self, blocks: &[u64], n_blocks: usize) -> u64 {
        let mut vals = vec![0u8; n_blocks];
        let rc = unsafe { sys::fsc_client_decrypt_blocks(self.inner.h, blocks.as_ptr(), n_blocks, vals.as_mut_ptr(), std::ptr::null_mut()) };
        assert!(rc == 0, "decryption failed: {}", cstr(unsafe { sys::fsc_client_last_error(self.inner.h) }));
        vals.iter().enumerate().fold(0u64, |acc, (i, v)| if 2 * i < 64 { acc | (((*v & 3) as u64) << (2 * i)) } else { acc })
    }
}

/// GPU server key: what `set_server_key` installs (replaces the thread-local tfhe ServerKey, src/biguint.rs:278).
pub struct GpuServerKey { ctx: *mut sys::fsc_ctx, words: usize }
impl Drop for GpuServerKey { fn drop(&mut self) { unsafe { sys::fsc_ctx_destroy(self.ctx); } } }
impl GpuServerKey {
    pub fn new(params: sys::fsc_params, device: i32, bsk: &[u64], ksk: &[u64]) -> Result<Self, Error> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::fsc_ctx_create(&params, device, 0, &mut ctx) };
        if rc != 0 { return Err(Error { status: rc, message: cstr(unsafe { sys::fsc_last_error(std::ptr::null()) }) }); }
        let key = GpuServerKey { ctx, words: params.glwe_dim as usize * params.poly_size as usize + 1 };
        let rc = unsafe { sys::fsc_keys_upload(ctx, bsk.as_ptr(), bsk.len(), ksk.as_ptr(), ksk.len()) };
        if rc != 0 { return Err(Error { status: rc, message: cstr(unsafe { sys::fsc_last_error(ctx) }) }); }
        Ok(key)
    }
    /// Server key from an FSCFILE1 container written by `fsc_server_keys_save` (no secrets inside).
    pub fn from_file(path: &str, device: i32) -> Result<Self, Error> {
        let c = std::ffi::CString::new(path).map_err(|e| Error { status: 1, message: e.to_string() })?;
        let mut p = ConfigBuilder::default().build().params;
        let (mut bsk, mut ksk) = (std::ptr::null_mut(), std::ptr::null_mut());
        let (mut nb, mut nk) = (0usize, 0usize);
        let rc = unsafe { sys::fsc_server_keys_load(c.as_ptr(), &mut p, &mut bsk, &mut nb, &mut ksk, &mut nk) };
        if rc != 0 { return Err(Error { status: rc, message: cstr(unsafe { sys::fsc_client_last_error(std::ptr::null()) }) }); }
        if p.acc_bits == 0 { p.acc_bits = 32; }
        let key = unsafe { Self::new(p, device, std::slice::from_raw_parts(bsk, nb), std::slice::from_raw_parts(ksk, nk)) };
        unsafe { sys::fsc_buffer_free(bsk); }
        key
    }
    pub fn raw(&self) -> *mut sys::fsc_ctx { self.ctx }
}

thread_local! { static SERVER: RefCell<Option<Rc<GpuServerKey>>> = RefCell::new(None); }

/// `tfhe::set_server_key(server_key)`: thread-local, used implicitly by every operator below.  Uploads the key to the GPU
/// (Fourier conversion on the device); panics where tfhe would (no usable device: there is no CPU fallback).
pub fn set_server_key(key: ServerKey) {
    let gpu = GpuServerKey::new(key.cfg.params, key.cfg.device, &key.bsk, &key.ksk).unwrap_or_else(|e| panic!("{}", e));
    set_gpu_server_key(gpu);
}
pub fn set_gpu_server_key(key: GpuServerKey) { SERVER.with(|s| *s.borrow_mut() = Some(Rc::new(key))); }
fn server() -> Rc<GpuServerKey> { SERVER.with(|s| s.borrow().clone()).expect("set_server_key was not called on this thread") }
fn check(key: &GpuServerKey, rc: i32) {
    if rc != 0 { panic!("fhe-sign-cuda error {}: {}", rc, cstr(unsafe { sys::fsc_last_error(key.ctx) })); }      // the reference unwrap()s / panics at the same places
}

// ---- FheUint look-alikes --------------------------------------------------------------------------------------------------
/// BLOCKS-block radix ciphertext living on the GPU; FheUint8/32/64 = 4/16/32 blocks of 2 message bits.
pub struct FheUint<const BLOCKS: usize> { h: *mut sys::fsc_radix, key: Rc<GpuServerKey> }
pub type FheUint8 = FheUint<4>;
pub type FheUint32 = FheUint<16>;
pub type FheUint64 = FheUint<32>;
pub type FheBool = FheUint<1>;

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
    fn cast<const A: usize>(x: &FheUint<A>) -> Self {
        let mut h = std::ptr::null_mut();
        check(&x.key, unsafe { sys::fsc_radix_cast(x.key.ctx, x.h, B, &mut h) });
        FheUint { h, key: x.key.clone() }
    }
    /// blocks encrypted by the client are handed to the device here (`blocks.len()` must be BLOCKS x (k N + 1) words)
    pub fn from_blocks(blocks: &[u64]) -> Self {
        let key = server();
        assert!(blocks.len() == B * key.words, "from_blocks: expected {} words, got {}", B * key.words, blocks.len());
        let mut h = std::ptr::null_mut();
        check(&key, unsafe { sys::fsc_radix_from_lwe(key.ctx, blocks.as_ptr(), B, &mut h) });
        FheUint { h, key }
    }
    pub fn to_blocks(&self) -> Vec<u64> {
        let mut out = vec![0u64; B * self.key.words];
        check(&self.key, unsafe { sys::fsc_radix_to_lwe(self.key.ctx, self.h, out.as_mut_ptr()) });
        out
    }
    /// a * b + addend over `OUT` blocks with one carry propagation: the fused form of `k + e * d` (src/schnorr.rs:274)
    pub fn mul_add_wide<const OUT: usize>(&self, b: &Self, addend: &Self) -> FheUint<OUT> {
        let mut h = std::ptr::null_mut();
        check(&self.key, unsafe { sys::fsc_radix_mul_add_wide(self.key.ctx, self.h, b.h, addend.h, OUT, &mut h) });
        FheUint { h, key: self.key.clone() }
    }
}

macro_rules! binop {
    ($tr:ident, $f:ident, $op:expr) => {
        impl<const B: usize> $tr for FheUint<B> { type Output = FheUint<B>; fn $f(self, r: Self) -> FheUint<B> { self.binary($op, &r) } }
        impl<const B: usize> $tr<&FheUint<B>> for FheUint<B> { type Output = FheUint<B>; fn $f(self, r: &FheUint<B>) -> FheUint<B> { self.binary($op, r) } }
        impl<const B: usize> $tr<FheUint<B>> for &FheUint<B> { type Output = FheUint<B>; fn $f(self, r: FheUint<B>) -> FheUint<B> { self.binary($op, &r) } }
        impl<const B: usize> $tr<&FheUint<B>> for &FheUint<B> { type Output = FheUint<B>; fn $f(self, r: &FheUint<B>) -> FheUint<B> { self.binary($op, r) } }
    };
}
binop!(Add, add, sys::OP_ADD);
binop!(Mul, mul, sys::OP_MUL);
binop!(Shr, shr, sys::OP_SHR);      // `&a >> &b`: encrypted amount, taken modulo the width (src/perf_test.rs:36)
macro_rules! scalarop {
    ($tr:ident, $f:ident, $op:expr, $($t:ty),+) => { $(
        impl<const B: usize> $tr<$t> for FheUint<B> { type Output = FheUint<B>; fn $f(self, r: $t) -> FheUint<B> { self.scalar($op, r as u64) } }
        impl<const B: usize> $tr<$t> for &FheUint<B> { type Output = FheUint<B>; fn $f(self, r: $t) -> FheUint<B> { self.scalar($op, r as u64) } }
    )+ };
}
scalarop!(Add, add, sys::OP_ADD, u8, u32, u64);      // src/schnorr.rs:604
scalarop!(Mul, mul, sys::OP_MUL, u8, u32, u64);      // src/schnorr.rs:588
scalarop!(Shr, shr, sys::OP_SHR, u8, u32, u64);      // src/biguint.rs:110,141
scalarop!(BitAnd, bitand, sys::OP_AND, u8, u32, u64);  // src/biguint.rs:116,143 ; src/perf_test.rs:48
scalarop!(Div, div, sys::OP_DIV, u8, u32, u64, i32); // src/perf_test.rs:54 (`/ 5`: an untyped literal)

pub mod prelude {
    pub use super::{CastFrom, CastInto, FheDecrypt, FheOrd, FheTryEncrypt};
}
/// `FheUint32::try_encrypt(value, &client_key)` (src/biguint.rs:26): encryption on the host, blocks handed to the GPU.
pub trait FheTryEncrypt<T, K>: Sized { fn try_encrypt(value: T, key: &K) -> Result<Self, Error>; }
/// `let v: u32 = ct.decrypt(&client_key)` (src/biguint.rs:70): download (synchronises) + decryption on the host.
pub trait FheDecrypt<T> { fn decrypt(&self, key: &ClientKey) -> T; }
pub trait CastFrom<T> { fn cast_from(x: T) -> Self; }
pub trait CastInto<T> { fn cast_into(self) -> T; }
pub trait FheOrd<Rhs = Self> { type Output; fn min(&self, rhs: Rhs) -> Self::Output; fn max(&self, rhs: Rhs) -> Self::Output; }

macro_rules! clear_types { ($($t:ty),+) => { $(
    impl<const B: usize> FheTryEncrypt<$t, ClientKey> for FheUint<B> {
        fn try_encrypt(value: $t, key: &ClientKey) -> Result<Self, Error> { Ok(FheUint::<B>::from_blocks(&key.encrypt(value as u64, B)?)) }
    }
    impl<const B: usize> FheDecrypt<$t> for FheUint<B> {
        fn decrypt(&self, key: &ClientKey) -> $t { key.decrypt(&self.to_blocks(), B) as $t }
    }
)+ }; }
clear_types!(u8, u16, u32, u64);
impl<const A: usize, const B: usize> CastFrom<FheUint<A>> for FheUint<B> { fn cast_from(x: FheUint<A>) -> Self { FheUint::<B>::cast(&x) } }
impl<const A: usize, const B: usize> CastInto<FheUint<B>> for FheUint<A> { fn cast_into(self) -> FheUint<B> { FheUint::<B>::cast(&self) } }
impl<const B: usize> FheOrd<&FheUint<B>> for FheUint<B> {
    type Output = FheUint<B>;
    fn min(&self, rhs: &FheUint<B>) -> FheUint<B> { self.binary(sys::OP_MIN, rhs) }
    fn max(&self, rhs: &FheUint<B>) -> FheUint<B> { self.binary(sys::OP_MAX, rhs) }
}
