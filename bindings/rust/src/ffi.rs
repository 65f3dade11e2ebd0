//! Raw declarations of include/slzw.h -- every symbol libslzw.so exports, one to one
//! (tests/test_rust_binding.py diffs names and arities against the header).
#![allow(dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct SlzwParams {
    pub flavour: u8,
    pub code_size: u8,
    pub big_endian: u8,
    pub tiff_early_change: u8,
}

#[repr(C)]
pub struct SlzwBatch {
    pub input: *const u8,
    pub in_off: *const u64,
    pub out: *mut u8,
    pub out_off: *const u64,
    pub out_len: *mut u64,
    pub status: *mut u32,
    pub detail: *mut u32,
    pub code_size: *const u8,
    pub n: u64,
}

#[repr(C)]
pub struct SlzwCtx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct SlzwMulti {
    _private: [u8; 0],
}

pub const SLZW_FLAVOUR_VARIABLE: u8 = 0;
pub const SLZW_FLAVOUR_FIXED: u8 = 1;
pub const SLZW_FLAVOUR_VARIABLE_LENIENT: u8 = 2;
pub const SLZW_PREDICTOR_DIFFERENCE: c_int = 0;
pub const SLZW_PREDICTOR_ACCUMULATE: c_int = 1;

pub const SLZW_OK: u32 = 0;
pub const SLZW_ERR_CODE_SIZE: u32 = 1;
pub const SLZW_ERR_UNEXPECTED_CODE: u32 = 2;
pub const SLZW_ERR_MISSING_CLEAR_CODE: u32 = 3;
pub const SLZW_ERR_IO_UNEXPECTED_EOF: u32 = 4;
pub const SLZW_ERR_IO_WRITE_ZERO: u32 = 5;
pub const SLZW_ERR_REFERENCE_PANIC: u32 = 6;

pub const SLZW_RC_OK: c_int = 0;
pub const SLZW_RC_CUDA: c_int = -1;
pub const SLZW_RC_INVALID: c_int = -2;
pub const SLZW_RC_NO_DEVICE: c_int = -3;
pub const SLZW_RC_NOMEM: c_int = -4;

extern "C" {
    pub fn slzw_create(device: c_int, ctx: *mut *mut SlzwCtx) -> c_int;
    pub fn slzw_destroy(ctx: *mut SlzwCtx);
    pub fn slzw_last_error(ctx: *const SlzwCtx) -> *const c_char;
    pub fn slzw_kernel_launches(ctx: *const SlzwCtx) -> u64;
    pub fn slzw_version() -> u32;
    pub fn slzw_last_deferred(ctx: *mut SlzwCtx, ids: *mut u32, cap: u64) -> u64;
    pub fn slzw_last_encode_shares(ctx: *mut SlzwCtx, bytes: *mut u64) -> c_int;
    pub fn slzw_encode_batch_device(ctx: *mut SlzwCtx, params: *const SlzwParams, batch: *const SlzwBatch, cuda_stream: *mut c_void) -> c_int;
    pub fn slzw_decode_batch_device(ctx: *mut SlzwCtx, params: *const SlzwParams, batch: *const SlzwBatch, cuda_stream: *mut c_void) -> c_int;
    pub fn slzw_encode_batch_host(ctx: *mut SlzwCtx, params: *const SlzwParams, batch: *const SlzwBatch) -> c_int;
    pub fn slzw_decode_batch_host(ctx: *mut SlzwCtx, params: *const SlzwParams, batch: *const SlzwBatch) -> c_int;
    pub fn slzw_encode_batch_host_dense(ctx: *mut SlzwCtx, params: *const SlzwParams, input: *const u8, in_off: *const u64, n: u64, code_size: *const u8, align: u64, out_dense: *mut u8, out_cap: u64, out_off: *mut u64, status: *mut u32, detail: *mut u32, needed: *mut u64) -> c_int;
    pub fn slzw_encode_batch_host_dense_begin(ctx: *mut SlzwCtx, params: *const SlzwParams, input: *const u8, in_off: *const u64, n: u64, code_size: *const u8, align: u64, out_off: *mut u64, status: *mut u32, detail: *mut u32, total: *mut u64) -> c_int;
    pub fn slzw_encode_batch_host_dense_finish(ctx: *mut SlzwCtx, out_dense: *mut u8, out_cap: u64) -> c_int;
    pub fn slzw_encode(ctx: *mut SlzwCtx, params: *const SlzwParams, input: *const u8, n: u64, out: *mut u8, cap: u64, out_len: *mut u64, detail: *mut u32) -> c_int;
    pub fn slzw_decode(ctx: *mut SlzwCtx, params: *const SlzwParams, input: *const u8, n: u64, out: *mut u8, cap: u64, out_len: *mut u64, detail: *mut u32) -> c_int;
    pub fn slzw_encode_bound(params: *const SlzwParams, n: u64) -> u64;
    pub fn slzw_decoded_sizes_batch_device(ctx: *mut SlzwCtx, params: *const SlzwParams, batch: *const SlzwBatch, cuda_stream: *mut c_void) -> c_int;
    pub fn slzw_decoded_sizes_batch_host(ctx: *mut SlzwCtx, params: *const SlzwParams, batch: *const SlzwBatch) -> c_int;
    pub fn slzw_compact_device(ctx: *mut SlzwCtx, src: *const u8, src_off: *const u64, len: *const u64, n: u64, align: u64, dst: *mut u8, dst_off: *mut u64, cuda_stream: *mut c_void) -> c_int;
    pub fn slzw_tiff_predictor_device(ctx: *mut SlzwCtx, direction: c_int, data: *mut u8, off: *const u64, len: *const u64, n: u64, row_bytes: u32, samples_per_pixel: u32, cuda_stream: *mut c_void) -> c_int;
    pub fn slzw_set_tiff_predictor(ctx: *mut SlzwCtx, row_bytes: u32, samples_per_pixel: u32) -> c_int;
    pub fn slzw_host_alloc(bytes: usize) -> *mut c_void;
    pub fn slzw_host_free(p: *mut c_void);
    pub fn slzw_status_message(is_decoder: c_int, status: u32, detail: u32, code_size: u8, buf: *mut c_char, buf_len: usize) -> c_int;
    pub fn slzw_multi_create(devices: *const c_int, n_devices: c_int, out: *mut *mut SlzwMulti) -> c_int;
    pub fn slzw_multi_destroy(m: *mut SlzwMulti);
    pub fn slzw_multi_device_count(m: *const SlzwMulti) -> c_int;
    pub fn slzw_multi_last_error(m: *const SlzwMulti) -> *const c_char;
    pub fn slzw_multi_kernel_launches(m: *const SlzwMulti) -> u64;
    pub fn slzw_partition_streams(weight_off: *const u64, n: u64, parts: c_int, bounds: *mut u64);
    pub fn slzw_multi_encode_batch_host(m: *mut SlzwMulti, params: *const SlzwParams, batch: *const SlzwBatch) -> c_int;
    pub fn slzw_multi_decode_batch_host(m: *mut SlzwMulti, params: *const SlzwParams, batch: *const SlzwBatch) -> c_int;
    pub fn slzw_multi_encode_batch_host_dense(m: *mut SlzwMulti, params: *const SlzwParams, input: *const u8, in_off: *const u64, n: u64, code_size: *const u8, align: u64, out_dense: *mut u8, out_cap: u64, out_off: *mut u64, status: *mut u32, detail: *mut u32, needed: *mut u64) -> c_int;
}
