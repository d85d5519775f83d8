//! `extern "C"` declarations of include/brotli_b200.h -- the same symbols the reference exports from
//! src/ffi/compressor.rs and src/ffi/multicompress/mod.rs, plus the device-resident additions.
#![allow(non_camel_case_types, non_snake_case)]
use std::os::raw::c_void;

#[repr(C)]
pub struct BrotliEncoderState {
    _private: [u8; 0],
}
#[repr(C)]
pub struct BrotliEncoderWorkPool {
    _private: [u8; 0],
}
#[repr(C)]
pub struct B200Encoder {
    _private: [u8; 0],
}

pub type brotli_alloc_func = Option<unsafe extern "C" fn(opaque: *mut c_void, size: usize) -> *mut c_void>;
pub type brotli_free_func = Option<unsafe extern "C" fn(opaque: *mut c_void, address: *mut c_void)>;

/// src/enc/encode.rs:1380-1385
#[repr(C)]
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum BrotliEncoderOperation {
    BROTLI_OPERATION_PROCESS = 0,
    BROTLI_OPERATION_FLUSH = 1,
    BROTLI_OPERATION_FINISH = 2,
    BROTLI_OPERATION_EMIT_METADATA = 3,
}

/// Numeric values of src/enc/parameters.rs:1-32 (passed as plain u32 keys).
pub mod param {
    pub const MODE: u32 = 0;
    pub const QUALITY: u32 = 1;
    pub const LGWIN: u32 = 2;
    pub const LGBLOCK: u32 = 3;
    pub const DISABLE_LITERAL_CONTEXT_MODELING: u32 = 4;
    pub const SIZE_HINT: u32 = 5;
    pub const LARGE_WINDOW: u32 = 6;
    pub const CATABLE: u32 = 167;
    pub const APPENDABLE: u32 = 168;
    pub const MAGIC_NUMBER: u32 = 169;
    pub const NO_DICTIONARY: u32 = 170;
    pub const BYTE_ALIGN: u32 = 172;
    pub const BARE_STREAM: u32 = 173;
}

extern "C" {
    // ---- src/ffi/compressor.rs ----
    pub fn BrotliEncoderCreateInstance(alloc: brotli_alloc_func, free: brotli_free_func, opaque: *mut c_void) -> *mut BrotliEncoderState;
    pub fn BrotliEncoderSetParameter(state: *mut BrotliEncoderState, p: u32, value: u32) -> i32;
    pub fn BrotliEncoderDestroyInstance(state: *mut BrotliEncoderState);
    pub fn BrotliEncoderIsFinished(state: *mut BrotliEncoderState) -> i32;
    pub fn BrotliEncoderHasMoreOutput(state: *mut BrotliEncoderState) -> i32;
    pub fn BrotliEncoderSetCustomDictionary(state: *mut BrotliEncoderState, size: usize, dict: *const u8);
    pub fn BrotliEncoderTakeOutput(state: *mut BrotliEncoderState, size: *mut usize) -> *const u8;
    pub fn BrotliEncoderVersion() -> u32;
    pub fn BrotliEncoderMaxCompressedSize(input_size: usize) -> usize;
    pub fn BrotliEncoderCompress(quality: i32, lgwin: i32, mode: i32, input_size: usize, input: *const u8,
                                 encoded_size: *mut usize, encoded: *mut u8) -> i32;
    pub fn BrotliEncoderCompressStreaming(state: *mut BrotliEncoderState, op: BrotliEncoderOperation, available_in: *mut usize,
                                          input: *const u8, available_out: *mut usize, output: *mut u8) -> i32;
    pub fn BrotliEncoderCompressStream(state: *mut BrotliEncoderState, op: BrotliEncoderOperation, available_in: *mut usize,
                                       next_in: *mut *const u8, available_out: *mut usize, next_out: *mut *mut u8,
                                       total_out: *mut usize) -> i32;
    pub fn BrotliEncoderMallocU8(state: *mut BrotliEncoderState, size: usize) -> *mut u8;
    pub fn BrotliEncoderFreeU8(state: *mut BrotliEncoderState, data: *mut u8, size: usize);
    pub fn BrotliEncoderMallocUsize(state: *mut BrotliEncoderState, size: usize) -> *mut usize;
    pub fn BrotliEncoderFreeUsize(state: *mut BrotliEncoderState, data: *mut usize, size: usize);
    // ---- src/ffi/multicompress/mod.rs ----
    pub fn BrotliEncoderMaxCompressedSizeMulti(input_size: usize, num_threads: usize) -> usize;
    pub fn BrotliEncoderCompressMulti(num_params: usize, keys: *const u32, values: *const u32, input_size: usize,
                                      input: *const u8, encoded_size: *mut usize, encoded: *mut u8, desired_num_threads: usize,
                                      alloc: brotli_alloc_func, free: brotli_free_func, opaque_per_thread: *mut *mut c_void) -> i32;
    pub fn BrotliEncoderCreateWorkPool(num_workers: usize, alloc: brotli_alloc_func, free: brotli_free_func,
                                       opaque_per_thread: *mut *mut c_void) -> *mut BrotliEncoderWorkPool;
    pub fn BrotliEncoderDestroyWorkPool(pool: *mut BrotliEncoderWorkPool);
    pub fn BrotliEncoderCompressWorkPool(pool: *mut BrotliEncoderWorkPool, num_params: usize, keys: *const u32, values: *const u32,
                                         input_size: usize, input: *const u8, encoded_size: *mut usize, encoded: *mut u8,
                                         desired_num_threads: usize, alloc: brotli_alloc_func, free: brotli_free_func,
                                         opaque_per_thread: *mut *mut c_void) -> i32;
    // ---- device-resident additions ----
    pub fn b200_device_count() -> i32;
    pub fn b200_effective_quality(requested_quality: i32) -> i32;
    pub fn b200_encoder_create(device: i32) -> *mut B200Encoder;
    pub fn b200_encoder_destroy(e: *mut B200Encoder);
    pub fn b200_max_compressed_size(n: usize) -> usize;
    pub fn b200_encoder_compress_range(e: *mut B200Encoder, quality: i32, lgwin: i32, size_hint: u64, input: *const u8, n: usize,
                                       range_start: usize, range_len: usize, first: i32, last: i32, byte_align: i32, out: *mut u8,
                                       out_cap: usize, out_size: *mut usize, device_io: i32) -> i32;
}
