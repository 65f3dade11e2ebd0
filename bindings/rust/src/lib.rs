//! `salzweg`-shaped API over the B200 codec (include/slzw.h).  Mirrors lzw/src/lib.rs:51-91,
//! lzw/src/encoder.rs:153-616 and lzw/src/decoder.rs:52-551 of redwarp/lzw: same type names, same
//! associated functions, same error enums, plus `encode_batch` / `decode_batch`.
//! NOT compiled in this repository's image (no Rust toolchain); see INTEGRATION.md.
use std::io::{self, Read, Write};

pub mod ffi;
use ffi::*;

/// lzw/src/lib.rs:59-65
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Endianness {
    BigEndian,
    LittleEndian,
}

/// lzw/src/lib.rs:71-91
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum CodeSizeStrategy {
    Default,
    Tiff,
}

/// lzw/src/encoder.rs:16-29
#[derive(Debug)]
pub enum EncodingError {
    Io(io::Error),
    CodeSize(u8),
    UnexpectedCode { code: u8, code_size: u8 },
}

/// lzw/src/decoder.rs:15-25
#[derive(Debug)]
pub enum DecodingError {
    Io(io::Error),
    CodeSize(u8),
    UnexpectedCode(u16),
    MissingClearCode,
}

thread_local! {
    // one context per thread: the reference's functions are stateless and re-entrant
    static CTX: *mut SlzwCtx = unsafe {
        let mut c = std::ptr::null_mut();
        assert_eq!(slzw_create(0, &mut c), 0, "no sm_100 device: this crate has no CPU path");
        c
    };
}

fn params(flavour: u8, cs: u8, e: Endianness, s: CodeSizeStrategy) -> SlzwParams {
    SlzwParams {
        flavour,
        code_size: cs,
        big_endian: (e == Endianness::BigEndian) as u8,
        tiff_early_change: (s == CodeSizeStrategy::Tiff) as u8,
    }
}

fn run_encode<R: Read, W: Write>(mut data: R, mut into: W, p: SlzwParams) -> Result<(), EncodingError> {
    let mut input = Vec::new();
    data.read_to_end(&mut input).map_err(EncodingError::Io)?;
    let cap = unsafe { slzw_encode_bound(&p, input.len() as u64) };
    let mut out = vec![0u8; cap as usize];
    let (mut len, mut detail) = (0u64, 0u32);
    let st = CTX.with(|c| unsafe {
        slzw_encode(*c, &p, input.as_ptr(), input.len() as u64, out.as_mut_ptr(), cap, &mut len, &mut detail)
    });
    // bytes produced before an error reach the writer, as in the reference
    into.write_all(&out[..len as usize]).map_err(EncodingError::Io)?;
    match st {
        0 => into.flush().map_err(EncodingError::Io),
        1 => Err(EncodingError::CodeSize(detail as u8)),
        2 => Err(EncodingError::UnexpectedCode { code: detail as u8, code_size: p.code_size }),
        5 => Err(EncodingError::Io(io::ErrorKind::WriteZero.into())),
        6 => panic!("index out of bounds (salzweg panics on this input, encoder.rs:99)"),
        _ => Err(EncodingError::Io(io::Error::new(io::ErrorKind::Other, "CUDA failure"))),
    }
}

pub struct VariableEncoder;
impl VariableEncoder {
    /// lzw/src/encoder.rs:199-220
    pub fn encode<R: Read, W: Write>(data: R, into: W, code_size: u8, endianness: Endianness,
                                     strategy: CodeSizeStrategy) -> Result<(), EncodingError> {
        run_encode(data, into, params(0, code_size, endianness, strategy))
    }
    /// lzw/src/encoder.rs:262-271
    pub fn encode_to_vec<R: Read>(data: R, code_size: u8, endianness: Endianness,
                                  strategy: CodeSizeStrategy) -> Result<Vec<u8>, EncodingError> {
        let mut v = Vec::new();
        Self::encode(data, &mut v, code_size, endianness, strategy)?;
        Ok(v)
    }
}

pub struct GifStyleEncoder;
impl GifStyleEncoder {
    /// lzw/src/encoder.rs:392-399
    pub fn encode<R: Read, W: Write>(data: R, into: W, code_size: u8) -> Result<(), EncodingError> {
        run_encode(data, into, params(0, code_size, Endianness::LittleEndian, CodeSizeStrategy::Default))
    }
    /// lzw/src/encoder.rs:435-439
    pub fn encode_to_vec<R: Read>(data: R, code_size: u8) -> Result<Vec<u8>, EncodingError> {
        let mut v = Vec::new();
        Self::encode(data, &mut v, code_size)?;
        Ok(v)
    }
}

pub struct TiffStyleEncoder;
impl TiffStyleEncoder {
    /// lzw/src/encoder.rs:479-487
    pub fn encode<R: Read, W: Write>(data: R, into: W) -> Result<(), EncodingError> {
        run_encode(data, into, params(0, 8, Endianness::BigEndian, CodeSizeStrategy::Tiff))
    }
    /// lzw/src/encoder.rs:519-523
    pub fn encode_to_vec<R: Read>(data: R) -> Result<Vec<u8>, EncodingError> {
        let mut v = Vec::new();
        Self::encode(data, &mut v)?;
        Ok(v)
    }
}

pub struct FixedEncoder;
impl FixedEncoder {
    /// lzw/src/encoder.rs:565-576
    pub fn encode<R: Read, W: Write>(data: R, into: W, endianness: Endianness) -> Result<(), EncodingError> {
        run_encode(data, into, params(1, 0, endianness, CodeSizeStrategy::Default))
    }
    /// lzw/src/encoder.rs:609-616
    pub fn encode_to_vec<R: Read>(data: R, endianness: Endianness) -> Result<Vec<u8>, EncodingError> {
        let mut v = Vec::new();
        Self::encode(data, &mut v, endianness)?;
        Ok(v)
    }
}

// The four decoder types (lzw/src/decoder.rs:99-551) follow the same pattern over slzw_decode:
// a size-only call (out = NULL) sizes the Vec, status 2 -> UnexpectedCode(detail as u16),
// 3 -> MissingClearCode, 4 -> Io(UnexpectedEof), 5 -> Io(WriteZero).

/// New relative to salzweg: many independent streams (TIFF strips, GIF frames, text chunks) in one
/// call.  `offsets` has n + 1 entries into `data`; returns one result per stream.
pub fn encode_batch(data: &[u8], offsets: &[u64], p: SlzwParams) -> Vec<Result<Vec<u8>, EncodingError>> {
    let n = offsets.len() - 1;
    let mut out_off = vec![0u64; n + 1];
    for i in 0..n {
        let b = unsafe { slzw_encode_bound(&p, offsets[i + 1] - offsets[i]) };
        out_off[i + 1] = out_off[i] + ((b + 15) & !15);
    }
    let mut out = vec![0u8; out_off[n] as usize];
    let (mut len, mut st, mut det) = (vec![0u64; n], vec![0u32; n], vec![0u32; n]);
    let b = SlzwBatch {
        input: data.as_ptr(), in_off: offsets.as_ptr(), out: out.as_mut_ptr(), out_off: out_off.as_ptr(),
        out_len: len.as_mut_ptr(), status: st.as_mut_ptr(), detail: det.as_mut_ptr(),
        code_size: std::ptr::null(), n: n as u64,
    };
    let rc = CTX.with(|c| unsafe { slzw_encode_batch_host(*c, &p, &b) });
    assert_eq!(rc, 0, "slzw_encode_batch_host failed");
    (0..n).map(|i| match st[i] {
        0 => Ok(out[out_off[i] as usize..(out_off[i] + len[i]) as usize].to_vec()),
        1 => Err(EncodingError::CodeSize(det[i] as u8)),
        2 => Err(EncodingError::UnexpectedCode { code: det[i] as u8, code_size: p.code_size }),
        _ => Err(EncodingError::Io(io::ErrorKind::WriteZero.into())),
    }).collect()
}
