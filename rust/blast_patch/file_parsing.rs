//! Replacement bodies for blast/src/file_parsing/{wav,aiff,mpeg}.rs::parse on top of blast-cuda-sys.
//! Signatures, AudioFile and DecodeError are unchanged (decode_helpers.rs:1-38).  NOT compiled in the build
//! image (no rustc); shown so a maintainer can see exactly what the drop-in looks like.
use std::ffi::{CStr, CString};
use std::sync::OnceLock;

use blast_cuda_sys as sys;
use super::decode_helpers::{AudioFile, DecodeError, DecodeResult};

struct Gpu(*mut sys::blast_ctx);
unsafe impl Send for Gpu {}
unsafe impl Sync for Gpu {}
static GPU: OnceLock<Gpu> = OnceLock::new();

/// The process-wide GPU context (the reference is single-threaded on both decode, main.rs:18-89, and render,
/// runtime.rs:320-380: one blast_ctx serves both).  Used by engine.rs's Conductor::prepare.
pub fn gpu_ctx() -> *mut sys::blast_ctx { ctx() }

fn ctx() -> *mut sys::blast_ctx {
    GPU.get_or_init(|| {
        let mut c = std::ptr::null_mut();
        let rc = unsafe { sys::blast_ctx_create(&mut c, 0) };
        assert_eq!(rc, sys::BLAST_OK, "{}", last_error());   // there is no CPU fallback
        Gpu(c)
    }).0
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::blast_last_error()) }.to_string_lossy().into_owned()
}

fn to_err(rc: i32) -> DecodeError {
    match rc {
        sys::BLAST_ERR_UNSUPPORTED_FORMAT => DecodeError::UnsupportedFormat(last_error()),
        sys::BLAST_ERR_UNEXPECTED_EOF => DecodeError::UnexpectedEof,
        sys::BLAST_ERR_INVALID_DATA => DecodeError::InvalidData(last_error()),
        sys::BLAST_ERR_REF_PANIC => panic!("{}", last_error()),          // the reference panics here too
        _ => DecodeError::Io(std::io::Error::new(std::io::ErrorKind::Other, last_error())),
    }
}

fn parse_pcm(path: &str, format: &str,
             probe: unsafe extern "C" fn(*const u8, usize, *mut sys::blast_pcm_desc) -> i32) -> DecodeResult<AudioFile> {
    let reader = std::fs::read(path)?;                                     // File::open + read_to_end
    let mut desc = sys::blast_pcm_desc::default();
    let rc = unsafe { probe(reader.as_ptr(), reader.len(), &mut desc) };   // header walk, wav.rs:69-138 / aiff.rs:99-154
    if rc != sys::BLAST_OK { return Err(to_err(rc)); }
    let n = unsafe { sys::blast_pcm_out_len(&desc) };
    let mut samples: Vec<i16> = Vec::with_capacity(n);
    let (file, len, out) = (reader.as_ptr(), reader.len(), samples.as_mut_ptr());
    let rc = unsafe { sys::blast_pcm_decode_batch(ctx(), 1, &file, &len, &desc, &out, std::ptr::null()) };
    if rc != sys::BLAST_OK { return Err(to_err(rc)); }
    unsafe { samples.set_len(n) };
    // the name rule runs after decoding, as in the reference (wav.rs:156-164)
    let c_path = CString::new(path).map_err(|_| DecodeError::InvalidData("File has no name".to_string()))?;
    let mut name = vec![0u8; path.len() + 1];
    let rc = unsafe { sys::blast_file_name(c_path.as_ptr(), name.as_mut_ptr() as *mut _, name.len()) };
    if rc != sys::BLAST_OK { return Err(to_err(rc)); }
    let name = CStr::from_bytes_until_nul(&name).unwrap().to_string_lossy();
    Ok(AudioFile::new(&name, format, desc.sample_rate, desc.num_channels, desc.bits_per_sample, samples))
}

/// AudioFile.samples resident in HBM: what Conductor::prepare hands to blast_conductor_create instead of the per-voice
/// clone of the sample Vec (engine.rs:309).  Owns its device allocation.
pub struct DeviceTrack { d_samples: *mut i16, n_samples: u64, num_channels: u32, sample_rate: u32 }

impl DeviceTrack {
    pub fn upload(ctx: *mut sys::blast_ctx, t: &AudioFile) -> Self {
        let bytes = t.samples.len() * std::mem::size_of::<i16>();
        let mut p: *mut std::ffi::c_void = std::ptr::null_mut();
        let rc = unsafe { sys::blast_dev_alloc(ctx, bytes.max(4), &mut p) };
        assert_eq!(rc, sys::BLAST_OK, "{}", last_error());
        let rc = unsafe { sys::blast_memcpy_h2d(ctx, p, t.samples.as_ptr() as *const _, bytes) };
        assert_eq!(rc, sys::BLAST_OK, "{}", last_error());
        assert_eq!(unsafe { sys::blast_ctx_sync(ctx) }, sys::BLAST_OK, "{}", last_error());
        Self { d_samples: p as *mut i16, n_samples: t.samples.len() as u64, num_channels: t.num_channels, sample_rate: t.sample_rate }
    }
    pub fn as_c(&self) -> sys::blast_track {
        sys::blast_track { d_samples: self.d_samples, n_samples: self.n_samples, num_channels: self.num_channels, sample_rate: self.sample_rate }
    }
}

impl Drop for DeviceTrack {
    fn drop(&mut self) { unsafe { sys::blast_dev_free(ctx(), self.d_samples as *mut _); } }
}

pub mod wav  { pub fn parse(path: &str) -> super::DecodeResult<super::AudioFile> { super::parse_pcm(path, "wav",  super::sys::blast_wav_probe) } }
pub mod aiff { pub fn parse(path: &str) -> super::DecodeResult<super::AudioFile> { super::parse_pcm(path, "aiff", super::sys::blast_aiff_probe) } }

pub mod mpeg {
    use super::*;
    pub fn parse(path: &str) -> DecodeResult<Vec<u8>> {
        let reader = std::fs::read(path)?;
        let (mut n_off, mut n_cand, mut plen, mut refh) = (0u64, 0u64, 0u64, 0u32);
        let rc = unsafe { sys::blast_mpeg_parse(ctx(), reader.as_ptr(), reader.len() as u64, 1, std::ptr::null_mut(), 0,
                                                &mut n_off, &mut refh, &mut n_cand, std::ptr::null_mut(), 0, &mut plen) };
        if rc != sys::BLAST_OK { return Err(to_err(rc)); }
        let mut data = vec![0u8; plen as usize];
        let rc = unsafe { sys::blast_mpeg_parse(ctx(), reader.as_ptr(), reader.len() as u64, 1, std::ptr::null_mut(), 0,
                                                &mut n_off, &mut refh, &mut n_cand, data.as_mut_ptr(), plen, &mut plen) };
        if rc != sys::BLAST_OK { return Err(to_err(rc)); }
        Ok(data)
    }
}
