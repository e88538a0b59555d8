//! Replacement body of blast/src/audio_processing/engine.rs's Conductor on top of blast-cuda-sys.
//! `prepare`, `apply` and `coordinate` keep their signatures (engine.rs:36, 83, 46); the Voice / Group / TempoState /
//! Seq object graph lives behind the `blast_conductor` handle and every sample, position step, Seq hit and RNG
//! draw is computed on the GPU.  NOT compiled in the build image (no rustc).
use std::collections::HashMap;

use alsa_sys::*;
use blast_cuda_sys as sys;

use crate::audio_processing::commands::*;
use crate::audio_processing::blast_time::{sample_rate, blast_time::{TempoMode, TempoUnit}};
use crate::file_parsing::decode_helpers::AudioFile;
use crate::file_parsing::DeviceTrack;       // defined in the file_parsing patch (file_parsing.rs in this directory)

pub struct Conductor {
    ctx: *mut sys::blast_ctx,
    h: *mut sys::blast_conductor,
    out_channels: usize,
    bus: Vec<i16>,                      // one period, interleaved S16 (runtime.rs:272-276)
    _tracks: Vec<DeviceTrack>,          // decoded samples resident in HBM (see file_parsing.rs: d_out)
}

fn tempo(tr: &TempoRepr) -> sys::blast_tempo_repr {
    sys::blast_tempo_repr {
        idx: tr.idx as u64,
        owned: tr.owned as u32,
        mode: match tr.mode { TempoMode::Process => 0, TempoMode::Voice => 1, TempoMode::Group => 2,
                              TempoMode::Context => 3, TempoMode::TBD => 4 },
        unit: match tr.unit { TempoUnit::Samples => 0, TempoUnit::Millis => 1, TempoUnit::Bpm => 2 },
        interval: tr.interval,
    }
}

fn idx(i: &Idx) -> (u32, u64) {
    match i { Idx::Tempo(k) => (0, *k as u64), Idx::Voice(k) => (1, *k as u64), Idx::Process(k) => (2, *k as u64),
              Idx::Group(k) => (3, *k as u64) }
}

impl Conductor {
    pub fn prepare(out_channels: usize, tracks: HashMap<String, AudioFile>) -> Self {
        let ctx = crate::file_parsing::gpu_ctx();
        let dev: Vec<DeviceTrack> = tracks.into_values().map(|t| DeviceTrack::upload(ctx, &t)).collect();
        let c: Vec<sys::blast_track> = dev.iter().map(|t| t.as_c()).collect();
        let mut h = std::ptr::null_mut();
        let rc = unsafe { sys::blast_conductor_create(ctx, out_channels as u32, sample_rate::get(), c.as_ptr(),
                                                      c.len() as u32, &mut h) };
        assert_eq!(rc, sys::BLAST_OK);
        Self { ctx, h, out_channels, bus: Vec::new(), _tracks: dev }
    }

    pub fn apply(&mut self, cmd: Command) {
        let mut c: sys::blast_command = unsafe { std::mem::zeroed() };
        // keep-alive storage for the pointer fields
        let (mut mv, mut mu, mut mn, mut mp): (Vec<u64>, Vec<u8>, Vec<u32>, Vec<u64>) = Default::default();
        let (steps, chance);
        match cmd {
            Command::Load(a) => { c.kind = sys::BLAST_CMD_LOAD; c.idx = a.track_idx as u64; c.tempo = tempo(&a.tempo_repr); }
            Command::Start(a) => { c.kind = sys::BLAST_CMD_START; (c.idx_kind, c.idx) = idx(&a.idx); }
            Command::Pause(a) => { c.kind = sys::BLAST_CMD_PAUSE; (c.idx_kind, c.idx) = idx(&a.idx); }
            Command::Resume(a) => { c.kind = sys::BLAST_CMD_RESUME; (c.idx_kind, c.idx) = idx(&a.idx); }
            Command::Stop(a) => { c.kind = sys::BLAST_CMD_STOP; (c.idx_kind, c.idx) = idx(&a.idx); }
            Command::Unload(a) => { c.kind = sys::BLAST_CMD_UNLOAD; c.idx = a.idx as u64; }
            Command::Velocity(a) => { c.kind = sys::BLAST_CMD_VELOCITY; c.idx = a.idx as u64; c.val = a.val; }
            Command::Group(a) => {
                c.kind = sys::BLAST_CMD_GROUP; c.tempo = tempo(&a.tempo);
                for (v, f, ps) in &a.vs_fs_ps {
                    mv.push(*v as u64); mu.push(*f as u8); mn.push(ps.len() as u32);
                    mp.extend(ps.iter().map(|p| *p as u64));
                }
                c.n_members = mv.len() as u32;
                c.member_voice = mv.as_ptr(); c.member_update_tempo = mu.as_ptr();
                c.member_n_procs = mn.as_ptr(); c.member_proc_ids = mp.as_ptr();
            }
            Command::Tc(a) => { c.kind = sys::BLAST_CMD_TC; c.tempo = tempo(&a.tempo); }
            Command::Seq(a) => {
                c.kind = sys::BLAST_CMD_SEQ; (c.idx_kind, c.idx) = idx(&a.idx); c.tempo = tempo(&a.tempo);
                c.period = a.period as u64; c.n_steps = a.steps.len() as u32;
                steps = a.steps; chance = a.chance;
                c.steps = steps.as_ptr(); c.chance = chance.as_ptr();
                // X128P's fields are private (blast_rand.rs:4-8): the patch adds the accessor in blast_rand.rs (this directory)
                (c.rng_s0, c.rng_s1) = a.rng.state();
            }
            Command::Quit(_) => { unsafe { libc::raise(libc::SIGTERM); } return; }
        }
        let rc = unsafe { sys::blast_conductor_apply(self.ctx, self.h, &c) };
        if rc == sys::BLAST_ERR_REF_PANIC { panic!("{}", crate::file_parsing::last_error()); }   // `.unwrap()` on a bad index
        assert_eq!(rc, sys::BLAST_OK);
    }

    /// One ALSA period: render `frames` frames on the GPU, then copy the interleaved S16 bus into the mmap areas
    /// (the ALSA write itself stays host-side I/O).
    pub fn coordinate(&mut self, areas_ptr: *const snd_pcm_channel_area_t, offset: snd_pcm_uframes_t, frames: snd_pcm_uframes_t) {
        self.bus.resize(frames as usize * self.out_channels, 0);
        let rc = unsafe { sys::blast_conductor_coordinate(self.ctx, self.h, frames as u64, self.bus.as_mut_ptr()) };
        assert_eq!(rc, sys::BLAST_OK);
        unsafe {
            let areas = std::slice::from_raw_parts(areas_ptr, self.out_channels);
            for f in 0..frames as usize {
                for ch in 0..self.out_channels {
                    let a = &areas[ch];
                    let bit = a.first as isize + (offset as usize + f) as isize * a.step as isize;   // engine.rs:56-59
                    *((a.addr as *mut u8).offset(bit / 8) as *mut i16) = self.bus[f * self.out_channels + ch];
                }
            }
        }
    }
}

impl Drop for Conductor {
    fn drop(&mut self) { unsafe { sys::blast_conductor_destroy(self.ctx, self.h) } }
}
