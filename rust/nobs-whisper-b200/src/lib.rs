//! The whisper-rs surface used by `src-tauri/src/whisper.rs` (lines 3, 39-45, 83-141), re-implemented
//! over the B200 library.  Names, argument meaning and error behaviour follow whisper-rs 0.15 so that
//! the reference compiles against this crate unchanged (Cargo.toml: replace the `whisper-rs`
//! dependency by `whisper-rs = { package = "nobs-whisper-b200", path = ".../rust/nobs-whisper-b200" }`).
use nobs_whisper_b200_sys as sys;
use std::borrow::Cow;
use std::ffi::{CStr, CString};
use std::fmt;
use std::os::raw::c_int;
use std::sync::Arc;

#[derive(Debug)]
pub enum WhisperError {
    InitError,
    NoSamples,
    NullByteInString,
    UnableToCalculateSpectrogram,
    FailedToEncode,
    FailedToDecode,
    GenericError(c_int),
}
impl fmt::Display for WhisperError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        let detail = unsafe { CStr::from_ptr(sys::whisper_b200_last_error()) }.to_string_lossy();
        write!(f, "{:?}: {}", self, detail)
    }
}
impl std::error::Error for WhisperError {}

#[derive(Clone, Copy)]
pub struct WhisperContextParameters {
    inner: sys::whisper_context_params,
}
impl Default for WhisperContextParameters {
    fn default() -> Self {
        Self { inner: unsafe { sys::whisper_context_default_params() } }
    }
}
impl WhisperContextParameters {
    pub fn use_gpu(&mut self, v: bool) -> &mut Self {
        self.inner.use_gpu = v; // accepted and ignored: the library has only the GPU path
        self
    }
    pub fn gpu_device(&mut self, d: c_int) -> &mut Self {
        self.inner.gpu_device = d;
        self
    }
}

struct CtxPtr(*mut sys::whisper_context);
unsafe impl Send for CtxPtr {}
unsafe impl Sync for CtxPtr {} // the C side serialises GPU work per context with a mutex
impl Drop for CtxPtr {
    fn drop(&mut self) {
        unsafe { sys::whisper_free(self.0) }
    }
}

#[derive(Clone)]
pub struct WhisperContext {
    ctx: Arc<CtxPtr>,
}
impl WhisperContext {
    pub fn new_with_params(path: &str, params: WhisperContextParameters) -> Result<Self, WhisperError> {
        let c = CString::new(path).map_err(|_| WhisperError::NullByteInString)?;
        let p = unsafe { sys::whisper_init_from_file_with_params_no_state(c.as_ptr(), params.inner) };
        if p.is_null() { Err(WhisperError::InitError) } else { Ok(Self { ctx: Arc::new(CtxPtr(p)) }) }
    }
    pub fn create_state(&self) -> Result<WhisperState, WhisperError> {
        let s = unsafe { sys::whisper_init_state(self.ctx.0) };
        if s.is_null() { Err(WhisperError::InitError) } else { Ok(WhisperState { ctx: self.ctx.clone(), ptr: s }) }
    }
}

pub enum SamplingStrategy {
    Greedy { best_of: c_int },
    BeamSearch { beam_size: c_int, patience: f32 },
}

pub struct FullParams<'a, 'b> {
    fp: sys::whisper_full_params,
    language: Option<CString>,
    initial_prompt: Option<CString>,
    _phantom: std::marker::PhantomData<(&'a (), &'b ())>,
}
impl<'a, 'b> FullParams<'a, 'b> {
    pub fn new(strategy: SamplingStrategy) -> Self {
        let mut fp = unsafe {
            sys::whisper_full_default_params(match strategy {
                SamplingStrategy::Greedy { .. } => sys::WHISPER_SAMPLING_GREEDY,
                SamplingStrategy::BeamSearch { .. } => sys::WHISPER_SAMPLING_BEAM_SEARCH,
            })
        };
        match strategy {
            SamplingStrategy::Greedy { best_of } => fp.greedy.best_of = best_of,
            SamplingStrategy::BeamSearch { beam_size, patience } => {
                fp.beam_search.beam_size = beam_size;
                fp.beam_search.patience = patience;
            }
        }
        Self { fp, language: None, initial_prompt: None, _phantom: std::marker::PhantomData }
    }
    pub fn set_language(&mut self, lang: Option<&str>) {
        self.language = lang.and_then(|l| CString::new(l).ok());
        self.fp.language = self.language.as_ref().map_or(std::ptr::null(), |c| c.as_ptr());
    }
    pub fn set_initial_prompt(&mut self, prompt: &str) {
        self.initial_prompt = CString::new(prompt).ok();
        self.fp.initial_prompt = self.initial_prompt.as_ref().map_or(std::ptr::null(), |c| c.as_ptr());
    }
    pub fn set_print_special(&mut self, v: bool) { self.fp.print_special = v }
    pub fn set_print_progress(&mut self, v: bool) { self.fp.print_progress = v }
    pub fn set_print_realtime(&mut self, v: bool) { self.fp.print_realtime = v }
    pub fn set_print_timestamps(&mut self, v: bool) { self.fp.print_timestamps = v }
    pub fn set_translate(&mut self, v: bool) { self.fp.translate = v }
    pub fn set_no_context(&mut self, v: bool) { self.fp.no_context = v }
    pub fn set_single_segment(&mut self, v: bool) { self.fp.single_segment = v }
    pub fn set_suppress_blank(&mut self, v: bool) { self.fp.suppress_blank = v }
    pub fn set_no_speech_thold(&mut self, v: f32) { self.fp.no_speech_thold = v }
    pub fn set_entropy_thold(&mut self, v: f32) { self.fp.entropy_thold = v }
    pub fn set_logprob_thold(&mut self, v: f32) { self.fp.logprob_thold = v }
}

pub struct WhisperState {
    ctx: Arc<CtxPtr>,
    ptr: *mut sys::whisper_state,
}
unsafe impl Send for WhisperState {}
impl Drop for WhisperState {
    fn drop(&mut self) {
        unsafe { sys::whisper_free_state(self.ptr) }
    }
}
impl WhisperState {
    pub fn full(&mut self, params: FullParams, data: &[f32]) -> Result<c_int, WhisperError> {
        if data.is_empty() {
            return Err(WhisperError::NoSamples);
        }
        let ret = unsafe { sys::whisper_full_with_state(self.ctx.0, self.ptr, params.fp, data.as_ptr(), data.len() as c_int) };
        match ret {
            0 => Ok(ret),
            -1 => Err(WhisperError::UnableToCalculateSpectrogram),
            7 => Err(WhisperError::FailedToEncode),
            8 => Err(WhisperError::FailedToDecode),
            e => Err(WhisperError::GenericError(e)),
        }
    }
    pub fn full_n_segments(&self) -> c_int {
        unsafe { sys::whisper_full_n_segments_from_state(self.ptr) }
    }
    pub fn get_segment(&self, i: c_int) -> Option<WhisperSegment<'_>> {
        if i >= 0 && i < self.full_n_segments() { Some(WhisperSegment { state: self, idx: i }) } else { None }
    }
}

pub struct WhisperSegment<'a> {
    state: &'a WhisperState,
    idx: c_int,
}
impl<'a> WhisperSegment<'a> {
    pub fn to_str_lossy(&self) -> Result<Cow<'a, str>, WhisperError> {
        let p = unsafe { sys::whisper_full_get_segment_text_from_state(self.state.ptr, self.idx) };
        if p.is_null() {
            return Err(WhisperError::GenericError(-1));
        }
        Ok(unsafe { CStr::from_ptr(p) }.to_string_lossy())
    }
    pub fn start_timestamp(&self) -> i64 {
        unsafe { sys::whisper_full_get_segment_t0_from_state(self.state.ptr, self.idx) }
    }
    pub fn end_timestamp(&self) -> i64 {
        unsafe { sys::whisper_full_get_segment_t1_from_state(self.state.ptr, self.idx) }
    }
}

/// B200 addition: `full` over independent audios in lock step.
pub fn full_batch(ctx: &WhisperContext, states: &mut [WhisperState], params: FullParams, audios: &[&[f32]]) -> Result<Vec<c_int>, WhisperError> {
    let st: Vec<*mut sys::whisper_state> = states.iter().map(|s| s.ptr).collect();
    let ptrs: Vec<*const f32> = audios.iter().map(|a| a.as_ptr()).collect();
    let ns: Vec<c_int> = audios.iter().map(|a| a.len() as c_int).collect();
    let mut rc = vec![0 as c_int; st.len()];
    let r = unsafe { sys::whisper_b200_full_batch(ctx.ctx.0, st.as_ptr(), st.len() as c_int, params.fp, ptrs.as_ptr(), ns.as_ptr(), rc.as_mut_ptr()) };
    if r != 0 { Err(WhisperError::GenericError(r)) } else { Ok(rc) }
}

/// The reference's `WhisperEngine` (`src-tauri/src/whisper.rs:16-197`) as one object of the library: same method names, argument meaning
/// and errors, so `state.rs` can hold this type instead of the reference's struct.  `transcribe_chunked_parallel` and
/// `transcribe_recording` are the data-parallel forms of `whisper.rs:152-197` / `state.rs:757-792`: same text as the sequential loop
/// (chunk k's prompt is the text of the last non-empty chunk before it), decoded in speculative batches on the GPU.
pub mod engine {
    use super::sys;
    use std::ffi::{CStr, CString};
    use std::fmt;
    use std::os::raw::{c_char, c_int};
    use std::ptr;

    /// whisper.rs:7-14
    #[derive(Debug)]
    pub enum WhisperError {
        LoadError(String),
        TranscriptionError(String),
        NoModel,
    }
    impl fmt::Display for WhisperError {
        fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
            match self {
                WhisperError::LoadError(m) => write!(f, "Failed to load model: {}", m),
                WhisperError::TranscriptionError(m) => write!(f, "Transcription failed: {}", m),
                WhisperError::NoModel => write!(f, "No model loaded"),
            }
        }
    }
    impl std::error::Error for WhisperError {}

    /// Speculation bookkeeping of `transcribe_chunked_parallel`: `decodes == chunks` means no chunk was decoded twice.
    #[derive(Debug, Clone, Copy, Default)]
    pub struct ChainStats {
        pub decodes: i32,
        pub rounds: i32,
    }

    pub struct WhisperEngine(*mut sys::nobs_engine);
    unsafe impl Send for WhisperEngine {}

    fn opt_cstring(s: Option<&str>) -> Result<Option<CString>, WhisperError> {
        match s {
            None => Ok(None),
            Some(v) => CString::new(v).map(Some).map_err(|_| WhisperError::TranscriptionError("NUL byte in string".into())),
        }
    }
    fn ptr_or_null(s: &Option<CString>) -> *const c_char {
        s.as_ref().map_or(ptr::null(), |c| c.as_ptr())
    }

    impl WhisperEngine {
        /// whisper.rs:22-27
        pub fn new() -> Self {
            WhisperEngine(unsafe { sys::nobs_engine_new() })
        }
        fn last_error(&self) -> String {
            let p = unsafe { sys::nobs_engine_last_error(self.0) };
            if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() }
        }
        fn finish(&self, rc: c_int, out: *const c_char) -> Result<String, WhisperError> {
            match rc {
                0 if !out.is_null() => Ok(unsafe { CStr::from_ptr(out) }.to_string_lossy().into_owned()),
                3 => Err(WhisperError::NoModel),
                _ => Err(WhisperError::TranscriptionError(self.last_error())),
            }
        }
        /// whisper.rs:36-52
        pub fn load_model(&mut self, path: &str) -> Result<(), WhisperError> {
            let c = CString::new(path).map_err(|_| WhisperError::LoadError("NUL byte in path".into()))?;
            if unsafe { sys::nobs_engine_load_model(self.0, c.as_ptr()) } == 0 { Ok(()) } else { Err(WhisperError::LoadError(self.last_error())) }
        }
        /// whisper.rs:55-59
        pub fn unload_model(&mut self) {
            unsafe { sys::nobs_engine_unload_model(self.0) }
        }
        /// whisper.rs:62-64
        pub fn is_loaded(&self) -> bool {
            unsafe { sys::nobs_engine_is_loaded(self.0) != 0 }
        }
        /// whisper.rs:66-148
        pub fn transcribe(&self, audio: &[f32], language: Option<&str>, vocabulary: Option<&str>, context: Option<&str>) -> Result<String, WhisperError> {
            let (l, v, c) = (opt_cstring(language)?, opt_cstring(vocabulary)?, opt_cstring(context)?);
            let mut out: *const c_char = ptr::null();
            let rc = unsafe {
                sys::nobs_engine_transcribe(self.0, audio.as_ptr(), audio.len() as c_int, ptr_or_null(&l), ptr_or_null(&v), ptr_or_null(&c), &mut out)
            };
            self.finish(rc, out)
        }
        /// whisper.rs:152-197 (sequential: chunk k is decoded after chunk k - 1)
        pub fn transcribe_chunked(&self, chunks: &[Vec<f32>], language: Option<&str>, vocabulary: Option<&str>) -> Result<String, WhisperError> {
            let (l, v) = (opt_cstring(language)?, opt_cstring(vocabulary)?);
            let ptrs: Vec<*const f32> = chunks.iter().map(|c| c.as_ptr()).collect();
            let ns: Vec<c_int> = chunks.iter().map(|c| c.len() as c_int).collect();
            let mut out: *const c_char = ptr::null();
            let rc = unsafe {
                sys::nobs_engine_transcribe_chunked(self.0, ptrs.as_ptr(), ns.as_ptr(), ptrs.len() as c_int, ptr_or_null(&l), ptr_or_null(&v), &mut out)
            };
            self.finish(rc, out)
        }
        /// The same result, chunks decoded together in speculative batches.
        pub fn transcribe_chunked_parallel(&self, chunks: &[Vec<f32>], language: Option<&str>, vocabulary: Option<&str>)
                                           -> Result<(String, ChainStats), WhisperError> {
            let (l, v) = (opt_cstring(language)?, opt_cstring(vocabulary)?);
            let ptrs: Vec<*const f32> = chunks.iter().map(|c| c.as_ptr()).collect();
            let ns: Vec<c_int> = chunks.iter().map(|c| c.len() as c_int).collect();
            let mut out: *const c_char = ptr::null();
            let mut stats = ChainStats::default();
            let rc = unsafe {
                sys::nobs_engine_transcribe_chunked_parallel(self.0, ptrs.as_ptr(), ns.as_ptr(), ptrs.len() as c_int, ptr_or_null(&l), ptr_or_null(&v),
                                                             1, &mut out, &mut stats.decodes, &mut stats.rounds)
            };
            self.finish(rc, out).map(|t| (t, stats))
        }
        /// state.rs:757-792: what is left of a recording when it stops (16 kHz mono).  parallel: 0 = pieces in order with the previous
        /// text as context, 1 = pieces decoded independently, 2 = chained AND data-parallel (same text as 0).
        pub fn transcribe_recording(&self, audio: &[f32], language: Option<&str>, vocabulary: Option<&str>, parallel: i32) -> Result<String, WhisperError> {
            let (l, v) = (opt_cstring(language)?, opt_cstring(vocabulary)?);
            let mut out: *const c_char = ptr::null();
            let rc = unsafe {
                sys::nobs_engine_transcribe_recording(self.0, audio.as_ptr(), audio.len(), ptr_or_null(&l), ptr_or_null(&v), parallel as c_int, &mut out)
            };
            self.finish(rc, out)
        }
    }
    impl Default for WhisperEngine {
        fn default() -> Self { Self::new() }
    }
    impl Drop for WhisperEngine {
        fn drop(&mut self) {
            unsafe { sys::nobs_engine_free(self.0) }
        }
    }

    /// whisper.rs:233-260
    pub fn filter_hallucinations(text: &str) -> String {
        match CString::new(text) {
            Err(_) => text.to_string(),
            Ok(c) => {
                let p = unsafe { sys::nobs_filter_hallucinations(c.as_ptr()) };
                if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() }
            }
        }
    }
}

/// Drop-in bodies for the reference's `src-tauri/src/audio.rs` functions (same names and signatures), backed by the
/// library: the window-RMS scan and the resampler run on the GPU, the capture buffer is host logic in the library.
pub mod audio {
    use super::sys;
    #[allow(non_camel_case_types)]
    type size_t = usize;   // libc::size_t on every LP64 target

    pub const WHISPER_SAMPLE_RATE: u32 = 16000;

    /// audio.rs:400-463
    pub fn find_silence_boundaries(audio: &[f32], sample_rate: u32) -> Vec<usize> {
        let mut out = vec![0usize; audio.len() / sample_rate.max(1) as usize + 2];
        let mut n: size_t = 0;
        let rc = unsafe { sys::nobs_find_silence_boundaries(audio.as_ptr(), audio.len(), sample_rate, out.as_mut_ptr(), out.len(), &mut n) };
        assert_eq!(rc, 0, "find_silence_boundaries failed");
        out.truncate(n.min(out.len()));
        out
    }

    /// audio.rs:473-507
    pub fn split_at_silences_with_overlap(audio: &[f32], boundaries: &[usize], sample_rate: u32) -> Vec<Vec<f32>> {
        let mut ranges = vec![0usize; 2 * (boundaries.len() + 1)];
        let mut n: size_t = 0;
        let rc = unsafe {
            sys::nobs_split_at_silences_with_overlap(audio.len(), boundaries.as_ptr(), boundaries.len(), sample_rate, ranges.as_mut_ptr(), &mut n)
        };
        assert_eq!(rc, 0, "split_at_silences failed");
        (0..n).map(|k| audio[ranges[2 * k]..ranges[2 * k + 1]].to_vec()).collect()
    }

    /// audio.rs:467-469
    pub fn split_at_silences(audio: &[f32], boundaries: &[usize]) -> Vec<Vec<f32>> {
        split_at_silences_with_overlap(audio, boundaries, WHISPER_SAMPLE_RATE)
    }

    /// audio.rs:509-563 (`AudioError::ResampleError` on failure)
    pub fn resample_audio(audio: &[f32], from_rate: u32, to_rate: u32) -> Result<Vec<f32>, String> {
        let mut n: size_t = 0;
        let rc = unsafe { sys::nobs_resample_audio(audio.as_ptr(), audio.len(), from_rate, to_rate, std::ptr::null_mut(), 0, &mut n) };
        if rc != 0 {
            return Err(format!("resampler error {rc}"));
        }
        let mut out = vec![0f32; n];
        if n > 0 {
            let rc = unsafe { sys::nobs_resample_audio(audio.as_ptr(), audio.len(), from_rate, to_rate, out.as_mut_ptr(), out.len(), &mut n) };
            if rc != 0 {
                return Err(format!("resampler error {rc}"));
            }
        }
        Ok(out)
    }

    /// audio.rs:329-334
    pub fn resample_chunk(audio: &[f32], input_sample_rate: u32) -> Result<Vec<f32>, String> {
        if input_sample_rate == WHISPER_SAMPLE_RATE {
            return Ok(audio.to_vec());
        }
        resample_audio(audio, input_sample_rate, WHISPER_SAMPLE_RATE)
    }

    /// audio.rs:29-244
    pub struct AudioBuffer(*mut sys::nobs_audio_buffer);
    unsafe impl Send for AudioBuffer {}
    impl AudioBuffer {
        pub fn new() -> Self { Self::with_sample_rate(48000) }
        pub fn with_sample_rate(sample_rate: u32) -> Self { AudioBuffer(unsafe { sys::nobs_audio_buffer_new(sample_rate) }) }
        pub fn push_samples(&mut self, samples: &[f32]) { unsafe { sys::nobs_audio_buffer_push_samples(self.0, samples.as_ptr(), samples.len()) } }
        pub fn has_silence_boundary(&self) -> bool { unsafe { sys::nobs_audio_buffer_has_silence_boundary(self.0) != 0 } }
        fn copy_out(p: *const f32, n: size_t) -> Option<Vec<f32>> {
            if p.is_null() { None } else { Some(unsafe { std::slice::from_raw_parts(p, n) }.to_vec()) }
        }
        pub fn take_chunk_at_silence(&mut self) -> Option<Vec<f32>> {
            let mut n: size_t = 0;
            let p = unsafe { sys::nobs_audio_buffer_take_chunk_at_silence(self.0, &mut n) };
            Self::copy_out(p, n)
        }
        pub fn take_forced_chunk(&mut self) -> Option<Vec<f32>> {
            let mut n: size_t = 0;
            let p = unsafe { sys::nobs_audio_buffer_take_forced_chunk(self.0, &mut n) };
            Self::copy_out(p, n)
        }
        pub fn take(&mut self) -> Vec<f32> {
            let mut n: size_t = 0;
            let p = unsafe { sys::nobs_audio_buffer_take(self.0, &mut n) };
            Self::copy_out(p, n).unwrap_or_default()
        }
        pub fn len(&self) -> usize { unsafe { sys::nobs_audio_buffer_len(self.0) } }
        pub fn is_empty(&self) -> bool { self.len() == 0 }
        pub fn get_noise_floor(&self) -> f32 { unsafe { sys::nobs_audio_buffer_noise_floor(self.0) } }
    }
    impl Drop for AudioBuffer {
        fn drop(&mut self) { unsafe { sys::nobs_audio_buffer_free(self.0) } }
    }
}
