"""nobs-whisper-b200: B200-native Whisper transcription engine behind the whisper-rs API.

`api` mirrors the whisper-rs types the reference uses (reference src-tauri/src/whisper.rs:3);
`engine` mirrors the reference's WhisperEngine wrapper (whisper.rs:16-260).  Both are thin
ctypes layers over the C ABI of libnobswhisper_b200.so (include/whisper_b200.h); all compute
runs in that library's sm_100a kernels.
"""
from .api import (FullParams, SamplingStrategy, WhisperContext, WhisperContextParameters, WhisperError, WhisperSegment,  # noqa: F401
                  WhisperState, decode_batch, full_batch)
from .engine import NoModel, LoadError, TranscriptionError, WhisperEngine, filter_hallucinations  # noqa: F401
