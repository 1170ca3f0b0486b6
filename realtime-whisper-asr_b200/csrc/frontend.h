// Launchers for the HBM-bound front-end kernels (frontend.cu).  All pointers are device pointers.
#pragma once
#include "common.h"

namespace b200w {

// log-mel of n_streams concatenated clips.  out holds, per stream s, a [n_mels][N_s/160 + 1] fp32
// matrix at element offset out_offsets[s].  tile_prefix[s] = sum_{j<s} ceil(frames_j / 32).
void logmel_launch(const float* audio, const long long* offsets, const long long* out_offsets,
                   const int* tile_prefix, int n_streams, int total_tiles, int n_mels, float* out,
                   int* stream_max, int max_frames, cudaStream_t stream);

// Cut 3000-frame encoder windows (bf16, time-major, 128 channels) out of the fp32 features.
void pack_windows_launch(const float* feats, const long long* feat_offsets, const int* n_frames,
                         const int* seeks, const int* win_stream, int n_windows, int n_mels,
                         bf16* windows, cudaStream_t stream);

// Audio-dedup features, one CTA per chunk.  out / raw_out: [n_chunks][5].
void dedup_features_launch(const float* audio, const long long* offsets, int n_chunks,
                           int max_points, int sample_rate, float* out, float* raw_out,
                           cudaStream_t stream);

// Zero-phase 8th-order IIR (scipy.signal.filtfilt defaults) of every chunk.  coeffs_host: b[9], a[9] (a[0] = 1),
// zi[8] as doubles.  rpad_scratch: [n_chunks][27] floats.  Chunks of <= 27 samples are copied unchanged.
void bandpass_filtfilt_launch(const float* audio, const long long* offsets, int n_chunks, const double* coeffs_host,
                              float* out, float* rpad_scratch, cudaStream_t stream);
static constexpr int BANDPASS_PAD = 27;

// Batched VAD gate (DynamicVADIterator / upstream VADIterator semantics): windows fp32 [n_windows][512];
// win_offsets int32 [n_streams + 1]; state int64 [n_streams][3]; events int64 [n_streams][4].
void vad_window_energy_launch(const float* windows, int n_windows, float k, float* probs, cudaStream_t stream);
void vad_gate_launch(const float* probs, const int* win_offsets, int n_streams, long long* state,
                     const double* min_silence, float threshold, long long speech_pad, long long* events,
                     cudaStream_t stream);

}  // namespace b200w
