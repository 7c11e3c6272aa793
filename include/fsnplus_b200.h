/*
 * fsnplus_b200 -- C ABI of the B200-native FullSubNet+/FullSubNet inference forward.
 *
 * This is the drop-in boundary for ONE path of RookieJunChen/FullSubNet-plus: the model forward
 *   speech_enhance/fullsubnet_plus/model/fullsubnet_plus.py:122-209  (FullSubNet_Plus.forward)
 *   speech_enhance/fullsubnet/model/fullsubnet.py:68-118             (Model.forward)
 * as called by the inferencer at
 *   speech_enhance/fullsubnet_plus/inferencer/inferencer.py:150 and :123.
 * The reference has no FFI of its own (it is pure Python/PyTorch); the Python mirror in
 * fullsubnet-plus_b200/fsnplus_b200/model.py binds these entry points with ctypes and exposes the
 * reference's nn.Module constructors / forward signatures / state_dict keys.
 *
 * Conventions: plain C types only; every pointer named d_* is a CUDA device pointer, h_* a host
 * pointer; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all entry points
 * return 0 on success or a negative FSN_E* code, with a message available from fsn_last_error().
 * There is no CPU fallback: without a CUDA device every compute entry point fails with FSN_ECUDA.
 */
#ifndef FSNPLUS_B200_H_
#define FSNPLUS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSN_OK 0
#define FSN_EINVAL (-1)  /* bad argument / unsupported configuration (the reference raises/asserts) */
#define FSN_ECUDA (-2)   /* CUDA runtime error (including "no device")                              */
#define FSN_ESTATE (-3)  /* call order error (e.g. forward before finalize, missing parameter)      */

/* model_kind */
#define FSN_KIND_PLUS 0 /* fullsubnet_plus.model.fullsubnet_plus.FullSubNet_Plus (fullsubnet_plus.py:16) */
#define FSN_KIND_FSN 1  /* fullsubnet.model.fullsubnet.Model                      (fullsubnet.py:12)      */

/* activations: audio_zen/model/module/sequence_model.py:84-93 */
#define FSN_ACT_NONE 0
#define FSN_ACT_RELU 1
#define FSN_ACT_TANH 2
#define FSN_ACT_RELU6 3

/* norm_type: audio_zen/model/base_model.py:318-330 (norm_wrapper) */
#define FSN_NORM_OFFLINE_LAPLACE 0
#define FSN_NORM_CUMULATIVE_LAPLACE 1
#define FSN_NORM_OFFLINE_GAUSSIAN 2
#define FSN_NORM_CUMULATIVE_LAYER 3

/* channel_attention_model: fullsubnet_plus.py:52-70 (0 keeps a zero-initialised config on the inference.toml default) */
#define FSN_ATTN_TSSE 0    /* ChannelTimeSenseSELayer  attention_model.py:43-104  */
#define FSN_ATTN_SE 1      /* ChannelSELayer           attention_model.py:6-40    */
#define FSN_ATTN_CBAM 2    /* ChannelCBAMLayer         attention_model.py:296-334 */
#define FSN_ATTN_ECA 3     /* ChannelECAlayer          attention_model.py:337-359 */

/* sequence_model of the recurrent modules: sequence_model.py:31-46 */
#define FSN_RNN_LSTM 0
#define FSN_RNN_GRU 1      /* runs on the LSTM kernels as pseudo-gates (r, z, n_x, n_h) with a GRU cell */

/* lstm_impl */
#define FSN_LSTM_AUTO 0
#define FSN_LSTM_MMA 1     /* generic mma.sync kernel (any hidden size / layer count)               */
#define FSN_LSTM_TCGEN05 2 /* persistent tcgen05/TMEM kernels: fused (2 layers, hidden <= 384) or layer-wise (hidden <= 512); input <= 64 */

typedef struct fsn_config {
    int32_t model_kind;
    int32_t num_freqs;        /* F                                   (config/inference.toml:33) */
    int32_t look_ahead;       /*                                      (:34)                      */
    int32_t sb_num_neighbors; /*                                      (:31)                      */
    int32_t fb_num_neighbors; /*                                      (:32)                      */
    int32_t fb_hidden;        /* fb_model_hidden_size (LSTM full band; ignored by the TCN)       */
    int32_t sb_hidden;        /* sb_model_hidden_size                 (:40)                      */
    int32_t num_layers;       /* 2 in the reference (fullsubnet_plus.py:76,106); additive knob   */
    int32_t output_size;      /* 2                                    (fullsubnet_plus.py:31)    */
    int32_t fb_act;           /* fb_output_activate_function          (:36)                      */
    int32_t sb_act;           /* sb_output_activate_function          (:37)                      */
    int32_t norm_type;        /*                                      (:42)                      */
    int32_t kersize[3];       /* TSSE kernel sizes                    (:44)                      */
    int32_t lstm_impl;        /* FSN_LSTM_*                                                      */
    int32_t fast_math;        /* 0: ex2/rcp gates; 1: tanh.approx gates (the Python mirror's default) */
    int32_t channel_attention;/* FSN_ATTN_*  channel_attention_model  (:38)                      */
    int32_t rnn_type;         /* FSN_RNN_*   sequence_model           (:35)                      */
    int32_t subband_num;      /* 0 or 1: off; > 1 needs FSN_ATTN_ECA (the only attention whose
                                 reference forward runs in that mode, fullsubnet_plus.py:146-163) */
    int32_t tcn_causal;       /* 0: TCNBlock(causal=False), what SequenceModel("TCN") builds (sequence_model.py:47-58);
                                 1: TCNBlock(causal=True) in the three full-band models (causal_conv.py:74-75,104-105):
                                 depth-wise taps t-2d, t-d, t instead of t-d, t, t+d.  Additive knob */
} fsn_config;

typedef struct fsn_model fsn_model;

int fsn_version(void);
const char* fsn_last_error(void);

/* Replaces FullSubNet_Plus.__init__ / Model.__init__ (fullsubnet_plus.py:17-120, fullsubnet.py:13-66):
 * validates the configuration and creates an empty parameter store on the current CUDA device. */
int fsn_model_create(const fsn_config* cfg, fsn_model** out);
void fsn_model_destroy(fsn_model* m);

/* Replaces nn.Module.load_state_dict for this module tree (audio_zen/inferencer/base_inferencer.py:104):
 * `key` is the reference state_dict key, `h_data` float32 host data in the reference's shape. */
int fsn_model_set_param(fsn_model* m, const char* key, const float* h_data, int64_t numel);
/* Number of keys the configuration expects / i-th expected key and its element count. */
int fsn_model_num_params(const fsn_model* m);
int fsn_model_param_info(const fsn_model* m, int index, const char** key, int64_t* numel);
/* Re-orders / converts the parameters into the kernel layouts (fp16 gate-interleaved LSTM images, ...). */
int fsn_model_finalize(fsn_model* m);

/* Replaces FullSubNet_Plus.forward(noisy_mag, noisy_real, noisy_imag) (fullsubnet_plus.py:122-209) and
 * Model.forward(noisy_mag) (fullsubnet.py:68-118; pass d_real = d_imag = NULL).
 * Inputs: [B, 1, F, T] float32, contiguous, device.  Output: [B, output_size, F, T] float32, contiguous,
 * device, caller-owned.  Every sample is processed independently (the eval semantics of the reference
 * called with batch 1; the training-only drop_band of fullsubnet_plus.py:192-196 is never applied).  Because of that a
 * batch whose workspace would exceed FSN_WS_CAP_GB (default 48) is run as equal sub-batches on `stream`, with identical results. */
int fsn_model_forward(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                      float* d_out, void* stream);
/* The model call PLUS the two lines that follow it in the reference's inferencer methods (fullsubnet_plus/inferencer/inferencer.py:152-157,
 * audio_zen/acoustics/mask.py:60-63): decompress_cIRM and the complex multiplication with the noisy spectrum, fused into the epilogue
 * of the sub-band LSTM kernel -- the mask never goes to memory.  d_real / d_imag are the planes the model takes anyway (required
 * here also for fullsubnet.Model, which does not read them in its forward); d_enh: [B, F, T] complex64 (interleaved re, im) =
 * the argument of the inferencer's iSTFT. */
int fsn_model_forward_enhance(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                              float* d_enh, void* stream);
/* Same call with HOST buffers (pinned or pageable): H2D copies, the forward and the D2H copy of the
 * mask are all enqueued on `stream` and the call returns after the result is in h_out. */
int fsn_model_forward_host(fsn_model* m, const float* h_mag, const float* h_real, const float* h_imag, int32_t B, int32_t T,
                           float* h_out, void* stream);

/* Pipelined variants for STREAMS OF BATCHES (the reference's inferencer loop, base_inferencer.py:133-160, issues one forward
 * after the other; the outputs of a batch are only needed when its files are written).  The forward is split at the one point
 * where it changes character: the full-band front end (norm, attention, TCN / full-band LSTM, sub-band statistics + packing) is
 * bandwidth/latency-bound and occupies the whole GPU for a short time, the sub-band LSTM is tensor-bound and occupies 130 of the
 * 148 SMs for a long time.  Both entry points run the front end of batch i+1 on an internal stream, into the second of two
 * workspace lanes, WHILE the sub-band LSTM of batch i runs on another internal stream: the front end fills the idle SMs.
 *
 * fsn_model_submit: device buffers.  The inputs must be complete in `stream` order at the time of the call and stay untouched,
 * d_out unread, until fsn_model_wait(m, stream) has been called (it makes `stream` wait for every submitted batch; it does not
 * block the host) or fsn_model_sync_host(m) has returned.  A plain fsn_model_forward on any stream first waits for the pipeline.
 *
 * fsn_model_forward_host_async: pinned host buffers; additionally the H2D copy of batch i+1 and the D2H copy of batch i-1
 * overlap on two copy streams.  `stream` is ignored (host memory carries no stream order).  The host buffers must stay
 * untouched (h_out unread) until fsn_model_sync_host() returns. */
int fsn_model_submit(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                     float* d_out, void* stream);
int fsn_model_submit_enhance(fsn_model* m, const float* d_mag, const float* d_real, const float* d_imag, int32_t B, int32_t T,
                             float* d_enh, void* stream);   /* pipelined fsn_model_forward_enhance */
int fsn_model_wait(fsn_model* m, void* stream);
/* Finer-grained completion for consumers that post-process batch i while batch i+1 runs: the workspace lane (0 / 1) the LAST
 * fsn_model_submit used, and a wait on the batch most recently submitted into one lane (valid until the next submit into it). */
int fsn_model_last_lane(const fsn_model* m);
int fsn_model_wait_lane(fsn_model* m, int32_t lane, void* stream);
int fsn_model_forward_host_async(fsn_model* m, const float* h_mag, const float* h_real, const float* h_imag, int32_t B, int32_t T,
                                 float* h_out, void* stream);
int fsn_model_sync_host(fsn_model* m);

/* Streaming (frame-by-frame) inference for the causal configuration (BASELINE config #4): fullsubnet.Model with
 * norm_type cumulative_laplace_norm or cumulative_layer_norm.  State carried between calls: running norm sums, (h, c) of the
 * full-band and sub-band LSTMs.  Call n consumes magnitude frame n ([B, F] float32, device) and writes the cIRM of frame
 * n - look_ahead to d_mask ([B, 2, F]); *h_valid = 0 for the first look_ahead calls.  Feeding look_ahead zero frames at the
 * end flushes the tail, exactly like the reference's right padding (fullsubnet.py:81). */
typedef struct fsn_stream fsn_stream;
int fsn_stream_create(fsn_model* m, int32_t B, fsn_stream** out);
int fsn_stream_step(fsn_stream* st, const float* d_mag_frame, float* d_mask, int32_t* h_valid, void* stream);
void fsn_stream_destroy(fsn_stream* st);

/* Fused post-processing of the inferencer method (fullsubnet_plus/inferencer/inferencer.py:152-157 and
 * audio_zen/acoustics/mask.py:60-63): decompress_cIRM (K = 10, limit = 9.9) + complex multiply with the noisy spectrum.
 * d_crm [B, 2, F, T] float32 (model output), d_noisy / d_enh [B, F, T] complex64 (interleaved re, im). */
int fsn_apply_cirm(const float* d_crm, const float* d_noisy, float* d_enh, int32_t B, int32_t F, int32_t T, void* stream);

/* Test hooks: copy an intermediate of the LAST forward to a device buffer.
 *   "fb_in"  [nbranch, B, F, T+look_ahead]  post-norm (and post-attention) full-band inputs
 *   "fb_out" [nbranch, B, F, T+look_ahead]  full-band model outputs
 *   "sb_mu"  [B]                            utterance mean of the (un-normalised) sub-band input */
int fsn_model_get_stage(fsn_model* m, const char* name, float* d_dst, int64_t numel, void* stream);
/* Kernels launched by the last forward (bench.py reports it as gpu_launches). */
int64_t fsn_model_last_launch_count(const fsn_model* m);
/* Device time (ms, CUDA events on the forward's stream) of the sub-band LSTM kernel of the last forward;
 * synchronises on the closing event.  < 0 if no forward has run. */
float fsn_model_last_lstm_ms(fsn_model* m);
/* Same for the last n forwards (n <= 32, oldest first); returns how many were written.  Synchronises. */
int fsn_model_lstm_ms_history(fsn_model* m, float* h_ms, int32_t n);
/* Device timeline of the last n forwards (n <= 32, oldest first): 4 floats each = front-end start, front-end end, sub-band
 * LSTM start, LSTM end in ms relative to the oldest one's front-end start (CUDA events on the streams the phases run on).  With
 * the pipelined entry points the front end of batch i+1 lies inside the LSTM interval of batch i.  Returns the count; synchronises. */
int fsn_model_timeline(fsn_model* m, float* h_ms4, int32_t n);
/* Which LSTM implementation the last forward used for the sub-band model (FSN_LSTM_MMA / _TCGEN05). */
int fsn_model_last_lstm_impl(const fsn_model* m);

/* Host-side packers exposed for CPU tests of the kernel layouts (no GPU needed). */
/* 128-row x 64-half tile with the SWIZZLE_128B layout: byte offset of element (row, k). */
uint32_t fsn_sw128_offset(uint32_t row, uint32_t k);
/* Size in bytes and content of the tcgen05 weight stream for a 2-layer LSTM. */
int64_t fsn_tc5_weight_stream_bytes(int32_t input_size, int32_t hidden);
/* weight row (in [W_i; W_f; W_g; W_o] order) that gate column n (0..127) of 32-unit chunk j maps to */
int32_t fsn_tc5_gate_row(int32_t hidden, int32_t chunk, int32_t n);
int fsn_tc5_pack_weights(int32_t input_size, int32_t hidden, const float* w_ih0, const float* w_hh0, const float* w_ih1,
                         const float* w_hh1, uint16_t* h_dst /* fp16 bits */);

/* Layer-wise tcgen05 path, one layer: recurrent stream (hidden/32 chunks x hidden/64 tiles of 16 KB), the
 * input-projection matrix for the GEMM ([4 hidden, k_pad] fp16, rows in chunk column order) and the pre-scaled biases
 * ([4 hidden] in the same order; gru != 0 scales pseudo-gate 3 like a tanh argument).  b_ih / b_hh in the nn.LSTM 4-block layout. */
int64_t fsn_tc5r_weight_stream_bytes(int32_t hidden);
int fsn_tc5r_pack_layer(int32_t hidden, int32_t k_in, int32_t k_pad, const float* w_ih, const float* w_hh, const float* b_ih,
                        const float* b_hh, int32_t gru, uint16_t* h_stream, uint16_t* h_wih, float* h_bias);

#ifdef __cplusplus
}
#endif
#endif /* FSNPLUS_B200_H_ */
