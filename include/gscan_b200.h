/*
 * gscan_b200.h - C ABI of libgscan_b200.so: the sm_100a kernels under the drop-in `Model`.
 *
 * The reference (LauraRuis/multimodal_seq2seq_gSCAN) has no FFI layer: its boundary is the Python
 * class seq2seq.model.Model (reference seq2seq/model.py:24-261).  This library is the layer that
 * sits UNDER a Model with that exact API (multimodal_seq2seq_gscan_b200/model.py); every entry
 * point below names the reference code it replaces.  Conventions for every function:
 *   - plain C: raw DEVICE pointers + explicit sizes, no torch types; returns 0 on success,
 *     a negative GSCAN_E_* code on bad arguments, or a positive cudaError_t value;
 *   - launches only on the given stream (a cudaStream_t passed as void*), never synchronises,
 *     allocates nothing: all memory (outputs, workspace) is provided by the caller;
 *   - all floating point is IEEE fp32 (the reference's dtype); tokens are int64 as in the
 *     reference's LongTensors; lengths are int32 on the device (the reference passes host lists);
 *   - threading: ONE host thread per device at a time.  gscan_forward / gscan_backward / gscan_greedy_decode fork onto
 *     library-owned helper streams (one set per device) and join back into the caller's stream before returning, so a
 *     call looks like ordinary work on `stream`; those helper streams, the stage-timing events, the cached TMA
 *     descriptors and the SM budgets of the persistent GEMMs are per-device library state without locks.  Calls for
 *     DIFFERENT devices may run concurrently from different threads (all caches are keyed by device); two threads
 *     driving the same device must serialise their calls.  This matches the reference, which is single-threaded
 *     and uses the default stream (SURVEY.md 8(b) "Threading").
 *
 * Parameter table: `params` / `grads` are host arrays of GSCAN_NUM_PARAMS device pointers in
 * model.parameters() order (= Adam state order), indexed by the GSCAN_P_* enum.
 */
#ifndef GSCAN_B200_H
#define GSCAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSCAN_ABI_VERSION 1

enum gscan_error {
  GSCAN_OK = 0,
  GSCAN_E_BADARG = -1,      /* null pointer / inconsistent dims                         */
  GSCAN_E_UNSUPPORTED = -2, /* shape outside what the kernels handle (see gscan_check_dims) */
  GSCAN_E_WORKSPACE = -3    /* workspace too small                                      */
};

/* index = position in the reference's model.parameters() (model.py:46-98) */
enum gscan_param {
  GSCAN_P_CONV1_W = 0, GSCAN_P_CONV1_B, GSCAN_P_CONV2_W, GSCAN_P_CONV2_B, GSCAN_P_CONV3_W, GSCAN_P_CONV3_B,
  GSCAN_P_VIS_KEY_W, GSCAN_P_VIS_QUERY_W, GSCAN_P_VIS_ENERGY_W,
  GSCAN_P_ENC_EMB,
  GSCAN_P_ENC_WIH, GSCAN_P_ENC_WHH, GSCAN_P_ENC_BIH, GSCAN_P_ENC_BHH,
  GSCAN_P_ENC_WIH_R, GSCAN_P_ENC_WHH_R, GSCAN_P_ENC_BIH_R, GSCAN_P_ENC_BHH_R,
  GSCAN_P_E2D_W, GSCAN_P_E2D_B,
  GSCAN_P_TXT_KEY_W, GSCAN_P_TXT_QUERY_W, GSCAN_P_TXT_ENERGY_W,
  GSCAN_P_COND_W, GSCAN_P_COND_B,           /* NULL when conditional_attention == 0 */
  GSCAN_P_DEC_EMB,
  GSCAN_P_DEC_WIH, GSCAN_P_DEC_WHH, GSCAN_P_DEC_BIH, GSCAN_P_DEC_BHH,
  GSCAN_P_O2H_W, GSCAN_P_H2O_W,
  GSCAN_NUM_PARAMS
};

/* Problem shape.  Names follow SURVEY.md section 8. */
typedef struct gscan_dims {
  int32_t B;     /* examples in this call                                                  */
  int32_t Ti;    /* padded command length = max(cmd_len); commands has row stride Ti_stride */
  int32_t Tt;    /* padded target length (all Tt steps are computed, as the reference does) */
  int32_t G;     /* grid size; M = G*G situation cells                                      */
  int32_t C;     /* situation channels (num_cnn_channels)                                   */
  int32_t F;     /* cnn_hidden_num_channels; D = 3F                                         */
  int32_t K3;    /* cnn_kernel_size of conv_3 (conv_1 is 1x1, conv_2 is 5x5)                */
  int32_t E;     /* command embedding dimension                                             */
  int32_t H;     /* encoder_hidden_size == decoder_hidden_size; must be a multiple of 4     */
  int32_t Vi;    /* input vocabulary size                                                   */
  int32_t V;     /* target vocabulary size                                                  */
  int32_t conditional_attention; /* 1: visual query = tanh(W_c [h; c_T] + b_c) (seq2seq_model.py:394-396) */
  int32_t auxiliary_task;        /* 1: aux_logp = log_softmax(sum_t beta_t) is produced (model.py:166-170) */
  int32_t pad_idx_in;  /* input_padding_idx  (embedding row without gradient)  */
  int32_t pad_idx_out; /* target_pad_idx                                       */
  int32_t Ti_stride;   /* row stride (elements) of `commands`; >= Ti           */
} gscan_dims;

int gscan_abi_version(void);

/* Number of kernels this library has launched in this process (memsets / copies not counted). */
unsigned long long gscan_launch_count(void);

/*
 * Stage timing for bench.py / profiling.  After gscan_profile(1), gscan_forward and gscan_backward
 * record CUDA events on the caller's stream at their stage boundaries; gscan_profile_read waits for
 * them and writes GSCAN_NUM_STAGES durations in milliseconds (-1 for a stage that did not run):
 *   0 CNN + command encoder + key projections     1 weight packing + target embeddings + gate pre-GEMM
 *   2 decoder forward sweep (all Tt steps)        3 output projection + log-softmax (+ aux head)
 *   4 (unused seam between forward and backward)  5 output-projection backward
 *   6 decoder backward sweep (BPTT)               7 decoder weight-gradient GEMMs + embedding scatter
 *   8 CNN / key / initial-state / encoder backward
 * Not thread safe; intended for single-stream measurement.
 */
#define GSCAN_NUM_STAGES 9
int gscan_profile(int enable);
int gscan_profile_read(float* stage_ms /* [GSCAN_NUM_STAGES] */);

/* 0 if the kernels support this shape, GSCAN_E_UNSUPPORTED otherwise. */
int gscan_check_dims(const gscan_dims* d);

/* Number of floats of workspace gscan_forward / gscan_backward need for this shape.  The same
 * workspace must be passed to the backward call that follows a forward call: it carries the
 * saved activations (the role autograd's graph plays in the reference). */
size_t gscan_workspace_floats(const gscan_dims* d);

/*
 * Teacher-forced forward of the whole model.
 * Replaces Model.forward = encode_input + decode_input_batched + log_softmax (+ aux)
 * (reference model.py:172-219; cnn_model.py:22-36; seq2seq_model.py:47-89, 359-490).
 *   commands   [B, Ti_stride] int64      cmd_len [B] int32 (valid tokens incl. SOS/EOS, >= 1)
 *   situations [B, G, G, C]  fp32, [row, col, channel]
 *   targets    [B, Tt] int64 (SOS ... EOS pad)
 *   drop_cnn [B, G*G, 3F], drop_enc [B, Ti, E], drop_dec [B, Tt, H]: dropout masks already
 *       scaled by 1/(1-p), or NULL for no dropout (eval mode / p = 0)
 *   logp     [B, Tt, V]  out: log-probabilities (Model.forward's first output)
 *   aux_logp [B, G*G]    out: only written when d->auxiliary_task (may be NULL otherwise)
 */
int gscan_forward(const gscan_dims* d, const float* const* params,
                  const int64_t* commands, const int32_t* cmd_len, const float* situations,
                  const int64_t* targets,
                  const float* drop_cnn, const float* drop_enc, const float* drop_dec,
                  float* workspace, size_t workspace_floats,
                  float* logp, float* aux_logp, void* stream);

/*
 * Backward of gscan_forward: replaces autograd over the reference graph (`loss.backward()`,
 * reference train.py:110).  d_logp [B,Tt,V] and d_aux_logp [B,G*G] (NULL = zero) are the upstream
 * gradients of the two outputs.  Every grads[i] receives the gradient (OVERWRITTEN, not
 * accumulated) of params[i]; grads of the two embedding padding rows are zero, as
 * nn.Embedding(padding_idx) gives (seq2seq_model.py:42,351).
 */
int gscan_backward(const gscan_dims* d, const float* const* params,
                   const int64_t* commands, const int32_t* cmd_len, const float* situations,
                   const int64_t* targets,
                   const float* drop_cnn, const float* drop_enc, const float* drop_dec,
                   float* workspace, size_t workspace_floats,
                   const float* d_logp, const float* d_aux_logp,
                   float* const* grads, void* stream);

/*
 * Model.encode_input (reference model.py:172-179).  Tt in `d` is ignored.
 *   feat [B, G*G, 3F], enc_out [Ti, B, H] (time-major, zero at padded positions), hidden [B, H].
 *   workspace: gscan_encode_workspace_floats(d) floats.
 */
size_t gscan_encode_workspace_floats(const gscan_dims* d);
int gscan_encode(const gscan_dims* d, const float* const* params,
                 const int64_t* commands, const int32_t* cmd_len, const float* situations,
                 const float* drop_cnn, const float* drop_enc,
                 float* workspace, size_t workspace_floats,
                 float* feat, float* enc_out, float* hidden, void* stream);

/*
 * One decoder step: Model.decode_input -> BahdanauAttentionDecoderRNN.forward_step
 * (reference model.py:181-188; seq2seq_model.py:359-428), eval-mode (no dropout) unless drop_dec.
 *   tokens [B] int64; h_in, c_in [B,H]; keys_text [Ti,B,H] and keys_vis [B,G*G,H] are the
 *   PROJECTED keys (predict.py:87-90); outputs logits [B,V], h_out, c_out [B,H],
 *   alpha [B,Ti], beta [B,G*G].   workspace: gscan_step_workspace_floats(d) floats.
 */
size_t gscan_step_workspace_floats(const gscan_dims* d);
int gscan_decoder_step(const gscan_dims* d, const float* const* params,
                       const int64_t* tokens, const float* h_in, const float* c_in,
                       const float* keys_text, const int32_t* cmd_len, const float* keys_vis,
                       const float* drop_dec /* [B,H] or NULL */,
                       float* workspace, size_t workspace_floats,
                       float* logits, float* h_out, float* c_out, float* alpha, float* beta,
                       void* stream);

/*
 * Batched greedy decoding: replaces the reference's batch-size-1 loop (predict.py:57-128) with
 * identical per-sequence results.  Each sequence starts from `sos`, takes the first-max argmax of
 * the logits, stops after producing `eos` or after max_decoding_steps + 1 tokens (`<=` at
 * predict.py:101).
 *   out_tokens [B, max_decoding_steps+1] int64: generated tokens WITHOUT the final eos; unused = -1
 *   out_len    [B] int32: number of tokens kept per sequence
 *   out_steps  [B] int32: decoder steps executed (= out_len, +1 if the sequence ended with eos)
 *   beta_sum   [B, G*G]: sum of visual attention over executed steps (predict.py:111,119)
 *   aux_logp   [B, G*G] or NULL: log_softmax(beta_sum) (Model.auxiliary_task_forward)
 *   alphas [B, max_decoding_steps+1, Ti], betas [B, max_decoding_steps+1, G*G]: per-step
 *       attention weights for the predictions JSON (predict.py:44-51); either may be NULL.
 *   workspace: gscan_greedy_workspace_floats(d) floats.
 */
size_t gscan_greedy_workspace_floats(const gscan_dims* d);
int gscan_greedy_decode(const gscan_dims* d, const float* const* params,
                        const int64_t* commands, const int32_t* cmd_len, const float* situations,
                        int32_t max_decoding_steps, int32_t sos, int32_t eos,
                        float* workspace, size_t workspace_floats,
                        int64_t* out_tokens, int32_t* out_len, int32_t* out_steps,
                        float* beta_sum, float* aux_logp, float* alphas, float* betas,
                        void* stream);

/*
 * NLLLoss(ignore_index = pad_idx), mean over the non-ignored entries, of logp [B,Tt,V] against
 * targets [B,Tt] shifted left by `shift`: entry (b,t) is scored against targets[b, t+shift], entries
 * with t+shift >= Tt hold the literal 0 the reference appends (ignored when pad_idx == 0, scored as class 0 otherwise).
 *   shift = 1: Model.get_loss (reference model.py:108-115,147-160: drop SOS, append a pad);
 *   shift = 0, Tt = 1, pad_idx = -100: Model.get_auxiliary_loss (model.py:59,162-164).
 * loss_out[0] = mean NLL, loss_out[1] = number of scored entries (as float); loss_out must hold
 * GSCAN_NLL_OUT_FLOATS floats (the rest is scratch of the fixed-order two-stage reduction).
 * gscan_nll_backward writes d_logp = d_loss * dNLL/dlogp (dense, overwritten).
 */
#define GSCAN_NLL_OUT_FLOATS 68
int gscan_nll_forward(const float* logp, const int64_t* targets, int32_t B, int32_t Tt, int32_t V,
                      int32_t pad_idx, int32_t shift, float* loss_out /* [GSCAN_NLL_OUT_FLOATS] */, void* stream);
int gscan_nll_backward(const int64_t* targets, int32_t B, int32_t Tt, int32_t V, int32_t pad_idx,
                       int32_t shift, const float* loss_out /* [2] from forward */,
                       const float* d_loss /* [1] */, float* d_logp /* [B,Tt,V], overwritten */,
                       void* stream);
/*
 * The number of scored entries alone (loss_out[1]; loss_out[0] = 0), from the targets: with it gscan_nll_backward can
 * form d_logp BEFORE the forward pass has produced logp (the gradient of a mean or sum of NLL terms does not depend on
 * their values).  The fused trainer does this so that nothing but the output head stands between the two sweeps;
 * reference semantics: model.py:100 (NLLLoss(ignore_index)) and train.py:102-107.
 */
int gscan_nll_count(const int64_t* targets, int32_t B, int32_t Tt, int32_t pad_idx, int32_t shift,
                    float* loss_out /* [2] */, void* stream);

/*
 * Model.get_metrics (reference model.py:117-137) without host syncs:
 * counts[0] = matching non-pad tokens, counts[1] = non-pad tokens, counts[2] = exactly matching
 * sequences (int32 each).  accuracy = 100*counts[0]/counts[1]; exact = 100*counts[2]/B.
 */
int gscan_metrics(const float* logp, const int64_t* targets, int32_t B, int32_t Tt, int32_t V,
                  int32_t pad_idx, int32_t* counts /* [3] */, void* stream);

/*
 * Dropout drawn inside the kernels (new; the reference draws its masks with three nn.Dropout modules per step:
 * cnn_model.py:31, seq2seq_model.py:59,385).  gscan_forward_rng / gscan_backward_rng are gscan_forward / gscan_backward
 * without mask arguments: every dropout site (0 CNN features [B, G*G, 3F], 1 command embeddings [B, Ti, E], 2 target
 * embeddings [B, Tt, H]) takes element i of its mask from a Philox4x32-10 stream keyed by (seed, offset, site) - keep
 * with probability 1 - p, scale by 1 / (1 - p).  The backward call must get the SAME gscan_dropout as its forward call.
 * Advance `offset` by one per training step.  gscan_dropout_mask writes the n mask values of one site (parity tests:
 * the materialised masks passed to gscan_forward / gscan_backward give bit-identical results).
 */
typedef struct gscan_dropout {
  float p_cnn, p_enc, p_dec;
  uint64_t seed, offset;
} gscan_dropout;
int gscan_forward_rng(const gscan_dims* d, const float* const* params,
                      const int64_t* commands, const int32_t* cmd_len, const float* situations,
                      const int64_t* targets, const gscan_dropout* rng,
                      float* workspace, size_t workspace_floats, float* logp, float* aux_logp, void* stream);
/*
 * Forward pass of a TRAINING step whose loss gradient with respect to the log-probabilities is known in advance (for the
 * reference's loss - NLLLoss of the shifted targets, train.py:102-107 - it depends on the targets only: gscan_nll_count +
 * gscan_nll_backward).  As gscan_forward (masks) / gscan_forward_rng (rng != NULL), and the output-head backward pass -
 * log-softmax backward, hidden_to_output and output_to_hidden, what model.py:188 and seq2seq_model.py:378-380 leave to
 * autograd - runs inside this call, most of it beside the decoder sweep; the gscan_backward[_rng] call that follows on
 * the same workspace with the SAME d_logp pointer (contents unchanged) then starts with the reverse-time sweep.  A
 * backward call with another d_logp recomputes that stage: results never depend on which path ran.
 * d_logp_ready: NULL, or a cudaEvent_t recorded after the work that produces d_logp (it may still be running on another
 * stream when this call is made; only the head-backward kernels wait for it).
 */
int gscan_forward_train(const gscan_dims* d, const float* const* params,
                        const int64_t* commands, const int32_t* cmd_len, const float* situations,
                        const int64_t* targets, const float* drop_cnn, const float* drop_enc, const float* drop_dec,
                        const gscan_dropout* rng, float* workspace, size_t workspace_floats, float* logp,
                        float* aux_logp, const float* d_logp /* [B,Tt,V] */, void* d_logp_ready, void* stream);
int gscan_backward_rng(const gscan_dims* d, const float* const* params,
                       const int64_t* commands, const int32_t* cmd_len, const float* situations,
                       const int64_t* targets, const gscan_dropout* rng,
                       float* workspace, size_t workspace_floats, const float* d_logp, const float* d_aux_logp,
                       float* const* grads, void* stream);
int gscan_dropout_mask(const gscan_dropout* rng, int32_t which, size_t n, float* out, void* stream);

/*
 * Fused Adam over a flat parameter/gradient buffer (what torch.optim.Adam does per tensor in
 * reference train.py:67-70,110-113, with the LambdaLR factor folded into `lr`):
 *   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v/(1-b2^t)) + eps)
 * grad_scale multiplies g first (used to turn summed data-parallel gradients into means).
 */
int gscan_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                    float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                    void* stream);

/*
 * The same step with the gradient divided by a number that lives on the DEVICE: g = grad * grad_scale / *grad_denom.
 * Data parallelism (new; the reference is single-device): every rank backpropagates the SUM form of its
 * shard's loss into the flat buffer and appends its [n_tok, n_examples]; one all-reduce sums gradients and
 * counts together, and grad_denom points at the summed token count inside that buffer - the normalisation of
 * NLLLoss(ignore_index) (reference model.py:100,159) applied after the collective without a host round trip.
 * grad_denom == NULL behaves like gscan_adam_step.
 */
int gscan_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                        float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                        const float* grad_denom, void* stream);

/*
 * Building blocks exported for the parity tests (each is also used inside the calls above).
 */
/* C[M,N] (ldc) = act(opA(A) . opB(B) + bias) with element (i,k) of opA at A[i*a_rs + k*a_cs] and
 * element (k,j) of opB at B[k*b_rs + j*b_cs]; act: 0 none, 1 tanh, 2 relu; accumulate: C += ... */
int gscan_sgemm(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                const float* bias, int32_t act, int32_t accumulate, void* stream);
/* Same product on an explicitly chosen kernel: path 0 = mma.sync 3xTF32 (gemm.cuh), path 1 = tcgen05 3xTF32 with
 * TMA-fed operands (gemm_tc.cuh; GSCAN_E_UNSUPPORTED unless both operands are 16-byte aligned with leading dimensions
 * that are multiples of 4 floats, M >= 64, N >= 32, K >= 32).  ksplit > 1 splits K over CTAs and ADDS the partial sums
 * into C atomically (C must be initialised; bias / act are then rejected).  gscan_sgemm and the training step choose
 * the path by size; this entry point exists so that both kernels can be checked against each other. */
int gscan_sgemm_path(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                     float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                     const float* bias, int32_t act, int32_t accumulate, int32_t ksplit, int32_t path, void* stream);
/* ConvolutionalNet.forward (reference cnn_model.py:22-36): feat [B,G*G,3F]. */
int gscan_cnn_forward(const gscan_dims* d, const float* const* params, const float* situations,
                      const float* drop_cnn, float* workspace, size_t workspace_floats,
                      float* feat, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSCAN_B200_H */
