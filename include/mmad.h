/*
 * mmad.h -- C ABI of libmmad.so, the B200 (sm_100a) implementation of the
 * autoencoder + RaPP reconstruction-aggregation hot path of
 * Yoo-Youngjae/ICRA2021_multimodal_ad.
 *
 * The reference is pure Python and has no FFI of its own; each
 * entry point below names the reference function (file:line) whose arithmetic it
 * replaces.  The Python package ``icra2021_multimodal_ad_b200`` binds these with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - every function returns 0 on success, a negative MMAD_E_* code on error;
 *    mmad_last_error() returns a thread-local message.  No C++ exception crosses.
 *  - pointers named d_* are DEVICE pointers, h_* are HOST pointers.  The caller owns
 *    every buffer; the library allocates device memory only in mmad_create /
 *    mmad_set_layer / mmad_nap_set_fit (packed weights), mmad_peer_create / mmad_peer_grad_alloc (exchange
 *    buffers), and lazily on first use: in the two host-buffer entry points that take no workspace
 *    (mmad_score_host, mmad_stream_*) and, for a per-modality model in the fp32 mode, the transposed
 *    weight copy of the fused whole-chain kernel at its first mmad_score.  Otherwise never in mmad_score.
 *  - scratch comes from a caller-provided workspace sized by the *_workspace_bytes
 *    queries; work is enqueued on the caller's cudaStream_t (passed as void*),
 *    asynchronously, with no implicit synchronisation unless stated.
 *  - a handle is bound to the device current at mmad_create and is not thread-safe.
 *  - there is no CPU fallback: without a CUDA device every compute call fails with
 *    MMAD_E_CUDA.
 */
#ifndef MMAD_H_
#define MMAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMAD_MAX_LAYERS 16

#define MMAD_OK 0
#define MMAD_E_ARG (-1)
#define MMAD_E_CUDA (-2)
#define MMAD_E_STATE (-3)
#define MMAD_E_WORKSPACE (-4)
#define MMAD_E_UNSUPPORTED (-5)

/* GEMM arithmetic.  FP32: CUDA-core fp32 FMA (bit-for-bit fp32 products, the parity
 * reference mode).  F16X3: tcgen05 tensor cores, every fp32 operand split into an fp16
 * pair (hi + lo), three MMAs per product, fp32 accumulation in TMEM -- 1e-4 score
 * tolerance on the golden models, see DESIGN.md section 3 for a trained model.  F16: one tcgen05 pass on the hi parts only (separately stated
 * tolerance).  F16F8: the hi*hi pass in fp16 plus ONE fp8 (e4m3) pass of twice the
 * contraction length carrying both cross terms, [lo8 | a8] . [Wh8 ; Wl8] -- two thirds of the
 * F16X3 tensor work; at D = 1728 the same measured score error as F16X3 (both bounded by the tensor
 * core's truncating accumulation: <= 1e-4 on 99.9 % of the windows of a trained model), 5e-4 for the
 * narrow sensors.  Training in this mode uses the F16X3 split. */
#define MMAD_PREC_FP32 0
#define MMAD_PREC_F16X3 1
#define MMAD_PREC_F16 2
#define MMAD_PREC_F16F8 3

typedef struct mmad_handle* mmad_t;

typedef struct {
    int n_enc;                              /* encoder layers (reference: config.n_layers) */
    int n_dec;                              /* decoder layers */
    int enc_widths[MMAD_MAX_LAYERS + 1];    /* [D, h1.., btl]   model_builder.py:21-28 */
    int dec_widths[MMAD_MAX_LAYERS + 1];    /* [btl, .., D]     model_builder.py:30-37 */
    float lrelu_slope;                      /* 0.2              modules/activation.py:37-38 */
    float bn_eps;                           /* 1e-5             layers/fc_layer.py:30 */
    int precision;                          /* MMAD_PREC_* */
} mmad_desc_t;

const char* mmad_last_error(void);
int mmad_version(void);

/* model_builder.py:6-53 (ae_wrapper/get_model): build the packed device model. */
int mmad_create(const mmad_desc_t* desc, mmad_t* out);
int mmad_destroy(mmad_t h);
int mmad_set_precision(mmad_t h, int precision);   /* invalidates an installed NAP fit (refit in the new arithmetic) */

/* Tuning knobs of a handle (none of them has a reference counterpart; they select between implementations of the
 * same reference arithmetic):
 *   "acc_comp"        relative shrink of a tensor-core accumulator per MMA instruction that the epilogue compensates
 *                     (first-order correction of the truncating fp32 accumulation); 0 = off
 *   "nap_passes"      F16X3 NAP rotation: 3 = full split (default), 4 = fp16 hi*hi + fp8 cross terms (the F16F8 arithmetic for
 *                     this one GEMM; <= 2.2e-5 on well-conditioned selections), 2 = whitening rows
 *                     rounded to fp16, 0 = default.  Changing it invalidates the installed NAP fit
 *   "require_pinned"  1: mmad_score_host returns MMAD_E_ARG for pageable bulk input instead of accepting it
 *   "smallnet"        0: FP32-mode models whose widths are all <= 128 use the per-layer kernels instead of the fused
 *                     whole-chain kernel (default 1)
 * Unknown names return MMAD_E_ARG. */
int mmad_set_option(mmad_t h, const char* name, double value);

/* Load one FCLayer (layers/fc_layer.py:23-35) from state_dict tensors.  module 0 =
 * encoder, 1 = decoder.  d_W [N,K] row-major fp32, d_b [N].  BatchNorm pointers are
 * NULL for the bare last layer (modules/fc_module.py:49-54).  Repacks (pads K to the
 * tile multiple, splits fp16 hi/lo, folds eval BatchNorm into scale/shift).  New weights invalidate the
 * installed NAP fit and the cached launch graphs: run the NAP fit again before asking for NAP scores. */
int mmad_set_layer(mmad_t h, int module, int index, const float* d_W, const float* d_b,
                   const float* d_gamma, const float* d_beta, const float* d_mean,
                   const float* d_var, void* stream);

/* layers/fc_layer.py:37-48 FCLayer.forward (eval mode) as a stand-alone fused op on raw
 * state_dict tensors: y = BN(lrelu(x W^T + b)) or y = x W^T + b when d_gamma == NULL.
 * This is what ``for layer in model.encoder.layer_list: x = layer(x)``
 * (reconstruction_aggregation.py:25-27) executes per layer. */
int mmad_fc_layer_forward(const float* d_x, int ldx, int n, int K, int N, const float* d_W, const float* d_b,
                          const float* d_gamma, const float* d_beta, const float* d_mean, const float* d_var,
                          float slope, float eps, float* d_y, int ldy, void* stream);

/* Bytes of workspace needed to process up to max_rows rows per call (scoring, forward
 * and NAP calls chunk internally to what the workspace allows). */
size_t mmad_workspace_bytes(mmad_t h, int max_rows);

/* models/auto_encoder.py:36-50 AutoEncoder.forward in eval mode: d_xhat[n,D] = dec(enc(x)).
 * d_z (optional, [n,btl]) receives AutoEncoder.encode (36-39). ldx = row stride of d_x in floats. */
int mmad_ae_forward(mmad_t h, const float* d_x, int ldx, int n, float* d_xhat, float* d_z,
                    void* d_ws, size_t ws_bytes, void* stream);

/* models/auto_encoder.py:79-91 validate / 52-55 get_loss_value in eval mode:
 * *d_loss (one float) = sum (xhat - x)^2. */
int mmad_recon_loss(mmad_t h, const float* d_x, int ldx, int n, float* d_loss,
                    void* d_ws, size_t ws_bytes, void* stream);

/* reconstruction_aggregation.py:6-37 get_diffs fused with utils/metric.py:132-133
 * (base score) and 145-171 (SAP):
 *   d_base[n] = mean_j d_0^2              (NULL to skip)
 *   d_sap[n]  = mean over concat(d_lo..d_hi-1) of d^2   (NULL to skip)
 *   d_diffs   = the concatenated diffs [n, sum_{l=lo}^{hi-1} w_l] row-major (NULL to skip)
 *   d_nap[n]  = NAP score (utils/metric.py:183-222) from the fit installed with
 *               mmad_nap_set_fit (NULL to skip; the fit's layer range must equal lo..hi)
 * layer_lo / layer_hi follow the python slice diffs[lo:hi] AFTER the reference's
 * clamping (utils/metric.py:155-162); 0 <= lo < hi <= n_enc+1.
 * encoder(x) is evaluated once, not twice as in the reference. */
int mmad_score(mmad_t h, const float* d_x, int ldx, int n, int layer_lo, int layer_hi,
               float* d_base, float* d_sap, float* d_nap, float* d_diffs,
               void* d_ws, size_t ws_bytes, void* stream);

/* Same as mmad_score with HOST buffers: rows go to the device in chunks, host->device copies
 * overlap compute on internal streams, scores come back through pinned staging, and the call
 * returns after everything has landed (synchronous).  This is the call the e2e benchmark times.
 * h_x should be PINNED (cudaMallocHost / cudaHostRegister / torch pin_memory): from pageable memory every
 * chunk copy blocks the calling thread while the driver stages it, and the copy/compute overlap is lost
 * (accepted, slower; mmad_set_option("require_pinned", 1) turns it into MMAD_E_ARG).
 * Exception to "the caller owns every buffer": this entry point has no workspace argument, so the library
 * allocates its device staging + workspace on the first call (and again only if a later call needs more). */
int mmad_score_host(mmad_t h, const float* h_x, int ldx, long long n, int layer_lo, int layer_hi,
                    float* h_base, float* h_sap, float* h_nap);

/* Realtime calls (test_file/realtime_tester.py:291-309: 10 windows per call, base + SAP, nap=False): mmad_score_host with
 * n <= 64 and h_nap == NULL runs the whole chain in ONE cooperative launch of exact-fp32 kernels, whatever the handle's
 * precision mode; input and scores travel through pinned, device-mapped memory and completion is a sequence flag the
 * host spins on (no H2D / D2H copies, no stream synchronise).  mmad_stream_input returns that pinned input buffer
 * ([*max_rows, D] floats, owned by the handle): a caller that assembles its windows there and passes the same pointer to
 * mmad_score_host saves the staging memcpy.  MMAD_E_UNSUPPORTED when the model does not fit the kernel (D % 4 != 0, a
 * layer wider than 16 columns per SM); such models keep the graph-replay path inside mmad_score_host. */
int mmad_stream_input(mmad_t h, int layer_lo, int layer_hi, float** h_in, int* max_rows);

/* modules/loss.py:31-32 nn.MSELoss(reduction='sum'): *d_out += sum_i (a_i - b_i)^2 (zero it first). */
int mmad_sq_diff_sum(const float* d_a, const float* d_b, long long n, float* d_out, void* stream);
/* models/auto_encoder.py:73 loss.backward(): the gradients of the fused step are d(loss)/d(param) for a unit seed;
 * d_buf[n] *= *d_scale for any other autograd seed, nothing is touched when *d_scale == 1 (n % 4 == 0, 16-byte aligned). */
int mmad_scale_unless_one(float* d_buf, long long n, const float* d_scale, void* stream);
/* utils/metric.py:133,171 (d**2).mean(axis=1) over a [n, cols] matrix with row stride ld. */
int mmad_row_mean_sq(const float* d_d, int ld, int n, int cols, float* d_out, void* stream);
/* decorators/variational_info_bottleneck.py:19-42 ('normal'): d_out [B, 2h] -> mu, logvar [B,h],
 * z[k,B,h] = eps * exp(logvar/2) + mu (d_eps [k,B,h]); d_eps == NULL => z = mu broadcast. */
int mmad_vib_reparam(const float* d_out, int ld, int B, int h, int k, const float* d_eps, float* d_z,
                     float* d_mu, float* d_logvar, void* stream);

/* ---- NAP fit: utils/normalize.py:47-70 (Rotater.fit) + 20-34 (Standardizer.fit) ----
 * Pass 1: d_sum[Dsel] (fp64) += column sums of the concatenated diffs of these rows.
 * Pass 2: d_gram[Dsel*Dsel] (fp64, row-major, full symmetric) += (d-mu)^T (d-mu),
 *         d_mu[Dsel] fp32 being sum/N after the caller has combined all shards
 *         (multi-GPU: all-reduce d_sum, then d_gram, between the calls).
 * Dsel = mmad_concat_width(h, lo, hi). */
int mmad_concat_width(mmad_t h, int layer_lo, int layer_hi);
int mmad_nap_accumulate_sum(mmad_t h, const float* d_x, int ldx, int n, int layer_lo, int layer_hi,
                            double* d_sum, void* d_ws, size_t ws_bytes, void* stream);
int mmad_nap_accumulate_gram(mmad_t h, const float* d_x, int ldx, int n, int layer_lo, int layer_hi,
                             const float* d_mu, double* d_gram, void* d_ws, size_t ws_bytes, void* stream);
/* Install a fit: d_mu[Dsel]; d_vt [K, Dsel] row-major = rows are right singular vectors
 * v_j (utils/normalize.py:67); d_var[K] (Standardizer.var), d_mu2[K] (Standardizer.mu).
 * Packs V*diag(var^-1/2) for the scoring GEMM. */
int mmad_nap_set_fit(mmad_t h, int layer_lo, int layer_hi, int K, const float* d_mu,
                     const float* d_vt, const float* d_var, const float* d_mu2, void* stream);

/* Standardizer.fit on the rotated train rows (utils/normalize.py:25-34), with the rotation done by
 * the same kernels that score: d_rsum[K] += sum_r rot[r,j], d_rsq[K] += sum_r rot[r,j]^2 (fp64),
 * rot = (d - mu) V from the installed fit.  After combining shards (all-reduce both), the caller
 * forms mu2 = rsum/N, var = (rsq - N mu2^2)/(N-1) and installs them with mmad_nap_set_standardizer,
 * so rounding noise of the rotation in near-null directions is normalised like the reference's. */
int mmad_nap_rotate_stats(mmad_t h, const float* d_x, int ldx, int n, int layer_lo, int layer_hi,
                          double* d_rsum, double* d_rsq, void* d_ws, size_t ws_bytes, void* stream);
int mmad_nap_set_standardizer(mmad_t h, const float* d_var, const float* d_mu2, void* stream);
/* Declare that the first `upper_triangular` rows of the installed d_vt are upper triangular in the concatenated-diff
 * coordinates (vt[j, c] == 0 for c < j; 0 = dense, K = all rows): a whitening factor R with R^T R = V diag(1/var) V^T
 * gives the same score sum_j ((d-mu).v_j)^2 / var_j = |R (d-mu)|^2 (utils/metric.py:220-222) with half the products;
 * the tensor-core kernels then skip the zero k-blocks of those rows.  A fit may mix both: a triangular factor of the
 * well-conditioned part of the spectrum followed by plain eigenvector rows for the weak directions (Engine.nap_fit,
 * factor="hybrid").  The caller guarantees the structure. */
int mmad_nap_set_structure(mmad_t h, int upper_triangular);

/* ---- stand-alone normaliser ops (utils/normalize.py API compatibility: Rotater / Standardizer on
 * arbitrary device matrices d[n, cols], row stride ld) ----
 * mmad_col_stats: d_mean[cols] = column means (fp64 accumulation, utils/normalize.py:31,61);
 *   d_var (optional) = diag(np.cov) with ddof=1 (normalize.py:34).
 * mmad_gram_accumulate: d_gram[cols*cols] (fp64) += (d-mu)^T (d-mu)  (basis of Rotater.fit: the right
 *   singular vectors of the centred matrix are the eigenvectors of this Gram matrix).
 * mmad_rotate: d_out[n, K] = (d - mu) V with d_vt = V^T [K, cols] row-major (normalize.py:72-103).
 * mmad_standardize: d_out = (d - mu) / sqrt(var) (normalize.py:36-45; no epsilon, like the reference).
 * ws: mmad_normalizer_workspace_bytes(cols). */
size_t mmad_normalizer_workspace_bytes(int cols);
int mmad_col_stats(const float* d_d, int ld, long long n, int cols, float* d_mean, float* d_var, void* d_ws,
                   size_t ws_bytes, void* stream);
int mmad_gram_accumulate(const float* d_d, int ld, long long n, int cols, const float* d_mu, double* d_gram,
                         void* d_ws, size_t ws_bytes, void* stream);
int mmad_rotate(const float* d_d, int ld, long long n, int cols, const float* d_mu, const float* d_vt, int K,
                float* d_out, int ldo, void* d_ws, size_t ws_bytes, void* stream);
int mmad_standardize(const float* d_d, int ld, long long n, int cols, const float* d_mu, const float* d_var,
                     float* d_out, int ldo, void* stream);
/* utils/normalize.py:52-70 across ranks: the centred Gram matrix is symmetric, so the cross-rank SUM moves its upper triangle
 * only.  d_packed holds D(D+1)/2 doubles, row i = columns i..D-1; unpack writes both triangles of d_full[D*D]. */
int mmad_tri_pack(const double* d_full, int D, double* d_packed, void* stream);
int mmad_tri_unpack(const double* d_packed, int D, double* d_full, void* stream);

/* ---- metrics: utils/metric.py:29-130 (sklearn roc_curve/auc, precision_recall_curve,
 * np.quantile, F1, confusion) on device.  d_score fp32, d_label uint8 (0/1).
 * h_out receives doubles; results are bit-identical to the reference given identical
 * scores.  These calls synchronise the stream (they return host scalars). */
size_t mmad_metric_workspace_bytes(long long n);
int mmad_auc_roc(const float* d_score, const uint8_t* d_label, long long n, double* h_out,
                 void* d_ws, size_t ws_bytes, void* stream);
int mmad_auc_prc(const float* d_score, const uint8_t* d_label, long long n, double* h_out,
                 void* d_ws, size_t ws_bytes, void* stream);
/* threshold = np.quantile(valid, q) in fp32 (NumPy 2 semantics). */
int mmad_quantile(const float* d_valid, long long n, float q, float* h_out,
                  void* d_ws, size_t ws_bytes, void* stream);
/* h_counts[4] = {tp, fp, fn, tn} with pred = score > thr (strict=1, get_f1_score) or
 * score >= thr (strict=0, get_confusion_matrix). */
int mmad_confusion(const float* d_score, const uint8_t* d_label, long long n, float thr, int strict,
                   long long* h_counts, void* d_ws, size_t ws_bytes, void* stream);

/* ---- training: models/auto_encoder.py:57-77 AutoEncoder.step ----
 * Train-mode forward (BatchNorm batch statistics), loss = sum (xhat-x)^2, backward.
 * Parameters, gradients, BN running stats live in caller-owned fp32 tensors, passed
 * as arrays of device pointers in state_dict order per layer.  If d_eps != NULL the
 * encoder output is treated as (mu, logvar) and reparameterised with the given noise
 * (decorators/variational_info_bottleneck.py:19-42) and beta_kl * KL is added. */
typedef struct {
    float* W; float* b; float* gamma; float* beta; float* run_mean; float* run_var;   /* params/buffers */
    long long* num_batches_tracked;                                                  /* int64 scalar, += 1 (may be NULL) */
    float* gW; float* gb; float* ggamma; float* gbeta;                               /* gradients */
} mmad_train_layer_t;

size_t mmad_train_workspace_bytes(mmad_t h, int batch);
/* allreduce hook: called (if non-NULL) on the BatchNorm statistic buffers (forward: column sum and
 * sum of squares; backward: sum g and sum g*xhat) so that N-GPU data parallel equals 1 GPU on the
 * concatenated batch; SUM-all-reduce count doubles at d_buf in place, ordered on stream. */
typedef int (*mmad_allreduce_fn)(void* ctx, double* d_buf, long long count, void* stream);
int mmad_train_fwd_bwd(mmad_t h, const float* d_x, int ldx, int batch, long long global_batch,
                       const mmad_train_layer_t* enc, const mmad_train_layer_t* dec,
                       const float* d_eps, float beta_kl, float bn_momentum,
                       float* d_loss, void* d_ws, size_t ws_bytes,
                       mmad_allreduce_fn allreduce, void* allreduce_ctx, void* stream);
/* The loss of the LAST mmad_train_fwd_bwd enqueued on this handle, read from a (value, sequence) pair in mapped pinned memory
 * that the step publishes as soon as its forward pass has produced the loss -- i.e. without waiting for the backward pass
 * and whatever else is queued behind it (AutoEncoder.step returns float(loss) every step, models/auto_encoder.py:77; through
 * the stream that read is a wait for the whole step).  Blocks (polling) until the value has arrived. */
int mmad_train_loss(mmad_t h, float* h_loss);
/* torch.optim.Adam (novelty_detection.py:90) over a flat list of tensors, one launch. */
int mmad_adam_step(int n_tensors, float* const* h_params, float* const* h_grads, float* const* h_m,
                   float* const* h_v, const long long* h_numel, int step, float lr, float beta1,
                   float beta2, float eps, float grad_scale, void* stream);

/* ---- communicator: the exchange steps of the path enqueued by the library itself ----
 * The reference has no distributed code; these back the data-parallel train step and the sharded NAP fit.
 * One rank calls mmad_comm_unique_id (ncclGetUniqueId), the 128-byte id is distributed by the host layer's own
 * rendezvous, every rank calls mmad_comm_init (ncclCommInitRank; collective).  With a communicator of world > 1
 * installed and allreduce == NULL, mmad_train_fwd_bwd all-reduces the BatchNorm statistics itself (and the step is
 * still captured into a CUDA graph); mmad_comm_allreduce_* are in-place SUM all-reduces on the given stream
 * (flat gradient buffer, NAP sums / Gram).  libnccl.so.2 is resolved from the process at run time. */
#define MMAD_UNIQUE_ID_BYTES 128
int mmad_comm_unique_id(unsigned char* h_id);
int mmad_comm_init(mmad_t h, const unsigned char* h_id, int rank, int world);
int mmad_comm_destroy(mmad_t h);
int mmad_comm_world(mmad_t h);
/* on: mmad_train_fwd_bwd also SUM-all-reduces the gradients in TWO buckets on its second stream -- the decoder's as soon
 * as the decoder's backward pass is done (overlapping the encoder's), the encoder's at the end of the step (needs a
 * communicator; the parameter gradients of a module must form one contiguous block, as the host layer's flat buffer does). */
int mmad_comm_set_grad_allreduce(mmad_t h, int on);
int mmad_comm_allreduce_f32(mmad_t h, float* d_buf, long long count, void* stream);
int mmad_comm_allreduce_f64(mmad_t h, double* d_buf, long long count, void* stream);

/* Latency-bound exchange of small fp64 vectors (the BatchNorm statistics of a data-parallel step: 32 strictly serialised
 * all-reduces of <= 2 x 1408 doubles per step) over NVLink peer memory instead of ncclAllReduce: every rank creates a
 * buffer (mmad_peer_create returns its 64-byte cudaIpc handle), the host layer gathers the handles of all ranks of the
 * node (rank order) and every rank maps them (mmad_peer_open; call a barrier afterwards).  From then on
 * mmad_comm_allreduce_f64 / the in-step exchanges of at most 4096 doubles run as ONE kernel (push (data, sequence)
 * pairs to every peer, poll, add in rank order -- identical bits on every rank), also inside captured CUDA graphs. */
#define MMAD_IPC_HANDLE_BYTES 64
int mmad_peer_create(mmad_t h, unsigned char* h_handle);
int mmad_peer_open(mmad_t h, const unsigned char* h_handles, int rank, int world);
int mmad_peer_close(mmad_t h);
int mmad_peer_allreduce_f64(mmad_t h, double* d_buf, long long count, void* stream);
/* Gradient all-reduce over peer memory.  The flat fp32 gradient buffer must be mapped by the peers, so the library owns it:
 * mmad_peer_grad_alloc (after mmad_peer_open) allocates n_floats (rounded up to 4) zeroed floats on this rank, returns the
 * device pointer and its cudaIpc handle; the host layer gathers the handles (rank order) and every rank calls
 * mmad_peer_grad_open, then a barrier.  From then on mmad_comm_allreduce_f32 on EXACTLY that buffer (pointer and rounded
 * count) runs as one kernel per rank: entry barrier over flags in peer memory, rank r sums chunk r of all ranks' buffers
 * (P2P loads, rank order) and stores it into every buffer, exit barrier.  Replaces torch.distributed.all_reduce of
 * novelty_detection.py's gradient step (the reference is single-GPU).  Freed by mmad_peer_close. */
int mmad_peer_grad_alloc(mmad_t h, long long n_floats, float** d_ptr, unsigned char* h_handle);
int mmad_peer_grad_open(mmad_t h, const unsigned char* h_handles);

/* ---- multimodal feature extractor in front of the autoencoder ----
 * utils/data_loaders.py:152-229 (HSR_Net.forward) / 601-674 (Multisensory_module.forward): per sample
 * conv stacks on the 32x32 RGB and depth images, broadcast force-torque scalar, two 1-d convolutions on the 13
 * MFCC coefficients, concatenated as [rgb 1024 | depth 512 | force-torque 64 | mic 128] -- one launch for the batch.
 * d_r [B,3,32,32], d_d [B,1,32,32], d_t [B], d_m [B,13]; a NULL input drops that modality (the reference's
 * unimodal variants) and the remaining blocks are packed in the same order.  Weights: nn.Conv2d / nn.Conv1d
 * tensors in PyTorch layout.  h_affine (host, 8 floats, may be NULL): (scale, shift) applied to r, d, t, m on load
 * = norm_vec (utils/data_loaders.py:703-712). */
typedef struct {
    const float *conv1r_w, *conv1r_b, *conv2r_w, *conv2r_b, *conv3r_w, *conv3r_b;
    const float *conv1d_w, *conv1d_b, *conv2d_w, *conv2d_b, *conv3d_w, *conv3d_b;
    const float *conv1l_w, *conv1l_b, *conv2l_w, *conv2l_b;
} mmad_feature_weights_t;
int mmad_multisensory_width(int has_r, int has_d, int has_t, int has_m);
int mmad_multisensory_forward(const float* d_r, const float* d_d, const float* d_t, const float* d_m, int batch,
                              const mmad_feature_weights_t* w, const float* h_affine, float* d_out, int ldo,
                              void* stream);

/* ---- measurement hooks (bench.py) ----
 * mmad_profile_begin: record a CUDA-event pair around every fused-GEMM launch of this handle.
 * mmad_profile_end: synchronise; h_out[0] = sum of GEMM kernel durations (ms), h_out[1] = their
 * algorithmic FLOPs (2*M*N*K per launch), h_out[2] = number of GEMM launches.
 * mmad_launch_count: kernels launched by the library since load (all handles). */
int mmad_profile_begin(mmad_t h);
int mmad_profile_end(mmad_t h, double* h_out);
unsigned long long mmad_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MMAD_H_ */
