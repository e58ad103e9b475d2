/* nif_b200 — C ABI of the B200-native NIF hot path (libnif_b200.so).
 *
 * The reference (pswpswpsw/nif) has no FFI: its hot path is Python/TF graph code.
 * Each entry point below names the reference interface it replaces
 * (file:line under /root/reference).  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (16-byte aligned,
 *     contiguous row-major fp32 unless stated); nothing is allocated or freed;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as
 *     void*); no call synchronises;
 *   - return 0 on success, a negative NIF_E_* otherwise; nif_last_error()
 *     returns a thread-local message for the last failure.
 *
 * ShapeNet weights are never materialised per sample.  The last linear layer of
 * the ParameterNet (w_h [K,P], b_h [P]) is re-laid once per step into a
 * kernel-friendly "packed" image (nif_pack) and all kernels work from that.
 */
#ifndef NIF_B200_H
#define NIF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  NIF_OK = 0,
  NIF_E_BAD_DESC = -1,    /* inconsistent / unsupported descriptor */
  NIF_E_BAD_ARG = -2,     /* null / misaligned pointer, negative size */
  NIF_E_UNSUPPORTED = -3, /* valid but not implemented for this build */
  NIF_E_CUDA = -4         /* CUDA runtime error; text in nif_last_error() */
};

enum { NIF_VARIANT_NIF = 0, NIF_VARIANT_SIREN = 1, NIF_VARIANT_SIREN_RES = 2 };
enum { NIF_ACT_LINEAR = 0, NIF_ACT_SINE = 1, NIF_ACT_SWISH = 2, NIF_ACT_TANH = 3,
       NIF_ACT_RELU = 4, NIF_ACT_SIGMOID = 5 };

/* Static description of one hyper-network head + ShapeNet.
 * variant: NIF._call_shape_net (nif/model.py:233-324, act + residual),
 *          NIFMultiScale._call_shape_net_mres plain (:880-954) / res-block (:767-879).
 * act    : tf.keras.activations.get(cfg_shape_net["activation"]) (:303); SIREN variants use sine.
 * si,so,n,l : cfg_shape_net input_dim/output_dim/units/nlayers (:84-87); K = latent_dim (:89).
 * omega0 : cfg_shape_net["omega_0"] (:533); 1.0 for NIF_VARIANT_NIF. */
typedef struct nif_desc {
  int32_t variant;
  int32_t act;
  int32_t si, so, n, l, K;
  float omega0;
  int32_t dtype_compute; /* 0 = fp32 on CUDA cores; 2 = tcgen05 tensor cores with the 3-product fp16 split
                            (fp32-grade, same parity gates; 32 < n <= 64, falls through to 0 for shapes it lacks) */
  int32_t acc_rows;      /* tensor-core batch reductions (weight gradients): rows per accumulation chain.  The tensor
                            core truncates its fp32 accumulator on every instruction, so the error of a batch-reduced
                            gradient grows with the chain; 0 = default (4096 rows: dw within 1e-5 of fp64 at 65 536
                            rows), larger = fewer partials, larger error (16 384: 2e-5).  Rounded up to 64, min 256. */
} nif_desc_t;

/* Sizes derived from a descriptor. */
typedef struct nif_sizes {
  int64_t po_dim;        /* P: columns of pnet_output (nif/model.py:169-173, 569-587) */
  int64_t n_layers;      /* ShapeNet matrices: l+2, or 2l+2 with res-blocks */
  int64_t np;            /* padded width used on chip (32/64/128) */
  int64_t packed_floats; /* floats of one packed weight image (nif_pack output) */
  int64_t save_floats_per_row;  /* floats/row of the activation stash written by forward for backward; the buffer
                                   holds save_floats_per_row * B64 floats, B64 = B rounded up to a multiple of 64
                                   (the tensor-core path tiles rows in groups) */
  int64_t grad_ws_floats;       /* floats of the partial-gradient workspace for batch size B (see nif_sizes B) */
  int64_t tile_rows;     /* rows per CTA tile */
  int64_t kernel_path;   /* which kernels serve this descriptor: 0 = fp32 CUDA-core tile kernels, 2 = tcgen05 FP16x3
                            (the value of dtype_compute that is actually honoured for this shape) */
} nif_sizes_t;

const char* nif_last_error(void);
int nif_version(void);

/* Fill `out` for batch size B (B only affects grad_ws_floats). */
int nif_query_sizes(const nif_desc_t* d, int64_t B, nif_sizes_t* out);

/* Re-lay the last linear layer for the kernels.  Replaces the slicing/reshape
 * of pnet_output columns (nif/model.py:253-300, 769-846, 883-933), applied once
 * to the shared [K,P] matrix instead of to the (B,P) activations.
 * w_h [K,P], b_h [P]  ->  packed [packed_floats].  G>1 packs G independent
 * heads: w_h may be NULL (K must then be 0) and b_h is [G,P] — used by the
 * grouped inference path where each group's full weight vector is given. */
int nif_pack(const nif_desc_t* d, int64_t G, const float* w_h, const float* b_h,
             float* packed, void* stream);

/* Fused forward: u[b,:] = ShapeNet(x[b,:]; weights = z[b,:] @ w_h + b_h).
 * Replaces HyperLinearForSIREN.call (nif/layers/siren.py:514-522) / Dense(po_dim)
 * (nif/model.py:220-230) + _call_shape_net / _call_shape_net_mres
 * (nif/model.py:233-324, 738-954) + EinsumLayer (nif/layers/mlp.py:209-219).
 * z [G*B,K] (ignored when K==0), x [G*B,si] or, if x_shared!=0, [B,si] shared by
 * all groups; packed [G*packed_floats]; u [G*B,so].
 * save: NULL for inference, else [save_floats_per_row * B64], B64 = B rounded up to 64 (G must be 1); its layout is
 * private to the library (row-major or tiled, depending on the kernels that serve the descriptor). */
int nif_forward(const nif_desc_t* d, int64_t G, int64_t B, const float* z, const float* x,
                int32_t x_shared, const float* packed, float* u, float* save, void* stream);

/* Forward with forward-mode tangents (JacobianLayer, nif/layers/gradient.py:36-49,
 * 207-231, realised as tangents instead of one reverse pass per output).
 * n_dir directions; zdot [n_dir,B,K] (may be NULL = 0), xdot [n_dir,B,si] (may be
 * NULL = 0); udot [n_dir,B,so]. */
int nif_forward_tangent(const nif_desc_t* d, int64_t B, const float* z, const float* x,
                        const float* packed, int32_t n_dir, const float* zdot, const float* xdot,
                        float* u, float* udot, void* stream);

/* Second-order forward mode for one pair of directions (a, b): HessianLayer (nif/layers/gradient.py:130-180, 234-261,
 * compute_output_and_grad_and_hessian) without the two nested tapes.  zdot [2][B][K] / xdot [2][B][si] hold the two
 * directions (either may be null), zddot [B][K] the second derivative of the latent code along (a, b) (null = 0;
 * coordinates have none).  Outputs: u [B][so], udot [2][B][so] = (du/da, du/db), uddot [B][so] = d2u / da db.
 * a == b (the same direction twice) gives a diagonal entry. */
int nif_forward_tangent2(const nif_desc_t* d, int64_t B, const float* z, const float* x, const float* packed,
                         const float* zdot, const float* xdot, const float* zddot, float* u, float* udot,
                         float* uddot, void* stream);

/* Sobolev training: JacobianLayer INSIDE the loss (tutorial/8_NIF_with_Sobolov_training.ipynb cell 20:
 * JacobianLayer(model, y_index, x_index) -> concat [u, du/dt, du/dx] -> Sobolov_MSE on u and du/dx), where Keras
 * differentiates the tape of nif/layers/gradient.py:207-231 a second time.  Here: forward-mode tangents with a stash,
 * then a reverse-over-forward pass over every stashed direction.  A direction is (xdot [B,si], zdot [B,K]): a ShapeNet
 * input column has xdot = e_c, zdot = 0; a ParameterNet input column (d/dt) has xdot = 0 and zdot = the trunk tangent
 * of the latent code, and the pass then also returns dL/dzdot for the caller's trunk.
 * nif_forward_tangent_save stashes all n_dir directions: save is [save_floats_per_row * B64] from
 * nif_sobolev_query_dirs(n_dir) (nif_sobolev_query = one direction), B64 = B rounded up to a multiple of 64; ws: [ws_floats].
 * SIREN ShapeNets served by the tensor-core kernels (kernel_path 2, no res-blocks) with zdot == NULL run the whole pair on
 * them: the tangent forward as a mode of the forward kernel, both adjoint passes through the reverse kernels. */
int nif_sobolev_query(const nif_desc_t* d, int64_t B, int64_t* save_floats_per_row, int64_t* ws_floats);
int nif_sobolev_query_dirs(const nif_desc_t* d, int64_t B, int32_t n_dir, int64_t* save_floats_per_row,
                           int64_t* ws_floats);
int nif_forward_tangent_save(const nif_desc_t* d, int64_t B, const float* z, const float* x,
                             const float* packed, int32_t n_dir, const float* zdot, const float* xdot,
                             float* u, float* udot, float* save, void* stream);
/* du [B,so] = dL/du, dudot [n_dir,B,so] = dL/d(udot), xdot [n_dir,B,si], zdot [n_dir,B,K] or NULL (no direction moves the
 * latent code); zdot_dirs: bit d set = direction d moves the latent code (the others skip that part of the pass and get
 * dzdot = 0).  dw_h / db_h written (beta == 0) or accumulated (beta == 1), dz [B,K] written, dzdot [n_dir,B,K] written
 * when zdot is given.  nif_sobolev_backward: one ShapeNet-input direction. */
int nif_sobolev_backward_dirs(const nif_desc_t* d, int64_t B, const float* z, const float* x, int32_t n_dir,
                              const float* zdot, uint32_t zdot_dirs, const float* xdot, const float* packed, const float* save,
                              const float* du, const float* dudot, float* dw_h, float* db_h, float beta, float* dz,
                              float* dzdot, float* ws, void* stream);
int nif_sobolev_backward(const nif_desc_t* d, int64_t B, const float* z, const float* x, const float* xdot,
                         const float* packed, const float* save, const float* du, const float* dudot,
                         float* dw_h, float* db_h, float beta, float* dz, float* ws, void* stream);

/* model_x_to_u_given_w (nif/model.py:435-464, 956-986): every row brings its own
 * full weight vector.  x [B,si], w [B,P] (reference column layout), u [B,so]. */
int nif_forward_given_w(const nif_desc_t* d, int64_t B, const float* x, const float* w,
                        float* u, void* stream);

/* Keras 'mse' (+ optional sample_weight) and its reverse pass through the fused
 * path (what Keras train_step does with GradientTape over the graph above).
 * loss   [1]  = sum_b sw[b] * mean_c (u-target)^2 * inv_global_batch   (accumulated, +=)
 * du seed     = 2 (u-target) sw[b] inv_global_batch / so
 * dw_h [K,P], db_h [P]: gradients in the REFERENCE layout; written if beta==0,
 * accumulated (+=) if beta==1.   dz [B,K]: written.
 * `u`, `save` come from nif_forward(..., save) on the same inputs.
 * ws: [grad_ws_floats] scratch. */
int nif_mse_backward(const nif_desc_t* d, int64_t B, const float* z, const float* x,
                     const float* packed, const float* u, const float* save,
                     const float* target, const float* sample_weight, float inv_global_batch,
                     float* loss, float* dw_h, float* db_h, float beta, float* dz,
                     float* ws, void* stream);
/* The same call, and dz_ready_event (a cudaEvent_t, may be NULL) is recorded on `stream` once the kernels that produce dz
 * have been enqueued: the caller can start the ParameterNet trunk's reverse pass on another stream while the
 * weight-gradient kernels of this call run. */
int nif_mse_backward_ev(const nif_desc_t* d, int64_t B, const float* z, const float* x, const float* packed,
                        const float* u, const float* save, const float* target, const float* sample_weight,
                        float inv_global_batch, float* loss, float* dw_h, float* db_h, float beta, float* dz,
                        float* ws, void* dz_ready_event, void* stream);

/* Same reverse pass with a caller-supplied seed du [B,so] (for custom losses). */
int nif_backward(const nif_desc_t* d, int64_t B, const float* z, const float* x,
                 const float* packed, const float* save, const float* du,
                 float* dw_h, float* db_h, float beta, float* dz, float* ws, void* stream);

/* tf.keras.optimizers.Adam update (third-party in the reference):
 * m += (g-m)(1-b1); v += (g*g-v)(1-b2); p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps).
 * Optional kernel regularisers (nif/model.py:107-125): g += l1*sign(p) + 2*l2*p.
 * g_scale multiplies g first (e.g. 1/world_size).  lr, b1, b2, eps are doubles because Keras forms
 * 1-b1, 1-b2 and the bias-corrected step in Python floats before casting to the variable dtype. */
int nif_adam_step(int64_t n, float* p, const float* g, float* m, float* v, double lr,
                  double b1, double b2, double eps, int64_t t, float l1, float l2, float g_scale,
                  void* stream);

/* The same update with the bias-corrected step size  alpha = lr*sqrt(1-b2^t)/(1-b1^t)  read from DEVICE memory
 * (one float, formed by the caller in double like nif_adam_step does).  Nothing that changes from step to step is a
 * launch argument, so a whole optimisation step (trunk, forward, loss, reverse passes, this update) can be captured
 * once in a CUDA graph and replayed: the caller rewrites *alpha_dev (stream-ordered) before every replay. */
int nif_adam_step_dev(int64_t n, float* p, const float* g, float* m, float* v, const float* alpha_dev,
                      double b1, double b2, double eps, float l1, float l2, float g_scale, void* stream);

/* AdaBeliefOptimizer (nif/optimizers/external_optimizers.py:321-628; dense update :458-528), the optimiser tutorials 1
 * and 3 train with:  m = b1 m + (1-b1) g;  v = b2 v + (1-b2)(g-m)^2 + eps;  [amsgrad: vhat = max(vhat, v)];
 * p -= lr_t * ( rectify ? (sma_t >= sma_threshold ? r_t m^/(sqrt(v^)+eps) : m^) : m^/(sqrt(v^)+eps)  + weight_decay * p ),
 * m^ = m/(1-b1^t), v^ = v/(1-b2^t).  lr_t is the caller's (warm-up adjusted) learning rate of step t (t >= 1); vhat may be
 * NULL (amsgrad off).  l1 / l2 / g_scale as in nif_adam_step.  Buffers need 4-byte alignment only. */
int nif_adabelief_step(int64_t n, float* p, const float* g, float* m, float* v, float* vhat, double lr_t, double b1,
                       double b2, double eps, int64_t t, int32_t rectify, double sma_threshold, double weight_decay,
                       float l1, float l2, float g_scale, void* stream);

/* Lion (nif/optimizers/external_optimizers.py:631-735; dense update :681-702):
 * p -= lr * (sign(b1 m + (1-b1) g) + wd * p);  m = b2 m + (1-b2) g. */
int nif_lion_step(int64_t n, float* p, const float* g, float* m, double lr, double b1, double b2, double wd, float l1,
                  float l2, float g_scale, void* stream);

/* Gradient centralisation (nif/optimizers/gtcf.py:27-32): for a gradient of rank >= 2 viewed as [rows, cols] (cols = its
 * last axis), subtract from every column its mean over the rows, in place. */
int nif_centralize_gradient(int64_t rows, int64_t cols, float* g, void* stream);

/* Data-parallel Adam fused with its collective over NVSwitch multicast memory (replaces tf.distribute.MirroredStrategy's
 * all-reduce + per-replica update, README.md:39-49): for this rank's 1/world slice of the flat buffers, the gradient is
 * read with multimem.ld_reduce.add from the MULTICAST address g_mc of the replicas' symmetric gradient buffers, the tf.keras
 * Adam update of nif_adam_step runs on that slice (moments m, v are only maintained for the own slice), and the new
 * parameters are written with multimem.st through p_mc into every replica.  p is this rank's own parameter buffer.
 * n: floats of the buffers (multiple of 4).  alpha_dev != NULL: bias-corrected step size from device memory (as
 * nif_adam_step_dev), else formed from lr and t.  The caller puts a cross-rank barrier before (every replica's gradient
 * is complete) and after (every slice has landed) the call. */
int nif_adam_step_multimem(int64_t n, int32_t rank, int32_t world, float* p_mc, const float* g_mc, const float* p, float* m,
                           float* v, double lr, const float* alpha_dev, double b1, double b2, double eps, int64_t t,
                           float l1, float l2, float g_scale, void* stream);

/* ParameterNet trunk (everything before the last linear layer): Dense(act) -> nlayers x MLP_SimpleShortCut ->
 * Dense(latent), i.e. _call_parameter_net up to the bottleneck (nif/model.py:326-343, 176-216, 668-720;
 * nif/layers/mlp.py:148-160).  theta is the trunk's weight vector in the column order
 *   [ W_first (pi x units) | W_hidden[i] (units x units) ... | W_bottleneck (units x latent)
 *     | b_first | b_hidden[i] ... | b_bottleneck ],  every matrix row-major [in, out].
 * act: NIF_ACT_* except SINE (SIREN trunks are outside these kernels); units <= 64. */
typedef struct nif_trunk_desc {
  int32_t pi, latent, units, nlayers, act;
} nif_trunk_desc_t;

/* n_theta: floats of theta; save_floats_per_row: stash per row for the reverse pass; packed_floats: scratch for the
 * re-laid weights (written by forward, read by backward of the same step); ws_floats: reverse scratch for batch B. */
int nif_trunk_query(const nif_trunk_desc_t* d, int64_t B, int64_t* n_theta, int64_t* save_floats_per_row,
                    int64_t* packed_floats, int64_t* ws_floats);
/* Which kernels serve this trunk: 3 = tcgen05 tensor cores, bf16x3 split (fp32-grade; units <= 64, latent <= 64, 1..5
 * hidden layers), 0 = fp32 CUDA-core tile kernels.  The stash then holds save_floats_per_row * B64 floats, B64 = B rounded
 * up to a multiple of 64. */
int nif_trunk_kernel_path(const nif_trunk_desc_t* d);
/* z [B,latent] = trunk(p_in [B,pi]); save may be NULL for inference. */
int nif_trunk_forward(const nif_trunk_desc_t* d, int64_t B, const float* p_in, const float* theta, float* z,
                      float* save, float* packed, void* stream);
/* g_theta (same layout as theta) = d loss / d theta given dz [B,latent]; written if beta == 0, accumulated if 1. */
int nif_trunk_backward(const nif_trunk_desc_t* d, int64_t B, const float* p_in, const float* theta,
                       const float* save, const float* dz, float* g_theta, float beta, const float* packed,
                       float* ws, void* stream);

/* Wide ParameterNet trunks (units > 64) under mixed_bfloat16: the Dense products are library GEMMs on bf16 operands;
 * these two entry points are everything between two GEMMs of a layer, fused, with the rounding points of the policy
 * (bf16 GEMM operands and products, fp32 elsewhere; nif/model.py:101-105, 326-343, nif/layers/mlp.py:148-160).
 *   forward :  h_out [B,n] = (h_in ? h_in : 0) + act( float(y_bf16 [B,n]) + bias [n] ),  h_out_bf16 = bf16(h_out) (optional)
 *   backward:  t = (dh_in ? dh_in : 0) + (pend_bf16 ? float(pend_bf16) : 0)      -- pend: the product g_above @ W_above^T
 *              g_bf16 = bf16(t * act'(float(y_bf16) + bias)),  db [n] = sum_b t * act'(...) (fp32),  dh_out = t (optional,
 *              may alias dh_in)
 * n: a multiple of 8 that divides 2048, <= 256; act: NIF_ACT_* except SINE; ws: nif_trunk_ew_ws_floats(n) floats. */
int nif_trunk_ew_forward(int64_t B, int32_t n, int32_t act, const void* y_bf16, const float* bias, const float* h_in,
                         float* h_out, void* h_out_bf16, void* stream);
int nif_trunk_ew_ws_floats(int32_t n);
int nif_trunk_ew_backward(int64_t B, int32_t n, int32_t act, const void* y_bf16, const float* bias, const float* dh_in,
                          const void* pend_bf16, float* dh_out, void* g_bf16, float* db, float* ws, void* stream);

/* Utility used by the benchmark: sustained FP32 FMA rate of this GPU (TFLOP/s),
 * measured with CUDA events; blocks until done. */
int nif_measure_fp32_peak(double* tflops);

/* Per-kernel timing for the benchmark's kernel table.  Between nif_profile_begin() and nif_profile_end() every kernel
 * the library launches is bracketed by two CUDA events on its own stream (eager launches only: do not open a profile
 * while capturing a CUDA graph).  nif_profile_end blocks until that work is done and writes one line per kernel name,
 * "name launches total_ms\n" in order of first launch, into the HOST buffer `out` (NUL-terminated, cut at cap). */
int nif_profile_begin(void);
int nif_profile_end(char* out, int64_t cap);

/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 for a fresh checksum): the checksum of the TFRecord
 * framing (nif/data/tfr_dataset.py:84-88, 159 use tf.io.TFRecordWriter / tf.data.TFRecordDataset, third-party).  Host
 * code only; the one entry point that takes host pointers. */
uint32_t nif_crc32c(const void* data, uint64_t n, uint32_t crc);

#ifdef __cplusplus
}
#endif
#endif
