/* bfilt.h -- C ABI of libbfilt.so: B200 (sm_100a) kernels for the gaussfiltax filtering hot path.
 *
 * The reference (kostastsa/BayesianFiltering, package `gaussfiltax`) is pure Python/JAX and has no
 * FFI of its own; the boundary these entry points replace is the body of its two filter functions:
 *
 *   bf_pf_*   <->  gaussfiltax/inference.py:1302-1380  bootstrap_particle_filter (lax.scan of _step,
 *                  :1330-1366) + gaussfiltax/utils.py:207-214 _resample + models.py:82-84
 *   bf_gsf_*  <->  gaussfiltax/inference.py:303-377   gaussian_sum_filter (lax.scan of _step, :333-365)
 *                  + _condition_on :72-105, _predict :51-70, utils.py:256-259 psd_solve
 *   bf_resample / bf_cumsum  <->  jax.random.choice(p=w) as called at utils.py:210
 *                  (cumsum in lax.associative_scan order, fixed-bisection searchsorted)
 *
 * Plain pointers and sizes only.  Unless an entry point's name ends in `_host`, every data pointer is
 * a DEVICE pointer owned by the caller, all work is enqueued on the caller's stream, and nothing
 * synchronises the device.  `*_host` variants take HOST pointers, do the H2D/D2H copies themselves and
 * return after the stream has drained.  All entry points return 0 or a negative bf_status; the message
 * of the last failure on the calling thread is bf_last_error().
 *
 * Python callables cannot run inside a kernel, so models come from a compiled registry
 * (bf_model_id); the arrays of the reference's ParamsNLSSM / ParamsBPF travel in bf_model_desc.
 */
#ifndef BFILT_H
#define BFILT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BF_ABI_VERSION 4

typedef enum bf_status {
  BF_OK = 0, BF_EINVAL = -1, BF_ECUDA = -2, BF_ENOMODEL = -3, BF_ENOMEM = -4, BF_ECOMM = -5
} bf_status;

/* compiled model registry; theta layouts are documented in csrc/models.cuh and oracle/models.py */
typedef enum bf_model_id {
  BF_MODEL_UNGM = 1,  /* dx=dq=dy=1           theta = a b c d e                       */
  BF_MODEL_BOT = 2,   /* dx=4 dq=2 dy=1|2     theta = dt acc omega_scale              */
  BF_MODEL_L63 = 3,   /* dx=dq=3 dy=1         theta = sigma rho beta dt c             */
  BF_MODEL_L96 = 4,   /* dx=dq=d dy=d/2       theta = F dt c1 c2                      */
  BF_MODEL_LG = 5,    /* dx=dq<=8 dy=dr<=8    theta = A[dx*dx] H[dy*dx] (row-major)   */
  BF_MODEL_SIN1D = 6, /* dx=dq<=3 dy=1        theta = w c                             */
  BF_MODEL_MSV = 7    /* dx=dq=dy=dr=d<=3     theta = Phi[d] beta[d] sigma H0[d] (diagonals): multiplicative emission
                         noise, state-dependent emission covariance (adaptive_experiment.py:47-57) */
} bf_model_id;

typedef enum bf_resampler { BF_RESAMPLE_MULTINOMIAL = 0, BF_RESAMPLE_SYSTEMATIC = 1 } bf_resampler;
typedef enum bf_pf_path { BF_PATH_AUTO = 0, BF_PATH_RESIDENT = 1, BF_PATH_STREAMED = 2 } bf_pf_path;

/* The array fields of ParamsNLSSM / ParamsBPF (gaussfiltax/models.py:26-84).  HOST pointers, fp32,
 * row-major, copied at create time. */
typedef struct bf_model_desc {
  int32_t model_id;
  int32_t dx, dq, dy, dr;
  int32_t n_theta;
  const float* theta;
  const float* q0;   /* [dq]     dynamics_noise_bias        */
  const float* Q;    /* [dq,dq]  dynamics_noise_covariance  */
  const float* r0;   /* [dr]     emission_noise_bias        */
  const float* R;    /* [dr,dr]  emission_noise_covariance  */
  const float* m0;   /* [dx]     initial_mean               */
  const float* P0;   /* [dx,dx]  initial_covariance         */
} bf_model_desc;

/* ---------------------------------------------------------------- bootstrap particle filter */
typedef struct bf_pf_s* bf_pf_handle;

typedef struct bf_pf_config {
  int64_t num_particles;     /* N per trial (per rank when world > 1)                          */
  int32_t num_trials;        /* B independent filters                                          */
  float ess_threshold;       /* resample when 1/sum(w^2) < ess_threshold*N  (inference.py:1356) */
  int32_t resampler;         /* bf_resampler                                                   */
  int32_t path;              /* bf_pf_path: RESIDENT = one CTA per trial, state in shared memory;
                                STREAMED = one persistent cooperative kernel, state in HBM      */
  uint64_t seed;             /* Philox key for every random input that is not supplied          */
  int32_t device;            /* CUDA device ordinal                                             */
  int32_t trial_offset;      /* global index of this handle's trial 0: Philox counters use trial_offset + b, so a block of trials
                                run on another rank draws the same numbers as in a single-handle run (0 = standalone) */
  /* One filter sharded over the GPUs of a box (one process per GPU): this rank owns num_particles consecutive
   * particles of a filter of shard_world * num_particles.  0/0 or x/1 = not sharded.  See bf_pf_shard_* below. */
  int32_t shard_rank, shard_world;
} bf_pf_config;

typedef struct bf_pf_io {
  int32_t T;                    /* steps in this call                                               */
  int32_t t_offset;             /* global index of the first step (Philox counters, chunked runs)   */
  int32_t resume;               /* 0: start from init_* ; 1: continue from the handle's carry        */
  int32_t emissions_per_trial;  /* 0: emissions is [T,dy] shared by all trials ; 1: [B,T,dy]         */
  const float* emissions;
  const float* inputs;          /* [T] or NULL (zeros, inference.py:23)                              */
  const float* init_particles;  /* [B,N,dx] or NULL (drawn from N(m0,P0) with Philox)                */
  const float* init_weights;    /* [B,N] or NULL (1/N, inference.py:1371)                            */
  const float* noise;           /* [T,B,N,dq] coloured process noise q0+L_Q z, or NULL (Philox)      */
  const float* resample_u;      /* multinomial [T,B,N], systematic [T,B]; or NULL (Philox)           */
  /* outputs, each may be NULL.  Time-major; the Python layer returns [N,T,...] views. */
  float* out_particles;         /* [B,T,N,dx] post-resample (inference.py:1359-1362)                */
  float* out_weights;           /* [B,T,N]    post-resample                                          */
  float* out_propagated;        /* [B,T,N,dx] pre-resample particles                                 */
  float* out_norm_weights;      /* [B,T,N]    pre-resample normalised weights                        */
  int32_t* out_ancestors;       /* [B,T,N]    identity on steps that did not resample                */
  float* out_mean;              /* [B,T,dx]   sum_i w_i x_i before resampling                        */
  float* out_loglik;            /* [B,T]      log p(y_t | y_1:t-1) estimate                          */
  float* out_ess;               /* [B,T]                                                             */
  int32_t* out_resampled;       /* [B,T]      1 if the step resampled                                */
  float* final_particles;       /* [B,N,dx]   carry after the last step                              */
  float* final_weights;         /* [B,N]                                                             */
} bf_pf_io;

int bf_pf_create(const bf_model_desc* model, const bf_pf_config* cfg, bf_pf_handle* out);
int bf_pf_run(bf_pf_handle h, const bf_pf_io* io, void* cuda_stream);
int bf_pf_step(bf_pf_handle h, const bf_pf_io* io, void* cuda_stream); /* bf_pf_run with T forced to 1 */
int bf_pf_run_host(bf_pf_handle h, const bf_pf_io* host_io);
int bf_pf_kernel_launches(bf_pf_handle h, int64_t* out);   /* launches enqueued so far by this handle */
/* Streamed path only: wall time (ms, device globaltimer as seen by CTA 0) accumulated in the phases of the
 * persistent kernel since the last reset -- [0] propagate+ancestor search, [1] statistics+weights+tile sums,
 * [2] scan of the tile sums, [3] in-tile scan -> cdf, [4] final gather.  Synchronises the device. */
int bf_pf_phase_times(bf_pf_handle h, double* out_ms5, int32_t reset);
int bf_pf_destroy(bf_pf_handle h);

/* ---- one particle filter sharded over `shard_world` GPUs (power of two <= 8, one process per GPU).
 * The reference has no multi-device code; this is `bootstrap_particle_filter` (inference.py:1302-1380) for an N that
 * one GPU should not hold alone.  num_particles (per rank) must be a power of two >= 2048, one trial, streamed path.
 * Per time step the ranks exchange (1) their CTA partials of (max, sum, sum^2, sum x) -- every rank folds all of them
 * in the same order, so the ESS test takes the same branch everywhere -- (2) the tile sums and first weights of their
 * tiles, from which every rank derives the same canonical scan offsets, and (3) a barrier before the next step reads
 * the peers' cdf (tau) and particle state through CUDA-IPC mapped pointers: ancestors that live on another rank are
 * gathered directly from its memory over NVLink.  Default transport (BF_SHARD_MODE=p2p): ONE persistent kernel per
 * run; a rank stores its contribution straight into every peer's exchange buffer and raises a sequence-numbered flag
 * there, readers spin on their own flags (bounded: a missing peer makes the run fail with BF_ECOMM instead of
 * hanging).  BF_SHARD_MODE=nccl: one launch per phase with ncclAllGather / ncclAllReduce between them.
 * Protocol: create on every rank -> bf_pf_shard_export -> exchange the blobs between the processes (any transport)
 * -> bf_pf_shard_import; rank 0 bf_nccl_unique_id -> broadcast -> bf_pf_shard_comm_init on every rank; then bf_pf_run /
 * bf_pf_run_host as usual with the LOCAL shard of every per-particle array (resume / t_offset are not supported). */
#define BF_SHARD_BLOB_BYTES 512
#define BF_NCCL_ID_BYTES 128
int bf_pf_shard_export(bf_pf_handle h, void* blob /* BF_SHARD_BLOB_BYTES */);
int bf_pf_shard_import(bf_pf_handle h, const void* blobs /* shard_world * BF_SHARD_BLOB_BYTES, rank order */);
int bf_nccl_unique_id(void* id /* BF_NCCL_ID_BYTES */);
int bf_pf_shard_comm_init(bf_pf_handle h, const void* id /* BF_NCCL_ID_BYTES */);
/* 1 if the handle has a communicator -- one created for an earlier handle of the same (device, world, rank) is reused
 * for the life of the process, so only the first sharded handle pays for ncclCommInitRank */
int bf_pf_shard_comm_ready(bf_pf_handle h);

/* ---------------------------------------------------------------- Gaussian-sum (EKF) filter */
typedef struct bf_gsf_s* bf_gsf_handle;

/* flags: the legacy NumPy classes (gaussfiltax/gausssumfilt.py:30-78, gaussfilt.py:88-130, :325-361) run the same
 * per-component EKF in a different order and with their own quirks */
#define BF_GSF_PREDICT_FIRST 1    /* init_* are FILTERED moments: predict once before the first update (gausssumfilt.py:55-59) */
#define BF_GSF_NO_JITTER 2        /* gain from S itself (legacy np.linalg.inv(Sy), :67), not S + 1e-6 (utils.py:258)           */
#define BF_GSF_COV_PLUS_P 4       /* predict with cov + F cov F^T -- adds cov, never Q (gausssumfilt.py:59, sic)               */

typedef struct bf_gsf_config {
  int32_t num_components;   /* M */
  int32_t num_trials;       /* B */
  int32_t device;
  int32_t flags;            /* BF_GSF_* ; 0 = gaussfiltax/inference.py:303-377 */
} bf_gsf_config;

typedef struct bf_gsf_io {
  int32_t T;
  int32_t resume;               /* 0: start from init_* ; 1: continue from the handle's carry */
  int32_t emissions_per_trial;  /* 0: [T,dy] shared ; 1: [B,T,dy] */
  int32_t reserved;
  const float* emissions;
  const float* inputs;          /* [T] or NULL */
  const float* init_means;      /* [B,M,dx]  (reference draws MVN(m0,P0), inference.py:367) */
  const float* init_covs;       /* [B,M,dx,dx] or NULL (P0 replicated, inference.py:368)    */
  const float* init_weights;    /* [B,M] or NULL (1/M, inference.py:369)                    */
  /* outputs, each may be NULL; time-major, the Python layer returns [M,T,...] views */
  float* out_weights;           /* [B,T,M]        */
  float* out_means;             /* [B,T,M,dx]     */
  float* out_covs;              /* [B,T,M,dx,dx]  */
  float* out_pred_means;        /* [B,T,M,dx]     */
  float* out_pred_covs;         /* [B,T,M,dx,dx]  */
  float* out_loglik;            /* [B,T]          */
  float* out_mix_mean;          /* [B,T,dx]  sum_m w_m mu_m (posterior mean of the mixture) */
  /* time-varying noise covariances: `_get_params` (inference.py:21, :337-340) indexes a covariance that carries a
     leading T axis by the step.  NULL = the constant matrix of the model description.  (The biases are never
     time-indexed by the reference: `_get_params(x, 2, t)` needs ndim 3.) */
  const float* Q_t;             /* [T,dq,dq] or NULL */
  const float* R_t;             /* [T,dr,dr] or NULL */
  /* legacy Monte-Carlo filter only (bf_gsf_set_legacy_moments, kind MCF): the draws of every moment approximation as
     offsets x - m ~ N(0, P): per step and component first the `num_particles` prediction draws, then the update draws.
     NULL = drawn in the kernel (Cholesky of P times Philox normals). */
  const float* mc_draws;        /* [B,T,M,2,num_particles,dx] or NULL */
} bf_gsf_io;

int bf_gsf_create(const bf_model_desc* model, const bf_gsf_config* cfg, bf_gsf_handle* out);
int bf_gsf_run(bf_gsf_handle h, const bf_gsf_io* io, void* cuda_stream);
int bf_gsf_step(bf_gsf_handle h, const bf_gsf_io* io, void* cuda_stream); /* bf_gsf_run with T forced to 1: one `_step` of the
                                  scan (inference.py:333-365); with BF_GSF_PREDICT_FIRST and one component it is `_kalman_step` (:107-120) */
int bf_gsf_run_host(bf_gsf_handle h, const bf_gsf_io* host_io);
/* Switch the handle to the UNSCENTED Gaussian-sum filter (gaussfiltax/inference.py:379-456 `unscented_gaussian_sum_filter`,
 * `_ukf_condition_on_nonadditive` :198-224, `_ukf_predict_nonadditive` :146-174, `utils._get_sigma_points` utils.py:247-254):
 * every later bf_gsf_run / _step / _run_host propagates 2 (dx + dr) and 2 (dx + dq) sigma points of the noise-augmented
 * state per component instead of linearising; alpha, beta, kappa are the fields of the reference's ParamsUKF (:41-49).
 * Models with dx <= 4; constant noise covariances (Q_t / R_t must be NULL). */
int bf_gsf_set_unscented(bf_gsf_handle h, float alpha, float beta, float kappa);
/* Switch the handle to one of the LEGACY moment approximations of gaussfiltax/gaussfilt.py behind GaussFilt.run (:88-130) and
 * gaussfilt.GaussSumFilt.run (:325-361): additive noise (dq = dx), single-argument f(x) = f(x,0,0), g(x) = h(x,0,0).
 *   BF_LEGACY_UKF  UKF.moment_approx (:146-170): 2 dx + 1 sigma points from chol(P) (utils.split_to_sigma_points, utils.py:52-65)
 *   BF_LEGACY_MCF  MCF.moment_approx (:186-198): num_particles Monte-Carlo draws of N(m, P) per moment approximation
 * (EKF.moment_approx is the default kernel with BF_GSF_PREDICT_FIRST | BF_GSF_NO_JITTER; MCLAF.moment_approx cannot run in the
 * reference either: gaussfilt.py:286 calls utils.sdp_opt with 8 arguments, utils.py:110 takes 7.)
 * Create the handle with BF_GSF_PREDICT_FIRST | BF_GSF_NO_JITTER for the classes' predict-then-update order and plain inverse. */
#define BF_LEGACY_UKF 1
#define BF_LEGACY_MCF 2
int bf_gsf_set_legacy_moments(bf_gsf_handle h, int32_t kind, float alpha, float beta, float kappa, int32_t num_particles, uint64_t seed);
int bf_gsf_kernel_launches(bf_gsf_handle h, int64_t* out);
int bf_gsf_destroy(bf_gsf_handle h);

/* ---------------------------------------------------------------- augmented Gaussian-sum filter
 * gaussfiltax/inference.py:621-812 `speedy_augmented_gaussian_sum_filter` (and, for the same supplied draws, :458-619
 * `augmented_gaussian_sum_filter` with gaussfiltax/containers.py:63-161 `_branches_from_tree1/2`): every component is split
 * into N1 Gaussian particles of covariance a0 P, each is predicted (EKF, :51-70), split again into N2 of covariance a1 P^-,
 * each of the M N1 N2 leaves is updated (EKF, :72-105), reweighted by its evidence, and M survivors are drawn with
 * jax.random.choice(p = weights) -- the particle filter's resampler, bit-exact ancestors. */
typedef struct bf_agsf_s* bf_agsf_handle;

typedef struct bf_agsf_config {
  int32_t num_components;   /* M   num_components[0] */
  int32_t num_branch1;      /* N1  num_components[1] */
  int32_t num_branch2;      /* N2  num_components[2] */
  int32_t num_trials;       /* B independent filters */
  int32_t device;
  int32_t reserved;
  float alpha1, alpha2;     /* opt_args: Delta = alpha1 P, Lambda = alpha2 P^-, both in [0,1)  (:666, :718) */
  uint64_t seed;            /* Philox key for draws that are not supplied */
} bf_agsf_config;

typedef struct bf_agsf_io {
  int32_t T;
  int32_t emissions_per_trial;  /* 0: [T,dy] shared ; 1: [B,T,dy] */
  int32_t rand_per_step;        /* 0: one set of draws reused at every step (the reference: its keys are closure constants,
                                   :677, :718, :764) ; 1: z_iid / s_iid / res_u carry a leading T axis */
  int32_t reserved;
  const float* emissions;
  const float* inputs;          /* [T] or NULL */
  const float* init_means;      /* [B,M,dx]  (reference: MVN(m0,P0).sample(M, PRNGKey(0)), :795) */
  const float* init_covs;       /* [B,M,dx,dx] or NULL (P0 replicated, :796) */
  const float* z_iid;           /* [(T,)M,dx,N1]     standard normals of branch 1 (:679), or NULL (Philox) */
  const float* s_iid;           /* [(T,)M*N1,dx,N2]  standard normals of branch 2 (:722), or NULL */
  const float* res_u;           /* [(T,)M]           uniforms of jr.choice (:764), or NULL */
  float* out_weights;           /* [B,T,M]        = 1/M (:768) */
  float* out_means;             /* [B,T,M,dx]     */
  float* out_covs;              /* [B,T,M,dx,dx]  */
  float* out_loglik;            /* [B,T]          max + log(sum) of the leaf evidences */
  float* out_pre_weights;       /* [B,T,M*N1*N2]  aux_outputs["pre_weights"]   (:776) */
  float* out_updated_means;     /* [B,T,M*N1*N2,dx] aux_outputs["updated_means"] (:775) */
  int32_t* out_ancestors;       /* [B,T,M]        leaf index of each survivor */
} bf_agsf_io;

int bf_agsf_create(const bf_model_desc* model, const bf_agsf_config* cfg, bf_agsf_handle* out);
int bf_agsf_run(bf_agsf_handle h, const bf_agsf_io* io, void* cuda_stream);
int bf_agsf_run_host(bf_agsf_handle h, const bf_agsf_io* host_io);
int bf_agsf_destroy(bf_agsf_handle h);

/* gaussfiltax/utils.py:216-244 `optimal_resampling` (Fearnhead & Clifford 2003): from M weights (DEVICE, 2 <= N <= M <= 4096)
 * return N (index, weight) pairs -- the largest weights kept exactly, the rest resampled with a common threshold weight.
 * u [M]: the uniforms of its jr.choice(key, M, (M,), p=res_weights). */
int bf_optimal_resampling(const float* w, int32_t M, int32_t N, const float* u, int32_t* idx_out, float* w_out, void* cuda_stream);

/* ---------------------------------------------------------------- stand-alone resampler pieces */
/* cdf[i] = inclusive prefix sum of w in lax.associative_scan order; bit-exact vs oracle/scan.py. */
int bf_cumsum(const float* w, float* cdf, int64_t n, void* cuda_stream);
/* anc[i] for supplied uniforms: multinomial u[n], systematic u[1]; bit-exact vs oracle/scan.py. */
int bf_resample(const float* w, const float* u, int32_t* anc, int64_t n, int32_t resampler, void* cuda_stream);

/* ---------------------------------------------------------------- model evaluation + data generator */
/* Evaluate a registry model on n states (DEVICE pointers).  what: 0 f(x,q,u) -> [n,dx] (aux = q [n,dq] or NULL=q0);
 * 1 h(x,r,u) -> [n,dy] (aux = r [n,dr] or NULL=r0); 2 dF/dx [n,dx,dx]; 3 dF/dq [n,dx,dq]; 4 dH/dx [n,dy,dx];
 * 5 emission log-prob of one y (aux = y [dy]) -> [n].  Replaces calling the Python f/h/jacfwd(f)
 * (gaussfiltax/inference.py:327-329) on host arrays. */
int bf_model_eval(const bf_model_desc* model, int32_t what, int64_t n, const float* x, const float* aux, float u,
                  float* out, void* cuda_stream);
/* NonlinearSSM.sample (gaussfiltax/models.py:240-289) for B independent trials (DEVICE pointers):
 * states [B,T,dx], emissions [B,T,dy].  z0 [B,dx], q_noise [B,T,dq] (row t=0 unused), r_noise [B,T,dr]
 * are already-coloured draws, or NULL to draw them with Philox(seed).  BF_MODEL_L96 (large state): Philox draws only. */
int bf_ssm_sample(const bf_model_desc* model, int32_t B, int32_t T, const float* inputs, uint64_t seed,
                  const float* z0, const float* q_noise, const float* r_noise, float* states, float* emissions,
                  void* cuda_stream);

/* FP32 (non-tensor) FMA throughput of the device, measured with a register-only FFMA kernel (TFLOP/s): the denominator of
 * the FP32-pipe fractions bench.py reports for the Gaussian-sum filters.  Synchronises the device. */
int bf_measure_fp32_peak(int32_t device, double* tflops);

/* Handles draw their device memory from the device's stream-ordered pool and leave it cached there when they are
 * destroyed (creating a handle per call stays cheap).  This returns the cached memory to the driver. */
int bf_trim_memory(int32_t device);

const char* bf_last_error(void);
int bf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BFILT_H */
