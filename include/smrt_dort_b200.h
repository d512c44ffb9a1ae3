/* smrt_dort_b200.h — C ABI of the B200-native DORT hot path (libsmrt_dort_b200.so).
 *
 * The reference (smrt-model/smrt, pure Python) has NO FFI: the hot path sits behind two Python plugin seams
 *   - the rtsolver seam  DORT(**options).solve(snowpack, emmodels, sensor, atmosphere)   smrt/rtsolver/dort.py:148-161,189
 *     called once per simulation by Model.run_single_simulation                          smrt/core/model.py:584-619
 *   - the runner seam    runner(function, argument_list) -> list[Result]                 smrt/core/model.py:395-398
 * This header is what a binding for that path binds (ctypes stub in INTEGRATION.md): one *batched* call that replaces
 * the per-simulation chain  IBA/DMRT.__init__ -> compute_stream -> compute_interface_properties ->
 * EigenValueSolver.solve -> dort_modem_banded -> sum_modes -> interpolate_intensity  for B problems at once.
 *
 * Conventions
 *   - plain C symbols, no C++ types, no exceptions across the boundary
 *   - every function returns int: 0 = ok, < 0 = API misuse / CUDA error; smrtb200_last_error() gives the message
 *     (thread-local)
 *   - the caller owns every buffer; the library owns only the opaque plan (workspace + streams)
 *   - *_device entry points take DEVICE pointers and are asynchronous on the given cudaStream_t (passed as void*)
 *   - *_host entry points take HOST pointers, stage through pinned memory and return when the results are in the
 *     output buffers
 *   - per-problem numerical failures never abort the batch: they are reported in status[b]
 *   - all floating point data is IEEE fp64; complex numbers are (re, im) pairs of doubles
 */
#ifndef SMRT_DORT_B200_H
#define SMRT_DORT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SMRTB200_ABI_VERSION 4

/* sensor mode (reference smrt/core/sensor.py:331-339) */
#define SMRTB200_MODE_PASSIVE 0
#define SMRTB200_MODE_ACTIVE 1

/* electromagnetic model of a layer */
#define SMRTB200_EM_IBA 0            /* smrt/emmodel/iba.py:53 */
#define SMRTB200_EM_DMRT_QCA_SR 1    /* smrt/emmodel/dmrt_qca_shortrange.py:55 */
#define SMRTB200_EM_NONSCATTERING 2  /* smrt/emmodel/nonscattering.py:17 */
#define SMRTB200_EM_DMRT_QCACP_SR 3  /* smrt/emmodel/dmrt_qcacp_shortrange.py:55 */
#define SMRTB200_EM_RAYLEIGH 4       /* smrt/emmodel/rayleigh.py:17-39; ms_p0 = radius of the microstructure */
#define SMRTB200_EM_PRESCRIBED_KSKAEPS 5 /* smrt/emmodel/prescribed_kskaeps.py:20-27: eps_bg = effective permittivity,
                                            ms_p0 = ks, ms_p1 = ka (ms_kind ignored); Rayleigh phase matrix */
#define SMRTB200_EM_IBA_ORIGINAL 6   /* smrt/emmodel/iba_original.py:15-44: IBA with the absorption of Maetzler 1998 */
#define SMRTB200_EM_IBA_MAXWELL_GARNETT 7 /* smrt/emmodel/iba_maxwell_garnett.py:24-52: Maxwell-Garnett effective
                                            permittivity, apparent permittivity = background */

/* microstructure model (FT of the autocorrelation function) */
#define SMRTB200_MS_EXPONENTIAL 0 /* p0 = corr_length                smrt/microstructure_model/exponential.py:53-58 */
#define SMRTB200_MS_SHS 1         /* p0 = radius, p1 = stickiness    smrt/microstructure_model/sticky_hard_spheres.py:63-167 */
#define SMRTB200_MS_HOMOGENEOUS 2 /* no scatterers (non-scattering layers only) */
#define SMRTB200_MS_INDEPENDENT_SPHERE 3 /* p0 = radius                    smrt/microstructure_model/independent_sphere.py:62-80 */
#define SMRTB200_MS_TEUBNER_STREY 4      /* p0 = corr_length, p1 = repeat_distance   smrt/microstructure_model/teubner_strey.py:53-62 */
#define SMRTB200_MS_UNIFIED_TS_1 5       /* p0 = zeta1, p1 = zeta2, polydispersity >= 1  smrt/microstructure_model/unified_teubner_strey.py:25-36, 64-80 */
#define SMRTB200_MS_UNIFIED_TS_2 6       /* p0 = zeta1, p1 = zeta2, polydispersity < 1 */
#define SMRTB200_MS_SHS_T 7              /* p0 = radius, p1 = t (Percus-Yevick parameter given directly)
                                            smrt/microstructure_model/unified_sticky_hard_spheres.py:21-106 */
/* (unified_scaled_exponential.py is SMRTB200_MS_EXPONENTIAL with corr_length = polydispersity * porod_length) */

/* interface above a layer */
#define SMRTB200_IF_FLAT 0        /* Fresnel, smrt/interface/flat.py:11-75 + smrt/core/fresnel.py:99-146,417-474 */
#define SMRTB200_IF_TRANSPARENT 1 /* smrt/interface/transparent.py:7-49 */
/* moderately rough interfaces: coherent reflection / transmission under the Kirchhoff approximation
 * (smrt/interface/interface_utils.py:16-64) and the IEM backscatter of Fung et al. 1992 as a DIAGONAL diffuse reflection
 * in every azimuth mode (smrt/interface/iem_fung92.py:88-214; no diffuse transmission); parameters in interface_params */
#define SMRTB200_IF_IEM_FUNG92 2
#define SMRTB200_IF_IEM_FUNG92_BRIOGONI10 3 /* Fresnel coefficients at normal incidence when ks kl > sqrt(eps_r):
                                               smrt/interface/iem_fung92_brogioni10.py:31-54 */

/* substrate */
#define SMRTB200_SUB_NONE 0
#define SMRTB200_SUB_FLAT 1 /* flat half-space of permittivity substrate_eps at substrate_temperature, smrt/substrate/flat.py:15-17 */
/* Fresnel coefficients of the half-space with a per-stream adjustment of the V and H power coefficients (the third
 * Stokes component keeps its Fresnel value, as in the reference); substrate_params holds the model parameters */
#define SMRTB200_SUB_SOIL_WEGMULLER 2  /* params[0] = roughness_rms (m)          smrt/substrate/soil_wegmuller.py:20-81 */
#define SMRTB200_SUB_SOIL_QNH 3        /* params = H, Q, Nv, Nh                  smrt/substrate/soil_qnh.py:22-89 */
#define SMRTB200_SUB_REFLECTOR 4       /* params = specular reflection V, H; passive only; substrate_eps unused
                                          smrt/substrate/reflector.py:51-111 (scalar / dict specifications) */
#define SMRTB200_SUB_ROUGH_CHOUDHURY 5 /* params[0] = roughness_rms (m)          smrt/substrate/rough_choudhury79.py:19-79 */
#define SMRTB200_SUB_REFLECTOR_BACKSCATTER 6 /* params = specular reflection V, H, backscattering coefficient VV, HH
                                          (linear); passive and active: the backscatter enters every azimuth mode as a
                                          DIAGONAL diffuse reflection +- sigma0 w / (2 mu (1 + 2 m_max)), absent from the
                                          coherent pass     smrt/substrate/reflector_backscatter.py:66-135,
                                          smrt/rtsolver/rtsolver_utils.py:690-709, 728-740 */
#define SMRTB200_SUB_IEM_FUNG92 7       /* params = roughness_rms (m), corr_length (m), autocorrelation (0 exponential /
                                          1 gaussian), series_truncation; substrate_eps = soil permittivity.  Coherent
                                          part under the Kirchhoff approximation (interface/interface_utils.py:21-64),
                                          IEM backscatter of Fung et al. 1992 as a DIAGONAL diffuse reflection spread
                                          over the azimuth modes       smrt/interface/iem_fung92.py:88-214 */
#define SMRTB200_SUB_IEM_FUNG92_BRIOGONI10 8 /* same parameters; Fresnel coefficients at normal incidence when
                                          ks kl > sqrt(eps_r)          smrt/interface/iem_fung92_brogioni10.py:31-54 */

/* phase_normalization option (smrt/rtsolver/dort.py:94-103, 782-819) */
#define SMRTB200_NORM_OFF 0
#define SMRTB200_NORM_ON 1     /* True / "auto": normalise, error if the correction exceeds 30 % */
#define SMRTB200_NORM_FORCED 2 /* "forced" */

/* status[b]: low 4 bits = error code, bit 4 = warning flag */
#define SMRTB200_OK 0
#define SMRTB200_ERR_NORMALIZATION 1 /* SMRTError of dort.py:792-801 */
#define SMRTB200_ERR_EIGEN 2         /* SMRTError of dort.py:826,844,852,1068-1085 (layer matrix not diagonalisable with real positive spectrum) */
#define SMRTB200_ERR_SINGULAR 3      /* singular boundary block (scipy.linalg.solve_banded would raise, dort.py:469) */
#define SMRTB200_ERR_INPUT 4         /* invalid per-problem input (fewer than 2 streams in a layer, bad enum, SHS t has no solution ...) */
#define SMRTB200_ERR_SUBSTRATE 5     /* substrate model outside its validity range (Warning of rough_choudhury79.py:29-31: k sigma > 0.1) */
#define SMRTB200_ERR_MASK 15
#define SMRTB200_WARN_SHALLOW 16     /* smrt_warn of dort.py:460-467: optically shallow snowpack without substrate */

typedef struct smrtb200_plan smrtb200_plan; /* opaque */

typedef struct {
  int abi_version;     /* SMRTB200_ABI_VERSION */
  int device;          /* CUDA device ordinal */
  int mode;            /* SMRTB200_MODE_* */
  int n_max_stream;    /* DORT option n_max_stream (dort.py:150), 2..256 */
  int m_max;           /* DORT option m_max (dort.py:151); ignored (0) in passive mode; 0..16 */
  int max_layers;      /* L: row stride of every [B, L] array */
  int max_batch;       /* largest B of one solve call (host staging is sized for it) */
  int n_theta;         /* number of viewing angles (passive) */
  int n_inc;           /* number of incidence angles (active; theta == theta_inc for backscatter), <= 8 */
  int normalization;   /* SMRTB200_NORM_* */
  int rayleigh_jeans;  /* 1 = Rayleigh-Jeans approximation: B(T) = T (dort.py:160, rtsolver_utils.py:411-419) */
  double prune_deep_snowpack; /* optical depth beyond which layers are dropped (dort.py:117-120, 444-452); <= 0: off */
  int chunk;           /* problems processed per kernel wave; 0 = automatic */
  int reserved;        /* flags; bit 0: run the chunks one after the other on a single stream (clean per-kernel timings) */
} smrtb200_options;

/* One batch of B independent (snowpack x frequency) problems.  Arrays marked [B, L] have row stride max_layers. */
typedef struct {
  int B;
  /* inputs */
  const double* frequency;            /* [B] Hz */
  const int* nlayer;                  /* [B] 1..L */
  const double* thickness;            /* [B, L] m */
  const double* temperature;          /* [B, L] K */
  const double* frac_volume;          /* [B, L] */
  const double* eps_bg;               /* [B, L, 2] background permittivity  layer.permittivity(0, f)  smrt/core/layer.py:120-156 */
  const double* eps_sc;               /* [B, L, 2] scatterer permittivity   layer.permittivity(1, f) */
  const int* emmodel;                 /* [B, L] SMRTB200_EM_* */
  const int* ms_kind;                 /* [B, L] SMRTB200_MS_* */
  const double* ms_p0;                /* [B, L] */
  const double* ms_p1;                /* [B, L] */
  const int* interface_kind;          /* [B, L] interface ABOVE layer l, SMRTB200_IF_* */
  const int* dense_snow_correction;   /* [B, L] 1 = emmodel option dense_snow_correction="auto" (iba.py:95-96) */
  const int* substrate_kind;          /* [B] SMRTB200_SUB_* */
  const double* substrate_eps;        /* [B, 2] */
  const double* substrate_temperature;/* [B] K; <= 0 means "no temperature" (dort.py:429-441) */
  const double* substrate_params;     /* [B, 4] parameters of SMRTB200_SUB_* kinds >= 2; may be NULL (all zero) */
  const double* atmosphere;           /* [B, 3] isotropic atmosphere (passive mode): tb_down, tb_up (K), transmittance
                                         smrt/atmosphere/simple_isotropic_atmosphere.py:49-77, rtsolver_utils.py:141-147,
                                         302-305; may be NULL = (0, 0, 1) = no atmosphere */
  const double* inclusion;            /* [B, L, 5] shape of the inclusions of a layer: weights of the "spheres" and
                                         "random_needles" solutions in the Polder - van Santen effective permittivity
                                         (layer.inclusion_shape / mixing_ratio: smrt/permittivity/
                                         generic_mixing_formula.py:88-141), then the three depolarisation factors of
                                         the IBA field ratio and of Maxwell-Garnett (layer.depolarization_factors or
                                         length_ratio: smrt/emmodel/iba.py:112-119, smrt/permittivity/
                                         depolarization_factors.py:9-46); may be NULL = (1, 0, 1/3, 1/3, 1/3) */
  const double* interface_params;     /* [B, L, 4] parameters of the interface ABOVE layer l for SMRTB200_IF_* kinds >= 2:
                                         roughness_rms (m), corr_length (m), autocorrelation (0 exponential / 1 gaussian),
                                         series_truncation; may be NULL when every interface is flat or transparent (a
                                         rough kind without parameters is treated as flat) */
  const double* theta;                /* [n_theta] rad, viewing angles (passive) */
  const double* theta_inc;            /* [n_inc] rad, incidence angles (active) */
  double phi;                         /* rad, relative azimuth (active; pi = backscatter) */
  /* outputs */
  double* values;        /* passive: [B, 2, n_theta] brightness temperature (V, H), K
                            active : [B, 3, 3, n_inc] intensity laid out exactly as the reference's Result
                                     (polarization_inc-label, polarization-label, theta_inc), rtsolver_utils.py:309-332;
                                     sigma0 = 4 pi cos(theta) * value (smrt/core/result.py:485) */
  double* ks;            /* [B, L] scattering coefficient   (Result.other_data["ks"], rtsolver_utils.py:373-398) */
  double* ka;            /* [B, L] absorption coefficient */
  double* eps_eff;       /* [B, L, 2] effective permittivity */
  int* n_streams_out;    /* [B] number of valid entries in stream_angles */
  double* stream_angles; /* [B, n_max_stream] degrees; passive: all air streams, active: the incident streams */
  double* optical_depth; /* [B] sum over the layers kept of min|beta| * thickness (dort.py:444) */
  int* status;           /* [B] SMRTB200_OK / SMRTB200_ERR_* | SMRTB200_WARN_* */
} smrtb200_batch;

/* Library / device information. */
int smrtb200_abi_version(void);
const char* smrtb200_last_error(void);
int smrtb200_device_count(int* count);

/* Plan lifetime.  A plan owns the device workspace, two CUDA streams and the pinned staging buffers. */
int smrtb200_plan_create(const smrtb200_options* options, smrtb200_plan** plan);
int smrtb200_plan_destroy(smrtb200_plan* plan);
/* bytes of device workspace held by the plan */
int smrtb200_plan_workspace_bytes(const smrtb200_plan* plan, unsigned long long* bytes);
/* number of kernels launched by the plan since creation (bench.py reports it as gpu_launches) */
int smrtb200_plan_launch_count(const smrtb200_plan* plan, unsigned long long* launches);

/* Solve with DEVICE pointers; asynchronous on `cuda_stream` (a cudaStream_t; NULL = default stream).
 * Replaces, for every b in [0, B): DORT.solve of smrt/rtsolver/dort.py:189-261 including the emmodel construction of
 * smrt/core/model.py:529-582. */
int smrtb200_solve_batch_device(smrtb200_plan* plan, const smrtb200_batch* batch, void* cuda_stream);

/* Solve with HOST pointers (pageable or pinned): H2D copies, kernels and D2H copies, then synchronises. */
int smrtb200_solve_batch_host(smrtb200_plan* plan, const smrtb200_batch* batch);

/* Synchronise `cuda_stream` (the stream given to smrtb200_solve_batch_device) and collect the CUDA-event timings of
 * the last solve.  smrtb200_solve_batch_host does this itself. */
int smrtb200_plan_sync_timing(smrtb200_plan* plan, void* cuda_stream);

/* Time of the last solve on the device, from CUDA events recorded on the launching streams around the kernels
 * (ms; H2D/D2H excluded): total (first launch to last completion), and the summed durations of the per-layer eigen
 * kernel launches and of the boundary kernel launches (one launch of each per chunk; n_chunks = number of launches). */
int smrtb200_plan_last_timing(const smrtb200_plan* plan, float* total_ms, float* eigen_kernel_ms,
                              float* boundary_kernel_ms, int* n_chunks);

/* Micro-benchmark used by bench.py to obtain the FP64 roofline denominator on the box (MEASURED_PEAKS.json has no
 * FP64 entry): runs an unrolled DFMA kernel for ~`ms` milliseconds and returns the achieved TFLOP/s. */
int smrtb200_measure_fp64_peak(int device, float ms, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SMRT_DORT_B200_H */
