/*
 * hhd.h — C ABI of the B200-native engine for the Hass–Hertäg–Durstewitz PFC network.
 *
 * This is the drop-in boundary.  The reference (Lab-SisNe/ReScience_Hass_Hertag_Durstewitz_2016,
 * paths relative to /root/reference) has no FFI: its façade classes hand equation strings and parameter arrays to
 * Brian 2 and call `Network.run`.  Each entry point below replaces one group of Brian-2 objects the façade builds;
 * the "replaces" note cites the reference call site.  A maintainer binds these with ctypes.
 *
 * Conventions
 *   - every function returns 0 (HHD_OK) or a negative hhd_status; the text is in hhd_last_error(handle)
 *     (hhd_last_error(NULL) gives the last error of a failed hhd_create on this thread);
 *   - all input arrays are HOST pointers owned by the caller and copied during the call; outputs are caller-allocated
 *     HOST buffers unless the function name ends in `_dev` (then: device pointer, e.g. a torch CUDA tensor's data_ptr());
 *   - units are the reference's pre-Brian units (codes/CortexNetwork.py:322-336): ms, mV, pA, nS, pF.  1 nS*mV = 1 pA,
 *     1 pF*mV/ms = 1 pA, so no scale factors appear in the equations;
 *   - one handle = one CUDA device + one stream; a handle is not thread-safe, distinct handles are independent;
 *   - arithmetic type on the path: IEEE fp64, no implicit FMA contraction, `exp` = hhd_exp (csrc/hhd_math.h) so that
 *     the CPU oracle (oracle/hhd_oracle.c) reproduces a neuron's trajectory bit for bit;
 *   - there is NO CPU fallback: hhd_create fails with HHD_E_CUDA when no sm_100 device is usable.
 */
#ifndef HHD_H_
#define HHD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HHD_ABI_VERSION 1

typedef struct hhd_engine* hhd_handle;

typedef enum {
  HHD_OK = 0,
  HHD_E_ARG = -1,      /* bad argument / call order */
  HHD_E_CUDA = -2,     /* CUDA runtime error or no usable device */
  HHD_E_NOMEM = -3,
  HHD_E_CAPACITY = -4, /* a device-side buffer overflowed (spike ring, monitor) */
  HHD_E_STATE = -5,    /* call not legal in this state (e.g. set_synapses after run) */
  HHD_E_COMM = -6      /* NCCL error */
} hhd_status;

/* Integrators.  The reference passes Brian's `method` string (codes/Main.py:32 'rk4', codes/Mainrk2.py:32 'gsl_rk2',
 * codes_revision/CortexSetup.py:36).  euler/rk2/rk4 are Brian's fixed-step ExplicitStateUpdaters
 * (rk2 = midpoint: k1=dt f(x,t); k2=dt f(x+k1/2,t+dt/2); x+=k2).  HHD_RK23_ADAPTIVE is the gsl_rk2 stand-in
 * (embedded RK2(3) with per-neuron sub-stepping inside each dt). */
typedef enum { HHD_EULER = 0, HHD_RK2 = 1, HHD_RK4 = 2, HHD_RK23_ADAPTIVE = 3 } hhd_method;

/* How synaptic increments are summed into the target conductances within one step.
 *   HHD_ACCUM_FP64  : fp64 atomicAdd in arrival order (reference semantics `g_post += ...`; order of simultaneous
 *                     events is unspecified, as it is in Brian across pathways) — trajectories agree with the oracle
 *                     to rounding until the chaotic divergence horizon.
 *   HHD_ACCUM_FIXED : each increment is rounded to a multiple of 2^-HHD_FIXED_FRAC_BITS nS and summed in int64 per
 *                     (target, channel, step), then added to g once — order independent, so spike lists are
 *                     bit-identical to the oracle run in the same mode for any run length. */
typedef enum { HHD_ACCUM_FP64 = 0, HHD_ACCUM_FIXED = 1 } hhd_accum;
#define HHD_FIXED_FRAC_BITS 40

/* Execution plan for hhd_run (same results, different launch structure). */
typedef enum {
  HHD_PLAN_AUTO = 0,
  HHD_PLAN_GRAPH = 1,     /* 2 grid-wide kernels per step (k_neuron, k_synapse), captured in CUDA graphs (any N, HBM-bound regime) */
  HHD_PLAN_RESIDENT = 2,  /* persistent thread-block-cluster kernel, state in registers (N small, latency-bound regime) */
  HHD_PLAN_PERSISTENT = 3,/* removed in round 2 (cooperative whole-run kernel: measured slower); hhd_run returns HHD_E_ARG */
  HHD_PLAN_PIPELINED = 4  /* removed in round 2 (one launch per step: measured no faster); hhd_run returns HHD_E_ARG */
} hhd_plan;

typedef struct {
  int32_t abi_version;        /* = HHD_ABI_VERSION */
  int32_t n_neurons;          /* neurons owned by this handle (local partition when n_ranks > 1) */
  int32_t n_global;           /* neurons in the whole network (= n_neurons when n_ranks == 1) */
  int32_t global_offset;      /* global id of local neuron 0 (contiguous partition) */
  double  dt_ms;              /* NeuronGroup(dt=...)            codes/CortexNetwork.py:309 */
  int32_t method;             /* hhd_method                      codes/CortexNetwork.py:309 */
  int32_t accum;              /* hhd_accum */
  int32_t plan;               /* hhd_plan */
  double  refractory_ms;      /* NeuronGroup(refractory=5*ms)    codes/CortexNetwork.py:309 */
  double  block_window_ms;    /* the `t - last_spike < 5*ms` gate codes/CortexNetwork.py:253 */
  double  last_spike0_ms;     /* group.last_spike = -1e8 ms      codes/CortexNetwork.py:320 (revision: -1000 ms) */
  uint64_t seed;              /* Brian seed(simseed)             codes/CortexNetwork.py:98  (Philox key here) */
  int32_t device;             /* CUDA device ordinal */
  int32_t rank, n_ranks;      /* column partition over GPUs; 0,1 for a single GPU */
  int32_t record_spikes;      /* SpikeMonitor active at start    codes/CortexNetwork.py:487 */
  int32_t instance_size;      /* > 0: the n_neurons are n_neurons/instance_size INDEPENDENT networks of this many neurons
                                 (no synapse crosses an instance): BASELINE config 2, one thread-block cluster each */
  int32_t reserved[6];
} hhd_config;

/* Per-neuron parameter rows of hhd_set_neurons (each a length-n_neurons fp64 array).
 * Same rows as NeuPar (codes/CortexNetwork.py:322-333) without the unused `a`, plus I_DC (:314-318). */
enum {
  HHD_P_C = 0, HHD_P_GL, HHD_P_EL, HHD_P_DELTAT, HHD_P_VUP, HHD_P_TAUW, HHD_P_B, HHD_P_VR, HHD_P_VT,
  HHD_P_IREF, HHD_P_VDEP, HHD_P_IDC, HHD_N_PARAMS
};

/* Channels, in the order of STypPar's columns (codes/ParamSetup.py:318-326). */
enum { HHD_AMPA = 0, HHD_GABA = 1, HHD_NMDA = 2, HHD_N_CHANNELS = 3 };

/* State / derived variables for monitors and hhd_read_state.  Names follow the equation strings
 * codes/CortexNetwork.py:233-302. */
enum {
  HHD_V = 0, HHD_W,
  HHD_G_AMPA_ON, HHD_G_AMPA_OFF, HHD_G_GABA_ON, HHD_G_GABA_OFF, HHD_G_NMDA_ON, HHD_G_NMDA_OFF,   /* 8 ODE states */
  HHD_G_AMPA, HHD_G_GABA, HHD_G_NMDA,                /* off - on */
  HHD_I_AMPA, HHD_I_GABA, HHD_I_NMDA, HHD_I_SYN, HHD_I_INJ, HHD_I_TOT, HHD_I_EXC, HHD_I_EXC_TOT, HHD_I_EXP,
  HHD_LAST_SPIKE,                                     /* the custom `last_spike` variable, ms */
  HHD_N_VARS
};
/* Per-synapse variables for hhd_read_syn_state (in the caller's original synapse order). */
enum { HHD_SYN_U = 0, HHD_SYN_R = 1, HHD_SYN_A = 2 /* amplitude of the next delivery, ((g_max*a_syn)*(1-failure))*Gsyn in nS; 0 before the first presynaptic spike */, HHD_SYN_DELAY_STEPS = 3 };

enum { HHD_REDUCE_NONE = 0, HHD_REDUCE_SUM = 1, HHD_REDUCE_MEAN = 2 };

typedef struct {
  int64_t steps;              /* time steps taken since create */
  int64_t neuron_updates;     /* steps * n_neurons */
  int64_t syn_events;         /* (presynaptic spike, synapse) pairs taken through STP + failure (failed releases count) */
  int64_t syn_deliveries;     /* of those, conductance increments actually delivered so far */
  int64_t ext_events;         /* external-source (spike, synapse) pairs delivered */
  int64_t spikes;             /* threshold crossings (all neurons, whether recorded or not) */
  int64_t w_crossings;        /* `w_crossing` events   codes/CortexNetwork.py:306,310 */
  int64_t nonfinite;          /* neuron-steps whose V or w was NaN/Inf after integration */
  int64_t kernel_launches;    /* GPU kernels launched by hhd_run so far */
  double  run_ms_device;      /* device time of all hhd_run calls (CUDA events on the handle's stream) */
  int64_t reserved[6];
} hhd_counters;

/* ---- lifetime ------------------------------------------------------------------------------------------------- */
int hhd_abi_version(void);
int hhd_create(const hhd_config* cfg, hhd_handle* out);
void hhd_destroy(hhd_handle h);
const char* hhd_last_error(hhd_handle h);

/* ---- model: replaces NeuronGroup(...) + attribute assignments, codes/CortexNetwork.py:308-336 ------------------ */
/* params: HHD_N_PARAMS pointers (rows above).  V0/w0: initial V, w (codes/CortexNetwork.py:335-336).            */
int hhd_set_neurons(hhd_handle h, const double* const* params, const double* V0, const double* w0);
/* has_out[j] != 0 iff GLOBAL neuron j has at least one outgoing synapse anywhere: only those get `last_spike`
 * written (pathway p5 `last_spike_pre = t`, codes/CortexNetwork.py:365).  Optional when n_ranks == 1
 * (derived from the synapse list). */
int hhd_set_has_outgoing(hhd_handle h, const uint8_t* has_out_global);
/* Channel kinetics, replaces the constants formatted into the equation string, codes/CortexNetwork.py:113-131,296.
 * gsyn[c] multiplies every increment of channel c (`Gsyn`, :129-131).  mg_*: NMDA Mg block (:125-127). */
int hhd_set_channels(hhd_handle h, const double tau_on[3], const double tau_off[3], const double E_rev[3],
                     const double gsyn[3], double mg_fac, double mg_slope, double mg_half);

/* Replace the background current row (I_DC, pA) between runs: the per-window drive of a long simulation
 * (`group.I_DC[...] = constant_current` codes/CortexNetwork.py:314-318; step-wise stand-in for TimedArray I_AC :137-151). */
int hhd_set_idc(hhd_handle h, const double* idc);

/* Time-varying injected current I_AC(t): replaces `I_AC = ta(t, i)` with ta = TimedArray(table, dt)
 * (codes/CortexNetwork.py:133-151, 726-727; revision CortexSetup.py:193-213).  table is [n_steps][n_cols] (pA), row s holds
 * for step s (row 0 before it, the last row after it, as TimedArray does); col_of_neuron[i] selects the column of LOCAL
 * neuron i (-1: no timed current).  Runge-Kutta stages evaluated at t + dt read row s + 1.  Call before the first hhd_run. */
int hhd_set_timed_current(hhd_handle h, int64_t n_steps, int32_t n_cols, const double* table, const int32_t* col_of_neuron);

/* Conductance noise: replaces the `noise` variant of the model equations (codes/CortexNetwork.py:157-230, 391-394),
 * `dg_X_off/dt = -g_X_off/tau_off_X + xi_X * sqrt(2 * noise_D * nCon) * nS/sqrt(ms)` for X = AMPA, GABA, NMDA with independent
 * white noises.  sigma[i] = sqrt(2 * noise_D * nCon[i]) in nS/sqrt(ms) per LOCAL neuron; as Brian's stochastic `euler`
 * updater does, every step adds sigma * sqrt(dt) * N(0,1) after the deterministic update.  The Gaussian numbers come from
 * Philox keyed by (seed, global neuron, step) — reproducible, not Brian's Mersenne-Twister stream.  Needs HHD_EULER (Brian
 * refuses the other methods for stochastic equations).  Call before the first hhd_run. */
int hhd_set_noise(hhd_handle h, const double* sigma);

/* ---- synapses: replaces Synapses(...).connect(i,j) + per-synapse assignments + delays/orders,
 *      codes/CortexNetwork.py:343-482 (revision: CortexSetup.py:62-91,239-308) ---------------------------------- */
/* src, tgt are GLOBAL neuron ids; every tgt must be local to this handle.  chan in {0,1,2}.  delay_ms -> steps as
 * floor(delay/dt + 0.5) (Brian's C++ SpikeQueue).  Synapse k keeps index k (= column k of SynPar) for RNG keys and
 * hhd_read_syn_state; syn_id_base is the global index of this handle's synapse 0 (0 for one GPU). */
int hhd_set_synapses(hhd_handle h, int64_t nsyn, const int32_t* src, const int32_t* tgt, const uint8_t* chan,
                     const double* g_max, const double* U, const double* tau_rec, const double* tau_fac,
                     const double* p_fail, const double* delay_ms, int64_t syn_id_base);

/* Same with DEVICE pointers (e.g. torch CUDA tensors): for networks whose synapse arrays are generated on the GPU. */
int hhd_set_synapses_dev(hhd_handle h, int64_t nsyn, const int32_t* src, const int32_t* tgt, const uint8_t* chan,
                         const double* g_max, const double* U, const double* tau_rec, const double* tau_fac,
                         const double* p_fail, const double* delay_ms, int64_t syn_id_base);

/* ---- external input: replaces SpikeGeneratorGroup + input Synapses (no STP, no delay, order 0),
 *      codes/CortexNetwork.py:525-719 (revision CortexSetup.py:370-500) ----------------------------------------- */
/* nspk spikes (source id, time in ms; binned like Brian: step = int((t + 1e-3*dt)/dt)); nsyn input synapses
 * (source id, LOCAL target, channel mask bit0 AMPA bit1 GABA bit2 NMDA, g_max in nS, failure probability).
 * One failure draw per (spike, synapse) gates all channels of that synapse (codes/CortexNetwork.py:588-594). */
int hhd_add_spike_source(hhd_handle h, int32_t n_sources, int64_t nspk, const int32_t* spk_src, const double* spk_t_ms,
                         int64_t nsyn, const int32_t* syn_src, const int32_t* syn_tgt, const uint8_t* syn_chanmask,
                         const double* syn_gmax, const double* syn_pfail);

/* ---- monitors: replaces StateMonitor / SpikeMonitor, codes/CortexNetwork.py:487,499-512;
 *      revision CortexSetup.py:501-607 (long-run monitors with population sum) ---------------------------------- */
/* Records variable `var` of the listed LOCAL neurons (idx == NULL: all) at the start of every step (Brian's `start`
 * slot, before the state update).  reduce != NONE stores one value per step (sum/mean over the listed neurons:
 * the ΣI_tot "LFP", codes_revision/Tests.py:29).  capacity_steps bounds the device buffer. */
int hhd_add_state_monitor(hhd_handle h, int32_t var, int64_t n_idx, const int32_t* idx, int32_t reduce,
                          int64_t capacity_steps, int32_t* mon_id);
/* mon_id >= 0: a state monitor; mon_id == -1: the spike monitor (revision toggles `spikemonitor.active` after the
 * transient, CortexSetup.py:175). */
int hhd_set_monitor_active(hhd_handle h, int32_t mon_id, int32_t active);
/* Samples are [n_samples][n_idx] row-major (n_idx = 1 when reduced).  Reading does not consume. */
int hhd_monitor_samples(hhd_handle h, int32_t mon_id, int64_t* n_samples, int64_t* first_step);
int hhd_read_monitor(hhd_handle h, int32_t mon_id, double* out, int64_t cap_values, int64_t* n_samples);
int hhd_read_monitor_dev(hhd_handle h, int32_t mon_id, const double** dev_ptr, int64_t* n_samples);
int hhd_clear_monitor(hhd_handle h, int32_t mon_id);

/* ---- run: replaces Network.run(t), codes/CortexNetwork.py:729; revision CortexSetup.py:172 -------------------- */
int hhd_run(hhd_handle h, int64_t n_steps);          /* continues from the current step; callable repeatedly */

/* Measurement aid (bench.py roofline): runs n_steps like hhd_run but launches the kernels one by one with a CUDA event
 * between them on the handle's stream and returns the mean device time per launch of each kernel in microseconds:
 * us[0] state update (k_neuron), us[1] the spike exchange when it is the NCCL all-gather fallback (else 0: over peer
 * memory it happens inside the two kernels), us[2] slot `synapses` (k_synapse: spike-time STP + delivery). */
int hhd_profile_kernels(hhd_handle h, int64_t n_steps, double us[3]);

/* ---- results ---------------------------------------------------------------------------------------------------- */
int hhd_spike_count(hhd_handle h, int64_t* n);       /* recorded spikes so far (SpikeMonitor.num_spikes) */
/* Sorted by (step, neuron) like SpikeMonitor.i / .t (t = step*dt).  Neuron ids are GLOBAL. */
int hhd_read_spikes(hhd_handle h, int32_t* neuron, int64_t* step, int64_t cap, int64_t* n_out);
int hhd_read_state(hhd_handle h, int32_t var, double* out /* n_neurons */);
int hhd_write_state(hhd_handle h, int32_t var /* one of the 8 ODE states or HHD_LAST_SPIKE */, const double* in);
int hhd_read_syn_state(hhd_handle h, int32_t which, double* out /* nsyn */);
int hhd_get_counters(hhd_handle h, hhd_counters* out);
int hhd_current_step(hhd_handle h, int64_t* step);

/* ---- multi-GPU: column partition, spike indices all-gathered once per step over NCCL --------- */
int hhd_comm_unique_id(uint8_t out_id[128]);                       /* rank 0, then broadcast by the host program */
int hhd_comm_init(hhd_handle h, const uint8_t id[128]);            /* collective over cfg.n_ranks handles */

/* ---- spike-train analyses on the device --------------------------------------------------------------
 * Stand-alone (no engine handle): spikes as (neuron, t_ms) arrays in time order — t_ms are the very floats the reference's
 * code would see (`SpikeMonitor.t / ms`; step * dt for hhd_read_spikes' steps).  All buffers are host memory.  The results are the integer sufficient statistics of the reference's analyses; the
 * coefficients follow on the host from the reference's formulas (rescience_hass_hertag_durstewitz_2016_b200/analysis.py). */
enum hhd_bin_rule {
  HHD_BIN_TRUNC = 0,    /* ((t - t0) / bin).astype(int)       codes/AuxiliarEquations.py:544 (set_binary_vector) */
  HHD_BIN_FLOORDIV = 1  /* ((t - t0) // bin).astype(int)      codes/AuxiliarEquations.py:434 (pop_rate), codes_revision/AnalysesTools.py:193, 218 */
};
int hhd_an_words(int64_t n_bins);                                  /* uint64 words per bit-packed row: ceil(n_bins / 64) + 1 */
/* Binary spike trains of the selected neurons (row q = neuron sel[q]): bit b of a row is set when the neuron has a spike with
 * t0 <= t < t1 in bin b.  out_bits [n_sel][hhd_an_words(n_bins)], out_bins [n_sel][n_bins] of 0/1 and
 * out_bin_counts [n_bins] (neurons active per bin: pop_rate's column sum) may each be NULL.
 * Replaces set_binary_vector / get_binned_spiking_trains / the loop of pop_rate (file:line above). */
int hhd_an_binned_trains(int32_t device, int64_t n_spikes, const int32_t* neuron, const double* t_ms,
                         int32_t n_sel, const int32_t* sel, double t0_ms, double t1_ms, double bin_ms, int32_t rule,
                         int64_t n_bins, uint64_t* out_bits, uint8_t* out_bins, int32_t* out_bin_counts);
/* For every lag L >= 0 (in bins) and every ORDERED pair (i, j): out_sxy[l][i][j] = sum_{t < n_bins - L} x_i[t] x_j[t + L];
 * out_head[l][i] = sum_{t < n_bins - L} x_i[t], out_tail[l][i] = sum_{t >= L} x_i[t].  Everything Pearson(t_vec1, t_vec2, lag)
 * (codes/AuxiliarEquations.py:571-595) and Series.corr(Series.shift(lag)) (codes_revision/AnalysesTools.py:231-240) need. */
int hhd_an_lagged_counts(int32_t device, int32_t n_sel, int64_t n_bins, const uint64_t* bits, int32_t n_lags, const int32_t* lags,
                         int32_t* out_sxy, int32_t* out_head, int32_t* out_tail);
/* Per selected neuron: number of inter-spike intervals, their mean [ms] and CV = np.std(ISI) / np.mean(ISI) (NaN without
 * intervals), of the spikes with t0 <= t < t1 when has_window.  codes_revision/AnalysesTools.py:421-429 (ISIstats). */
int hhd_an_isi_stats(int32_t device, int64_t n_spikes, const int32_t* neuron, const double* t_ms,
                     int32_t n_sel, const int32_t* sel, int32_t has_window, double t0_ms, double t1_ms,
                     double* out_mean, double* out_cv, int32_t* out_count);

#ifdef __cplusplus
}
#endif
#endif /* HHD_H_ */
