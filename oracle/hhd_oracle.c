/*
 * hhd_oracle.c — CPU restatement of the reference's simulation loop.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (rescience_hass_hertag_durstewitz_2016_b200/) never does and has no CPU fallback.
 *
 * What it restates.  The reference (paths relative to /root/reference) contains no simulation loop of its own: it
 * hands equation strings to Brian 2 (pinned 2.3, codes/README.md:24; not installable here) and calls Network.run
 * (codes/CortexNetwork.py:729).  This file executes those strings with Brian 2's documented per-step schedule
 * (SURVEY.md §3.3): start(monitors) -> groups(state update) -> thresholds(spike, w_crossing, generators)
 * -> synapses(pathways by order, SpikeQueue) -> resets -> after_resets(run_on_event).
 *   model equations            codes/CortexNetwork.py:233-302   (deriv() below, term by term)
 *   threshold / reset / event  codes/CortexNetwork.py:304-311
 *   synapse pathways p0..p9b   codes/CortexNetwork.py:360-372, delays/orders :463-482
 *   external input pathways    codes/CortexNetwork.py:588-594, 686-692
 *   monitors                   codes/CortexNetwork.py:487, 499-512
 *
 * PARITY STATUS: step-level parity with Brian 2 is UNPINNED — the reference ships no golden trajectories and Brian 2
 * cannot run in this environment.  What pins this oracle: (i) analytic known answers (closed-form simpAdEx f-I curve
 * of codes/NetworkSetup.py:444-553, hand-computed STP sequences, delay bookkeeping) in tests/test_oracle_*.py,
 * (ii) the reference's statistical targets (content.tex:239,318,334-335,344; codes/Original_spiketime.txt) over 7 network
 * realisations in the long-run tests, and a second, independent restatement (oracle/numpy_oracle.py, tests/test_numpy_oracle.py),
 * (iii) the network inputs themselves, which are bit-exact outputs of the reference's own NetworkSetup (tests/golden).
 *
 * Deliberately written the way Brian works (per-synapse event queue, values read at delivery time, one pass per
 * schedule slot) and NOT the way the CUDA engine works (amplitudes added to a circular per-delay conductance buffer at spike
 * time, corrected when a neuron spikes again while an older spike is in flight), so that
 * agreement between the two is evidence.  Exports the ABI of include/hhd.h under the prefix `hhdo_`.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <time.h>

#define hhd_abi_version hhdo_abi_version
#define hhd_create hhdo_create
#define hhd_destroy hhdo_destroy
#define hhd_last_error hhdo_last_error
#define hhd_set_neurons hhdo_set_neurons
#define hhd_set_has_outgoing hhdo_set_has_outgoing
#define hhd_set_channels hhdo_set_channels
#define hhd_set_synapses hhdo_set_synapses
#define hhd_set_synapses_dev hhdo_set_synapses_dev
#define hhd_add_spike_source hhdo_add_spike_source
#define hhd_add_state_monitor hhdo_add_state_monitor
#define hhd_set_monitor_active hhdo_set_monitor_active
#define hhd_monitor_samples hhdo_monitor_samples
#define hhd_read_monitor hhdo_read_monitor
#define hhd_read_monitor_dev hhdo_read_monitor_dev
#define hhd_clear_monitor hhdo_clear_monitor
#define hhd_run hhdo_run
#define hhd_set_timed_current hhdo_set_timed_current
#define hhd_set_noise hhdo_set_noise
#define hhd_set_idc hhdo_set_idc
#define hhd_profile_kernels hhdo_profile_kernels
#define hhd_spike_count hhdo_spike_count
#define hhd_read_spikes hhdo_read_spikes
#define hhd_read_state hhdo_read_state
#define hhd_write_state hhdo_write_state
#define hhd_read_syn_state hhdo_read_syn_state
#define hhd_get_counters hhdo_get_counters
#define hhd_current_step hhdo_current_step
#define hhd_comm_unique_id hhdo_comm_unique_id
#define hhd_comm_init hhdo_comm_init
#include "../include/hhd.h"
#include "../rescience_hass_hertag_durstewitz_2016_b200/csrc/hhd_math.h"

typedef struct { int64_t* v; int64_t n, cap; } vec64;
typedef struct { double* v; int64_t n, cap; } vecd;

static int vec64_push(vec64* a, int64_t x) {
  if (a->n == a->cap) {
    int64_t nc = a->cap ? a->cap * 2 : 64;
    int64_t* nv = (int64_t*)realloc(a->v, (size_t)nc * sizeof(int64_t));
    if (!nv) return -1;
    a->v = nv; a->cap = nc;
  }
  a->v[a->n++] = x;
  return 0;
}
static int vecd_push(vecd* a, double x) {
  if (a->n == a->cap) {
    int64_t nc = a->cap ? a->cap * 2 : 64;
    double* nv = (double*)realloc(a->v, (size_t)nc * sizeof(double));
    if (!nv) return -1;
    a->v = nv; a->cap = nc;
  }
  a->v[a->n++] = x;
  return 0;
}

typedef struct {
  int32_t var, reduce, active;
  int64_t n_idx;
  int32_t* idx;        /* local neuron ids */
  int64_t first_step;  /* step of sample 0 (-1 if none) */
  vecd samples;        /* [n_samples][n_idx or 1] */
} monitor;

typedef struct {
  int32_t n_sources;
  int64_t nspk; int32_t* spk_src; int64_t* spk_step;   /* sorted by step (stable) */
  int64_t nsyn; int32_t* syn_src; int32_t* syn_tgt; uint8_t* syn_mask; double* syn_gmax; double* syn_pfail;
  int64_t* src_ptr; int64_t* src_syn;                   /* CSR by source */
  int64_t cursor;                                       /* next spike to emit */
  int64_t syn_id_base;                                  /* RNG index base of this source set */
} ext_source;

struct hhd_engine {
  hhd_config cfg;
  int N, NG;
  int64_t step;
  int refr_steps;
  char err[512];
  /* neurons */
  double* par[HHD_N_PARAMS];
  double* x[8];
  double* last_spike;       /* [NG], SECONDS (k * dt_s, as Brian holds it): the model's custom variable, written by pathway p5 */
  double* noise_sigma;      /* [N] nS/sqrt(ms), NULL = no conductance noise */
  double* hstep;            /* [N] adaptive integrator: step size carried from one dt to the next (use_last_timestep) */
  int64_t* lastspike_step;  /* [N], Brian's internal `lastspike` as a step index */
  uint8_t* has_out;         /* [NG] */
  int has_out_given;
  int neurons_set, channels_set, started;
  double tau_on[3], tau_off[3], E[3], gsyn[3], mg_fac, mg_slope, mg_half;
  /* synapses in caller order */
  int64_t nsyn, syn_id_base;
  int32_t *src, *tgt; uint8_t* chan;
  double *gmax, *U, *tau_rec, *tau_fac, *pfail;
  int64_t* delay_steps;
  double *u, *R, *a_syn; uint8_t* failure; uint8_t* touched;
  int64_t* row_ptr; int64_t* row_syn;   /* CSR by GLOBAL source, synapse ids ascending within a row */
  int64_t max_delay;
  vec64* queue;                         /* [max_delay+1] slots of synapse ids: Brian's SpikeQueue */
  double* pendd;                        /* HHD_ACCUM_FP64: [3*N] per-step sums of increments */
  int64_t* pend[2];                     /* HHD_ACCUM_FIXED: [3*N] int64 sums for on/off (identical, kept for clarity) */
  /* inputs */
  ext_source* ext; int n_ext; int64_t ext_syn_total;
  /* monitors */
  monitor* mon; int n_mon;
  int spikes_active;
  vec64 spk_neuron, spk_step;
  /* scratch of the current step */
  int32_t* cur_spikes; int n_cur_spikes;      /* local ids */
  int32_t* cur_wcross; int n_cur_wcross;
  uint8_t* flag_scratch;
  double* iac_table; int32_t* iac_col; int64_t iac_steps; int iac_cols;
  int step_open;
  hhd_counters cnt;
};

static char g_create_err[512];

static int fail(hhd_handle h, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h ? h->err : g_create_err, 512, fmt, ap);
  va_end(ap);
  return code;
}

int hhd_abi_version(void) { return HHD_ABI_VERSION; }
const char* hhd_last_error(hhd_handle h) { return h ? h->err : g_create_err; }

int hhd_create(const hhd_config* cfg, hhd_handle* out) {
  if (!cfg || !out) return fail(NULL, HHD_E_ARG, "null argument");
  if (cfg->abi_version != HHD_ABI_VERSION) return fail(NULL, HHD_E_ARG, "abi_version %d != %d", cfg->abi_version, HHD_ABI_VERSION);
  if (cfg->n_neurons <= 0 || cfg->dt_ms <= 0) return fail(NULL, HHD_E_ARG, "n_neurons and dt_ms must be positive");
  if (cfg->method < HHD_EULER || cfg->method > HHD_RK23_ADAPTIVE) return fail(NULL, HHD_E_ARG, "unknown method %d", cfg->method);
  hhd_handle h = (hhd_handle)calloc(1, sizeof(*h));
  if (!h) return fail(NULL, HHD_E_NOMEM, "out of memory");
  h->cfg = *cfg;
  if (h->cfg.n_ranks <= 0) h->cfg.n_ranks = 1;
  if (h->cfg.n_global <= 0) h->cfg.n_global = cfg->n_neurons;
  h->N = cfg->n_neurons;
  h->NG = h->cfg.n_global;
  if (h->cfg.global_offset < 0 || h->cfg.global_offset + h->N > h->NG) { free(h); return fail(NULL, HHD_E_ARG, "partition out of range"); }
  /* NeuronGroup(refractory=5*ms): not_refractory = timestep(t-lastspike,dt) >= timestep(refractory,dt)  [Brian 2] */
  h->refr_steps = (int)hhd_timestep(cfg->refractory_ms, cfg->dt_ms);
  for (int i = 0; i < HHD_N_PARAMS; ++i) h->par[i] = (double*)calloc((size_t)h->N, sizeof(double));
  for (int i = 0; i < 8; ++i) h->x[i] = (double*)calloc((size_t)h->N, sizeof(double));
  h->last_spike = (double*)malloc((size_t)h->NG * sizeof(double));
  h->hstep = (double*)malloc((size_t)h->N * sizeof(double));
  for (int i = 0; i < h->N; ++i) h->hstep[i] = cfg->dt_ms;
  h->lastspike_step = (int64_t*)malloc((size_t)h->N * sizeof(int64_t));
  h->has_out = (uint8_t*)calloc((size_t)h->NG, 1);
  h->cur_spikes = (int32_t*)malloc((size_t)h->N * sizeof(int32_t));
  h->cur_wcross = (int32_t*)malloc((size_t)h->N * sizeof(int32_t));
  for (int i = 0; i < h->NG; ++i) h->last_spike[i] = cfg->last_spike0_ms * 1e-3;   /* codes/CortexNetwork.py:320 */
  for (int i = 0; i < h->N; ++i) h->lastspike_step[i] = INT64_MIN / 4;       /* Brian: lastspike = -inf */
  h->spikes_active = cfg->record_spikes;
  h->row_ptr = (int64_t*)calloc((size_t)h->NG + 1, sizeof(int64_t));
  h->queue = (vec64*)calloc(1, sizeof(vec64));
  *out = h;
  return HHD_OK;
}

static void free_ext(ext_source* e) {
  free(e->spk_src); free(e->spk_step); free(e->syn_src); free(e->syn_tgt); free(e->syn_mask);
  free(e->syn_gmax); free(e->syn_pfail); free(e->src_ptr); free(e->src_syn);
}

void hhd_destroy(hhd_handle h) {
  if (!h) return;
  for (int i = 0; i < HHD_N_PARAMS; ++i) free(h->par[i]);
  for (int i = 0; i < 8; ++i) free(h->x[i]);
  free(h->last_spike); free(h->hstep); free(h->noise_sigma); free(h->lastspike_step); free(h->has_out); free(h->cur_spikes); free(h->cur_wcross); free(h->flag_scratch); free(h->iac_table); free(h->iac_col);
  free(h->src); free(h->tgt); free(h->chan); free(h->gmax); free(h->U); free(h->tau_rec); free(h->tau_fac);
  free(h->pfail); free(h->delay_steps); free(h->u); free(h->R); free(h->a_syn); free(h->failure); free(h->touched);
  free(h->row_ptr); free(h->row_syn);
  if (h->queue) { for (int64_t s = 0; s <= h->max_delay; ++s) free(h->queue[s].v); free(h->queue); }
  free(h->pend[0]); free(h->pend[1]); free(h->pendd);
  for (int i = 0; i < h->n_ext; ++i) free_ext(&h->ext[i]);
  free(h->ext);
  for (int i = 0; i < h->n_mon; ++i) { free(h->mon[i].idx); free(h->mon[i].samples.v); }
  free(h->mon);
  free(h->spk_neuron.v); free(h->spk_step.v);
  free(h);
}

int hhd_set_neurons(hhd_handle h, const double* const* params, const double* V0, const double* w0) {
  if (!h || !params || !V0 || !w0) return fail(h, HHD_E_ARG, "null argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_neurons after run");
  for (int p = 0; p < HHD_N_PARAMS; ++p) {
    if (!params[p]) return fail(h, HHD_E_ARG, "params[%d] is null", p);
    memcpy(h->par[p], params[p], (size_t)h->N * sizeof(double));
  }
  for (int i = 0; i < h->N; ++i) {
    if (!(h->par[HHD_P_C][i] > 0) || !(h->par[HHD_P_GL][i] > 0) || !(h->par[HHD_P_DELTAT][i] > 0) || !(h->par[HHD_P_TAUW][i] > 0))
      return fail(h, HHD_E_ARG, "neuron %d: C, g_L, delta_T and tau_w must be positive", i);
  }
  memcpy(h->x[HHD_V], V0, (size_t)h->N * sizeof(double));   /* codes/CortexNetwork.py:335 */
  memcpy(h->x[HHD_W], w0, (size_t)h->N * sizeof(double));   /* :336 */
  h->neurons_set = 1;
  return HHD_OK;
}

int hhd_set_idc(hhd_handle h, const double* idc) {
  if (!h || !idc) return fail(h, HHD_E_ARG, "null argument");
  memcpy(h->par[HHD_P_IDC], idc, (size_t)h->N * sizeof(double));
  return HHD_OK;
}

int hhd_set_timed_current(hhd_handle h, int64_t n_steps, int32_t n_cols, const double* table, const int32_t* col_of_neuron) {
  if (!h || n_steps <= 0 || n_cols <= 0 || !table || !col_of_neuron) return fail(h, HHD_E_ARG, "bad argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_timed_current after run");
  for (int i = 0; i < h->N; ++i) if (col_of_neuron[i] >= n_cols) return fail(h, HHD_E_ARG, "column index out of range");
  free(h->iac_table); free(h->iac_col);
  h->iac_table = (double*)malloc((size_t)n_steps * n_cols * 8);
  h->iac_col = (int32_t*)malloc((size_t)h->N * 4);
  memcpy(h->iac_table, table, (size_t)n_steps * n_cols * 8);
  memcpy(h->iac_col, col_of_neuron, (size_t)h->N * 4);
  h->iac_steps = n_steps; h->iac_cols = n_cols;
  return HHD_OK;
}

int hhd_set_noise(hhd_handle h, const double* sigma) {
  if (!h || !sigma) return fail(h, HHD_E_ARG, "null argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_noise after run");
  if (h->cfg.method != HHD_EULER) return fail(h, HHD_E_ARG, "conductance noise needs the euler method (stochastic equations)");
  free(h->noise_sigma);
  h->noise_sigma = (double*)malloc((size_t)h->N * 8);
  memcpy(h->noise_sigma, sigma, (size_t)h->N * 8);
  return HHD_OK;
}

int hhd_set_has_outgoing(hhd_handle h, const uint8_t* has_out_global) {
  if (!h || !has_out_global) return fail(h, HHD_E_ARG, "null argument");
  memcpy(h->has_out, has_out_global, (size_t)h->NG);
  h->has_out_given = 1;
  return HHD_OK;
}

int hhd_set_channels(hhd_handle h, const double tau_on[3], const double tau_off[3], const double E_rev[3],
                     const double gsyn[3], double mg_fac, double mg_slope, double mg_half) {
  if (!h || !tau_on || !tau_off || !E_rev || !gsyn) return fail(h, HHD_E_ARG, "null argument");
  for (int c = 0; c < 3; ++c) {
    if (!(tau_on[c] > 0) || !(tau_off[c] > 0)) return fail(h, HHD_E_ARG, "channel %d: time constants must be positive", c);
    h->tau_on[c] = tau_on[c]; h->tau_off[c] = tau_off[c]; h->E[c] = E_rev[c]; h->gsyn[c] = gsyn[c];
  }
  h->mg_fac = mg_fac; h->mg_slope = mg_slope; h->mg_half = mg_half;
  h->channels_set = 1;
  return HHD_OK;
}

int hhd_set_synapses(hhd_handle h, int64_t nsyn, const int32_t* src, const int32_t* tgt, const uint8_t* chan,
                     const double* g_max, const double* U, const double* tau_rec, const double* tau_fac,
                     const double* p_fail, const double* delay_ms, int64_t syn_id_base) {
  if (!h || nsyn < 0) return fail(h, HHD_E_ARG, "bad argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_synapses after run");
  if (h->nsyn) return fail(h, HHD_E_STATE, "synapses already set");
  if (nsyn && (!src || !tgt || !chan || !g_max || !U || !tau_rec || !tau_fac || !p_fail || !delay_ms))
    return fail(h, HHD_E_ARG, "null argument");
  const int lo = h->cfg.global_offset, hi = lo + h->N;
  int64_t maxd = 0;
  for (int64_t k = 0; k < nsyn; ++k) {
    if (src[k] < 0 || src[k] >= h->NG) return fail(h, HHD_E_ARG, "synapse %lld: source %d out of range", (long long)k, src[k]);
    if (tgt[k] < lo || tgt[k] >= hi) return fail(h, HHD_E_ARG, "synapse %lld: target %d is not local", (long long)k, tgt[k]);
    if (chan[k] > 2) return fail(h, HHD_E_ARG, "synapse %lld: channel %d", (long long)k, chan[k]);
    if (!(delay_ms[k] >= 0)) return fail(h, HHD_E_ARG, "synapse %lld: negative delay", (long long)k);
    if (!(tau_rec[k] > 0) || !(tau_fac[k] > 0)) return fail(h, HHD_E_ARG, "synapse %lld: tau_rec/tau_fac must be positive", (long long)k);
    int64_t d = hhd_delay_steps(delay_ms[k], h->cfg.dt_ms);
    if (d > 65535) return fail(h, HHD_E_ARG, "synapse %lld: delay of %lld steps exceeds 65535", (long long)k, (long long)d);
    if (d > maxd) maxd = d;
  }
  size_t n = (size_t)(nsyn ? nsyn : 1);
  h->src = (int32_t*)malloc(n * 4); h->tgt = (int32_t*)malloc(n * 4); h->chan = (uint8_t*)malloc(n);
  h->gmax = (double*)malloc(n * 8); h->U = (double*)malloc(n * 8); h->tau_rec = (double*)malloc(n * 8);
  h->tau_fac = (double*)malloc(n * 8); h->pfail = (double*)malloc(n * 8); h->delay_steps = (int64_t*)malloc(n * 8);
  h->u = (double*)malloc(n * 8); h->R = (double*)malloc(n * 8); h->a_syn = (double*)malloc(n * 8);
  h->failure = (uint8_t*)calloc(n, 1); h->touched = (uint8_t*)calloc(n, 1); h->row_syn = (int64_t*)malloc(n * 8);
  for (int64_t k = 0; k < nsyn; ++k) {
    h->src[k] = src[k]; h->tgt[k] = tgt[k]; h->chan[k] = chan[k]; h->gmax[k] = g_max[k]; h->U[k] = U[k];
    h->tau_rec[k] = tau_rec[k] * 1e-3; h->tau_fac[k] = tau_fac[k] * 1e-3; h->pfail[k] = p_fail[k];   /* tau in seconds, see hhd_time_s */
    h->delay_steps[k] = hhd_delay_steps(delay_ms[k], h->cfg.dt_ms);
    h->u[k] = U[k]; h->R[k] = 1.0; h->a_syn[k] = U[k];     /* codes/CortexNetwork.py:421-426 */
    h->row_ptr[src[k] + 1]++;
    if (!h->has_out_given) h->has_out[src[k]] = 1;
  }
  for (int j = 0; j < h->NG; ++j) h->row_ptr[j + 1] += h->row_ptr[j];
  int64_t* fill = (int64_t*)malloc(((size_t)h->NG + 1) * 8);
  memcpy(fill, h->row_ptr, ((size_t)h->NG + 1) * 8);
  for (int64_t k = 0; k < nsyn; ++k) h->row_syn[fill[src[k]]++] = k;
  free(fill);
  free(h->queue);
  h->max_delay = maxd;
  h->queue = (vec64*)calloc((size_t)maxd + 1, sizeof(vec64));
  h->nsyn = nsyn;
  h->syn_id_base = syn_id_base;
  return HHD_OK;
}

int hhd_set_synapses_dev(hhd_handle h, int64_t nsyn, const int32_t* src, const int32_t* tgt, const uint8_t* chan,
                         const double* g_max, const double* U, const double* tau_rec, const double* tau_fac,
                         const double* p_fail, const double* delay_ms, int64_t syn_id_base) {
  (void)nsyn; (void)src; (void)tgt; (void)chan; (void)g_max; (void)U; (void)tau_rec; (void)tau_fac; (void)p_fail; (void)delay_ms; (void)syn_id_base;
  return fail(h, HHD_E_STATE, "the CPU oracle takes host arrays only");
}

int hhd_add_spike_source(hhd_handle h, int32_t n_sources, int64_t nspk, const int32_t* spk_src, const double* spk_t_ms,
                         int64_t nsyn, const int32_t* syn_src, const int32_t* syn_tgt, const uint8_t* syn_chanmask,
                         const double* syn_gmax, const double* syn_pfail) {
  if (!h || n_sources <= 0 || nspk < 0 || nsyn < 0) return fail(h, HHD_E_ARG, "bad argument");
  for (int64_t k = 0; k < nspk; ++k) {
    if (spk_src[k] < 0 || spk_src[k] >= n_sources) return fail(h, HHD_E_ARG, "spike %lld: source out of range", (long long)k);
    if (hhd_timestep(spk_t_ms[k], h->cfg.dt_ms) < h->step) return fail(h, HHD_E_ARG, "spike %lld lies in the past", (long long)k);
  }
  for (int64_t k = 0; k < nsyn; ++k) {
    if (syn_src[k] < 0 || syn_src[k] >= n_sources) return fail(h, HHD_E_ARG, "input synapse %lld: source out of range", (long long)k);
    if (syn_tgt[k] < 0 || syn_tgt[k] >= h->N) return fail(h, HHD_E_ARG, "input synapse %lld: target out of range", (long long)k);
    if (syn_chanmask[k] == 0 || syn_chanmask[k] > 7) return fail(h, HHD_E_ARG, "input synapse %lld: channel mask", (long long)k);
  }
  ext_source* ne = (ext_source*)realloc(h->ext, (size_t)(h->n_ext + 1) * sizeof(ext_source));
  if (!ne) return fail(h, HHD_E_NOMEM, "out of memory");
  h->ext = ne;
  ext_source* e = &h->ext[h->n_ext];
  memset(e, 0, sizeof *e);
  e->n_sources = n_sources; e->nspk = nspk; e->nsyn = nsyn;
  size_t ns = (size_t)(nspk ? nspk : 1), ny = (size_t)(nsyn ? nsyn : 1);
  e->spk_src = (int32_t*)malloc(ns * 4); e->spk_step = (int64_t*)malloc(ns * 8);
  e->syn_src = (int32_t*)malloc(ny * 4); e->syn_tgt = (int32_t*)malloc(ny * 4); e->syn_mask = (uint8_t*)malloc(ny);
  e->syn_gmax = (double*)malloc(ny * 8); e->syn_pfail = (double*)malloc(ny * 8);
  e->src_ptr = (int64_t*)calloc((size_t)n_sources + 1, 8); e->src_syn = (int64_t*)malloc(ny * 8);
  /* stable insertion sort by step keeps (step, input order): SpikeGeneratorGroup bins int((t+1e-3dt)/dt) [Brian 2] */
  for (int64_t k = 0; k < nspk; ++k) {
    int64_t st = hhd_timestep(spk_t_ms[k], h->cfg.dt_ms);
    int64_t p = k;
    while (p > 0 && e->spk_step[p - 1] > st) { e->spk_step[p] = e->spk_step[p - 1]; e->spk_src[p] = e->spk_src[p - 1]; --p; }
    e->spk_step[p] = st; e->spk_src[p] = spk_src[k];
  }
  for (int64_t k = 0; k < nsyn; ++k) {
    e->syn_src[k] = syn_src[k]; e->syn_tgt[k] = syn_tgt[k]; e->syn_mask[k] = syn_chanmask[k];
    e->syn_gmax[k] = syn_gmax[k]; e->syn_pfail[k] = syn_pfail[k];
    e->src_ptr[syn_src[k] + 1]++;
  }
  for (int j = 0; j < n_sources; ++j) e->src_ptr[j + 1] += e->src_ptr[j];
  int64_t* fill = (int64_t*)malloc(((size_t)n_sources + 1) * 8);
  memcpy(fill, e->src_ptr, ((size_t)n_sources + 1) * 8);
  for (int64_t k = 0; k < nsyn; ++k) e->src_syn[fill[syn_src[k]]++] = k;
  free(fill);
  e->syn_id_base = h->ext_syn_total;
  h->ext_syn_total += nsyn;
  h->n_ext++;
  return HHD_OK;
}

/* ---- the model --------------------------------------------------------------------------------------------------- */
typedef struct {
  double C, gL, EL, dT, Vup, tauw, b, Vr, VT, Iref, Vdep, IDC;
} npar;

typedef struct {
  double I_AMPA, I_GABA, I_NMDA, I_syn, I_inj, I_tot, I_exp, w_V, D0, dD0, ex, g[3];
} subexpr;

static npar load_par(const hhd_handle h, int i) {
  npar p;
  p.C = h->par[HHD_P_C][i]; p.gL = h->par[HHD_P_GL][i]; p.EL = h->par[HHD_P_EL][i]; p.dT = h->par[HHD_P_DELTAT][i];
  p.Vup = h->par[HHD_P_VUP][i]; p.tauw = h->par[HHD_P_TAUW][i]; p.b = h->par[HHD_P_B][i]; p.Vr = h->par[HHD_P_VR][i];
  p.VT = h->par[HHD_P_VT][i]; p.Iref = h->par[HHD_P_IREF][i]; p.Vdep = h->par[HHD_P_VDEP][i]; p.IDC = h->par[HHD_P_IDC][i];
  return p;
}

/* Subexpressions of codes/CortexNetwork.py:235-248,265-272 at state x. */
static void subexpressions(const hhd_handle h, const npar* p, const double x[8], double iac, subexpr* s) {
  const double V = x[HHD_V];
  s->g[0] = x[HHD_G_AMPA_OFF] - x[HHD_G_AMPA_ON];                               /* g_AMPA = g_AMPA_off - g_AMPA_on   :265 */
  s->g[1] = x[HHD_G_GABA_OFF] - x[HHD_G_GABA_ON];                               /* :266 */
  s->g[2] = x[HHD_G_NMDA_OFF] - x[HHD_G_NMDA_ON];                               /* :267 */
  s->I_AMPA = s->g[0] * (h->E[0] - V);                                          /* :269 */
  const double Mg = 1.0 / (1.0 + h->mg_fac * hhd_exp(h->mg_slope * (h->mg_half - V)));  /* :270 */
  s->I_GABA = s->g[1] * (h->E[1] - V);                                          /* :271 */
  s->I_NMDA = (s->g[2] * (h->E[2] - V)) * Mg;                                   /* :272 */
  s->I_syn = (s->I_AMPA + s->I_NMDA) + s->I_GABA;                               /* :238 */
  s->I_inj = p->IDC + iac;                                                      /* :236-237 */
  s->I_tot = s->I_syn + s->I_inj;                                               /* :239 */
  s->ex = hhd_exp((V - p->VT) / p->dT);
  s->I_exp = (p->gL * p->dT) * s->ex;                                           /* :245 */
  s->w_V = (s->I_tot + s->I_exp) - p->gL * (V - p->EL);                         /* :246 */
  s->D0 = (p->C / p->gL) * s->w_V;                                              /* :248 */
  s->dD0 = p->C * (s->ex - 1.0);                                                /* :249 */
}

/* Right-hand sides, codes/CortexNetwork.py:251-263.  ts = stage time and ls = this neuron's last_spike, both in SECONDS:
 * `int(t - last_spike < 5*ms)` is evaluated on Brian's SI values (t = k * dt with dt = 0.05 * 0.001) — on millisecond values the
 * comparison falls on the other side for one spike step in ten (k*0.05 - k0*0.05 vs 5.0 at k - k0 = 100). */
static void deriv(const hhd_handle h, const npar* p, const double x[8], double ts, double ls, double iac, double d[8]) {
  subexpr s;
  subexpressions(h, p, x, iac, &s);
  const double V = x[HHD_V], w = x[HHD_W];
  const double a = ((s.I_tot >= p->Iref) ? 1.0 : 0.0) * ((ts - ls < h->cfg.block_window_ms * 1e-3) ? 1.0 : 0.0);
  const double dV = (a * (-p->gL / p->C)) * (V - p->Vdep) + ((1.0 - a) * (s.w_V - w)) / p->C;      /* :251 */
  const double gate = ((w > s.w_V - s.D0 / p->tauw) ? 1.0 : 0.0) * ((w < s.w_V + s.D0 / p->tauw) ? 1.0 : 0.0) *
                      ((V <= p->VT) ? 1.0 : 0.0) * ((s.I_tot < p->Iref) ? 1.0 : 0.0);
  d[HHD_V] = dV;                                                                                   /* :252 */
  d[HHD_W] = (gate * -(p->gL * (1.0 - s.ex) + s.dD0 / p->tauw)) * dV;                              /* :253 */
  for (int c = 0; c < 3; ++c) {
    d[HHD_G_AMPA_ON + 2 * c] = -((1.0 / h->tau_on[c]) * x[HHD_G_AMPA_ON + 2 * c]);                 /* :256-263 */
    d[HHD_G_AMPA_OFF + 2 * c] = -((1.0 / h->tau_off[c]) * x[HHD_G_AMPA_OFF + 2 * c]);
  }
}

/* Brian 2 ExplicitStateUpdater definitions: euler, rk2 (midpoint), rk4. */
static void integrate_fixed(const hhd_handle h, const npar* p, double x[8], double ts, double ls, double iac0, double iac1) {
  const double dt = h->cfg.dt_ms, dts = hhd_dt_s(h->cfg.dt_ms);
  double k1[8], k2[8], k3[8], k4[8], y[8];
  deriv(h, p, x, ts, ls, iac0, k1);
  for (int v = 0; v < 8; ++v) k1[v] = dt * k1[v];
  if (h->cfg.method == HHD_EULER) {
    for (int v = 0; v < 8; ++v) x[v] = x[v] + k1[v];
    return;
  }
  for (int v = 0; v < 8; ++v) y[v] = x[v] + k1[v] * 0.5;
  deriv(h, p, y, ts + dts * 0.5, ls, iac0, k2);
  for (int v = 0; v < 8; ++v) k2[v] = dt * k2[v];
  if (h->cfg.method == HHD_RK2) {
    for (int v = 0; v < 8; ++v) x[v] = x[v] + k2[v];
    return;
  }
  for (int v = 0; v < 8; ++v) y[v] = x[v] + k2[v] * 0.5;
  deriv(h, p, y, ts + dts * 0.5, ls, iac0, k3);
  for (int v = 0; v < 8; ++v) k3[v] = dt * k3[v];
  for (int v = 0; v < 8; ++v) y[v] = x[v] + k3[v];
  deriv(h, p, y, ts + dts, ls, iac1, k4);      /* TimedArray read at t + dt: the next row */
  for (int v = 0; v < 8; ++v) k4[v] = dt * k4[v];
  for (int v = 0; v < 8; ++v) x[v] = x[v] + (((k1[v] + 2.0 * k2[v]) + 2.0 * k3[v]) + k4[v]) / 6.0;
}

/*
 * gsl_rk2 stand-in (codes/Mainrk2.py:32, codes_revision/Tests.py:126): an embedded RK2(3) pair with per-neuron
 * sub-stepping inside each dt, in the spirit of GSL's gsl_odeiv2_step_rk2 + standard controller that Brian's GSL
 * state updater drives (k1 = f(y), k2 = f(y + h/2 k1), k3 = f(y + h(2 k2 - k1)); y+ = y + h(k1 + 4 k2 + k3)/6; error =
 * y+ minus the second-order solution y + h k2), absolute tolerance 1e-6 in SI units on every variable (1e-3 mV for V;
 * the pA / nS variables never bind), reject when err > 1.1 tol and shrink by max(0.9 err^-1/2, 0.2), grow by
 * min(0.9 err^-1/4, 5) when err < 0.5 tol, last sub-step clipped to the end of the step.  GSL itself is not available
 * here, so this is a STATISTICAL stand-in, written with IEEE sqrt only so that the CUDA engine
 * reproduces it bit for bit (the growth exponent is -1/4 instead of GSL's -1/3 for that reason).
 */
#define RK23_MAX_SUBSTEPS 4096
#define OMP_MIN_NEURONS 8192   /* below this a parallel region costs more than the loop it splits */
/*
 * `gsl_rk2`: what Brian 2's GSL state updater runs per neuron and dt —
 *   stepper   gsl_odeiv2_step_rk2: k1 = f(y), k2 = f(y + h/2 k1), k3 = f(y + h(-k1 + 2 k2)), y+ = y + h (k1 + 4 k2 + k3)/6,
 *             yerr = h (k2 - (k1 + 4 k2 + k3)/6)   [|.| = h |k1 - 2 k2 + k3| / 6]
 *   control   gsl_odeiv2_control_standard, eps_abs = 1e-6 in SI units of every variable, eps_rel = 0, order 2:
 *             r = max |yerr_i| / eps_abs;  r > 1.1: h <- h max(0.2, 0.9 r^(-1/2)), retry;  r < 0.5: h <- h clamp(0.9 r^(-1/3), 1, 5)
 *   evolve    gsl_odeiv2_evolve_apply: a step that would pass the end of the dt interval is truncated ("final step") and
 *             its size is NOT taken as the suggestion for the next interval; a rejected final step is retried as a normal one
 *   carry     the suggestion survives from one dt to the next (Brian's use_last_timestep) in h->hstep.
 * The only liberties: hhd_cbrt for r^(-1/3) (within 2 ulp of pow; the same code runs on the GPU), a floor of 2^-20 dt on the
 * step size instead of GSL's "step size too small" error, and NaN errors accept the step.
 */
static void integrate_rk23(const hhd_handle h, const npar* p, double x[8], double t0, double ls, double iac0, double iac1, double* hcar) {
  const double dt = h->cfg.dt_ms;
  static const double inv_tol[8] = {1e3, 1e-6, 1e-3, 1e-3, 1e-3, 1e-3, 1e-3, 1e-3};
  double t = 0.0, h0 = *hcar;
  for (int it = 0; it < RK23_MAX_SUBSTEPS && t < dt; ++it) {
    double hs = h0;
    int final_step = 0;
    if (hs > dt - t) { hs = dt - t; final_step = 1; }     /* evolve.c: if (h0 > t1 - t) { h0 = t1 - t; final_step = 1; } */
    double k1[8], k2[8], k3[8], y[8], ynew[8];
    deriv(h, p, x, t0 + t * 1e-3, ls, iac0, k1);
    for (int v = 0; v < 8; ++v) y[v] = x[v] + (0.5 * hs) * k1[v];
    deriv(h, p, y, t0 + (t + 0.5 * hs) * 1e-3, ls, iac0, k2);
    for (int v = 0; v < 8; ++v) y[v] = x[v] + hs * (2.0 * k2[v] - k1[v]);
    deriv(h, p, y, t0 + (t + hs) * 1e-3, ls, final_step ? iac1 : iac0, k3);
    double rmax = 0.0;
    for (int v = 0; v < 8; ++v) {
      ynew[v] = x[v] + hs * (((k1[v] + 4.0 * k2[v]) + k3[v]) / 6.0);
      const double err = hs * (((k1[v] - 2.0 * k2[v]) + k3[v]) / 6.0);
      const double r = fabs(err) * inv_tol[v];
      if (r > rmax) rmax = r;                      /* NaN never wins: a non-finite error accepts the step */
    }
    if (rmax > 1.1 && hs > dt * 9.5367431640625e-07) {     /* 2^-20 dt floor */
      double f = 0.9 / sqrt(rmax);
      if (f < 0.2) f = 0.2;
      h0 = hs * f;                                 /* retried as a normal step */
      continue;
    }
    for (int v = 0; v < 8; ++v) x[v] = ynew[v];
    t = final_step ? dt : t + hs;
    double hn = hs;
    if (rmax < 0.5) {
      double f = 5.0;
      if (rmax > 9.765625e-4) {                    /* below 2^-10 the factor is clamped to 5 anyway (0.9 / 5)^3 = 0.0058 */
        f = 0.9 / hhd_cbrt(rmax);
        if (f > 5.0) f = 5.0;
        if (f < 1.0) f = 1.0;
      }
      hn = hs * f;
    }
    if (!final_step) h0 = hn;                      /* no suggestion from a truncated step */
  }
  *hcar = h0;
}

/* I_AC of local neuron i at step s: TimedArray semantics (first row before the table, last row after it). */
static double iac_at(const hhd_handle h, int i, int64_t s) {
  if (!h->iac_table || h->iac_col[i] < 0) return 0.0;
  if (s < 0) s = 0;
  if (s >= h->iac_steps) s = h->iac_steps - 1;
  return h->iac_table[(size_t)s * h->iac_cols + h->iac_col[i]];
}

static double monitor_value(const hhd_handle h, int var, int i) {
  double x[8];
  for (int v = 0; v < 8; ++v) x[v] = h->x[v][i];
  if (var < 8) return x[var];
  if (var == HHD_LAST_SPIKE) return h->last_spike[h->cfg.global_offset + i] * 1e3;
  npar p = load_par(h, i);
  subexpr s;
  subexpressions(h, &p, x, iac_at(h, i, h->step), &s);
  switch (var) {
    case HHD_G_AMPA: return s.g[0];
    case HHD_G_GABA: return s.g[1];
    case HHD_G_NMDA: return s.g[2];
    case HHD_I_AMPA: return s.I_AMPA;
    case HHD_I_GABA: return s.I_GABA;
    case HHD_I_NMDA: return s.I_NMDA;
    case HHD_I_SYN: return s.I_syn;
    case HHD_I_INJ: return s.I_inj;
    case HHD_I_TOT: return s.I_tot;
    case HHD_I_EXC: return s.I_AMPA + s.I_NMDA;                 /* codes/CortexNetwork.py:241 */
    case HHD_I_EXC_TOT: return (s.I_AMPA + s.I_NMDA) + s.I_inj; /* :242 */
    case HHD_I_EXP: return s.I_exp;
  }
  return 0.0;
}

/* ---- one step, first half: start + groups + thresholds (local) ---------------------------------------------------- */
static int check_ready(hhd_handle h) {
  if (!h->neurons_set) return fail(h, HHD_E_STATE, "hhd_set_neurons has not been called");
  if (!h->channels_set) return fail(h, HHD_E_STATE, "hhd_set_channels has not been called");
  if (!h->pendd) h->pendd = (double*)calloc((size_t)3 * h->N, sizeof(double));
  if (h->cfg.accum == HHD_ACCUM_FIXED && !h->pend[0]) {
    h->pend[0] = (int64_t*)calloc((size_t)3 * h->N, 8);
    h->pend[1] = (int64_t*)calloc((size_t)3 * h->N, 8);
  }
  h->started = 1;
  return HHD_OK;
}

int hhdo_step_begin(hhd_handle h, int32_t* local_spikes_global_ids, int32_t* n_spikes) {
  if (!h) return HHD_E_ARG;
  if (h->step_open) return fail(h, HHD_E_STATE, "step already open");
  int rc = check_ready(h);
  if (rc) return rc;
  const int64_t k = h->step;
  const double ts = hhd_time_s(k, h->cfg.dt_ms);
  const int off = h->cfg.global_offset;

  /* slot `start`: StateMonitors sample the state before the update (codes/CortexNetwork.py:510) */
  for (int m = 0; m < h->n_mon; ++m) {
    monitor* mo = &h->mon[m];
    if (!mo->active) continue;
    if (mo->first_step < 0) mo->first_step = k;
    double acc = 0.0;
    for (int64_t q = 0; q < mo->n_idx; ++q) {
      double v = monitor_value(h, mo->var, mo->idx[q]);
      if (mo->reduce == HHD_REDUCE_NONE) { if (vecd_push(&mo->samples, v)) return fail(h, HHD_E_NOMEM, "monitor out of memory"); }
      else acc += v;
    }
    if (mo->reduce == HHD_REDUCE_MEAN && mo->n_idx) acc /= (double)mo->n_idx;
    if (mo->reduce != HHD_REDUCE_NONE) vecd_push(&mo->samples, acc);
  }

  /* slot `groups`: state update of all 8 variables (codes/CortexNetwork.py:308-309) */
  int64_t bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : bad) if (h->N >= OMP_MIN_NEURONS)
#endif
  for (int i = 0; i < h->N; ++i) {
    npar p = load_par(h, i);
    double x[8];
    for (int v = 0; v < 8; ++v) x[v] = h->x[v][i];
    const double iac0 = iac_at(h, i, k), iac1 = iac_at(h, i, k + 1);
    if (h->cfg.method == HHD_RK23_ADAPTIVE) integrate_rk23(h, &p, x, ts, h->last_spike[off + i], iac0, iac1, &h->hstep[i]);
    else integrate_fixed(h, &p, x, ts, h->last_spike[off + i], iac0, iac1);
    if (h->noise_sigma && h->noise_sigma[i] != 0.0) {
      /* Euler-Maruyama term of Brian's stochastic `euler` updater: x + dt f(x) + xi sigma, xi = sqrt(dt) N(0,1) */
      const double sg = h->noise_sigma[i] * sqrt(h->cfg.dt_ms);
      double z0, z1, z2, z3;
      hhd_normal2(h->cfg.seed, (uint64_t)(off + i), (uint64_t)k, 2, &z0, &z1);
      hhd_normal2(h->cfg.seed, (uint64_t)(off + i), (uint64_t)k, 3, &z2, &z3);
      x[HHD_G_AMPA_OFF] = x[HHD_G_AMPA_OFF] + z0 * sg;
      x[HHD_G_GABA_OFF] = x[HHD_G_GABA_OFF] + z1 * sg;
      x[HHD_G_NMDA_OFF] = x[HHD_G_NMDA_OFF] + z2 * sg;
    }
    for (int v = 0; v < 8; ++v) h->x[v][i] = x[v];
    if (!isfinite(x[HHD_V]) || !isfinite(x[HHD_W])) bad++;
  }
  h->cnt.nonfinite += bad;

  /* slot `thresholds`: spike = V > V_up and not_refractory (:304); w_crossing (:306) evaluated on the same state */
  h->n_cur_spikes = 0;
  h->n_cur_wcross = 0;
  if (!h->flag_scratch) h->flag_scratch = (uint8_t*)malloc((size_t)h->N);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (h->N >= OMP_MIN_NEURONS)
#endif
  for (int i = 0; i < h->N; ++i) {
    const int not_refractory = (k - h->lastspike_step[i]) >= h->refr_steps;
    uint8_t f = (h->x[HHD_V][i] > h->par[HHD_P_VUP][i] && not_refractory) ? 1 : 0;
    npar p = load_par(h, i);
    double x[8];
    for (int v = 0; v < 8; ++v) x[v] = h->x[v][i];
    subexpr s;
    subexpressions(h, &p, x, iac_at(h, i, k), &s);      /* slot `thresholds`: the clock still reads t_k */
    const double w = x[HHD_W];
    if (w > s.w_V - s.D0 / p.tauw && w < s.w_V + s.D0 / p.tauw && x[HHD_V] <= p.VT) f |= 2;
    h->flag_scratch[i] = f;
  }
  for (int i = 0; i < h->N; ++i) {                      /* lists in ascending neuron order, as Brian's spike space */
    if (h->flag_scratch[i] & 1) {
      h->cur_spikes[h->n_cur_spikes++] = i;
      h->lastspike_step[i] = k;
    }
    if (h->flag_scratch[i] & 2) h->cur_wcross[h->n_cur_wcross++] = i;
  }
  h->cnt.spikes += h->n_cur_spikes;
  h->cnt.w_crossings += h->n_cur_wcross;
  if (h->spikes_active)
    for (int q = 0; q < h->n_cur_spikes; ++q) {                       /* SpikeMonitor (:487) */
      if (vec64_push(&h->spk_neuron, off + h->cur_spikes[q]) || vec64_push(&h->spk_step, k)) return fail(h, HHD_E_NOMEM, "spike monitor out of memory");
    }
  if (local_spikes_global_ids) for (int q = 0; q < h->n_cur_spikes; ++q) local_spikes_global_ids[q] = off + h->cur_spikes[q];
  if (n_spikes) *n_spikes = h->n_cur_spikes;
  h->step_open = 1;
  return HHD_OK;
}

static void add_increment(hhd_handle h, int c, int tgt_local, double inc) {
  if (h->cfg.accum == HHD_ACCUM_FIXED) {
    const int64_t q = llrint(inc * 1099511627776.0);   /* 2^40 */
    h->pend[0][(size_t)c * h->N + tgt_local] += q;
  } else {
    /* p7a/p7b ... p9b: `g_X_on_post += inc; g_X_off_post += inc`.  The increments of one step are summed per (target,
     * channel) first and the sum is added to both conductances at the end of the slot — g + (a + b) instead of
     * (g + a) + b — because that is the only order a parallel engine can reproduce; identical whenever a target receives
     * at most two increments per channel in a step, a rounding-level difference otherwise. */
    h->pendd[(size_t)c * h->N + tgt_local] += inc;
  }
}

/* ---- second half: synapses + resets + after_resets.  all_spikes = GLOBAL ids of every neuron that spiked this step
 *      anywhere in the network, ascending. ---------------------------------------------------------------------------- */
int hhdo_step_end(hhd_handle h, const int32_t* all_spikes, int32_t n_all) {
  if (!h) return HHD_E_ARG;
  if (!h->step_open) return fail(h, HHD_E_STATE, "no open step");
  const int64_t k = h->step;
  const double t = hhd_time_s(k, h->cfg.dt_ms);      /* seconds: t, last_spike_pre and the STP time constants as Brian holds them */
  const int off = h->cfg.global_offset;

  /* slot `synapses`, pathways p0..p6 (order 0..6, no delay) for every synapse of every spiking neuron (:360-366) */
  for (int q = 0; q < n_all; ++q) {
    const int j = all_spikes[q];
    const double last = h->last_spike[j];                 /* p1, p2 read last_spike_pre before p5 overwrites it */
    for (int64_t r = h->row_ptr[j]; r < h->row_ptr[j + 1]; ++r) {
      const int64_t s = h->row_syn[r];
      h->failure[s] = hhd_philox_u01(h->cfg.seed, (uint64_t)(h->syn_id_base + s), (uint64_t)k, 0) < h->pfail[s] ? 1 : 0;  /* p0 */
      const double u1 = h->U[s] + (h->u[s] * (1.0 - h->U[s])) * hhd_exp(-(t - last) / h->tau_fac[s]);                    /* p1 */
      const double R1 = 1.0 + ((h->R[s] - h->u[s] * h->R[s]) - 1.0) * hhd_exp(-(t - last) / h->tau_rec[s]);               /* p2 */
      h->u[s] = u1;                                                                                                       /* p3 */
      h->R[s] = R1;                                                                                                       /* p4 */
      h->touched[s] = 1;
      h->a_syn[s] = h->u[s] * h->R[s];                                                                                    /* p6 */
      /* p7a..p9b carry delay dtax (:463-468): Brian pushes the synapse index into the queue slot k + delay_steps */
      if (vec64_push(&h->queue[(k + h->delay_steps[s]) % (h->max_delay + 1)], s)) return fail(h, HHD_E_NOMEM, "queue out of memory");
      h->cnt.syn_events++;
    }
    if (h->has_out[j]) h->last_spike[j] = t;                                                                              /* p5 */
  }
  /* order-7 pathways: every synapse index due in this step, values read NOW (:367-372) */
  {
    vec64* slot = &h->queue[k % (h->max_delay + 1)];
    for (int64_t e = 0; e < slot->n; ++e) {
      const int64_t s = slot->v[e];
      const int c = h->chan[s];
      const double inc = ((h->gmax[s] * h->a_syn[s]) * (1.0 - (double)h->failure[s])) * h->gsyn[c];
      add_increment(h, c, h->tgt[s] - off, inc);
      h->cnt.syn_deliveries++;
    }
    slot->n = 0;
  }
  /* SpikeGeneratorGroup + input synapses: no STP, no delay (:588-594) */
  for (int x = 0; x < h->n_ext; ++x) {
    ext_source* e = &h->ext[x];
    while (e->cursor < e->nspk && e->spk_step[e->cursor] < k) e->cursor++;
    while (e->cursor < e->nspk && e->spk_step[e->cursor] == k) {
      const int sj = e->spk_src[e->cursor++];
      for (int64_t r = e->src_ptr[sj]; r < e->src_ptr[sj + 1]; ++r) {
        const int64_t s = e->src_syn[r];
        const int failure = hhd_philox_u01(h->cfg.seed, (uint64_t)(e->syn_id_base + s), (uint64_t)k, 1) < e->syn_pfail[s] ? 1 : 0;
        for (int c = 0; c < 3; ++c)
          if (e->syn_mask[s] & (1 << c)) add_increment(h, c, e->syn_tgt[s], (e->syn_gmax[s] * (1.0 - (double)failure)) * h->gsyn[c]);
        h->cnt.ext_events++;
      }
    }
  }
  if (h->cfg.accum != HHD_ACCUM_FIXED) {
    for (int c = 0; c < 3; ++c)
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (h->N >= OMP_MIN_NEURONS)
#endif
      for (int i = 0; i < h->N; ++i) {
        const double inc = h->pendd[(size_t)c * h->N + i];
        if (inc != 0.0) {
          h->x[HHD_G_AMPA_ON + 2 * c][i] += inc;
          h->x[HHD_G_AMPA_OFF + 2 * c][i] += inc;
          h->pendd[(size_t)c * h->N + i] = 0.0;
        }
      }
  }
  if (h->cfg.accum == HHD_ACCUM_FIXED) {
    for (int c = 0; c < 3; ++c)
#ifdef _OPENMP
#pragma omp parallel for schedule(static) if (h->N >= OMP_MIN_NEURONS)
#endif
      for (int i = 0; i < h->N; ++i) {
        const int64_t q = h->pend[0][(size_t)c * h->N + i];
        if (q) {
          const double inc = (double)q * (1.0 / 1099511627776.0);
          h->x[HHD_G_AMPA_ON + 2 * c][i] += inc;
          h->x[HHD_G_AMPA_OFF + 2 * c][i] += inc;
          h->pend[0][(size_t)c * h->N + i] = 0;
        }
      }
  }
  /* slot `resets`: V = V_r; w += b (:305) */
  for (int q = 0; q < h->n_cur_spikes; ++q) {
    const int i = h->cur_spikes[q];
    h->x[HHD_V][i] = h->par[HHD_P_VR][i];
    h->x[HHD_W][i] += h->par[HHD_P_B][i];
  }
  /* slot `after_resets`: run_on_event('w_crossing', 'w = w_V - D0/tau_w') (:311) on the current state */
  for (int q = 0; q < h->n_cur_wcross; ++q) {
    const int i = h->cur_wcross[q];
    npar p = load_par(h, i);
    double x[8];
    for (int v = 0; v < 8; ++v) x[v] = h->x[v][i];
    subexpr s;
    subexpressions(h, &p, x, iac_at(h, i, k), &s);
    h->x[HHD_W][i] = s.w_V - s.D0 / p.tauw;
  }
  h->step++;
  h->cnt.steps++;
  h->cnt.neuron_updates += h->N;
  h->step_open = 0;
  return HHD_OK;
}

int hhd_run(hhd_handle h, int64_t n_steps) {
  if (!h || n_steps < 0) return fail(h, HHD_E_ARG, "bad argument");
  if (h->cfg.n_ranks > 1) return fail(h, HHD_E_STATE, "partitioned oracle: drive it with hhdo_step_begin/hhdo_step_end");
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int32_t* buf = (int32_t*)malloc((size_t)h->N * 4);
  for (int64_t s = 0; s < n_steps; ++s) {
    int32_t n = 0;
    int rc = hhdo_step_begin(h, buf, &n);
    if (rc) { free(buf); return rc; }
    rc = hhdo_step_end(h, buf, n);
    if (rc) { free(buf); return rc; }
  }
  free(buf);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  h->cnt.run_ms_device += (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
  return HHD_OK;
}

int hhd_profile_kernels(hhd_handle h, int64_t n_steps, double us[3]) {
  (void)n_steps; (void)us;
  return fail(h, HHD_E_STATE, "the CPU oracle has no kernels to profile");
}

/* ---- monitors & results ------------------------------------------------------------------------------------------ */
int hhd_add_state_monitor(hhd_handle h, int32_t var, int64_t n_idx, const int32_t* idx, int32_t reduce,
                          int64_t capacity_steps, int32_t* mon_id) {
  (void)capacity_steps;
  if (!h || var < 0 || var >= HHD_N_VARS || reduce < 0 || reduce > 2) return fail(h, HHD_E_ARG, "bad argument");
  if (!idx) n_idx = h->N;
  for (int64_t q = 0; idx && q < n_idx; ++q) if (idx[q] < 0 || idx[q] >= h->N) return fail(h, HHD_E_ARG, "monitor index out of range");
  monitor* nm = (monitor*)realloc(h->mon, (size_t)(h->n_mon + 1) * sizeof(monitor));
  if (!nm) return fail(h, HHD_E_NOMEM, "out of memory");
  h->mon = nm;
  monitor* mo = &h->mon[h->n_mon];
  memset(mo, 0, sizeof *mo);
  mo->var = var; mo->reduce = reduce; mo->active = 1; mo->n_idx = n_idx; mo->first_step = -1;
  mo->idx = (int32_t*)malloc((size_t)(n_idx ? n_idx : 1) * 4);
  for (int64_t q = 0; q < n_idx; ++q) mo->idx[q] = idx ? idx[q] : (int32_t)q;
  if (mon_id) *mon_id = h->n_mon;
  h->n_mon++;
  return HHD_OK;
}

int hhd_set_monitor_active(hhd_handle h, int32_t mon_id, int32_t active) {
  if (!h || mon_id < -1 || mon_id >= h->n_mon) return fail(h, HHD_E_ARG, "bad monitor id");
  if (mon_id == -1) h->spikes_active = active; else h->mon[mon_id].active = active;
  return HHD_OK;
}

static int64_t mon_width(const monitor* mo) { return mo->reduce == HHD_REDUCE_NONE ? mo->n_idx : 1; }

int hhd_monitor_samples(hhd_handle h, int32_t mon_id, int64_t* n_samples, int64_t* first_step) {
  if (!h || mon_id < 0 || mon_id >= h->n_mon) return fail(h, HHD_E_ARG, "bad monitor id");
  const monitor* mo = &h->mon[mon_id];
  const int64_t w = mon_width(mo);
  if (n_samples) *n_samples = w ? mo->samples.n / w : 0;
  if (first_step) *first_step = mo->first_step;
  return HHD_OK;
}

int hhd_read_monitor(hhd_handle h, int32_t mon_id, double* out, int64_t cap_values, int64_t* n_samples) {
  if (!h || mon_id < 0 || mon_id >= h->n_mon) return fail(h, HHD_E_ARG, "bad monitor id");
  const monitor* mo = &h->mon[mon_id];
  if (mo->samples.n > cap_values) return fail(h, HHD_E_CAPACITY, "output buffer too small: need %lld values", (long long)mo->samples.n);
  if (mo->samples.n) memcpy(out, mo->samples.v, (size_t)mo->samples.n * 8);
  const int64_t w = mon_width(mo);
  if (n_samples) *n_samples = w ? mo->samples.n / w : 0;
  return HHD_OK;
}

int hhd_read_monitor_dev(hhd_handle h, int32_t mon_id, const double** dev_ptr, int64_t* n_samples) {
  (void)mon_id; (void)dev_ptr; (void)n_samples;
  return fail(h, HHD_E_STATE, "the CPU oracle has no device buffers");
}

int hhd_clear_monitor(hhd_handle h, int32_t mon_id) {
  if (!h || mon_id < -1 || mon_id >= h->n_mon) return fail(h, HHD_E_ARG, "bad monitor id");
  if (mon_id == -1) { h->spk_neuron.n = 0; h->spk_step.n = 0; }
  else { h->mon[mon_id].samples.n = 0; h->mon[mon_id].first_step = -1; }
  return HHD_OK;
}

int hhd_spike_count(hhd_handle h, int64_t* n) {
  if (!h || !n) return fail(h, HHD_E_ARG, "null argument");
  *n = h->spk_neuron.n;
  return HHD_OK;
}

int hhd_read_spikes(hhd_handle h, int32_t* neuron, int64_t* step, int64_t cap, int64_t* n_out) {
  if (!h) return HHD_E_ARG;
  if (h->spk_neuron.n > cap) return fail(h, HHD_E_CAPACITY, "output buffer too small: need %lld", (long long)h->spk_neuron.n);
  for (int64_t q = 0; q < h->spk_neuron.n; ++q) { neuron[q] = (int32_t)h->spk_neuron.v[q]; step[q] = h->spk_step.v[q]; }
  if (n_out) *n_out = h->spk_neuron.n;
  return HHD_OK;
}

int hhd_read_state(hhd_handle h, int32_t var, double* out) {
  if (!h || !out || var < 0 || var >= HHD_N_VARS) return fail(h, HHD_E_ARG, "bad argument");
  for (int i = 0; i < h->N; ++i) out[i] = monitor_value(h, var, i);
  return HHD_OK;
}

int hhd_write_state(hhd_handle h, int32_t var, const double* in) {
  if (!h || !in) return fail(h, HHD_E_ARG, "null argument");
  if (var >= 0 && var < 8) memcpy(h->x[var], in, (size_t)h->N * 8);
  else if (var == HHD_LAST_SPIKE) for (int i = 0; i < h->N; ++i) h->last_spike[h->cfg.global_offset + i] = in[i] * 1e-3;
  else return fail(h, HHD_E_ARG, "variable %d is not writable", var);
  return HHD_OK;
}

int hhd_read_syn_state(hhd_handle h, int32_t which, double* out) {
  if (!h || !out) return fail(h, HHD_E_ARG, "null argument");
  for (int64_t s = 0; s < h->nsyn; ++s) {
    switch (which) {
      case HHD_SYN_U: out[s] = h->u[s]; break;
      case HHD_SYN_R: out[s] = h->R[s]; break;
      case HHD_SYN_A: out[s] = h->touched[s] ? ((h->gmax[s] * h->a_syn[s]) * (1.0 - (double)h->failure[s])) * h->gsyn[h->chan[s]] : 0.0; break;
      case HHD_SYN_DELAY_STEPS: out[s] = (double)h->delay_steps[s]; break;
      default: return fail(h, HHD_E_ARG, "unknown synapse variable %d", which);
    }
  }
  return HHD_OK;
}

int hhd_get_counters(hhd_handle h, hhd_counters* out) {
  if (!h || !out) return fail(h, HHD_E_ARG, "null argument");
  *out = h->cnt;
  return HHD_OK;
}

int hhd_current_step(hhd_handle h, int64_t* step) {
  if (!h || !step) return fail(h, HHD_E_ARG, "null argument");
  *step = h->step;
  return HHD_OK;
}

int hhd_comm_unique_id(uint8_t out_id[128]) { memset(out_id, 0, 128); return HHD_OK; }
int hhd_comm_init(hhd_handle h, const uint8_t id[128]) { (void)id; return fail(h, HHD_E_STATE, "the CPU oracle exchanges spikes through hhdo_step_begin/hhdo_step_end"); }

/* Plain exports of the shared math so tests can compare them with libm / a reference Philox. */
double hhdo_exp(double x) { return hhd_exp(x); }
double hhdo_cbrt(double x) { return hhd_cbrt(x); }
double hhdo_log(double x) { return hhd_log(x); }
void hhdo_normal2(uint64_t seed, uint64_t index, uint64_t step, uint32_t stream, double* z) { hhd_normal2(seed, index, step, stream, &z[0], &z[1]); }
double hhdo_philox_u01(uint64_t seed, uint64_t index, uint64_t step, uint32_t stream) { return hhd_philox_u01(seed, index, step, stream); }
int64_t hhdo_delay_steps(double delay_ms, double dt_ms) { return hhd_delay_steps(delay_ms, dt_ms); }
int64_t hhdo_timestep(double t_ms, double dt_ms) { return hhd_timestep(t_ms, dt_ms); }
