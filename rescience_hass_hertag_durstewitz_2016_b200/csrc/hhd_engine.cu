/*
 * hhd_engine.cu — sm_100a engine behind the C ABI of include/hhd.h.  DESIGN.md §3-5 has the full picture.
 *
 * One time step of the grid-wide (graph) plan = two kernels, captured STEPS_PER_GRAPH at a time into a CUDA graph:
 *   k_neuron      persistent CTAs, every WARP its own TMA pipeline over 32-neuron blocks (one 5 KB bulk copy per tile):
 *                 finish the previous step (w_crossing test on the stored state, pending synaptic increments, reset
 *                 `V=V_r; w+=b`, w_crossing projection), sample the monitors (Brian's `start` slot), integrate the 8 ODEs
 *                 (euler / rk2 / rk4 / gsl_rk2; TimedArray current and conductance noise in the TIMED variants),
 *                 threshold + refractoriness, append spikes to the spike ring with one atomic per warp (ballot
 *                 compaction) and note what the synaptic kernel needs of a spike.  Replaces NeuronGroup's
 *                 stateupdater/thresholder/resetter + StateMonitor + SpikeMonitor (codes/CortexNetwork.py:304-311, 487, 510).
 *                 On a partitioned network every spiking lane also stores (id, step tag) as one self-validating 64-bit
 *                 word into every peer's receive buffer, and the last CTA sends the count the same way (send_spike,
 *                 push_spikes).
 *   k_synapse     slot `synapses`.  Pathways p0..p6 for the spikes of THIS step in 128-synapse tasks — failure draw (Philox
 *                 keyed by synapse and step), Tsodyks–Markram update of (u, R), the amplitude (codes/CortexNetwork.py:360-366)
 *                 — and the delayed pathways p7a..p9b (:367-372, `delay = dtax` :463-468) as ONE atomic add of that amplitude
 *                 into slot (k + delay) of a circular per-delay conductance buffer, which k_neuron(k + delay + 1) drains:
 *                 no delivery pass, no per-event queue memory.  Brian reads the amplitude when the delayed pathway executes
 *                 (SURVEY F5); that only differs if the neuron spikes again while an older spike is in flight (delays beyond
 *                 the refractory period, template LONG), and then the new spike replaces the amplitude in every slot not yet
 *                 consumed (add_correction).  SpikeGeneratorGroup events (:570-622) as well.  On a partitioned network every
 *                 CTA then takes its share of the OTHER ranks' spikes whole (remote_spike_phase).
 * The other plan on the same arithmetic: k_resident (one thread-block cluster per small independent network, whole
 * segments in one launch).  Round 1's cooperative whole-run kernel and one-launch-per-step schedule were measured slower
 * (profiles/r1_notes.md) and are gone.
 *
 * Data layout in HBM: per block of 32 neurons 17 rows of 32 doubles (8 state + the 9 parameters every step needs, 4352 B
 * contiguous = one bulk copy; b, V_r and V_dep are only read by the neurons that spike / sit in their 5 ms window), the
 * step's slot of the conductance buffer (3 words) and the meta word, last_spike per global neuron — 216 B moved per
 * neuron-update against the 240 B roofline figure of SURVEY §8d; per synapse (CSR by presynaptic neuron, sorted by delay) a 16-byte record {target, delay,
 * channel, amplitude} and a 64-byte record {U, tau_rec, tau_fac, g_max, p_fail, u, R, original index}.
 */
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#ifdef HHD_WITH_NCCL
#include <nccl.h>
#endif
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/hhd.h"
#include "hhd_model.cuh"

#define MAX_MON 12   /* with 12 the static shared memory of k_neuron still lets five CTAs share an SM */
#define MAX_RANKS 16
#define MAX_RED 4   /* reduced (sum/mean) monitors: their per-thread accumulators live in registers */
#ifndef NEURON_BLOCK
#define NEURON_BLOCK 128
#endif
#ifndef STP_BLOCK
#define STP_BLOCK 128
#endif
#define DELIVER_BLOCK 128
#define STEPS_PER_GRAPH 50
#define RES_BLOCK 128          /* threads (= neurons) per CTA of the cluster-resident plan */
#define RES_MAX_CLUSTER 16
#define DELAY_BUCKETS 16

enum { CNT_SYN_EVENTS = 0, CNT_DELIVERIES, CNT_EXT, CNT_SPIKES, CNT_WCROSS, CNT_NONFINITE, CNT_OVERFLOW, CNT_TIMEOUT, CNT_N };
#define CNT_SLOTS 256   /* the delivery counter is spread over 256 addresses: thousands of warps bump it every step, and
                           same-address atomics serialise in one L2 slice */
enum { FLAG_SPIKED = 1, FLAG_WCROSS = 2, FLAG_HAS_PREV = 2, FLAG_OUT = 4, FLAG_ZERO = 8, META_SHIFT = 4 };   /* FLAG_OUT / FLAG_ZERO: copies of the
   has_out bits (outgoing synapses: p5 exists; the row starts with zero-delay synapses), so that the thread that detects a
   spike needs no load — a dependent load there stalls its warp, and with it the CTA, for a microsecond.  */
/* FLAG_HAS_PREV (meta bit 1): the stored state is the result of an integrated step whose w_crossing test is still due (done
 * at the start of the next pass) */

struct MonDesc {
  int var, reduce, active, red_slot;
  long long n_idx;
  long long k0;          /* sample index of step k is k - k0 */
  long long capacity;    /* samples */
  const int* slot;       /* [N] position of neuron in the monitor's list, or -1; NULL = identity */
  double* buf;           /* [capacity][width] */
};

/* Synapses are stored row-contiguous (CSR by presynaptic neuron, rows sorted by delay) as two arrays of structs so that
 * one event touches one 16-byte and one 64-byte record instead of twelve separate arrays: the synapse data of a
 * 1M-neuron network spans 23 GB, far beyond TLB reach, and every separate array costs its own page walk. */
struct __align__(16) SynHot {   /* read by every delivery */
  int tgt;                      /* local target neuron */
  unsigned short delay;         /* steps */
  unsigned char chan, pad;
  double amp;                   /* ((g_max*a_syn)*(1-failure))*Gsyn of the latest presynaptic spike */
};
struct __align__(16) SynStp {   /* read and written once per presynaptic spike (pathways p0..p6) */
  double U, tau_rec, tau_fac, gmax, pfail, u, R;
  unsigned int orig, pad;       /* original synapse index: RNG key and hhd_read_syn_state order */
};

struct Params {
  int N, NG, off, refr_steps, accum, n_mon, max_delay, W;
  double dt, dt_s;             /* ms (integration) and seconds (Brian's clock: t, last_spike, the STP decays — hhd_time_s) */
  unsigned long long seed;
  ChannelConst cc;
  double* nb;                  /* [n_pad / 32][NB_FIELDS][32] neuron blocks: 8 state rows + 9 streamed parameter rows (nb_index) */
  const double* rare;          /* [N_RARE][n_pad] b, V_r, V_dep: fetched only by the lanes that need them */
  double* last_spike;          /* [NG] seconds */
  double* hstep;               /* [n_pad] adaptive method: the step size a neuron carries from one dt to the next */
  int* meta;                   /* [n_pad] lastspike_step << META_SHIFT | flags (of the previous step; FLAG_OUT, FLAG_ZERO static) */
  const unsigned char* has_out;/* [NG] */
  long long* pend;             /* [n_slots][n_pad / 32][3][32] (pend_index): the per-delay circular conductance buffer.  Slot
                                  (s mod n_slots) sums, per (channel, target), the synaptic increments that Brian's `synapses`
                                  slot of step s delivers: fp64 bits (HHD_ACCUM_FP64) or int64 multiples of 2^-40 nS
                                  (HHD_ACCUM_FIXED).  A presynaptic spike of step k adds its synapses' amplitudes to slots
                                  k + delay right away (k_synapse); the state-update pass of step s + 1 consumes and clears slot s. */
  int n_slots;                 /* max_delay + 1 */
  long long n_pad;             /* neurons rounded up to whole 32-neuron blocks */
  /* synapses, CSR by global presynaptic neuron, rows sorted by (delay, original index) */
  const long long* row_ptr;    /* [NG+1] */
  SynHot* hot;
  SynStp* stp;
  const unsigned short* row_maxd; /* [NG] largest delay (steps) in the row */
  const unsigned int* row_bkt; /* [NG][DELAY_BUCKETS] offset inside the row of the first synapse with delay >= b*bkt_width */
  int bkt_width;
  unsigned long long* dbg_times;   /* HHD_DEBUG_TIMES: [4096 steps][8] %globaltimer stamps (dbg_stamp) */
  int stp_chunks;                  /* spike-time tasks per spike */
  long long syn_id_base;
  /* spike ring */
  int* spk_neuron;             /* global ids */
  int* spk_step;
  unsigned int cap_mask;
  int* spk_hist;               /* [NG][hist_depth] steps of the latest spikes of every neuron, newest first (note_spike); NULL while
                                  max_delay < refractory steps */
  int hist_depth;
  int* spk_last;               /* [NG] step of the latest spike of every (global) neuron, written when the spike is appended */
  unsigned long long* head;    /* monotone count of spikes appended */
  unsigned long long* step_end;/* [W] head value at the end of step (k % W) */
  double* prev_spike;          /* [NG] last_spike of a neuron before its latest spike (read by the STP update of that spike) */
  unsigned int* pass_done;     /* CTAs of k_neuron that finished the current pass (the last one closes it) */
  double* mon_part;            /* [MAX_RED][grid] per-CTA partial sums of the reduced monitors of the current pass */
  /* external input events, grouped by step */
  const int* ext_ptr;          /* [ext_last - ext_first + 2] */
  long long ext_first, ext_last;
  const int* ext_tgt;
  const unsigned char* ext_mask;
  const double *ext_gmax, *ext_pfail;
  const unsigned int* ext_id;
  MonDesc* mon;                /* [MAX_MON] device */
  unsigned long long* counters;
  unsigned long long* del_slots;  /* [CNT_SLOTS] */
  const long long* base_step;
  /* TimedArray current I_AC: table[iac_steps][iac_cols], column of local neuron i = iac_col[i] (-1: none); NULL = unused */
  const double* iac_table;
  const double* noise_sigma;   /* [N] conductance noise, nS/sqrt(ms) (hhd_set_noise); NULL = none.  Handled by the TIMED variants. */
  double sqrt_dt;
  const int* iac_col;
  long long iac_steps;
  int iac_cols;
  /* multi-GPU spike exchange (n_ranks > 1): k_neuron appends local spikes to xchg_send = {count, ids...}; after the
   * all-gather k_append copies every rank's list into the ring */
  int n_ranks, xchg_cap;
  int* xchg_send;              /* [1 + xchg_cap] */
  int* xchg_recv;              /* [n_ranks][1 + xchg_cap] */
  /* peer-memory exchange (send_spike / push_spikes / remote_spike_phase): every rank's receive buffer [2 parities][n_ranks][1 + xchg_cap] 64-bit
   * words, mapped into this process (CUDA IPC over NVLink).  Every word validates itself: (payload << 32) | tag, tag = step + 1 —
   * word 0 of a rank's slot carries its spike count, words 1.. one neuron id each — so the sender needs no fence and no separate
   * flag, and the receiver polls the word it is about to use (the "LL" idea of NCCL).  peer_buf[rank] is the own buffer; NULL =
   * the NCCL all-gather is used instead */
  int rank;
  unsigned long long* peer_buf[MAX_RANKS];
  unsigned int* xchg_dead;     /* set when a peer did not arrive within the time limit: later steps do not wait again */
  unsigned long long* xchg_done;/* [0]: k + 1 once CTA 0 of k_synapse has appended the other ranks' spikes of step k; [2 + (k & 1)]: the ring
                                 * head after k_neuron(k) appended this rank's own spikes (written by its last CTA, push_spikes) */
  /* cluster-resident plan: the network is n_inst independent instances of inst_size neurons, one thread-block cluster each */
  int inst_size, n_inst, res_cap, res_cluster;
  unsigned int* res_store;     /* [n_inst][2 + W + 2*res_cap] head, pad, end[W], neuron[cap], step[cap]: the instance's spike list between launches */
};

/* ---------------------------------------------------------------------------------------------------------------- */

/* I_AC of a neuron whose column is `col` at step s: TimedArray semantics (first row before the table, last row after). */
__device__ __forceinline__ double iac_at(const Params& P, int col, long long s) {
  if (col < 0) return 0.0;
  s = s < 0 ? 0 : (s >= P.iac_steps ? P.iac_steps - 1 : s);
  return __ldg(P.iac_table + (size_t)s * P.iac_cols + col);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

/* ---- cp.async (LDGSTS) helpers: global -> shared without staging through registers ----------------------------- */
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

/* ---- TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP) + mbarrier transaction counting -------------------------- */
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* The same with an L2 eviction-priority hint (createpolicy): the state rows of a 1M-neuron network (64 MB) fit in the
 * 126 MB L2 and are read again one step later, the parameter rows (96 MB) only stream through. */
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(double* p, double v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;\n" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

/* Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may be
 * scheduled while its predecessor in the stream is still draining; pdl_wait() blocks until that predecessor has
 * completed and its writes are visible, pdl_trigger() lets the successor start being scheduled.  Both are no-ops for a
 * normal launch. */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifndef HHD_EXPERIMENT
#define HHD_EXPERIMENT 0          /* timing experiments only (2: no integration; 4: no evaluation of the stored state either) */
#endif
#ifndef HHD_L2_HINTS
#define HHD_L2_HINTS 1            /* L2 eviction priorities: keep the state rows, stream the parameter rows */
#endif
#ifndef HHD_NEURON_MINBLOCKS_EULER
#define HHD_NEURON_MINBLOCKS_EULER 4   /* CTAs per SM the state-update kernel is compiled for (registers: 128 at 4, 96 at 5) */
#endif
#ifndef HHD_NEURON_MINBLOCKS_RK2
#define HHD_NEURON_MINBLOCKS_RK2 4
#endif
#ifndef HHD_NEURON_MINBLOCKS_RK4
#define HHD_NEURON_MINBLOCKS_RK4 3     /* also the adaptive method */
#endif

/*
 * Neuron data in HBM: blocks of 32 neurons ("warp tiles"), each block = NB_FIELDS rows of 32 doubles — 8 state rows
 * (V, w, g_X_on/off) and 12 parameter rows — contiguous (5120 B), so that ONE bulk copy (cp.async.bulk, SASS UBLKCP)
 * brings everything a warp needs for its tile, and the eight state rows a warp writes back are eight full 256-byte lines.
 * The pending synaptic increments (3 x 32 x 8 B per block and parity) and the meta words (32 x 4 B) are blocked the same
 * way and arrive on the same mbarrier.  (Round 1 kept 20 separate arrays: 20 copies per 128-neuron tile on ONE mbarrier
 * per CTA, which tied the four warps of a CTA together tile by tile — 20 % of all warp stalls were that barrier.)
 */
#define NB_LANES 32
#define NB_STATE 8
#define NB_PARAMS 9                                         /* parameter rows that are streamed every step */
#define NB_FIELDS (NB_STATE + NB_PARAMS)                    /* 17 */
#define NB_BYTES (NB_FIELDS * NB_LANES * 8)                 /* 4352 */
/* Three of the twelve parameters are needed by almost no neuron-step and are NOT streamed: b and V_r only by the reset of a
 * neuron that spiked (0.015 % of the neuron-steps at 3 Hz), V_dep only while the depolarisation-block window of 5 ms after a
 * spike is open (1.5 %).  They live in a plain side array (P.rare) and are fetched by the lanes that need them — 24 of the
 * 240 algorithmic bytes per neuron-update are not moved at all. */
enum { ROW_C = 0, ROW_GL, ROW_EL, ROW_DELTAT, ROW_VUP, ROW_TAUW, ROW_VT, ROW_IREF, ROW_IDC };
enum { RARE_B = 0, RARE_VR, RARE_VDEP, N_RARE };
__device__ __host__ __forceinline__ int param_row(int p) {      /* HHD_P_* -> streamed row (>= 0) or -1 - rare row */
  switch (p) {
    case HHD_P_C: return ROW_C;          case HHD_P_GL: return ROW_GL;      case HHD_P_EL: return ROW_EL;
    case HHD_P_DELTAT: return ROW_DELTAT; case HHD_P_VUP: return ROW_VUP;    case HHD_P_TAUW: return ROW_TAUW;
    case HHD_P_VT: return ROW_VT;        case HHD_P_IREF: return ROW_IREF;  case HHD_P_IDC: return ROW_IDC;
    case HHD_P_B: return -1 - RARE_B;    case HHD_P_VR: return -1 - RARE_VR;
  }
  return -1 - RARE_VDEP;
}
#define PEND_BYTES (3 * NB_LANES * 8)                       /* 768 */
#define META_BYTES (NB_LANES * 4)                           /* 128 */

__device__ __host__ __forceinline__ size_t nb_index(int field, long long i) {
  return ((size_t)(i >> 5) * NB_FIELDS + (size_t)field) * NB_LANES + (size_t)(i & 31);
}
__device__ __host__ __forceinline__ size_t pend_index(int c, long long i) {
  return ((size_t)(i >> 5) * 3 + (size_t)c) * NB_LANES + (size_t)(i & 31);
}

__device__ __forceinline__ NeuronPar load_par(const Params& P, int i) {
  NeuronPar p;
  const double* b = P.nb + nb_index(NB_STATE, i);
  p.C = __ldg(b + ROW_C * NB_LANES); p.gL = __ldg(b + ROW_GL * NB_LANES); p.EL = __ldg(b + ROW_EL * NB_LANES);
  p.dT = __ldg(b + ROW_DELTAT * NB_LANES); p.Vup = __ldg(b + ROW_VUP * NB_LANES); p.tauw = __ldg(b + ROW_TAUW * NB_LANES);
  p.VT = __ldg(b + ROW_VT * NB_LANES); p.Iref = __ldg(b + ROW_IREF * NB_LANES); p.IDC = __ldg(b + ROW_IDC * NB_LANES);
  p.b = __ldg(P.rare + (size_t)RARE_B * P.n_pad + i); p.Vr = __ldg(P.rare + (size_t)RARE_VR * P.n_pad + i);
  p.Vdep = __ldg(P.rare + (size_t)RARE_VDEP * P.n_pad + i);
  return p;
}

__device__ __forceinline__ void dbg_stamp(const Params& P, const long long k, const int q) {
  if (P.dbg_times) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    P.dbg_times[(size_t)(k & 4095) * 8 + q] = g;
  }
}

__device__ __forceinline__ void push_spikes(const Params& P, const long long k);
__device__ __forceinline__ void send_spike(const Params& P, const long long k, const int pos, const int id);

/* What the appender of a spike (k_neuron for the rank's own neurons, remote_spike_phase / k_append for the other ranks') records besides the
 * ring entry: the neuron's previous spike time for the STP update that follows (p0..p6 read it, possibly from many CTAs),
 * the new one (p5: last_spike_pre = t, only neurons with outgoing synapses have that pathway), the step of its latest
 * spike, and — when delays can outlast the refractory period — its spike history.  Returns the has_out flags.
 * Two spikes of a neuron are at least refr_steps apart, and hist_depth = max_delay / refr_steps + 2: spikes that can both
 * be in flight never share a slot, and a reader just skips entries whose age is not in (0, max_delay]. */
__device__ __forceinline__ int hist_slot(const Params& P, const int k) { return (k / max(P.refr_steps, 1)) % P.hist_depth; }

__device__ __forceinline__ unsigned char note_spike(const Params& P, const int j, const int k, const double t) {
  const unsigned char ho = P.has_out[j];
  P.spk_last[j] = k;
  P.prev_spike[j] = P.last_spike[j];
  if (ho & 1) P.last_spike[j] = t;
  if (P.spk_hist) P.spk_hist[(size_t)j * P.hist_depth + hist_slot(P, k)] = k;
  return ho;
}

/* One warp's pipeline in shared memory: two stages of the 5 KB block, ONE buffer for the block's pending increments and
 * meta words (they are taken into registers first thing, so the next tile's can be requested right away).  11 136 B per
 * warp: five 128-thread CTAs per SM fit (5 x (44.5 KB + 1 KB) < 228 KB). */
struct __align__(128) WarpStage {
  double f[NB_FIELDS][NB_LANES];
};
struct __align__(128) WarpAux {
  long long pend[3][NB_LANES];
  int meta[NB_LANES];
};
#define NEURON_STAGES 2
#define NEURON_WARPS (NEURON_BLOCK / 32)
#define NEURON_WARP_SMEM (NEURON_STAGES * (int)sizeof(WarpStage) + (int)sizeof(WarpAux))
#define NEURON_SMEM (NEURON_WARPS * NEURON_WARP_SMEM)

template <bool MON>
struct NeuronShared {            /* static shared memory of a CTA that runs neuron_phase */
  double red[MON ? MAX_RED : 1][NEURON_WARPS];
  MonDesc mon[MON ? MAX_MON : 1];/* read once per CTA instead of once per tile from global memory */
  unsigned long long cnt[3];
  int is_last;
  unsigned long long full_bar[NEURON_WARPS][NEURON_STAGES + 1];   /* [..][NEURON_STAGES]: the pend/meta buffer */
};

/*
 * State update of time step k, persistent over warp tiles.  Every WARP runs its own two-stage pipeline: lane 0 issues
 * three bulk copies per tile (block, pending increments, meta words) onto the warp's mbarrier of that stage two tiles
 * ahead; the warp waits on the barrier, takes its tile into registers, hands the stage back (__syncwarp + proxy fence) and
 * integrates.  No CTA-wide barrier inside the loop: a warp never waits for another warp.
 * METHOD: 0 euler, 1 rk2, 2 rk4, 3 adaptive rk2(3) (gsl_rk2).  STEP = false: only finish the previous step (once at the end
 * of hhd_run, so that the state read back is the state after `resets` / `after_resets`).  TIMED: TimedArray current and
 * conductance noise.  MON: state monitors are sampled (a variant without any monitor code runs when none is active).
 * meta = lastspike_step << META_SHIFT | flags (FLAG_SPIKED | FLAG_HAS_PREV of the previous step; FLAG_OUT, FLAG_ZERO static).
 */
template <int METHOD, bool STEP, bool TIMED, bool MON>
__device__ __forceinline__ void neuron_phase(const Params& P, const long long k, unsigned char* smem, NeuronShared<MON>& sh) {
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  const int warp = tid >> 5;
  const double t = (double)k * P.dt_s;
  const int nblocks = (int)(P.n_pad >> 5);
  if (STEP && blockIdx.x == 0 && tid == 0) dbg_stamp(P, k, 0);
  if (tid < 3) sh.cnt[tid] = 0ull;
  if (MON && tid < P.n_mon) sh.mon[tid] = P.mon[tid];
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s <= NEURON_STAGES; ++s) mbar_init(&sh.full_bar[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  WarpStage* const st = reinterpret_cast<WarpStage*>(smem + (size_t)warp * NEURON_WARP_SMEM);
  WarpAux* const aux = reinterpret_cast<WarpAux*>(st + NEURON_STAGES);
  unsigned long long* const bar = sh.full_bar[warp];
  long long* const pin = P.pend + (size_t)((k + P.n_slots - 1) % P.n_slots) * 3 * (size_t)P.n_pad;     /* slot k-1 */
  const int gwarp = blockIdx.x * NEURON_WARPS + warp, nwarps = gridDim.x * NEURON_WARPS;

#if HHD_L2_HINTS
  const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
#endif
  auto issue = [&](int blk, int s) {           /* lane 0 only */
    if (blk < nblocks) {
      mbar_expect_tx(&bar[s], NB_BYTES);
#if HHD_L2_HINTS
      const double* src = P.nb + (size_t)blk * (NB_FIELDS * NB_LANES);
      tma_load_1d_hint(&st[s].f[0][0], src, NB_STATE * NB_LANES * 8, &bar[s], pol_keep);
      tma_load_1d_hint(&st[s].f[NB_STATE][0], src + NB_STATE * NB_LANES, NB_PARAMS * NB_LANES * 8, &bar[s], pol_stream);
#else
      tma_load_1d(&st[s].f[0][0], P.nb + (size_t)blk * (NB_FIELDS * NB_LANES), NB_BYTES, &bar[s]);
#endif
    }
  };
  auto issue_aux = [&](int blk) {              /* lane 0 only */
    if (blk < nblocks) {
      mbar_expect_tx(&bar[NEURON_STAGES], PEND_BYTES + META_BYTES);
      tma_load_1d(&aux->pend[0][0], pin + (size_t)blk * (3 * NB_LANES), PEND_BYTES, &bar[NEURON_STAGES]);
      tma_load_1d(&aux->meta[0], P.meta + (size_t)blk * NB_LANES, META_BYTES, &bar[NEURON_STAGES]);
    }
  };

  double mon_acc[MAX_RED];
#pragma unroll
  for (int r = 0; r < MAX_RED; ++r) mon_acc[r] = 0.0;
  unsigned n_spk = 0, n_wc = 0, n_bad = 0;
  unsigned parity = 0;                          /* bit s: phase of the stage's barrier */

  int blk = gwarp;
  if (lane == 0) {
    issue_aux(blk);
    issue(blk, 0);
    issue(blk + nwarps, 1);
  }
  /* last_spike lives per GLOBAL neuron (other ranks' neurons too) and is fetched TWO tiles ahead into a register: the
   * tile in between is the one whose V_dep is requested — only by the lanes whose block window is open — from the
   * last_spike that arrived during the previous tile.  (One tile ahead, the gate `t - last_spike` waited for the load it
   * had just issued: 19 % of the kernel's stall samples, profiles/r2_notes.md.) */
  auto ls_load = [&](int b) -> double {
    const int in = b * NB_LANES + (int)lane;
    return (b < nblocks && in < P.N) ? P.last_spike[P.off + in] : 0.0;
  };
  auto vdep_load = [&](int b, double lsb) -> double {
    const int in = b * NB_LANES + (int)lane;
    return (b < nblocks && in < P.N && (t - lsb < P.cc.window_s)) ? __ldg(P.rare + (size_t)RARE_VDEP * P.n_pad + in) : 0.0;
  };
  double ls_next = ls_load(blk);
  double ls_next2 = ls_load(blk + nwarps);
  double vdep_next = vdep_load(blk, ls_next);
  int col_next = -1;
  if (TIMED && P.iac_table && blk < nblocks && blk * NB_LANES + (int)lane < P.N) col_next = P.iac_col[blk * NB_LANES + lane];

  for (int it = 0; blk < nblocks; ++it, blk += nwarps) {
    const int s = it & 1;
    const int i = blk * NB_LANES + lane;
    const bool live = i < P.N;
    mbar_wait(&bar[NEURON_STAGES], (parity >> NEURON_STAGES) & 1u);
    mbar_wait(&bar[s], (parity >> s) & 1u);
    parity ^= (1u << s) | (1u << NEURON_STAGES);
    double x[8];
    NeuronPar p;
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] = st[s].f[v][lane];
    p.C = st[s].f[NB_STATE + ROW_C][lane]; p.gL = st[s].f[NB_STATE + ROW_GL][lane]; p.EL = st[s].f[NB_STATE + ROW_EL][lane];
    p.dT = st[s].f[NB_STATE + ROW_DELTAT][lane]; p.Vup = st[s].f[NB_STATE + ROW_VUP][lane]; p.tauw = st[s].f[NB_STATE + ROW_TAUW][lane];
    p.VT = st[s].f[NB_STATE + ROW_VT][lane]; p.Iref = st[s].f[NB_STATE + ROW_IREF][lane]; p.IDC = st[s].f[NB_STATE + ROW_IDC][lane];
    p.b = 0.0; p.Vr = 0.0;            /* fetched by the reset below */
    p.Vdep = vdep_next;               /* fetched one tile ahead for the lanes whose block window is open */
    long long pd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) pd[c] = aux->pend[c][lane];
    const int meta = aux->meta[lane];
    const double ls = ls_next;
    const int col = col_next;
    const int i_cur = i;
    /* the warp holds its tile in registers: hand the stage back and request the tile after next */
    __syncwarp();
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      issue_aux(blk + nwarps);
      issue(blk + 2 * nwarps, s);
    }
    {
      const int in = (blk + nwarps) * NB_LANES + (int)lane;
      const bool nlive = blk + nwarps < nblocks && in < P.N;
      if (TIMED && P.iac_table) col_next = nlive ? P.iac_col[in] : -1;
      /* V_dep multiplies the gate int(t - last_spike < 5 ms), which is closed at every stage time if it is closed at t */
      ls_next = ls_next2;                               /* requested one tile ago */
      vdep_next = vdep_load(blk + nwarps, ls_next);
      ls_next2 = ls_load(blk + 2 * nwarps);
    }

    SubExpr S;
    bool wcross = false;
    double iac0 = 0.0, iac1 = 0.0;
#if HHD_EXPERIMENT & 4
    S.I_tot = x[0] + p.C + p.gL + p.EL + p.dT + p.Vup + p.tauw + p.VT + p.Iref + p.Vdep + p.IDC + (double)pd[0] + (double)pd[1] + (double)pd[2];   /* timing experiment: data movement only */
    x[1] += S.I_tot * 1e-300;
#else
    {
      /* ---- the part of step k-1 that had to wait for that step's synaptic deliveries ------------------------------
       * x is the state right after the integration of step k-1.  Its subexpressions S decide the `w_crossing` event of
       * step k-1 (slot `thresholds`, codes/CortexNetwork.py:306) and — unless this neuron received input or was reset,
       * and then only their linear part changes — ARE the subexpressions of the first stage of step k: one evaluation
       * (two exponentials, three divisions) per neuron-step instead of two. */
      double iac_prev = 0.0, iac_now = 0.0, iac_nxt = 0.0;      /* I_AC at steps k-1, k, k+1 */
      if (TIMED && P.iac_table) {
        iac_prev = iac_at(P, col, k - 1);
        iac_now = iac_at(P, col, k);
        iac_nxt = iac_at(P, col, k + 1);
      }
      hhd_subexpr(P.cc, p, x, S, iac_prev);
      if (meta & FLAG_HAS_PREV) {
        const double band = hhd_div(S.D0, p.tauw);
        wcross = (x[1] > S.w_V - band) && (x[1] < S.w_V + band) && (x[0] <= p.VT);
      }
      bool new_g = false;
#pragma unroll
      for (int c = 0; c < 3; ++c) {          /* slot `synapses`: g_X_on_post += ..., g_X_off_post += ... (:367-372) */
        if (pd[c]) {
          const double inc = P.accum == HHD_ACCUM_FIXED ? (double)pd[c] * (1.0 / 1099511627776.0) : __longlong_as_double(pd[c]);
          x[2 + 2 * c] += inc;
          x[3 + 2 * c] += inc;
          pin[pend_index(c, i)] = 0;
          new_g = true;
        }
      }
      if (meta & FLAG_SPIKED) {        /* slot `resets`: V = V_r; w += b   (codes/CortexNetwork.py:305) */
        p.Vr = __ldg(P.rare + (size_t)RARE_VR * P.n_pad + i_cur);
        p.b = __ldg(P.rare + (size_t)RARE_B * P.n_pad + i_cur);
        x[0] = p.Vr;
        x[1] += p.b;
        hhd_subexpr(P.cc, p, x, S, iac_prev);
      } else if (new_g) {
        hhd_subexpr_new_g(P.cc, p, x, S);
      }
      if (wcross) x[1] = S.w_V - hhd_div(S.D0, p.tauw);      /* slot `after_resets`: w = w_V - D0/tau_w   (:311) */
      if (TIMED && iac_now != iac_prev) {  /* the clock moves on: I_inj of step k; again only the linear part changes */
        S.I_inj = p.IDC + iac_now;
        hhd_subexpr_new_g(P.cc, p, x, S);
      }
      iac0 = iac_now;
      iac1 = iac_nxt;
    }
#endif
    double* const xo = P.nb + nb_index(0, i);
    if (!STEP) {
#pragma unroll
      for (int v = 0; v < 8; ++v) xo[v * NB_LANES] = x[v];
      P.meta[i] = meta & ~3;
      n_wc += wcross && live;
      continue;
    }

    /* ---- slot `start`: monitors sample the pre-update state (codes/CortexNetwork.py:510) ----------------------------- */
    if (MON) {
      for (int m = 0; m < P.n_mon; ++m) {
        const MonDesc& md = sh.mon[m];
        if (!md.active) continue;
        const long long sample = k - md.k0;
        if (sample < 0 || sample >= md.capacity) {
          if (i == 0) atomicAdd(&P.counters[CNT_OVERFLOW], 1ull);
          continue;
        }
        int slot = -1;
        if (live) slot = md.slot ? md.slot[i] : i;
        if (slot >= 0) {
          const double v = hhd_var_is_derived(md.var) ? hhd_var_from_subexpr(S, md.var) : hhd_var_state(x, md.var, ls);
          if (md.reduce == HHD_REDUCE_NONE) md.buf[sample * md.n_idx + slot] = v;
          else {                 /* branch-free, so that the accumulators stay in registers (an indexed update put them in local memory) */
#pragma unroll
            for (int r = 0; r < MAX_RED; ++r) mon_acc[r] += (r == md.red_slot) ? v : 0.0;
          }
        }
      }
    }

    /* ---- slot `groups`: state update (:308-309) ------------------------------------------------------------ */
#if !(HHD_EXPERIMENT & 2)
    double hcar = METHOD == 3 ? P.hstep[i] : 0.0;
    hhd_integrate<METHOD>(P.cc, p, x, t, ls, P.dt, S, iac0, iac1, &hcar);
    if (METHOD == 3) P.hstep[i] = hcar;
    if (TIMED && METHOD == 0 && P.noise_sigma && live) {
      /* Euler-Maruyama term of Brian's stochastic `euler` updater: x + dt f(x) + xi sigma, xi = sqrt(dt) N(0,1) */
      const double sg = P.noise_sigma[i] * P.sqrt_dt;
      if (sg != 0.0) {
        double z0, z1, z2, z3;
        hhd_normal2(P.seed, (unsigned long long)(P.off + i), (unsigned long long)k, 2, &z0, &z1);
        hhd_normal2(P.seed, (unsigned long long)(P.off + i), (unsigned long long)k, 3, &z2, &z3);
        x[3] = x[3] + z0 * sg;
        x[5] = x[5] + z1 * sg;
        x[7] = x[7] + z2 * sg;
      }
    }
#endif
    const bool bad = live && !(isfinite(x[0]) && isfinite(x[1]));
    /* ---- slot `thresholds` (:304); the w_crossing test of this step runs at the start of the next pass ------------ */
    const int lastk = meta >> META_SHIFT;
    const bool spiked = live && (x[0] > p.Vup) && ((k - (long long)lastk) >= (long long)P.refr_steps);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
#if HHD_L2_HINTS
      st_hint(xo + v * NB_LANES, x[v], pol_keep);
#else
      xo[v * NB_LANES] = x[v];
#endif
    }
    P.meta[i] = ((spiked ? (int)k : lastk) << META_SHIFT) | (meta & (FLAG_OUT | FLAG_ZERO)) | (spiked ? FLAG_SPIKED : 0) | FLAG_HAS_PREV;
    /* ---- warp-ballot spike compaction: one atomic per warp that has spikes ----------------------------------- */
    const unsigned bal = __ballot_sync(0xffffffffu, spiked);
    if (bal) {
      if (P.n_ranks > 1) {          /* partitioned network: the step's local spikes also go to the exchange's send list */
        int base = 0;
        if (lane == 0) base = atomicAdd(&P.xchg_send[0], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (spiked) {
          const int pos = base + __popc(bal & ((1u << lane) - 1u));
          if (pos < P.xchg_cap) {
            P.xchg_send[1 + pos] = P.off + i;                       /* the list the NCCL fallback gathers */
            if (P.peer_buf[P.rank]) send_spike(P, k, pos, P.off + i);
          } else atomicAdd(&P.counters[CNT_OVERFLOW], 1ull);
        }
      }
      {                             /* the ring entry and what note_spike records, from registers only */
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.head, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (spiked) {
          const unsigned pos = (unsigned)(base + __popc(bal & ((1u << lane) - 1u))) & P.cap_mask;
          const int j = P.off + i;
          P.spk_neuron[pos] = j;
          P.spk_step[pos] = (int)k;
          P.spk_last[j] = (int)k;
          P.prev_spike[j] = ls;
          if (meta & FLAG_OUT) P.last_spike[j] = t;
          if (P.spk_hist) P.spk_hist[(size_t)j * P.hist_depth + hist_slot(P, (int)k)] = (int)k;
        }
      }
    }
    n_spk += spiked;
    n_wc += wcross && live;
    n_bad += bad;
  }
  if (!STEP) {                      /* only the w_crossing events found while finishing the last step are left to count */
    n_wc = __reduce_add_sync(0xffffffffu, n_wc);
    if (lane == 0 && n_wc) atomicAdd(&P.counters[CNT_WCROSS], (unsigned long long)n_wc);
    return;
  }

  /* ---- once per CTA: reduced monitors (LFP) and counters ------------------------------------------------------- */
  if (MON) {
#pragma unroll
    for (int r = 0; r < MAX_RED; ++r) {
      const double ws = warp_sum(mon_acc[r]);
      if (lane == 0) sh.red[r][warp] = ws;
    }
  }
  n_spk = __reduce_add_sync(0xffffffffu, n_spk);
  n_wc = __reduce_add_sync(0xffffffffu, n_wc);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if (lane == 0) {
    if (n_spk) atomicAdd(&sh.cnt[0], (unsigned long long)n_spk);
    if (n_wc) atomicAdd(&sh.cnt[1], (unsigned long long)n_wc);
    if (n_bad) atomicAdd(&sh.cnt[2], (unsigned long long)n_bad);
  }
  __syncthreads();
  const bool closes = MON || P.n_ranks > 1;      /* someone has to know when the whole grid is done */
  if (tid == 0) {
    if (sh.cnt[0]) atomicAdd(&P.counters[CNT_SPIKES], sh.cnt[0]);
    if (sh.cnt[1]) atomicAdd(&P.counters[CNT_WCROSS], sh.cnt[1]);
    if (sh.cnt[2]) atomicAdd(&P.counters[CNT_NONFINITE], sh.cnt[2]);
    if (MON) {
      /* per-CTA partial sums of the reduced monitors; the CTA that finishes last adds them in CTA order, so the trace does
       * not depend on the order in which CTAs finish (round 1 used one fp64 atomicAdd per CTA: run-to-run different bits) */
#pragma unroll
      for (int r = 0; r < MAX_RED; ++r) {
        double bsum = 0.0;
#pragma unroll
        for (int q = 0; q < NEURON_WARPS; ++q) bsum += sh.red[r][q];
        P.mon_part[(size_t)r * gridDim.x + blockIdx.x] = bsum;
      }
    }
    if (closes) {
      __threadfence();
      sh.is_last = atomicAdd(P.pass_done, 1u) == gridDim.x - 1;
    }
  }
  if (!closes) return;
  __syncthreads();
  if (!sh.is_last) return;
  __threadfence();
  if (MON) {
    for (int m = warp; m < P.n_mon; m += NEURON_WARPS) {
      const MonDesc& md = sh.mon[m];
      if (!md.active || md.reduce == HHD_REDUCE_NONE) continue;
      const long long sample = k - md.k0;
      if (sample < 0 || sample >= md.capacity) continue;
      /* fixed order: lane l adds CTAs l, l + 32, ...; then the butterfly */
      double acc = 0.0;
      const volatile double* part = P.mon_part + (size_t)md.red_slot * gridDim.x;
      for (unsigned b = lane; b < gridDim.x; b += 32) acc += part[b];
      acc = warp_sum(acc);
      if (lane == 0) md.buf[sample] = md.buf[sample] + acc;
    }
  }
  if (P.n_ranks > 1 && P.peer_buf[P.rank]) push_spikes(P, k);     /* the exchange's send half (exchange over peer memory) */
  if (tid == 0) *P.pass_done = 0u;
}

/*
 * METHOD: 0 euler, 1 rk2, 2 rk4, 3 adaptive rk2(3).  STEP = false: only finish the previous step (used once at the end of
 * hhd_run so that the state read back is the state after `resets`/`after_resets`).
 */
template <int METHOD, bool STEP, bool TIMED = false, bool MON = false>
__global__ void __launch_bounds__(NEURON_BLOCK, (METHOD == 0 ? HHD_NEURON_MINBLOCKS_EULER : METHOD == 1 ? HHD_NEURON_MINBLOCKS_RK2 : HHD_NEURON_MINBLOCKS_RK4)) k_neuron(const __grid_constant__ Params P, const int step_offset) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ NeuronShared<MON> sh;
  pdl_wait();
  neuron_phase<METHOD, STEP, TIMED, MON>(P, *P.base_step + step_offset, smem_raw, sh);
  pdl_trigger();     /* at the END: a dependent launched earlier would take the SM resources this grid's own CTAs still need */
}


/* buffer that collects the deliveries of slot s (Brian's `synapses` slot of step s) */
__device__ __forceinline__ long long* pend_of_slot(const Params& P, long long s) {
  return P.pend + (size_t)(s % P.n_slots) * 3 * (size_t)P.n_pad;
}
/* an amplitude that replaces an earlier one in a slot that has not been consumed yet (see spike_new_phase) */
__device__ __forceinline__ void add_correction(const Params& P, long long* pout, int c, int tgt, double amp_new, double amp_old) {
  if (P.accum == HHD_ACCUM_FIXED) {
    const long long q = __double2ll_rn(amp_new * 1099511627776.0) - __double2ll_rn(amp_old * 1099511627776.0);
    if (q) atomicAdd((unsigned long long*)&pout[pend_index(c, tgt)], (unsigned long long)q);
  } else if (amp_new != amp_old) {
    atomicAdd(reinterpret_cast<double*>(&pout[pend_index(c, tgt)]), amp_new - amp_old);
  }
}
__device__ __forceinline__ void add_increment(const Params& P, long long* pout, int c, int tgt, double inc) {
  /* summed per (channel, target) for this step; the target's next state-update pass adds the sum to g_on and g_off */
  if (P.accum == HHD_ACCUM_FIXED) {
    const long long q = __double2ll_rn(inc * 1099511627776.0);
    atomicAdd((unsigned long long*)&pout[pend_index(c, tgt)], (unsigned long long)q);
  } else {
    atomicAdd(reinterpret_cast<double*>(&pout[pend_index(c, tgt)]), inc);
  }
}

__device__ __forceinline__ void ext_events(const Params& P, const long long k, long long* const pout, const unsigned gtid, const unsigned gthreads);

/*
 * Slot `synapses` of step k (k_synapse).  Pathways p0..p6 (`u`, `R`, `a_syn`, failure draw at the presynaptic spike,
 * codes/CortexNetwork.py:351-366) for the spikes of THIS step, and the delayed pathways p7a..p9b (:367-372, `delay = dtax`
 * :463-468) as a write into the per-delay circular conductance buffer: the synapse's amplitude is added to slot k + delay
 * right away, so there is no delivery pass at all — an event costs one read of its two records (80 B), one write of
 * (u, R, amplitude) and one atomic, the 96 B of SURVEY 8d.  A task is one block of STP_BLOCK consecutive synapses of one
 * spike's row, stp_chunks tasks per spike, so a spike's row is rewritten by several CTAs at once; the previous spike time it
 * needs was saved by note_spike.
 * Brian reads `a_syn` and `failure` when the delayed pathway EXECUTES.  That only differs from the value at push
 * time if the neuron spikes again while an older spike of it is still in flight, which needs a delay beyond the refractory
 * period (LONG): the new spike then replaces, in every slot that has not been consumed yet, the amplitude the older spike
 * put there (old amplitude = the record's current one; slots s = k_older + delay >= k).  In fixed-point accumulation that
 * is exact; in fp64 accumulation the replacement is one rounded difference.
 * Also the SpikeGeneratorGroup events of the step (no delay).
 */
/* one task: block `chunk` of `chunks` of the outgoing row of neuron j, which spiked in step k (whole CTA) */
template <bool LONG>
__device__ __forceinline__ void spike_task(const Params& P, const long long k, const int j, const unsigned chunk, const unsigned chunks,
                                           int* const older_sh, unsigned long long& pushed) {
  const double t = (double)k * P.dt_s;
  const size_t slot_stride = (size_t)3 * (size_t)P.n_pad;
  /* everything the task needs of neuron j is requested in one go (row bounds, previous spike time, spike history, longest
   * delay): one global round trip instead of two before the synapse records can be requested */
  const long long r0 = P.row_ptr[j], r1 = P.row_ptr[j + 1];
  const double last = P.prev_spike[j];
  int hist_k = 0, maxd = 0;
  if (LONG && (int)threadIdx.x < P.hist_depth) {
    hist_k = P.spk_hist[(size_t)j * P.hist_depth + threadIdx.x];
    maxd = (int)P.row_maxd[j];
  }
  if (chunk == 0 && threadIdx.x == 0 && r1 > r0) atomicAdd(&P.counters[CNT_SYN_EVENTS], (unsigned long long)(r1 - r0));
  /* ages of the neuron's older spikes that can still have deliveries in flight (else -1): the same for every thread of the
   * task, so they are worked out once per task into shared memory */
  if (LONG) {
    __syncthreads();                                         /* the previous task's readers are done */
    if (threadIdx.x < 17) {
      int a = -1;
      if ((int)threadIdx.x < P.hist_depth) {
        const int age = (int)(k - (long long)hist_k);
        if (age > 0 && age <= maxd) a = age;
      }
      older_sh[threadIdx.x] = a;
    }
    __syncthreads();
  }
  for (long long r = r0 + (long long)chunk * blockDim.x + threadIdx.x; r < r1; r += (long long)chunks * blockDim.x) {
    const SynHot sh = P.hot[r];
    SynStp sy = P.stp[r];
    const double amp = hhd_stp_update(sy.u, sy.R, sy.U, sy.tau_rec, sy.tau_fac, sy.gmax, P.cc.gsyn[sh.chan], sy.pfail, t, last,
                                      P.seed, (unsigned long long)(P.syn_id_base + (long long)sy.orig), (unsigned long long)k);
    P.stp[r].u = sy.u;
    P.stp[r].R = sy.R;
    P.hot[r].amp = amp;
    const int d = (int)sh.delay;
    ++pushed;
    if (amp != 0.0) add_increment(P, P.pend + (size_t)((k + d) % P.n_slots) * slot_stride, sh.chan, sh.tgt, amp);
    if (LONG) {
      for (int q = 0; q < P.hist_depth; ++q) {
        const int a = older_sh[q];
        if (a > 0 && a <= d)                     /* the older spike's delivery of this synapse is still to come: slot k - a + d */
          add_correction(P, P.pend + (size_t)((k - a + d) % P.n_slots) * slot_stride, sh.chan, sh.tgt, amp, sh.amp);
      }
    }
  }
}

__device__ __forceinline__ void count_pushed(const Params& P, unsigned long long pushed, const unsigned bid) {
  pushed = (unsigned long long)warp_sum((double)pushed);
  if ((threadIdx.x & 31) == 0 && pushed) atomicAdd(&P.del_slots[(bid * 4 + (threadIdx.x >> 5)) & (CNT_SLOTS - 1)], pushed);
}

/* the spikes at ring positions [lo_new, hi): stp_chunks tasks each, dealt out to the CTAs */
template <bool LONG>
__device__ __forceinline__ void spike_new_phase(const Params& P, const long long k, const unsigned long long lo_new, const unsigned long long hi,
                                                const unsigned bid, const unsigned nblk, const bool with_ext) {
  __shared__ int older_sh[17];
  unsigned long long pushed = 0;
  const unsigned chunks = (unsigned)P.stp_chunks;
  const unsigned ntasks = (unsigned)(hi - lo_new) * chunks;
  for (unsigned task = bid; task < ntasks; task += nblk)
    spike_task<LONG>(P, k, P.spk_neuron[(unsigned)(lo_new + task / chunks) & P.cap_mask], task % chunks, chunks, older_sh, pushed);
  count_pushed(P, pushed, bid);
  if (with_ext) ext_events(P, k, pend_of_slot(P, k), bid * blockDim.x + threadIdx.x, nblk * blockDim.x);
}

/* Deliveries that were pushed into the conductance buffer but whose slot lies beyond the last completed step `last` (one warp
 * per spike of the last max_delay steps): hhd_get_counters reports syn_deliveries as Brian counts them, at execution. */
__global__ void k_inflight(const __grid_constant__ Params P, const long long last, const unsigned long long lo, const unsigned long long hi,
                           unsigned long long* out) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long n = 0;
  for (unsigned long long sp = lo + (((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); sp < hi; sp += warps) {
    const unsigned pos = (unsigned)sp & P.cap_mask;
    const int j = P.spk_neuron[pos];
    const long long ks = P.spk_step[pos];
    for (long long r = P.row_ptr[j] + lane; r < P.row_ptr[j + 1]; r += 32) n += (ks + (long long)P.hot[r].delay > last) ? 1ull : 0ull;
  }
  n = (unsigned long long)warp_sum((double)n);
  if (lane == 0 && n) atomicAdd(out, n);
}

/* SpikeGeneratorGroup events of step k: no STP, no delay, one failure draw per (spike, synapse) (:588-594) */
__device__ __forceinline__ void ext_events(const Params& P, const long long k, long long* const pout, const unsigned gtid, const unsigned gthreads) {
  if (P.ext_ptr && k >= P.ext_first && k <= P.ext_last) {
    const int e0 = P.ext_ptr[k - P.ext_first], e1 = P.ext_ptr[k - P.ext_first + 1];
    for (int e = e0 + (int)gtid; e < e1; e += (int)gthreads) {
      const double failure = hhd_philox_u01(P.seed, (unsigned long long)P.ext_id[e], (unsigned long long)k, 1) < P.ext_pfail[e] ? 1.0 : 0.0;
      const unsigned char mask = P.ext_mask[e];
      const double g = P.ext_gmax[e] * (1.0 - failure);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (mask & (1 << c)) add_increment(P, pout, c, P.ext_tgt[e], g * P.cc.gsyn[c]);
    }
    if (gtid == 0 && e1 > e0) atomicAdd(&P.counters[CNT_EXT], (unsigned long long)(e1 - e0));
  }
}

/* ================================================================================================================== *
 * Spike exchange of a partitioned network over peer memory (CUDA IPC over NVLink), split over the two launches of a step
 * so that it costs no launch of its own:
 *   send_spike    a lane of k_neuron whose neuron spikes stores (id, step tag) as ONE 64-bit word into its position of this
 *                 rank's list in EVERY peer's receive buffer — straight from the tile loop, no staging, no fence;
 *   push_spikes   the CTA of k_neuron that finishes LAST stores (count, step tag) into word 0 of that list on every peer;
 *   remote_spike_phase   every CTA of k_synapse first does its share of the LOCAL spikes' work (k_neuron put those into the
 *                 ring itself), then reads the other ranks' count words (tag = this step) and takes remote spike q = its index,
 *                 index + grid, ... whole: polls the id word (usually long there), files the spike in the ring behind the local
 *                 ones, note_spike, the row's spike-time work.  No CTA waits for another CTA.
 * Two parities suffice: a rank can only write step k + 2 after it has consumed everybody's step k + 1, which a peer only
 * sends after it has consumed step k; a stale word carries the tag of step k - 2 and never validates.  A peer that does not
 * arrive within 20 s (it died) is reported through CNT_TIMEOUT instead of hanging the device.  If the buffers cannot be
 * mapped the engine falls back to ncclAllGather + k_append.
 * HHD_DEBUG_TIMES: %globaltimer stamps per step — 0 k_neuron's first CTA starts, 1 its last CTA is done, 2 counts sent,
 * 3 k_synapse's CTA 0 starts, 4 its share of the local spikes done and all counts seen, 5 its share of the remote spikes done, 6 CTA 0 of k_synapse done (tools/step_stamps.py).
 * ================================================================================================================== */
__device__ __forceinline__ void st_peer(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_peer(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
/* spin until the word carries this step's tag; false if the peer never delivers (20 s: it died) */
__device__ __forceinline__ bool wait_peer_word(const Params& P, const unsigned long long* p, const unsigned tag, unsigned long long& v) {
  v = ld_peer(p);
  if ((unsigned)v == tag) return true;
  if (*(volatile unsigned int*)P.xchg_dead) return false;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    v = ld_peer(p);
    if ((unsigned)v == tag) return true;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull || *(volatile unsigned int*)P.xchg_dead) {
      if (!atomicExch(P.xchg_dead, 1u)) atomicAdd(&P.counters[CNT_TIMEOUT], 1ull);
      return false;
    }
  }
}

/* a spiking lane of k_neuron: its id straight into every peer's slot, position `pos` of this rank's list of step k */
__device__ __forceinline__ void send_spike(const Params& P, const long long k, const int pos, const int id) {
  const size_t slot = (size_t)1 + P.xchg_cap;
  const unsigned long long w = ((unsigned long long)(unsigned)id << 32) | (unsigned long long)(unsigned)(k + 1);
  for (int r = 0; r < P.n_ranks; ++r)
    if (r != P.rank) st_peer(P.peer_buf[r] + ((size_t)(k & 1) * P.n_ranks + P.rank) * slot + 1 + pos, w);
}

/* whole CTA (the last one of k_neuron to finish): the ids are on their way already, what is left is the count */
__device__ __forceinline__ void push_spikes(const Params& P, const long long k) {
  const int par = (int)(k & 1), me = P.rank, tid = threadIdx.x;
  const size_t slot = (size_t)1 + P.xchg_cap;
  if (tid == 0) dbg_stamp(P, k, 1);
  if (tid < P.n_ranks && tid != me) {
    const int n_mine = min(__ldcg(P.xchg_send), P.xchg_cap);
    st_peer(P.peer_buf[tid] + ((size_t)par * P.n_ranks + me) * slot, ((unsigned long long)(unsigned)n_mine << 32) | (unsigned long long)(unsigned)(k + 1));
  }
  __syncthreads();
  if (tid == 0) {
    P.xchg_send[0] = 0;
    P.xchg_done[2 + par] = *P.head;          /* every CTA of this pass has appended (pass_done): the rank's own spikes end here */
    dbg_stamp(P, k, 2);
  }
}

/* The other ranks' spikes of step k (every CTA of k_synapse): CTA q mod nblk takes remote spike q WHOLE — it polls the id word,
 * files the spike in the ring behind the local ones, does note_spike and then the spike-time work of the row, which holds local
 * targets only and is short or empty.  No CTA waits for another one: every CTA reads the ranks' count words itself. */
template <bool LONG>
__device__ __forceinline__ void remote_spike_phase(const Params& P, const long long k, const unsigned long long hi_local, const unsigned bid,
                                                   const unsigned nblk) {
  __shared__ int older_sh[17];
  __shared__ int count_sh[MAX_RANKS], base_sh[MAX_RANKS + 1], j_sh;
  const int par = (int)(k & 1), me = P.rank, tid = threadIdx.x;
  const size_t slot = (size_t)1 + P.xchg_cap;
  const unsigned tag = (unsigned)(k + 1);
  const unsigned long long* recv = P.peer_buf[me] + (size_t)par * P.n_ranks * slot;
  if (tid < P.n_ranks) {
    unsigned long long hdr = 0;
    count_sh[tid] = (tid != me && wait_peer_word(P, recv + (size_t)tid * slot, tag, hdr)) ? min((int)(hdr >> 32), P.xchg_cap) : 0;
  }
  __syncthreads();
  if (tid == 0) {
    if (bid == 0) dbg_stamp(P, k, 4);
    int b = 0;
    for (int r = 0; r < P.n_ranks; ++r) { base_sh[r] = b; b += count_sh[r]; }
    base_sh[P.n_ranks] = b;
  }
  __syncthreads();
  const int total = base_sh[P.n_ranks];
  unsigned long long pushed = 0;
  /* remote spike q goes to CTA nblk - 1 - q: the local spikes' tasks were dealt out from CTA 0 upwards, so with the usual
   * handful of spikes per step the two kinds of work land on different CTAs and run side by side */
  for (int q = (int)(nblk - 1u - bid); q < total; q += (int)nblk) {
    if (tid == 0) {
      int r = 0;
      while (r + 1 < P.n_ranks && q >= base_sh[r + 1]) ++r;
      unsigned long long w = 0;
      int j = -1;
      if (wait_peer_word(P, recv + (size_t)r * slot + 1 + (q - base_sh[r]), tag, w)) {
        j = (int)(w >> 32);
        const unsigned pos = (unsigned)(hi_local + (unsigned long long)q) & P.cap_mask;
        P.spk_neuron[pos] = j;
        P.spk_step[pos] = (int)k;
        note_spike(P, j, (int)k, (double)k * P.dt_s);
      }
      j_sh = j;
    }
    __syncthreads();
    const int j = j_sh;
    if (j >= 0) spike_task<LONG>(P, k, j, 0u, 1u, older_sh, pushed);
    __syncthreads();
  }
  count_pushed(P, pushed, bid);
  if (bid == 0 && tid == 0) {
    const unsigned long long hi = hi_local + (unsigned long long)total;
    *P.head = hi;
    P.step_end[(int)k % P.W] = hi;
    dbg_stamp(P, k, 5);
  }
}

#ifndef HHD_SYNAPSE_MINBLOCKS
#define HHD_SYNAPSE_MINBLOCKS 4    /* one wave of 4 CTAs per SM does every spike-time task of a step */
#endif
template <bool LONG>
__global__ void __launch_bounds__(STP_BLOCK, HHD_SYNAPSE_MINBLOCKS) k_synapse(const __grid_constant__ Params P, const int step_offset, const int stp_blocks) {
  pdl_wait();
  const long long k = *P.base_step + step_offset;
  (void)stp_blocks;
  const unsigned long long lo = P.step_end[((int)k + P.W - 1) % P.W];
  if (P.n_ranks > 1 && P.peer_buf[P.rank]) {
    /* Partitioned network, exchange over peer memory.  k_neuron put this rank's own spikes into the ring already: every CTA
     * does its share of their spike-time work first (by then the other ranks' words have usually landed), then its share of
     * the remote spikes. */
    const unsigned long long hi_local = __ldcg(P.xchg_done + 2 + (int)(k & 1));
    if (blockIdx.x == 0 && threadIdx.x == 0) dbg_stamp(P, k, 3);
    spike_new_phase<LONG>(P, k, lo, hi_local, blockIdx.x, gridDim.x, true);
    remote_spike_phase<LONG>(P, k, hi_local, blockIdx.x, gridDim.x);
  } else {
    const unsigned long long hi = *P.head;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.step_end[(int)k % P.W] = hi;
    spike_new_phase<LONG>(P, k, lo, hi, blockIdx.x, gridDim.x, true);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) dbg_stamp(P, k, 6);
  pdl_trigger();
}


/* ================================================================================================================== *
 * Cluster-resident plan (HHD_PLAN_RESIDENT): one thread-block cluster simulates one independent network instance of up
 * to 16 x 128 neurons for a whole segment of time steps inside ONE launch.  State and parameters live in registers,
 * pending conductance increments and the instance's active-spike list in (distributed) shared memory, and the two
 * phases of a step are separated by the hardware cluster barrier (~0.2 us) instead of kernel boundaries (~3 us each).
 * This is the latency-bound regime: the reference's own 1000-neuron column (BASELINE configs 0-1) and batches of
 * independent instances (config 2, one cluster per instance).  Same arithmetic, same slot order, same results as the
 * grid-wide plan.  Requires max_delay < refractory steps (true for any single column).
 * ================================================================================================================== */
namespace cg = cooperative_groups;

struct ResList {                 /* lives in the shared memory of CTA 0 of the cluster */
  unsigned int head;             /* spikes appended so far (monotone) */
  unsigned int pad;
};

template <int METHOD>
__global__ void __launch_bounds__(RES_BLOCK, 1) k_resident(const __grid_constant__ Params P, const int n_steps) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  /* layout: pend[3][128] (8 B each) | MonDesc[MAX_MON] | ResList | end[W] | neuron[cap] | step[cap] */
  unsigned long long* pend_sh = reinterpret_cast<unsigned long long*>(smem_raw);
  MonDesc* mon_sh = reinterpret_cast<MonDesc*>(pend_sh + 3 * RES_BLOCK);
  ResList* list_sh = reinterpret_cast<ResList*>(mon_sh + MAX_MON);
  unsigned int* end_sh = reinterpret_cast<unsigned int*>(list_sh + 1);
  int* lneuron_sh = reinterpret_cast<int*>(end_sh + P.W);
  int* lstep_sh = lneuron_sh + P.res_cap;

  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  const int CS = P.res_cluster;
  const int rank = (int)cluster.block_rank();
  const int inst = blockIdx.x / CS;
  const int il = rank * RES_BLOCK + tid;                 /* index inside the instance */
  const bool live = il < P.inst_size;
  const int i = inst * P.inst_size + il;                 /* local neuron index of this handle */
  const int cthreads = CS * RES_BLOCK, ctid = rank * RES_BLOCK + tid;
  const int cwarps = cthreads >> 5, cwid = ctid >> 5;
  const unsigned cap_mask = (unsigned)P.res_cap - 1u;
  const long long k0 = *P.base_step;

  /* remote views of CTA 0's list */
  ResList* L = cluster.map_shared_rank(list_sh, 0);
  unsigned int* Lend = cluster.map_shared_rank(end_sh, 0);
  int* Lneuron = cluster.map_shared_rank(lneuron_sh, 0);
  int* Lstep = cluster.map_shared_rank(lstep_sh, 0);

  for (int q = tid; q < 3 * RES_BLOCK; q += RES_BLOCK) pend_sh[q] = 0ull;
  if (tid < P.n_mon) mon_sh[tid] = P.mon[tid];
  unsigned int* store = P.res_store + (size_t)inst * (2 + P.W + 2 * P.res_cap);
  if (rank == 0) {
    if (tid == 0) { list_sh->head = store[0]; list_sh->pad = 0; }
    for (int q = tid; q < P.W; q += RES_BLOCK) end_sh[q] = store[2 + q];
    for (int q = tid; q < P.res_cap; q += RES_BLOCK) { lneuron_sh[q] = (int)store[2 + P.W + q]; lstep_sh[q] = (int)store[2 + P.W + P.res_cap + q]; }
  }

  double x[8];
  NeuronPar p;
  double ls = 0.0, hcar = P.dt;    /* hcar: the adaptive method's carried step size */
  int meta = 0;
  bool has_out = false;
  if (live) {
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] = P.nb[nb_index(v, i)];
    p = load_par(P, i);
    ls = P.last_spike[P.off + i];
    meta = P.meta[i];
    if (METHOD == 3) hcar = P.hstep[i];
    has_out = (P.has_out[P.off + i] & 1) != 0;
  }
  unsigned long long n_events = 0, n_deliv = 0, n_ext = 0;
  unsigned n_spk = 0, n_wc = 0, n_bad = 0;
  cluster.sync();

  for (int s = 0; s <= n_steps; ++s) {
    const long long k = k0 + s;
    const double t = (double)k * P.dt_s;
    /* ---------------- phase 1: finish step k-1, then (if s < n_steps) slots start / groups / thresholds of step k ------ */
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const unsigned long long raw = pend_sh[c * RES_BLOCK + tid];
        if (raw) {
          const double inc = P.accum == HHD_ACCUM_FIXED ? (double)(long long)raw * (1.0 / 1099511627776.0) : __longlong_as_double((long long)raw);
          x[2 + 2 * c] += inc;
          x[3 + 2 * c] += inc;
          pend_sh[c * RES_BLOCK + tid] = 0ull;
        }
      }
      if (meta & FLAG_SPIKED) {
        x[0] = p.Vr;
        x[1] += p.b;
        if (has_out) {                                   /* p5: last_spike_pre = t of the spike's step */
          ls = (double)(k - 1) * P.dt_s;
          P.last_spike[P.off + i] = ls;
        }
      }
      if (meta & FLAG_WCROSS) {
        SubExpr sx;
        hhd_subexpr(P.cc, p, x, sx);
        x[1] = sx.w_V - hhd_div(sx.D0, p.tauw);
      }
      meta &= ~3;
    }
    if (s == n_steps) break;

    for (int m = 0; m < P.n_mon; ++m) {
      const MonDesc& md = mon_sh[m];
      if (!md.active) continue;
      const long long sample = k - md.k0;
      if (sample < 0 || sample >= md.capacity) {
        if (i == 0) atomicAdd(&P.counters[CNT_OVERFLOW], 1ull);
        continue;
      }
      int slot = -1;
      if (live) slot = md.slot ? md.slot[i] : i;
      double v = 0.0;
      if (slot >= 0) v = hhd_var_value(P.cc, p, x, md.var, ls);
      if (md.reduce == HHD_REDUCE_NONE) {
        if (slot >= 0) md.buf[sample * md.n_idx + slot] = v;
      } else {
        const double ws = warp_sum(v);
        if (lane == 0 && ws != 0.0) atomicAdd(&md.buf[sample], ws);
      }
    }

    bool spiked = false, wcross = false;
    if (live) {
      SubExpr s1;
      hhd_subexpr(P.cc, p, x, s1);
      hhd_integrate<METHOD>(P.cc, p, x, t, ls, P.dt, s1, 0.0, 0.0, &hcar);
      n_bad += !(isfinite(x[0]) && isfinite(x[1]));
      const int lastk = meta >> META_SHIFT;
      spiked = (x[0] > p.Vup) && ((k - (long long)lastk) >= (long long)P.refr_steps);
      SubExpr sx;
      hhd_subexpr(P.cc, p, x, sx);
      const double band = hhd_div(sx.D0, p.tauw);
      wcross = (x[1] > sx.w_V - band) && (x[1] < sx.w_V + band) && (x[0] <= p.VT);
      meta = ((spiked ? (int)k : lastk) << META_SHIFT) | (meta & (FLAG_OUT | FLAG_ZERO)) | (spiked ? FLAG_SPIKED : 0) | (wcross ? FLAG_WCROSS : 0);
      n_wc += wcross;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, spiked);
    if (bal) {
      /* the instance's list (delivery) and the global ring (SpikeMonitor) */
      unsigned lbase = 0;
      unsigned long long gbase = 0;
      if (lane == 0) {
        lbase = atomicAdd(&L->head, (unsigned)__popc(bal));
        gbase = atomicAdd(P.head, (unsigned long long)__popc(bal));
      }
      lbase = __shfl_sync(0xffffffffu, lbase, 0);
      gbase = __shfl_sync(0xffffffffu, gbase, 0);
      if (spiked) {
        const unsigned o = __popc(bal & ((1u << lane) - 1u));
        Lneuron[(lbase + o) & cap_mask] = P.off + i;
        Lstep[(lbase + o) & cap_mask] = (int)k;
        const unsigned gpos = (unsigned)(gbase + o) & P.cap_mask;
        P.spk_neuron[gpos] = P.off + i;
        P.spk_step[gpos] = (int)k;
        ++n_spk;
      }
    }
    cluster.sync();

    /* ---------------- phase 2: slot `synapses` of step k, all warps of the cluster ---------------------------------- */
    const unsigned hi = L->head;
    const int kw = (int)((unsigned long long)k % (unsigned)P.W);          /* k mod W once; neighbours by wrap */
    const unsigned lo_new = Lend[kw == 0 ? P.W - 1 : kw - 1];
    const unsigned lo = Lend[kw + 1 == P.W ? 0 : kw + 1];
    /* spikes of this step: p0..p6 for the whole row (threads of the cluster stride over it), zero-delay deliveries */
    for (unsigned sp = lo_new; sp != hi; ++sp) {
      const int j = Lneuron[sp & cap_mask];
      const double last = P.last_spike[j];
      const long long r0 = P.row_ptr[j], r1 = P.row_ptr[j + 1];
      for (long long r = r0 + ctid; r < r1; r += cthreads) {
        SynStp sy = P.stp[r];
        const SynHot sh = P.hot[r];
        const double amp = hhd_stp_update(sy.u, sy.R, sy.U, sy.tau_rec, sy.tau_fac, sy.gmax, P.cc.gsyn[sh.chan], sy.pfail, t,
                                          last, P.seed, (unsigned long long)(P.syn_id_base + (long long)sy.orig), (unsigned long long)k);
        P.stp[r].u = sy.u;
        P.stp[r].R = sy.R;
        P.hot[r].amp = amp;
        if (sh.delay == 0) {
          ++n_deliv;
          if (amp != 0.0) {
            const int tl = sh.tgt - inst * P.inst_size;
            unsigned long long* dst = cluster.map_shared_rank(pend_sh, tl / RES_BLOCK) + sh.chan * RES_BLOCK + (tl % RES_BLOCK);
            if (P.accum == HHD_ACCUM_FIXED) atomicAdd(dst, (unsigned long long)__double2ll_rn(amp * 1099511627776.0));
            else atomicAdd(reinterpret_cast<double*>(dst), amp);
          }
        }
      }
      if (ctid == 0) n_events += (unsigned long long)(r1 - r0);
    }
    /* older spikes: one warp each; bucket index -> due segment of the delay-sorted row */
    for (unsigned sp = lo + cwid; (int)(lo_new - sp) > 0; sp += cwarps) {
      const int j = Lneuron[sp & cap_mask];
      const int age = (int)(k - (long long)Lstep[sp & cap_mask]);
      if (age > P.max_delay) continue;
      const long long r0 = P.row_ptr[j];
      const int b = age / P.bkt_width;
      const long long s0 = r0 + P.row_bkt[(size_t)j * DELAY_BUCKETS + b];
      const long long s1 = b + 1 < DELAY_BUCKETS ? r0 + P.row_bkt[(size_t)j * DELAY_BUCKETS + b + 1] : P.row_ptr[j + 1];
      for (long long rb = s0; rb < s1; rb += 32) {
        const long long r = rb + lane;
        int d = 0x7fffffff;
        SynHot sh;
        if (r < s1) { sh = P.hot[r]; d = sh.delay; }
        if (d == age) {
          ++n_deliv;
          if (sh.amp != 0.0) {
            const int tl = sh.tgt - inst * P.inst_size;
            unsigned long long* dst = cluster.map_shared_rank(pend_sh, tl / RES_BLOCK) + sh.chan * RES_BLOCK + (tl % RES_BLOCK);
            if (P.accum == HHD_ACCUM_FIXED) atomicAdd(dst, (unsigned long long)__double2ll_rn(sh.amp * 1099511627776.0));
            else atomicAdd(reinterpret_cast<double*>(dst), sh.amp);
          }
        }
        if (__all_sync(0xffffffffu, d > age)) break;
      }
    }
    /* SpikeGeneratorGroup events of this step that target this instance */
    if (P.ext_ptr && k >= P.ext_first && k <= P.ext_last) {
      const int e0 = P.ext_ptr[k - P.ext_first], e1 = P.ext_ptr[k - P.ext_first + 1];
      for (int e = e0 + ctid; e < e1; e += cthreads) {
        const int tl = P.ext_tgt[e] - inst * P.inst_size;
        if (tl < 0 || tl >= P.inst_size) continue;
        const double failure = hhd_philox_u01(P.seed, (unsigned long long)P.ext_id[e], (unsigned long long)k, 1) < P.ext_pfail[e] ? 1.0 : 0.0;
        const unsigned char mask = P.ext_mask[e];
        const double g = P.ext_gmax[e] * (1.0 - failure);
        ++n_ext;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (mask & (1 << c)) {
            const double inc = g * P.cc.gsyn[c];
            unsigned long long* dst = cluster.map_shared_rank(pend_sh, tl / RES_BLOCK) + c * RES_BLOCK + (tl % RES_BLOCK);
            if (P.accum == HHD_ACCUM_FIXED) atomicAdd(dst, (unsigned long long)__double2ll_rn(inc * 1099511627776.0));
            else atomicAdd(reinterpret_cast<double*>(dst), inc);
          }
      }
    }
    if (rank == 0 && tid == 0) end_sh[kw] = hi;
    cluster.sync();
  }

  /* ---------------- write back ---------------------------------------------------------------------------------------- */
  if (live) {
#pragma unroll
    for (int v = 0; v < 8; ++v) P.nb[nb_index(v, i)] = x[v];
    P.meta[i] = meta;
    if (METHOD == 3) P.hstep[i] = hcar;
  }
  if (rank == 0) {
    if (tid == 0) store[0] = list_sh->head;
    for (int q = tid; q < P.W; q += RES_BLOCK) store[2 + q] = end_sh[q];
    for (int q = tid; q < P.res_cap; q += RES_BLOCK) { store[2 + P.W + q] = (unsigned)lneuron_sh[q]; store[2 + P.W + P.res_cap + q] = (unsigned)lstep_sh[q]; }
  }
  n_events = (unsigned long long)warp_sum((double)n_events);
  n_deliv = (unsigned long long)warp_sum((double)n_deliv);
  n_ext = (unsigned long long)warp_sum((double)n_ext);
  n_spk = __reduce_add_sync(0xffffffffu, n_spk);
  n_wc = __reduce_add_sync(0xffffffffu, n_wc);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if (lane == 0) {
    if (n_events) atomicAdd(&P.counters[CNT_SYN_EVENTS], n_events);
    if (n_deliv) atomicAdd(&P.del_slots[(blockIdx.x * 4 + (tid >> 5)) & (CNT_SLOTS - 1)], n_deliv);
    if (n_ext) atomicAdd(&P.counters[CNT_EXT], n_ext);
    if (n_spk) atomicAdd(&P.counters[CNT_SPIKES], (unsigned long long)n_spk);
    if (n_wc) atomicAdd(&P.counters[CNT_WCROSS], (unsigned long long)n_wc);
    if (n_bad) atomicAdd(&P.counters[CNT_NONFINITE], (unsigned long long)n_bad);
  }
}

/* After the all-gather (NCCL fallback of the exchange): append the OTHER ranks' spikes of this step to the ring — k_neuron
 * appended this rank's own — and clear the send counter for the next step.  One CTA. */
__global__ void __launch_bounds__(256) k_append(const __grid_constant__ Params P, const int step_offset) {
  const long long k = *P.base_step + step_offset;
  __shared__ unsigned long long base_sh[17];
  if (threadIdx.x == 0) {
    unsigned long long b = *P.head;
    for (int r = 0; r < P.n_ranks; ++r) {
      base_sh[r] = b;
      if (r != P.rank) b += (unsigned long long)min(P.xchg_recv[(size_t)r * (1 + P.xchg_cap)], P.xchg_cap);
    }
    base_sh[P.n_ranks] = b;
  }
  __syncthreads();
  for (int r = 0; r < P.n_ranks; ++r) {
    if (r == P.rank) continue;
    const int* src = P.xchg_recv + (size_t)r * (1 + P.xchg_cap);
    const int n = min(src[0], P.xchg_cap);
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const unsigned pos = (unsigned)(base_sh[r] + q) & P.cap_mask;
      P.spk_neuron[pos] = src[1 + q];
      P.spk_step[pos] = (int)k;
      note_spike(P, src[1 + q], (int)k, (double)k * P.dt_s);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *P.head = base_sh[P.n_ranks];
    P.xchg_send[0] = 0;
  }
}

__global__ void k_advance(long long* base_step, int n) { *base_step += n; }

__global__ void k_read_var(const __grid_constant__ Params P, int var, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.N) return;
  double x[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) x[v] = P.nb[nb_index(v, i)];
  const NeuronPar p = load_par(P, i);
  out[i] = hhd_var_value(P.cc, p, x, var, P.last_spike[P.off + i], P.iac_table ? iac_at(P, P.iac_col[i], *P.base_step) : 0.0);
}

/* one field (0..7 state, 8.. parameters) of the neuron blocks from / to a plain [N] array */
__global__ void k_scatter_field(const __grid_constant__ Params P, int field, const double* in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.N) P.nb[nb_index(field, i)] = in[i];
}

/* ================================================== host side =================================================== */
struct Monitor {
  int var, reduce, active;
  long long n_idx, width, capacity, n_samples, first_step;
  int* d_slot;
  double* d_buf;
};

struct ExtSet {
  int n_sources;
  std::vector<int> spk_src;
  std::vector<long long> spk_step;
  std::vector<int> syn_src, syn_tgt;
  std::vector<unsigned char> syn_mask;
  std::vector<double> syn_gmax, syn_pfail;
  long long id_base;
};

struct hhd_engine {
  hhd_config cfg;
  std::string err;
  int N = 0, NG = 0;
  long long step = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Params P;
  bool neurons_set = false, channels_set = false, started = false, graph_dirty = true, ext_dirty = false;
  long long nsyn = 0;
  std::vector<unsigned char> has_out_host;
  bool has_out_given = false;
  std::vector<void*> allocs;
  std::vector<Monitor> mons;
  std::vector<ExtSet> exts;
  long long ext_syn_total = 0;
  void *d_ext_ptr = nullptr, *d_ext_tgt = nullptr, *d_ext_mask = nullptr, *d_ext_gmax = nullptr, *d_ext_pfail = nullptr, *d_ext_id = nullptr;
  MonDesc* d_mon = nullptr;
  double* d_stage = nullptr;        /* [N] staging buffer of scatter_field */
  long long* d_base_step = nullptr;
  unsigned long long* d_counters = nullptr;
  unsigned long long ring_cap = 0;
  long long seg_steps = 0;
  unsigned long long drained = 0;
  bool record_spikes = true;
  std::vector<int> rec_neuron;
  std::vector<long long> rec_step;
  size_t rec_sorted = 0;            /* rec_*[0, rec_sorted) are in (step, neuron) order */
  cudaStream_t stream_copy = nullptr;
  cudaEvent_t ev_seg[2] = {nullptr, nullptr};
  unsigned long long* pin_head = nullptr;   /* [2] pinned */
  int *pin_nb = nullptr, *pin_sb = nullptr; /* pinned staging of one segment's ring entries */
  size_t pin_cap = 0;
  unsigned long long drained_prev = 0;
  cudaGraphExec_t graph_exec = nullptr;
  hhd_counters cnt;
  int stp_grid = 0, deliver_grid = 0, sm_count = 0, neuron_grid = 0;
  bool mon_variant = false;         /* some state monitor is active: the k_neuron variant that samples them is launched */
  bool p2p = false;                 /* spike exchange over mapped peer memory (send_spike / remote_spike_phase) instead of NCCL */
  void* peer_mapped[MAX_RANKS] = {};/* cudaIpcOpenMemHandle results, closed in hhd_destroy */
#ifdef HHD_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  bool comm_ready = false;
  bool cross_instance = false; /* some synapse joins two instances: the resident plan is not applicable */
  bool resident = false;       /* HHD_PLAN_RESIDENT chosen */
  int res_smem = 0;
  bool pdl = false;            /* programmatic dependent launch between the kernels of the main stream */
  bool fused = false;          /* max_delay < refractory steps: no neuron has a new spike and an undelivered older one */
  int kernels_per_step = 3;
};

static thread_local std::string g_create_err;

static int fail(hhd_handle h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_err = buf;
  return code;
}

#define CU(call)                                                                                                \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess) return fail(h, HHD_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename T>
static int dev_alloc(hhd_handle h, T** p, size_t n, bool zero = true) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) return fail(h, HHD_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
  if (zero) cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream);
  h->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}
static void dev_free(hhd_handle h, void* p) {
  if (!p) return;
  auto it = std::find(h->allocs.begin(), h->allocs.end(), p);
  if (it != h->allocs.end()) h->allocs.erase(it);
  cudaFree(p);
}
template <typename T>
static int upload(hhd_handle h, T* dst, const T* src, size_t n) {
  if (!n) return 0;
  CU(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));   /* caller owns src: it may be freed right after the call */
  return 0;
}

extern "C" int hhd_abi_version(void) { return HHD_ABI_VERSION; }
extern "C" const char* hhd_last_error(hhd_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int hhd_create(const hhd_config* cfg, hhd_handle* out) {
  hhd_handle h = nullptr;
  if (!cfg || !out) return fail(h, HHD_E_ARG, "null argument");
  if (cfg->abi_version != HHD_ABI_VERSION) return fail(h, HHD_E_ARG, "abi_version %d != %d", cfg->abi_version, HHD_ABI_VERSION);
  if (cfg->n_neurons <= 0 || !(cfg->dt_ms > 0)) return fail(h, HHD_E_ARG, "n_neurons and dt_ms must be positive");
  if (cfg->method < HHD_EULER || cfg->method > HHD_RK23_ADAPTIVE) return fail(h, HHD_E_ARG, "unknown method %d", cfg->method);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(h, HHD_E_CUDA, "no CUDA device: %s (this engine has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(h, HHD_E_ARG, "device %d out of range (%d visible)", cfg->device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(h, HHD_E_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(h, HHD_E_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", cfg->device, prop.major, prop.minor);
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(h, HHD_E_CUDA, "cudaSetDevice failed");
  h = new hhd_engine();
  h->cfg = *cfg;
  if (h->cfg.n_ranks <= 0) h->cfg.n_ranks = 1;
  if (h->cfg.n_global <= 0) h->cfg.n_global = cfg->n_neurons;
  h->N = cfg->n_neurons;
  h->NG = h->cfg.n_global;
  if (h->cfg.global_offset < 0 || h->cfg.global_offset + h->N > h->NG) { delete h; return fail(nullptr, HHD_E_ARG, "partition out of range"); }
  h->sm_count = prop.multiProcessorCount;
  {
    /* L2 set-aside for lines loaded / stored with the evict_last priority (the state rows, k_neuron): the size of the
     * state rows, as far as the device allows (measured on 1 003 000 neurons: 0 MB 65.2, 64 MB 63.2, the maximum of
     * 79 MB 64.0 us per step), or HHD_L2_PERSIST_MB */
    const size_t n_pad0 = (size_t)((cfg->n_neurons + 31) / 32) * 32;
    size_t want = std::min((size_t)prop.persistingL2CacheMaxSize, n_pad0 * (size_t)(8 * 8));
    if (const char* env = getenv("HHD_L2_PERSIST_MB")) want = std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)atoll(env) << 20);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    if (getenv("HHD_VERBOSE")) fprintf(stderr, "[hhd] L2 %d MB, persisting set-aside %zu MB (max %d MB)\n", prop.l2CacheSize >> 20, want >> 20, prop.persistingL2CacheMaxSize >> 20);
  }
  memset(&h->cnt, 0, sizeof h->cnt);
  memset(&h->P, 0, sizeof h->P);
  h->record_spikes = cfg->record_spikes != 0;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(nullptr, HHD_E_CUDA, "cudaStreamCreate failed"); }
  cudaStreamCreateWithFlags(&h->stream_copy, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&h->ev_seg[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_seg[1], cudaEventDisableTiming);
  cudaMallocHost(&h->pin_head, 16);
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  Params& P = h->P;
  P.N = h->N; P.NG = h->NG; P.off = h->cfg.global_offset; P.dt = cfg->dt_ms; P.dt_s = hhd_dt_s(cfg->dt_ms); P.accum = cfg->accum; P.seed = cfg->seed;
  P.refr_steps = (int)hhd_timestep(cfg->refractory_ms, cfg->dt_ms);
  P.cc.window_s = cfg->block_window_ms * 1e-3;
  int rc = 0;
  const size_t n_pad = (size_t)((h->N + NB_LANES - 1) / NB_LANES) * NB_LANES;   /* whole 32-neuron blocks for the bulk copies */
  rc = dev_alloc(h, &P.nb, n_pad * NB_FIELDS);
  if (!rc) { double* r = nullptr; rc = dev_alloc(h, &r, n_pad * N_RARE); P.rare = r; }
  if (!rc) rc = dev_alloc(h, &P.last_spike, h->NG, false);
  if (!rc && cfg->method == HHD_RK23_ADAPTIVE) rc = dev_alloc(h, &P.hstep, n_pad, false);
  if (!rc) rc = dev_alloc(h, &P.meta, n_pad, false);
  unsigned char* ho = nullptr;
  if (!rc) rc = dev_alloc(h, &ho, h->NG);
  P.has_out = ho;
  P.n_pad = (long long)n_pad;
  P.pend = nullptr;            /* the conductance buffer is sized in finalize_setup, when the largest delay is known */
  if (!rc) rc = dev_alloc(h, &P.prev_spike, h->NG);
  if (!rc) rc = dev_alloc(h, &P.pass_done, 1);
  long long* rp = nullptr;
  if (!rc) rc = dev_alloc(h, &rp, (size_t)h->NG + 1);
  P.row_ptr = rp;
  if (!rc) rc = dev_alloc(h, &P.head, 1);
  if (!rc) rc = dev_alloc(h, &h->d_base_step, 1);
  if (!rc) rc = dev_alloc(h, &h->d_counters, CNT_N);
  if (!rc) rc = dev_alloc(h, &P.del_slots, CNT_SLOTS);
  if (!rc) rc = dev_alloc(h, &h->d_mon, MAX_MON);
  if (rc) { std::string m = h->err; hhd_destroy(h); g_create_err = m; return rc; }
  P.base_step = h->d_base_step; P.counters = h->d_counters; P.mon = h->d_mon;
  {
    std::vector<double> ls((size_t)h->NG, cfg->last_spike0_ms * 1e-3);     /* seconds */
    std::vector<int> lk(n_pad, -(1 << 30));          /* lastspike = -2^26 steps, no flags */
    cudaMemcpyAsync(P.last_spike, ls.data(), ls.size() * 8, cudaMemcpyHostToDevice, h->stream);
    if (P.hstep) {
      std::vector<double> h0(n_pad, cfg->dt_ms);
      cudaMemcpy(P.hstep, h0.data(), h0.size() * 8, cudaMemcpyHostToDevice);
    }
    cudaMemcpyAsync(P.meta, lk.data(), lk.size() * 4, cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
  }
  h->has_out_host.assign((size_t)h->NG, 0);
  *out = h;
  return HHD_OK;
}

extern "C" void hhd_destroy(hhd_handle h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
#ifdef HHD_WITH_NCCL
  if (h->comm) ncclCommDestroy(h->comm);
#endif
  for (int r = 0; r < MAX_RANKS; ++r) if (h->peer_mapped[r]) cudaIpcCloseMemHandle(h->peer_mapped[r]);
  for (void* p : h->allocs) cudaFree(p);
  if (h->stream_copy) { cudaStreamSynchronize(h->stream_copy); cudaStreamDestroy(h->stream_copy); }
  for (int q = 0; q < 2; ++q) if (h->ev_seg[q]) cudaEventDestroy(h->ev_seg[q]);
  if (h->pin_head) cudaFreeHost(h->pin_head);
  if (h->pin_nb) cudaFreeHost(h->pin_nb);
  if (h->pin_sb) cudaFreeHost(h->pin_sb);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

/* host array [N] -> one field of the neuron blocks (staged through a device buffer kept for the next call) */
static int scatter_field(hhd_handle h, int field, const double* in) {
  if (!h->d_stage) { int rc = dev_alloc(h, &h->d_stage, (size_t)h->N, false); if (rc) return rc; }
  CU(cudaMemcpyAsync(h->d_stage, in, (size_t)h->N * 8, cudaMemcpyHostToDevice, h->stream));
  k_scatter_field<<<(h->N + 255) / 256, 256, 0, h->stream>>>(h->P, field, h->d_stage);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));   /* caller owns `in` */
  return 0;
}

extern "C" int hhd_set_neurons(hhd_handle h, const double* const* params, const double* V0, const double* w0) {
  if (!h || !params || !V0 || !w0) return fail(h, HHD_E_ARG, "null argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_neurons after run");
  cudaSetDevice(h->cfg.device);
  for (int p = 0; p < HHD_N_PARAMS; ++p) if (!params[p]) return fail(h, HHD_E_ARG, "params[%d] is null", p);
  for (int i = 0; i < h->N; ++i)
    if (!(params[HHD_P_C][i] > 0) || !(params[HHD_P_GL][i] > 0) || !(params[HHD_P_DELTAT][i] > 0) || !(params[HHD_P_TAUW][i] > 0))
      return fail(h, HHD_E_ARG, "neuron %d: C, g_L, delta_T and tau_w must be positive", i);
  /* blocked layout (nb_index); the padding lanes of the last block get parameters with a stable fixed point at V = w = 0
   * (no exponential current, leak towards 0, a threshold that is never reached) so that they compute harmlessly */
  const size_t n_pad = (size_t)h->P.n_pad;
  std::vector<double> blk(n_pad * NB_FIELDS, 0.0), rare(n_pad * N_RARE, 0.0);
  const double pad_par[HHD_N_PARAMS] = {1.0, 1.0, 0.0, 1.0, 1e300, 1.0, 0.0, 0.0, 1000.0, 1e300, 0.0, 0.0};   /* C gL EL dT Vup tauw b Vr VT Iref Vdep IDC */
  for (size_t i = 0; i < n_pad; ++i) {
    const bool live = i < (size_t)h->N;
    for (int p = 0; p < HHD_N_PARAMS; ++p) {
      const int row = param_row(p);
      const double v = live ? params[p][i] : pad_par[p];
      if (row >= 0) blk[nb_index(NB_STATE + row, (long long)i)] = v;
      else rare[(size_t)(-1 - row) * n_pad + i] = v;
    }
    blk[nb_index(0, (long long)i)] = live ? V0[i] : 0.0;
    blk[nb_index(1, (long long)i)] = live ? w0[i] : 0.0;
  }
  int rc = upload(h, h->P.nb, blk.data(), blk.size());
  if (!rc) rc = upload(h, const_cast<double*>(h->P.rare), rare.data(), rare.size());
  if (rc) return rc;
  h->neurons_set = true;
  return HHD_OK;
}

extern "C" int hhd_set_idc(hhd_handle h, const double* idc) {
  if (!h || !idc) return fail(h, HHD_E_ARG, "null argument");
  if (!h->neurons_set) return fail(h, HHD_E_STATE, "hhd_set_neurons has not been called");
  cudaSetDevice(h->cfg.device);
  return scatter_field(h, NB_STATE + ROW_IDC, idc);
}

extern "C" int hhd_set_timed_current(hhd_handle h, int64_t n_steps, int32_t n_cols, const double* table, const int32_t* col_of_neuron) {
  if (!h || n_steps <= 0 || n_cols <= 0 || !table || !col_of_neuron) return fail(h, HHD_E_ARG, "bad argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_timed_current after run");
  for (int i = 0; i < h->N; ++i) if (col_of_neuron[i] >= n_cols) return fail(h, HHD_E_ARG, "column index out of range");
  cudaSetDevice(h->cfg.device);
  double* dt = nullptr;
  int* dc = nullptr;
  int rc = dev_alloc(h, &dt, (size_t)n_steps * n_cols, false);
  if (!rc) rc = dev_alloc(h, &dc, (size_t)h->N, false);
  if (!rc) rc = upload(h, dt, table, (size_t)n_steps * n_cols);
  if (!rc) rc = upload(h, dc, col_of_neuron, (size_t)h->N);
  if (rc) return rc;
  h->P.iac_table = dt; h->P.iac_col = dc; h->P.iac_steps = n_steps; h->P.iac_cols = n_cols;
  h->graph_dirty = true;
  return HHD_OK;
}

extern "C" int hhd_set_noise(hhd_handle h, const double* sigma) {
  if (!h || !sigma) return fail(h, HHD_E_ARG, "null argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_noise after run");
  if (h->cfg.method != HHD_EULER) return fail(h, HHD_E_ARG, "conductance noise needs the euler method (stochastic equations)");
  cudaSetDevice(h->cfg.device);
  double* ds = nullptr;
  int rc = dev_alloc(h, &ds, (size_t)h->N, false);
  if (!rc) rc = upload(h, ds, sigma, (size_t)h->N);
  if (rc) return rc;
  h->P.noise_sigma = ds;
  h->P.sqrt_dt = sqrt(h->cfg.dt_ms);
  h->graph_dirty = true;
  return HHD_OK;
}

extern "C" int hhd_set_has_outgoing(hhd_handle h, const uint8_t* has_out_global) {
  if (!h || !has_out_global) return fail(h, HHD_E_ARG, "null argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_has_outgoing after run");
  cudaSetDevice(h->cfg.device);
  h->has_out_host.assign(has_out_global, has_out_global + h->NG);
  h->has_out_given = true;
  return upload(h, const_cast<unsigned char*>(h->P.has_out), h->has_out_host.data(), (size_t)h->NG);
}

extern "C" int hhd_set_channels(hhd_handle h, const double tau_on[3], const double tau_off[3], const double E_rev[3],
                                const double gsyn[3], double mg_fac, double mg_slope, double mg_half) {
  if (!h || !tau_on || !tau_off || !E_rev || !gsyn) return fail(h, HHD_E_ARG, "null argument");
  for (int c = 0; c < 3; ++c) {
    if (!(tau_on[c] > 0) || !(tau_off[c] > 0)) return fail(h, HHD_E_ARG, "channel %d: time constants must be positive", c);
    h->P.cc.inv_tau_on[c] = 1.0 / tau_on[c];
    h->P.cc.inv_tau_off[c] = 1.0 / tau_off[c];
    h->P.cc.E[c] = E_rev[c];
    h->P.cc.gsyn[c] = gsyn[c];
  }
  h->P.cc.mg_fac = mg_fac; h->P.cc.mg_slope = mg_slope; h->P.cc.mg_half = mg_half;
  h->channels_set = true;
  h->graph_dirty = true;
  return HHD_OK;
}

/* ---- synapse setup on the device: validate, sort by (source, delay) with a stable radix sort, gather into CSR ----- */
enum { SYNERR_SRC = 1, SYNERR_TGT = 2, SYNERR_CHAN = 4, SYNERR_DELAY = 8, SYNERR_TAU = 16, SYNERR_DELAY_BIG = 32, SYNWARN_CROSS = 64 };

__global__ void k_syn_keys(long long nsyn, const int* src, const int* tgt, const unsigned char* chan, const double* tau_rec,
                           const double* tau_fac, const double* delay_ms, double dt, int NG, int lo, int hi, int inst_size,
                           unsigned long long* keys, unsigned int* vals, unsigned int* err, int* maxd) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsyn) return;
  unsigned int e = 0;
  const int s = src[k];
  if (s < 0 || s >= NG) e |= SYNERR_SRC;
  if (tgt[k] < lo || tgt[k] >= hi) e |= SYNERR_TGT;
  if (chan[k] > 2) e |= SYNERR_CHAN;
  if (!(delay_ms[k] >= 0)) e |= SYNERR_DELAY;
  if (!(tau_rec[k] > 0) || !(tau_fac[k] > 0)) e |= SYNERR_TAU;
  long long d = e & SYNERR_DELAY ? 0 : hhd_delay_steps(delay_ms[k], dt);
  if (d > 65535) { e |= SYNERR_DELAY_BIG; d = 65535; }
  if (!e && inst_size > 0 && (s - lo) / inst_size != (tgt[k] - lo) / inst_size) e |= SYNWARN_CROSS;
  if (e) atomicOr(err, e);
  atomicMax(maxd, (int)d);
  keys[k] = ((unsigned long long)(unsigned int)(e & SYNERR_SRC ? 0 : s) << 16) | (unsigned long long)d;
  vals[k] = (unsigned int)k;
}

__global__ void k_syn_gather(long long nsyn, const unsigned long long* keys, const unsigned int* perm, const int* tgt,
                             const unsigned char* chan, const double* gmax, const double* U, const double* tau_rec,
                             const double* tau_fac, const double* pfail, int lo, SynHot* hot, SynStp* stp) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nsyn) return;
  const unsigned int o = perm[r];
  SynHot h;
  h.tgt = tgt[o] - lo;
  h.delay = (unsigned short)(keys[r] & 0xffffull);
  h.chan = chan[o];
  h.pad = 0;
  h.amp = 0.0;
  hot[r] = h;
  SynStp y;
  y.U = U[o]; y.tau_rec = tau_rec[o] * 1e-3; y.tau_fac = tau_fac[o] * 1e-3;   /* seconds (hhd_time_s) */ y.gmax = gmax[o]; y.pfail = pfail[o];
  y.u = y.U;       /* u = U   (codes/CortexNetwork.py:422) */
  y.R = 1.0;       /* R = 1   (:424) */
  y.orig = o; y.pad = 0;
  stp[r] = y;
}

/* one field of the synapse records, scattered back to the caller's synapse order (hhd_read_syn_state) */
__global__ void k_syn_field(long long nsyn, const SynHot* hot, const SynStp* stp, int which, double* out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nsyn) return;
  const unsigned int o = stp[r].orig;
  out[o] = which == HHD_SYN_U ? stp[r].u : which == HHD_SYN_R ? stp[r].R : which == HHD_SYN_A ? hot[r].amp : (double)hot[r].delay;
}

/* row_ptr[j] = first sorted position whose source is >= j */
__global__ void k_syn_rows(long long nsyn, const unsigned long long* keys, int NG, long long* row_ptr, unsigned char* has_out, int set_has_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > NG) return;
  long long lo = 0, hi = nsyn;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)(keys[mid] >> 16) < (long long)j) lo = mid + 1; else hi = mid;
  }
  row_ptr[j] = lo;
  if (set_has_out && j < NG) {
    long long lo2 = lo, hi2 = nsyn;
    const bool any = lo2 < hi2 && (long long)(keys[lo2] >> 16) == (long long)j;
    has_out[j] = any ? 1 : 0;
  }
}

/* row_bkt[j][b] = offset inside row j of the first synapse whose delay is >= b * width (rows are sorted by delay) */
__global__ void k_syn_buckets(int NG, const long long* row_ptr, const SynHot* hot, int width, unsigned int* row_bkt, unsigned short* row_maxd) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)NG * DELAY_BUCKETS) return;
  const int j = (int)(q / DELAY_BUCKETS), b = (int)(q % DELAY_BUCKETS);
  const long long r0 = row_ptr[j], r1 = row_ptr[j + 1];
  const int want = b * width;
  long long lo = r0, hi = r1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((int)hot[mid].delay < want) lo = mid + 1; else hi = mid;
  }
  row_bkt[q] = (unsigned int)(lo - r0);
  if (b == 0) row_maxd[j] = r1 > r0 ? hot[r1 - 1].delay : 0;      /* rows are sorted by delay */
}

/* has_out bit 1: the row starts with zero-delay synapses (rows are sorted by delay); both bits are copied into the
 * local neurons' meta words */
__global__ void k_mark_flags(int NG, int N, int off, const long long* row_ptr, const SynHot* hot, unsigned char* has_out, int* meta) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= NG) return;
  const long long r0 = row_ptr[j], r1 = row_ptr[j + 1];
  const unsigned char z = (r1 > r0 && hot[r0].delay == 0) ? 2 : 0;
  const unsigned char ho = (unsigned char)((has_out[j] & 1) | z);
  has_out[j] = ho;
  if (j >= off && j < off + N) meta[j - off] = (meta[j - off] & ~(FLAG_OUT | FLAG_ZERO)) | ((ho & 1) ? FLAG_OUT : 0) | ((ho & 2) ? FLAG_ZERO : 0);
}

static int set_synapses_device(hhd_handle h, int64_t nsyn, const int32_t* src, const int32_t* tgt, const uint8_t* chan,
                               const double* g_max, const double* U, const double* tau_rec, const double* tau_fac,
                               const double* p_fail, const double* delay_ms, int64_t syn_id_base) {
  Params& P = h->P;
  const int lo = h->cfg.global_offset, hi = lo + h->N;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  unsigned int *vals = nullptr, *vals2 = nullptr, *derr = nullptr;
  int* dmaxd = nullptr;
  int rc = dev_alloc(h, &keys, (size_t)nsyn, false);
  if (!rc) rc = dev_alloc(h, &keys2, (size_t)nsyn, false);
  if (!rc) rc = dev_alloc(h, &vals, (size_t)nsyn, false);
  if (!rc) rc = dev_alloc(h, &vals2, (size_t)nsyn, false);
  if (!rc) rc = dev_alloc(h, &derr, 1, true);
  if (!rc) rc = dev_alloc(h, &dmaxd, 1, true);
  if (rc) return rc;
  const int T = 256;
  const unsigned grid = (unsigned)((nsyn + T - 1) / T);
  if (nsyn) k_syn_keys<<<grid, T, 0, h->stream>>>(nsyn, src, tgt, chan, tau_rec, tau_fac, delay_ms, h->cfg.dt_ms, h->NG, lo, hi, h->cfg.instance_size, keys, vals, derr, dmaxd);
  unsigned int err = 0;
  int maxd = 0;
  CU(cudaMemcpyAsync(&err, derr, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(&maxd, dmaxd, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (err & SYNWARN_CROSS) { h->cross_instance = true; err &= ~SYNWARN_CROSS; }
  if (err) {
    dev_free(h, keys); dev_free(h, keys2); dev_free(h, vals); dev_free(h, vals2); dev_free(h, derr); dev_free(h, dmaxd);
    return fail(h, HHD_E_ARG, "invalid synapse arrays:%s%s%s%s%s%s", err & SYNERR_SRC ? " source out of range;" : "",
                err & SYNERR_TGT ? " target is not local;" : "", err & SYNERR_CHAN ? " channel > 2;" : "",
                err & SYNERR_DELAY ? " negative delay;" : "", err & SYNERR_TAU ? " tau_rec/tau_fac must be positive;" : "",
                err & SYNERR_DELAY_BIG ? " delay exceeds 65535 steps;" : "");
  }
  int bits = 16;
  while ((1ll << (bits - 16)) < (long long)h->NG) ++bits;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (long long)nsyn, 0, bits, h->stream);
  unsigned char* tmp = nullptr;
  rc = dev_alloc(h, &tmp, tmp_bytes, false);
  if (rc) return rc;
  if (nsyn) {
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (long long)nsyn, 0, bits, h->stream);
    if (e != cudaSuccess) return fail(h, HHD_E_CUDA, "radix sort failed: %s", cudaGetErrorString(e));
  }
  rc = dev_alloc(h, &P.hot, (size_t)nsyn, false);
  if (!rc) rc = dev_alloc(h, &P.stp, (size_t)nsyn, false);
  if (rc) return rc;
  if (nsyn) k_syn_gather<<<grid, T, 0, h->stream>>>(nsyn, keys2, vals2, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, lo, P.hot, P.stp);
  k_syn_rows<<<(h->NG + 1 + T - 1) / T, T, 0, h->stream>>>(nsyn, keys2, h->NG, const_cast<long long*>(P.row_ptr),
                                                            const_cast<unsigned char*>(P.has_out), h->has_out_given ? 0 : 1);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  {
    unsigned int* bkt = nullptr;
    unsigned short* rmax = nullptr;
    rc = dev_alloc(h, &bkt, (size_t)h->NG * DELAY_BUCKETS, true);
    if (!rc) rc = dev_alloc(h, &rmax, (size_t)h->NG, true);
    if (rc) return rc;
    P.row_maxd = rmax;
    P.bkt_width = (maxd + 1 + DELAY_BUCKETS - 1) / DELAY_BUCKETS;
    const long long nq = (long long)h->NG * DELAY_BUCKETS;
    k_syn_buckets<<<(unsigned)((nq + T - 1) / T), T, 0, h->stream>>>(h->NG, P.row_ptr, P.hot, P.bkt_width, bkt, rmax);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    P.row_bkt = bkt;
  }
  {
    /* a spike-time task is one block of a spiking neuron's row: as many tasks per spike as 1.25 mean rows need */
    const double mean_row = (double)nsyn / std::max(1, h->NG);
    P.stp_chunks = std::max(1, (int)std::ceil(1.25 * mean_row / NEURON_BLOCK));
  }
  dev_free(h, keys); dev_free(h, keys2); dev_free(h, vals); dev_free(h, vals2); dev_free(h, derr); dev_free(h, dmaxd); dev_free(h, tmp);
  P.max_delay = maxd;
  P.syn_id_base = syn_id_base;
  h->nsyn = nsyn;
  h->graph_dirty = true;
  return HHD_OK;
}

static int check_set_synapses(hhd_handle h, int64_t nsyn, const void* a, const void* b, const void* c, const void* d, const void* e,
                              const void* f, const void* g, const void* i, const void* j) {
  if (!h || nsyn < 0) return fail(h, HHD_E_ARG, "bad argument");
  if (h->started) return fail(h, HHD_E_STATE, "set_synapses after run");
  if (h->nsyn) return fail(h, HHD_E_STATE, "synapses already set");
  if (nsyn && (!a || !b || !c || !d || !e || !f || !g || !i || !j)) return fail(h, HHD_E_ARG, "null argument");
  if (nsyn >= (1ll << 32)) return fail(h, HHD_E_ARG, "at most 2^32-1 synapses per handle");
  if (h->NG > (1 << 30)) return fail(h, HHD_E_ARG, "too many neurons");
  return 0;
}

extern "C" int hhd_set_synapses_dev(hhd_handle h, int64_t nsyn, const int32_t* src, const int32_t* tgt, const uint8_t* chan,
                                    const double* g_max, const double* U, const double* tau_rec, const double* tau_fac,
                                    const double* p_fail, const double* delay_ms, int64_t syn_id_base) {
  int rc = check_set_synapses(h, nsyn, src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms);
  if (rc) return rc;
  cudaSetDevice(h->cfg.device);
  CU(cudaDeviceSynchronize());   /* the caller's producer kernels (e.g. torch ops on another stream) must be complete */
  return set_synapses_device(h, nsyn, src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms, syn_id_base);
}

template <typename T>
static int stage(hhd_handle h, const T** slot, const T* src, size_t n) {
  T* d = nullptr;
  int rc = dev_alloc(h, &d, n, false);
  if (rc) return rc;
  *slot = d;
  return upload(h, d, src, n);
}

extern "C" int hhd_set_synapses(hhd_handle h, int64_t nsyn, const int32_t* src, const int32_t* tgt, const uint8_t* chan,
                                const double* g_max, const double* U, const double* tau_rec, const double* tau_fac,
                                const double* p_fail, const double* delay_ms, int64_t syn_id_base) {
  int rc = check_set_synapses(h, nsyn, src, tgt, chan, g_max, U, tau_rec, tau_fac, p_fail, delay_ms);
  if (rc) return rc;
  cudaSetDevice(h->cfg.device);
  const int32_t *dsrc = nullptr, *dtgt = nullptr; const uint8_t* dchan = nullptr;
  const double *dg = nullptr, *dU = nullptr, *dtr = nullptr, *dtf = nullptr, *dpf = nullptr, *ddl = nullptr;
  const size_t n = (size_t)nsyn;
  rc = stage(h, &dsrc, src, n);
  if (!rc) rc = stage(h, &dtgt, tgt, n);
  if (!rc) rc = stage(h, &dchan, chan, n);
  if (!rc) rc = stage(h, &dg, g_max, n);
  if (!rc) rc = stage(h, &dU, U, n);
  if (!rc) rc = stage(h, &dtr, tau_rec, n);
  if (!rc) rc = stage(h, &dtf, tau_fac, n);
  if (!rc) rc = stage(h, &dpf, p_fail, n);
  if (!rc) rc = stage(h, &ddl, delay_ms, n);
  if (!rc) rc = set_synapses_device(h, nsyn, dsrc, dtgt, dchan, dg, dU, dtr, dtf, dpf, ddl, syn_id_base);
  dev_free(h, (void*)dsrc); dev_free(h, (void*)dtgt); dev_free(h, (void*)dchan); dev_free(h, (void*)dg); dev_free(h, (void*)dU);
  dev_free(h, (void*)dtr); dev_free(h, (void*)dtf); dev_free(h, (void*)dpf); dev_free(h, (void*)ddl);
  return rc;
}

extern "C" int hhd_add_spike_source(hhd_handle h, int32_t n_sources, int64_t nspk, const int32_t* spk_src, const double* spk_t_ms,
                                    int64_t nsyn, const int32_t* syn_src, const int32_t* syn_tgt, const uint8_t* syn_chanmask,
                                    const double* syn_gmax, const double* syn_pfail) {
  if (!h || n_sources <= 0 || nspk < 0 || nsyn < 0) return fail(h, HHD_E_ARG, "bad argument");
  ExtSet e;
  e.n_sources = n_sources;
  for (int64_t k = 0; k < nspk; ++k) {
    if (spk_src[k] < 0 || spk_src[k] >= n_sources) return fail(h, HHD_E_ARG, "spike %lld: source out of range", (long long)k);
    const long long st = hhd_timestep(spk_t_ms[k], h->cfg.dt_ms);
    if (st < h->step) return fail(h, HHD_E_ARG, "spike %lld lies in the past", (long long)k);
    e.spk_src.push_back(spk_src[k]);
    e.spk_step.push_back(st);
  }
  for (int64_t k = 0; k < nsyn; ++k) {
    if (syn_src[k] < 0 || syn_src[k] >= n_sources) return fail(h, HHD_E_ARG, "input synapse %lld: source out of range", (long long)k);
    if (syn_tgt[k] < 0 || syn_tgt[k] >= h->N) return fail(h, HHD_E_ARG, "input synapse %lld: target out of range", (long long)k);
    if (syn_chanmask[k] == 0 || syn_chanmask[k] > 7) return fail(h, HHD_E_ARG, "input synapse %lld: channel mask", (long long)k);
  }
  e.syn_src.assign(syn_src, syn_src + nsyn);
  e.syn_tgt.assign(syn_tgt, syn_tgt + nsyn);
  e.syn_mask.assign(syn_chanmask, syn_chanmask + nsyn);
  e.syn_gmax.assign(syn_gmax, syn_gmax + nsyn);
  e.syn_pfail.assign(syn_pfail, syn_pfail + nsyn);
  e.id_base = h->ext_syn_total;
  h->ext_syn_total += nsyn;
  h->exts.push_back(std::move(e));
  h->ext_dirty = true;
  return HHD_OK;
}

/* Expand (spike, input synapse) pairs of all source sets, grouped by step, and upload them. */
static int rebuild_ext(hhd_handle h) {
  struct Ev { long long step; int tgt; unsigned char mask; double gmax, pfail; unsigned int id; };
  std::vector<Ev> evs;
  h->exts.erase(std::remove_if(h->exts.begin(), h->exts.end(), [&](const ExtSet& e) {
                  for (long long st : e.spk_step) if (st >= h->step) return false;
                  return true; }), h->exts.end());
  for (const ExtSet& e : h->exts) {
    std::vector<std::vector<int>> by_src((size_t)e.n_sources);
    for (size_t s = 0; s < e.syn_src.size(); ++s) by_src[e.syn_src[s]].push_back((int)s);
    for (size_t q = 0; q < e.spk_src.size(); ++q) {
      if (e.spk_step[q] < h->step) continue;
      for (int s : by_src[e.spk_src[q]])
        evs.push_back({e.spk_step[q], e.syn_tgt[s], e.syn_mask[s], e.syn_gmax[s], e.syn_pfail[s], (unsigned int)(e.id_base + s)});
    }
  }
  std::stable_sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.step < b.step; });
  dev_free(h, h->d_ext_ptr); dev_free(h, h->d_ext_tgt); dev_free(h, h->d_ext_mask);
  dev_free(h, h->d_ext_gmax); dev_free(h, h->d_ext_pfail); dev_free(h, h->d_ext_id);
  h->d_ext_ptr = h->d_ext_tgt = h->d_ext_mask = h->d_ext_gmax = h->d_ext_pfail = h->d_ext_id = nullptr;
  Params& P = h->P;
  P.ext_ptr = nullptr;
  h->ext_dirty = false;
  h->graph_dirty = true;
  if (evs.empty()) return 0;
  if (evs.size() >= (1ull << 31)) return fail(h, HHD_E_ARG, "too many external events");
  const long long first = evs.front().step, last = evs.back().step;
  std::vector<int> ptr((size_t)(last - first + 2), 0);
  for (const Ev& e : evs) ptr[(size_t)(e.step - first + 1)]++;
  for (size_t q = 1; q < ptr.size(); ++q) ptr[q] += ptr[q - 1];
  std::vector<int> tgt(evs.size());
  std::vector<unsigned char> mask(evs.size());
  std::vector<double> gmax(evs.size()), pfail(evs.size());
  std::vector<unsigned int> id(evs.size());
  for (size_t q = 0; q < evs.size(); ++q) { tgt[q] = evs[q].tgt; mask[q] = evs[q].mask; gmax[q] = evs[q].gmax; pfail[q] = evs[q].pfail; id[q] = evs[q].id; }
  int *dp = nullptr, *dt = nullptr; unsigned char* dm = nullptr; double *dg = nullptr, *df = nullptr; unsigned int* di = nullptr;
  int rc = dev_alloc(h, &dp, ptr.size(), false);
  if (!rc) rc = dev_alloc(h, &dt, evs.size(), false);
  if (!rc) rc = dev_alloc(h, &dm, evs.size(), false);
  if (!rc) rc = dev_alloc(h, &dg, evs.size(), false);
  if (!rc) rc = dev_alloc(h, &df, evs.size(), false);
  if (!rc) rc = dev_alloc(h, &di, evs.size(), false);
  if (!rc) rc = upload(h, dp, ptr.data(), ptr.size());
  if (!rc) rc = upload(h, dt, tgt.data(), tgt.size());
  if (!rc) rc = upload(h, dm, mask.data(), mask.size());
  if (!rc) rc = upload(h, dg, gmax.data(), gmax.size());
  if (!rc) rc = upload(h, df, pfail.data(), pfail.size());
  if (!rc) rc = upload(h, di, id.data(), id.size());
  if (rc) return rc;
  h->d_ext_ptr = dp; h->d_ext_tgt = dt; h->d_ext_mask = dm; h->d_ext_gmax = dg; h->d_ext_pfail = df; h->d_ext_id = di;
  P.ext_ptr = dp; P.ext_first = first; P.ext_last = last; P.ext_tgt = dt; P.ext_mask = dm; P.ext_gmax = dg; P.ext_pfail = df; P.ext_id = di;
  return 0;
}

/* ---- monitors ---------------------------------------------------------------------------------------------------- */
extern "C" int hhd_add_state_monitor(hhd_handle h, int32_t var, int64_t n_idx, const int32_t* idx, int32_t reduce,
                                     int64_t capacity_steps, int32_t* mon_id) {
  if (!h || var < 0 || var >= HHD_N_VARS || reduce < 0 || reduce > 2) return fail(h, HHD_E_ARG, "bad argument");
  if ((int)h->mons.size() >= MAX_MON) return fail(h, HHD_E_CAPACITY, "at most %d state monitors", MAX_MON);
  if (reduce != HHD_REDUCE_NONE) {
    int nred = 0;
    for (const Monitor& q : h->mons) nred += q.reduce != HHD_REDUCE_NONE;
    if (nred >= MAX_RED) return fail(h, HHD_E_CAPACITY, "at most %d reduced (sum/mean) monitors", MAX_RED);
  }
  cudaSetDevice(h->cfg.device);
  Monitor m;
  memset(&m, 0, sizeof m);
  m.var = var; m.reduce = reduce; m.active = 1; m.first_step = -1;
  m.n_idx = idx ? n_idx : h->N;
  m.width = reduce == HHD_REDUCE_NONE ? m.n_idx : 1;
  if (idx) {
    std::vector<int> slot((size_t)h->N, -1);
    for (int64_t q = 0; q < n_idx; ++q) {
      if (idx[q] < 0 || idx[q] >= h->N) return fail(h, HHD_E_ARG, "monitor index out of range");
      if (slot[idx[q]] >= 0) return fail(h, HHD_E_ARG, "monitor index %d listed twice", idx[q]);
      slot[idx[q]] = (int)q;
    }
    int rc = dev_alloc(h, &m.d_slot, (size_t)h->N, false);
    if (!rc) rc = upload(h, m.d_slot, slot.data(), slot.size());
    if (rc) return rc;
  }
  m.capacity = std::max<long long>(capacity_steps, 0);
  if (m.capacity) { int rc = dev_alloc(h, &m.d_buf, (size_t)(m.capacity * m.width), true); if (rc) return rc; }
  h->mons.push_back(m);
  h->graph_dirty = true;
  if (mon_id) *mon_id = (int)h->mons.size() - 1;
  return HHD_OK;
}

extern "C" int hhd_set_monitor_active(hhd_handle h, int32_t mon_id, int32_t active) {
  if (!h || mon_id < -1 || mon_id >= (int)h->mons.size()) return fail(h, HHD_E_ARG, "bad monitor id");
  if (mon_id == -1) h->record_spikes = active != 0; else h->mons[mon_id].active = active != 0;
  return HHD_OK;
}

extern "C" int hhd_monitor_samples(hhd_handle h, int32_t mon_id, int64_t* n_samples, int64_t* first_step) {
  if (!h || mon_id < 0 || mon_id >= (int)h->mons.size()) return fail(h, HHD_E_ARG, "bad monitor id");
  if (n_samples) *n_samples = h->mons[mon_id].n_samples;
  if (first_step) *first_step = h->mons[mon_id].first_step;
  return HHD_OK;
}

extern "C" int hhd_read_monitor(hhd_handle h, int32_t mon_id, double* out, int64_t cap_values, int64_t* n_samples) {
  if (!h || mon_id < 0 || mon_id >= (int)h->mons.size()) return fail(h, HHD_E_ARG, "bad monitor id");
  cudaSetDevice(h->cfg.device);
  Monitor& m = h->mons[mon_id];
  const long long nv = m.n_samples * m.width;
  if (nv > cap_values) return fail(h, HHD_E_CAPACITY, "output buffer too small: need %lld values", nv);
  if (nv) {
    CU(cudaMemcpyAsync(out, m.d_buf, (size_t)nv * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (m.reduce == HHD_REDUCE_MEAN && m.n_idx) for (long long q = 0; q < nv; ++q) out[q] /= (double)m.n_idx;
  }
  if (n_samples) *n_samples = m.n_samples;
  return HHD_OK;
}

extern "C" int hhd_read_monitor_dev(hhd_handle h, int32_t mon_id, const double** dev_ptr, int64_t* n_samples) {
  if (!h || mon_id < 0 || mon_id >= (int)h->mons.size() || !dev_ptr) return fail(h, HHD_E_ARG, "bad argument");
  *dev_ptr = h->mons[mon_id].d_buf;
  if (n_samples) *n_samples = h->mons[mon_id].n_samples;
  return HHD_OK;
}

extern "C" int hhd_clear_monitor(hhd_handle h, int32_t mon_id) {
  if (!h || mon_id < -1 || mon_id >= (int)h->mons.size()) return fail(h, HHD_E_ARG, "bad monitor id");
  cudaSetDevice(h->cfg.device);
  if (mon_id == -1) { h->rec_neuron.clear(); h->rec_step.clear(); h->rec_sorted = 0; return HHD_OK; }
  Monitor& m = h->mons[mon_id];
  if (m.d_buf && m.n_samples) CU(cudaMemsetAsync(m.d_buf, 0, (size_t)(m.n_samples * m.width) * 8, h->stream));
  m.n_samples = 0;
  m.first_step = -1;
  return HHD_OK;
}

/* ---- run --------------------------------------------------------------------------------------------------------- */
/* The TimedArray and monitor-sampling variants are separate instantiations: no cost when unused.  The pass that only
 * finishes the last step (STEP = false) samples nothing. */
template <int M, bool STEP>
static const void* neuron_fn(bool timed, bool mon) {
  if (timed) return (mon && STEP) ? (const void*)k_neuron<M, STEP, true, STEP> : (const void*)k_neuron<M, STEP, true, false>;
  return (mon && STEP) ? (const void*)k_neuron<M, STEP, false, STEP> : (const void*)k_neuron<M, STEP, false, false>;
}
template <bool STEP>
static const void* neuron_fn_of(int method, bool timed, bool mon) {
  switch (method) {
    case HHD_EULER: return neuron_fn<0, STEP>(timed, mon);
    case HHD_RK2: return neuron_fn<1, STEP>(timed, mon);
    case HHD_RK4: return neuron_fn<2, STEP>(timed, mon);
    default: return neuron_fn<3, STEP>(timed, mon);
  }
}

static unsigned long long next_pow2(unsigned long long v) { unsigned long long p = 1; while (p < v) p <<= 1; return p; }

#ifdef HHD_WITH_NCCL
/* Map every rank's receive buffer into this process (CUDA IPC; the handles travel through the NCCL communicator) so that
 * push_spikes can write spike lists and flags straight into peer memory.  Collective: every rank calls it from its first
 * hhd_run.  Any failure on any rank (no peer access, IPC not permitted in this container, HHD_NO_P2P set) leaves all ranks
 * on the NCCL all-gather. */
static int setup_p2p(hhd_handle h) {
  Params& P = h->P;
  const int R = P.n_ranks;
  h->p2p = false;
  if (!h->comm_ready || R > MAX_RANKS) return 0;
  const size_t n_words = (size_t)2 * R * (1 + (size_t)P.xchg_cap);
  unsigned long long* buf = nullptr;
  int rc = dev_alloc(h, &buf, n_words, true);          /* zeroed: tag 0 is never a step's tag (step + 1 >= 1) */
  if (rc) return rc;
  int ok = getenv("HHD_NO_P2P") ? 0 : 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (ok && cudaIpcGetMemHandle(&mine, buf) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  unsigned char *d_mine = nullptr, *d_all = nullptr;
  int* d_ok = nullptr;
  rc = dev_alloc(h, &d_mine, 64, true);
  if (!rc) rc = dev_alloc(h, &d_all, (size_t)64 * R, true);
  if (!rc) rc = dev_alloc(h, &d_ok, 1, true);
  if (rc) return rc;
  CU(cudaMemcpyAsync(d_mine, &mine, 64, cudaMemcpyHostToDevice, h->stream));
  if (ncclAllGather(d_mine, d_all, 64, ncclUint8, h->comm, h->stream) != ncclSuccess) return fail(h, HHD_E_COMM, "all-gather of the IPC handles failed");
  std::vector<cudaIpcMemHandle_t> all((size_t)R);
  CU(cudaMemcpyAsync(all.data(), d_all, (size_t)64 * R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < R; ++r) P.peer_buf[r] = nullptr;
  P.peer_buf[P.rank] = buf;
  for (int r = 0; r < R && ok; ++r) {
    if (r == P.rank) continue;
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    h->peer_mapped[r] = ptr;
    P.peer_buf[r] = (unsigned long long*)ptr;
  }
  CU(cudaMemcpyAsync(d_ok, &ok, 4, cudaMemcpyHostToDevice, h->stream));
  if (ncclAllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, h->comm, h->stream) != ncclSuccess) return fail(h, HHD_E_COMM, "all-reduce failed");
  int all_ok = 0;
  CU(cudaMemcpyAsync(&all_ok, d_ok, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  dev_free(h, d_mine); dev_free(h, d_all); dev_free(h, d_ok);
  h->p2p = all_ok != 0;
  if (getenv("HHD_VERBOSE")) fprintf(stderr, "[hhd] rank %d: spike exchange over %s\n", P.rank, h->p2p ? "peer memory (send_spike / remote_spike_phase)" : "NCCL all-gather");
  if (!h->p2p) {
    for (int r = 0; r < R; ++r) {
      if (h->peer_mapped[r]) { cudaIpcCloseMemHandle(h->peer_mapped[r]); h->peer_mapped[r] = nullptr; }
      P.peer_buf[r] = nullptr;
    }
  }
  return 0;
}
#endif

static int finalize_setup(hhd_handle h) {
  Params& P = h->P;
  if (!h->neurons_set) return fail(h, HHD_E_STATE, "hhd_set_neurons has not been called");
  if (!h->channels_set) return fail(h, HHD_E_STATE, "hhd_set_channels has not been called");
  if (h->cfg.n_ranks > 1 && !h->comm_ready) return fail(h, HHD_E_STATE, "n_ranks > 1: call hhd_comm_init on every rank before hhd_run");
  if (h->cfg.n_ranks > 16) return fail(h, HHD_E_ARG, "at most 16 ranks");
  P.n_ranks = h->cfg.n_ranks;
  if (!h->nsyn) {   /* empty network: arrays must still be valid pointers for the kernels */
    P.max_delay = 0;
  }
  P.W = P.max_delay + 2;
  int rc = dev_alloc(h, &P.step_end, (size_t)P.W, true);
  if (rc) return rc;
  P.n_slots = P.max_delay + 1;
  if ((rc = dev_alloc(h, &P.pend, (size_t)P.n_slots * 3 * (size_t)P.n_pad, true)))      /* [n_slots][3][n_pad] x 8 B, zeroed */
    return fail(h, HHD_E_NOMEM, "conductance buffer: %d delay slots x %lld neurons x 24 B do not fit", P.n_slots, (long long)P.n_pad);
  /* Spike ring: a neuron fires at most once per refractory period, so between two host drains (seg_steps apart) plus
   * the delivery window (max_delay) at most NG * per_neuron entries are live. */
  h->seg_steps = 2000;
  const long long refr = std::max(P.refr_steps, 1);
  const long long per_neuron = (2 * h->seg_steps + P.max_delay + 1 + refr - 1) / refr + 2;   /* two segments: see segment_collect */
  h->ring_cap = next_pow2((unsigned long long)h->NG * (unsigned long long)per_neuron);
  if (h->ring_cap > (1ull << 31)) return fail(h, HHD_E_NOMEM, "spike ring too large");
  {
    std::vector<int> never((size_t)h->NG, -(1 << 30));
    rc = dev_alloc(h, &P.spk_last, (size_t)h->NG, false);
    if (!rc) rc = upload(h, P.spk_last, never.data(), never.size());
    if (rc) return rc;
  }
  P.hist_depth = (int)(P.max_delay / refr) + 2;      /* the spike being processed + every older one that can still be in flight */
  if (P.max_delay >= P.refr_steps) {
    if (P.hist_depth > 17) return fail(h, HHD_E_ARG, "delays of more than 16 refractory periods are not supported");
    std::vector<int> never((size_t)h->NG * P.hist_depth, -(1 << 30));
    rc = dev_alloc(h, &P.spk_hist, never.size(), false);
    if (!rc) rc = upload(h, P.spk_hist, never.data(), never.size());
    if (rc) return rc;
  }
  rc = dev_alloc(h, &P.spk_neuron, (size_t)h->ring_cap, false);
  if (!rc) rc = dev_alloc(h, &P.spk_step, (size_t)h->ring_cap, false);
  if (rc) return rc;
  P.cap_mask = (unsigned int)(h->ring_cap - 1);
  if (P.n_ranks > 1) {
    const char* env = getenv("HHD_XCHG_CAP");
    /* must be identical on every rank (all-gather count), so it is derived from global quantities only */
    P.xchg_cap = env && atoi(env) > 0 ? atoi(env) : std::min(h->NG, 4096 + h->NG / (16 * P.n_ranks));
    rc = dev_alloc(h, &P.xchg_send, (size_t)1 + P.xchg_cap, true);
    if (!rc) rc = dev_alloc(h, &P.xchg_recv, (size_t)P.n_ranks * (1 + P.xchg_cap), true);
    if (!rc) rc = dev_alloc(h, &P.xchg_dead, 1, true);
    if (!rc) rc = dev_alloc(h, &P.xchg_done, 4, true);     /* [0] flag, [2 + parity] ring head after the rank's own spikes */
    if (rc) return rc;
    P.rank = h->cfg.rank;
#ifdef HHD_WITH_NCCL
    if ((rc = setup_p2p(h))) return rc;
#endif
  }
  {
    /* persistent state-update grid: as many CTAs as fit at once (shared-memory and register limited), never more than
     * there are warp tiles */
    int occ = 1 << 30;
    for (int v = 0; v < 8; ++v) {           /* every variant of this method may be launched: opt in to the shared memory */
      const bool timed = v & 1, mon = v & 2, step = v & 4;
      const void* fn = step ? neuron_fn_of<true>(h->cfg.method, timed, mon) : neuron_fn_of<false>(h->cfg.method, timed, mon);
      CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, NEURON_SMEM));
      int q = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, fn, NEURON_BLOCK, NEURON_SMEM));
      if (timed == (P.iac_table != nullptr || P.noise_sigma != nullptr)) occ = std::min(occ, q);
    }
    const int ctas = (int)((P.n_pad / NB_LANES + NEURON_WARPS - 1) / NEURON_WARPS);
    const char* env = getenv("HHD_NEURON_CTAS_PER_SM");
    int cps = std::max(1, occ);
    if (env && atoi(env) > 0) cps = std::min(cps, atoi(env));
    h->neuron_grid = std::max(1, std::min(ctas, h->sm_count * cps));
    rc = dev_alloc(h, &P.mon_part, (size_t)MAX_RED * h->neuron_grid, true);
    if (rc) return rc;
    if (getenv("HHD_VERBOSE")) fprintf(stderr, "[hhd] k_neuron: %d CTAs per SM, grid %d, %d bytes of dynamic shared memory per CTA\n", cps, h->neuron_grid, NEURON_SMEM);
  }
  /* k_synapse: spike-time CTAs + delivery CTAs in one launch; on a partitioned network every CTA must be resident while
   * CTA 0 receives the peers' spikes (the others spin), which the occupancy of the kernel bounds */
  h->stp_grid = h->sm_count * 4;
  h->deliver_grid = 0;                 /* no delivery pass: amplitudes go into the per-delay conductance buffer at spike time */
  if (getenv("HHD_STP_CTAS_PER_SM") && atoi(getenv("HHD_STP_CTAS_PER_SM")) > 0) h->stp_grid = h->sm_count * atoi(getenv("HHD_STP_CTAS_PER_SM"));
  if (!P.row_bkt) {                            /* no synapses: a valid (all-zero) bucket index */
    unsigned int* bkt = nullptr;
    unsigned short* rmax = nullptr;
    rc = dev_alloc(h, &bkt, (size_t)h->NG * DELAY_BUCKETS, true);
    if (!rc) rc = dev_alloc(h, &rmax, (size_t)h->NG, true);
    if (rc) return rc;
    P.row_bkt = bkt;
    P.row_maxd = rmax;
    P.bkt_width = 1;
    P.stp_chunks = 1;
  }
  if (P.stp_chunks < 1) return fail(h, HHD_E_STATE, "internal: stp_chunks not set");
  h->fused = P.max_delay < P.refr_steps && !getenv("HHD_NO_FUSE");      /* no delay outlasts the refractory period (k_synapse<false>) */
  h->kernels_per_step = 2;
  if (h->cfg.plan == HHD_PLAN_PERSISTENT || h->cfg.plan == HHD_PLAN_PIPELINED)
    return fail(h, HHD_E_ARG, "plan %d was removed in round 2 (measured slower than the graph plan, profiles/r1_notes.md): use auto, graph or resident", h->cfg.plan);
  {
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, h->fused ? (const void*)k_synapse<false> : (const void*)k_synapse<true>, STP_BLOCK, 0));
    const int cap = std::max(1, per_sm) * h->sm_count;
    if (h->cfg.n_ranks > 1 && h->stp_grid + h->deliver_grid > cap) {
      h->stp_grid = std::max(1, std::min(h->stp_grid, cap / 3));
      h->deliver_grid = std::max(1, cap - h->stp_grid);
    }
  }
  h->pdl = getenv("HHD_PDL") ? atoi(getenv("HHD_PDL")) != 0 : true;     /* on: -1.6 % per step on 1 GPU, -3 % on 2 (profiles/r2_notes.md); HHD_PDL=0 turns it off */
  {
    const int inst = h->cfg.instance_size > 0 ? h->cfg.instance_size : h->N;
    const bool can = h->cfg.n_ranks == 1 && P.max_delay < P.refr_steps && inst <= RES_MAX_CLUSTER * RES_BLOCK && h->N % inst == 0 &&
                     !h->cross_instance && h->NG == h->N && !P.iac_table && !P.noise_sigma;
    if (h->cfg.plan == HHD_PLAN_RESIDENT && !can)
      return fail(h, HHD_E_ARG, "HHD_PLAN_RESIDENT needs independent instances of at most %d neurons (instance_size), no synapse "
                  "between instances, max delay < refractory period and a single rank", RES_MAX_CLUSTER * RES_BLOCK);
    /* AUTO: resident only while every cluster is on the chip at once (it is the latency-bound regime's plan; with many
     * instances the grid-wide kernels have more parallelism per step — 256 x 1003 neurons: 29 vs 47 us per step) */
    int cs0 = 1;
    while (cs0 * RES_BLOCK < inst) cs0 <<= 1;
    const bool one_wave = (long long)(h->N / inst) * cs0 <= (long long)h->sm_count;
    h->resident = can && (h->cfg.plan == HHD_PLAN_RESIDENT || (h->cfg.plan == HHD_PLAN_AUTO && one_wave && !getenv("HHD_NO_RESIDENT")));
    if (h->resident) {
      P.inst_size = inst;
      P.n_inst = h->N / inst;
      int cs = 1;
      while (cs * RES_BLOCK < inst) cs <<= 1;
      P.res_cluster = cs;
      P.res_cap = (int)next_pow2((unsigned long long)std::max(64, 2 * inst));
      rc = dev_alloc(h, &P.res_store, (size_t)P.n_inst * (2 + P.W + 2 * (size_t)P.res_cap), true);
      if (rc) return rc;
      h->res_smem = 3 * RES_BLOCK * 8 + MAX_MON * (int)sizeof(MonDesc) + (int)sizeof(ResList) + P.W * 4 + 2 * P.res_cap * 4 + 16;
      const void* fn[4] = {(const void*)k_resident<0>, (const void*)k_resident<1>, (const void*)k_resident<2>, (const void*)k_resident<3>};
      for (int q = 0; q < 4; ++q) {
        CU(cudaFuncSetAttribute(fn[q], cudaFuncAttributeMaxDynamicSharedMemorySize, h->res_smem));
        if (cs > 8) CU(cudaFuncSetAttribute(fn[q], cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      }
    }
  }
  if (getenv("HHD_DEBUG_TIMES")) { rc = dev_alloc(h, &P.dbg_times, (size_t)8 * 4096, true); if (rc) return rc; }
  k_mark_flags<<<(h->NG + 255) / 256, 256, 0, h->stream>>>(h->NG, h->N, P.off, P.row_ptr, P.hot, const_cast<unsigned char*>(P.has_out), P.meta);
  CU(cudaGetLastError());
  h->started = true;
  return 0;
}

static int launch_resident(hhd_handle h, int n_steps) {
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof lc);
  lc.gridDim = dim3((unsigned)(h->P.n_inst * h->P.res_cluster));
  lc.blockDim = dim3(RES_BLOCK);
  lc.dynamicSmemBytes = (size_t)h->res_smem;
  lc.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)h->P.res_cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  cudaError_t e;
  switch (h->cfg.method) {
    case HHD_EULER: e = cudaLaunchKernelEx(&lc, k_resident<0>, h->P, n_steps); break;
    case HHD_RK2: e = cudaLaunchKernelEx(&lc, k_resident<1>, h->P, n_steps); break;
    case HHD_RK4: e = cudaLaunchKernelEx(&lc, k_resident<2>, h->P, n_steps); break;
    default: e = cudaLaunchKernelEx(&lc, k_resident<3>, h->P, n_steps); break;
  }
  if (e != cudaSuccess) return fail(h, HHD_E_CUDA, "launch of the cluster-resident kernel failed: %s", cudaGetErrorString(e));
  k_advance<<<1, 1, 0, h->stream>>>(h->d_base_step, n_steps);
  h->cnt.kernel_launches += 2;
  return 0;
}

template <typename K>
static void launch_pdl(hhd_handle h, K kernel, int grid, int block, size_t smem, cudaStream_t stream, int offset) {
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof lc);
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3((unsigned)block);
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = h->pdl ? 1 : 0;
  cudaLaunchKernelEx(&lc, kernel, h->P, offset);
}

template <bool STEP>
static void launch_neuron(hhd_handle h, int offset) {
  typedef void (*kern_t)(const Params, const int);
  const kern_t fn = (kern_t)neuron_fn_of<STEP>(h->cfg.method, h->P.iac_table != nullptr || h->P.noise_sigma != nullptr, h->mon_variant);
  launch_pdl(h, fn, h->neuron_grid, NEURON_BLOCK, NEURON_SMEM, h->stream, offset);
}

static void launch_synapse(hhd_handle h, int offset) {
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof lc);
  lc.gridDim = dim3((unsigned)std::max(1, h->stp_grid));
  lc.blockDim = dim3(STP_BLOCK);
  lc.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = h->pdl ? 1 : 0;
  if (h->fused) cudaLaunchKernelEx(&lc, k_synapse<false>, h->P, offset, h->stp_grid);
  else cudaLaunchKernelEx(&lc, k_synapse<true>, h->P, offset, h->stp_grid);
}

/* One time step: k_neuron, k_synapse.  On a partitioned network the step's spike indices travel every rank to every rank
 * in between — inside those two launches over peer memory (send_spike / remote_spike_phase), or as ncclAllGather + k_append when
 * the peers' buffers could not be mapped. */
static void launch_step(hhd_handle h, int offset) {
  launch_neuron<true>(h, offset);
#ifdef HHD_WITH_NCCL
  if (h->cfg.n_ranks > 1 && !h->p2p) {
    ncclAllGather(h->P.xchg_send, h->P.xchg_recv, (size_t)1 + h->P.xchg_cap, ncclInt32, h->comm, h->stream);
    k_append<<<1, 256, 0, h->stream>>>(h->P, offset);
  }
#endif
  launch_synapse(h, offset);
}

static int build_graph(hhd_handle h) {
  if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
  cudaGraph_t g = nullptr;
  CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  for (int s = 0; s < STEPS_PER_GRAPH; ++s) launch_step(h, s);
  k_advance<<<1, 1, 0, h->stream>>>(h->d_base_step, STEPS_PER_GRAPH);
  CU(cudaStreamEndCapture(h->stream, &g));
  cudaError_t e = cudaGraphInstantiate(&h->graph_exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(h, HHD_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  h->graph_dirty = false;
  return 0;
}

/*
 * Spike recording.  A run is cut into segments of seg_steps; at the end of a segment the ring head is copied to pinned
 * memory behind an event (segment_mark) and the NEXT segment is enqueued before the host looks at the previous one
 * (segment_collect): the host copies that segment's ring entries on a second stream and files them while the device
 * computes, so the device never waits for the host (it used to: ~13 ms of copying and sorting per 2000 steps, 10 % of a
 * 1M-neuron run).  The ring therefore holds two segments plus the delivery window.  Entries are filed in ring order
 * (step, arrival); SpikeMonitor order (step, neuron) is established when the spikes are read (finish_spike_order).
 */
static int segment_mark(hhd_handle h, int slot) {
  CU(cudaMemcpyAsync(&h->pin_head[slot], h->P.head, 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaEventRecord(h->ev_seg[slot], h->stream));
  return 0;
}

static int segment_collect(hhd_handle h, int slot) {
  CU(cudaEventSynchronize(h->ev_seg[slot]));
  const unsigned long long head = h->pin_head[slot];
  const unsigned long long n = head - h->drained;
  /* while the host read [drained_prev, drained) the device appended up to head */
  if (head - h->drained_prev > h->ring_cap)
    return fail(h, HHD_E_CAPACITY, "spike ring overflow (%llu spikes in two segments, capacity %llu)", head - h->drained_prev, h->ring_cap);
  if (h->record_spikes && n) {
    if (h->pin_cap < n) {
      if (h->pin_nb) cudaFreeHost(h->pin_nb);
      if (h->pin_sb) cudaFreeHost(h->pin_sb);
      h->pin_nb = h->pin_sb = nullptr;
      h->pin_cap = (size_t)n + (size_t)n / 2 + 1024;
      CU(cudaMallocHost(&h->pin_nb, h->pin_cap * 4));
      CU(cudaMallocHost(&h->pin_sb, h->pin_cap * 4));
    }
    const unsigned long long p0 = h->drained & (h->ring_cap - 1);
    const unsigned long long first = std::min(n, h->ring_cap - p0);
    CU(cudaMemcpyAsync(h->pin_nb, h->P.spk_neuron + p0, (size_t)first * 4, cudaMemcpyDeviceToHost, h->stream_copy));
    CU(cudaMemcpyAsync(h->pin_sb, h->P.spk_step + p0, (size_t)first * 4, cudaMemcpyDeviceToHost, h->stream_copy));
    if (first < n) {
      CU(cudaMemcpyAsync(h->pin_nb + first, h->P.spk_neuron, (size_t)(n - first) * 4, cudaMemcpyDeviceToHost, h->stream_copy));
      CU(cudaMemcpyAsync(h->pin_sb + first, h->P.spk_step, (size_t)(n - first) * 4, cudaMemcpyDeviceToHost, h->stream_copy));
    }
    CU(cudaStreamSynchronize(h->stream_copy));
    const int lo = h->cfg.global_offset, hi = lo + h->N;
    const int* nb = h->pin_nb;
    const int* sb = h->pin_sb;
    for (size_t q = 0; q < (size_t)n; ++q)
      if (nb[q] >= lo && nb[q] < hi) { h->rec_step.push_back((long long)sb[q]); h->rec_neuron.push_back(nb[q]); }
  }
  h->drained_prev = h->drained;
  h->drained = head;
  return 0;
}

/* End of a piece of work whose spikes are wanted now. */
static int drain_spikes(hhd_handle h) {
  int rc = segment_mark(h, 0);
  if (!rc) rc = segment_collect(h, 0);
  return rc;
}

/* (step, neuron) order for everything filed since the last call.  Segments are filed in time order, and inside one the
 * ring is ordered by step unless a plan appends per instance: sort each step's run, or the whole tail if it is not. */
static void finish_spike_order(hhd_handle h) {
  const size_t n = h->rec_neuron.size();
  size_t a = h->rec_sorted;
  if (a >= n) { h->rec_sorted = n; return; }
  bool monotone = true;
  for (size_t q = a + 1; q < n && monotone; ++q) monotone = h->rec_step[q] >= h->rec_step[q - 1];
  if (monotone) {
    while (a < n) {
      size_t b = a + 1;
      while (b < n && h->rec_step[b] == h->rec_step[a]) ++b;
      std::sort(h->rec_neuron.begin() + a, h->rec_neuron.begin() + b);
      a = b;
    }
  } else {
    std::vector<std::pair<long long, int>> tmp;
    tmp.reserve(n - a);
    for (size_t q = a; q < n; ++q) tmp.emplace_back(h->rec_step[q], h->rec_neuron[q]);
    std::sort(tmp.begin(), tmp.end());
    for (size_t q = a; q < n; ++q) { h->rec_step[q] = tmp[q - a].first; h->rec_neuron[q] = tmp[q - a].second; }
  }
  h->rec_sorted = n;
}

static int grow_monitors(hhd_handle h, long long n_steps) {
  for (Monitor& m : h->mons) {
    if (!m.active) continue;
    const long long need = m.n_samples + n_steps;
    if (need > m.capacity) {
      const long long ncap = std::max(need, m.capacity * 2);
      double* nb = nullptr;
      int rc = dev_alloc(h, &nb, (size_t)(ncap * m.width), true);
      if (rc) return rc;
      if (m.n_samples) CU(cudaMemcpyAsync(nb, m.d_buf, (size_t)(m.n_samples * m.width) * 8, cudaMemcpyDeviceToDevice, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      dev_free(h, m.d_buf);
      m.d_buf = nb;
      m.capacity = ncap;
    }
  }
  return 0;
}

static int push_monitor_descs(hhd_handle h) {
  MonDesc d[MAX_MON];
  memset(d, 0, sizeof d);
  int nred = 0;
  for (size_t q = 0; q < h->mons.size(); ++q) {
    const Monitor& m = h->mons[q];
    d[q].red_slot = m.reduce != HHD_REDUCE_NONE ? nred++ : -1;
    d[q].var = m.var; d[q].reduce = m.reduce; d[q].active = m.active; d[q].n_idx = m.n_idx;
    d[q].k0 = h->step - m.n_samples; d[q].capacity = m.capacity; d[q].slot = m.d_slot; d[q].buf = m.d_buf;
  }
  CU(cudaMemcpyAsync(h->d_mon, d, sizeof d, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->P.n_mon = (int)h->mons.size();
  bool any = false;
  for (const Monitor& m : h->mons) any = any || m.active;
  if (any != h->mon_variant) { h->mon_variant = any; h->graph_dirty = true; }
  return 0;
}

extern "C" int hhd_run(hhd_handle h, int64_t n_steps) {
  if (!h || n_steps < 0) return fail(h, HHD_E_ARG, "bad argument");
  cudaSetDevice(h->cfg.device);
  int rc = 0;
  if (!h->started && (rc = finalize_setup(h))) return rc;
  if (n_steps == 0) return HHD_OK;
  if (h->step + n_steps >= (1ll << 27) - 1) return fail(h, HHD_E_ARG, "step index would exceed 2^27");
  if (h->ext_dirty && (rc = rebuild_ext(h))) return rc;
  if ((rc = grow_monitors(h, n_steps))) return rc;
  const int prev_n_mon = h->P.n_mon;
  if ((rc = push_monitor_descs(h))) return rc;
  if (prev_n_mon != h->P.n_mon) h->graph_dirty = true;
  for (Monitor& m : h->mons) if (m.active && m.first_step < 0) m.first_step = h->step;
  CU(cudaMemcpyAsync(h->d_base_step, &h->step, 8, cudaMemcpyHostToDevice, h->stream));
  if (!h->resident && h->graph_dirty && n_steps >= STEPS_PER_GRAPH && (rc = build_graph(h))) return rc;

  CU(cudaEventRecord(h->ev0, h->stream));
  long long done = 0;
  int pending = -1, slot = 0;          /* segment whose spikes the host has not filed yet */
  auto segment_done = [&]() -> int {   /* mark this segment, then file the previous one while this one runs */
    int r = segment_mark(h, slot);
    if (!r && pending >= 0) r = segment_collect(h, pending);
    pending = slot;
    slot ^= 1;
    return r;
  };
  while (done < n_steps) {
    const long long seg = std::min<long long>(h->seg_steps, n_steps - done);
    if (h->resident) {
      if ((rc = launch_resident(h, (int)seg))) return rc;
      done += seg;
      CU(cudaGetLastError());
      if ((rc = segment_done())) return rc;
      continue;
    }
    long long s = 0;
    while (seg - s >= STEPS_PER_GRAPH && h->graph_exec && !h->graph_dirty) {
      CU(cudaGraphLaunch(h->graph_exec, h->stream));
      h->cnt.kernel_launches += (h->kernels_per_step + (h->cfg.n_ranks > 1 && !h->p2p ? 2 : 0)) * STEPS_PER_GRAPH + 1;
      s += STEPS_PER_GRAPH;
    }
    const int rem = (int)(seg - s);
    for (int q = 0; q < rem; ++q) launch_step(h, q);
    if (rem) { k_advance<<<1, 1, 0, h->stream>>>(h->d_base_step, rem); h->cnt.kernel_launches += h->kernels_per_step * rem + 1; }
    done += seg;
    if (done == n_steps) {
      launch_neuron<false>(h, 0);
      h->cnt.kernel_launches += 1;
    }
    CU(cudaGetLastError());
    if ((rc = segment_done())) return rc;
  }
  CU(cudaEventRecord(h->ev1, h->stream));
  if (pending >= 0 && (rc = segment_collect(h, pending))) return rc;
  CU(cudaStreamSynchronize(h->stream));
  if (h->P.dbg_times) {          /* HHD_DEBUG_TIMES=path: the last 4096 steps' stamps of this rank -> path.rank<r> (tools/step_stamps.py) */
    std::vector<unsigned long long> tt((size_t)8 * 4096);
    CU(cudaMemcpy(tt.data(), h->P.dbg_times, tt.size() * 8, cudaMemcpyDeviceToHost));
    const std::string path = std::string(getenv("HHD_DEBUG_TIMES")) + ".rank" + std::to_string(h->cfg.rank);
    if (FILE* f = fopen(path.c_str(), "w")) {
      for (int q = 0; q < 4096; ++q) {
        fprintf(f, "%d", q);
        for (int j = 0; j < 8; ++j) fprintf(f, " %llu", tt[(size_t)q * 8 + j]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  h->cnt.run_ms_device += ms;
  for (Monitor& m : h->mons) if (m.active) m.n_samples += n_steps;
  h->step += n_steps;
  h->cnt.steps += n_steps;
  h->cnt.neuron_updates += n_steps * (long long)h->N;
  unsigned long long c[CNT_N];
  CU(cudaMemcpy(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost));
  if (c[CNT_TIMEOUT]) return fail(h, HHD_E_COMM, "a peer rank did not deliver its spikes within 20 s (spike exchange over peer memory)");
  if (c[CNT_OVERFLOW]) return fail(h, HHD_E_CAPACITY, "a device buffer overflowed (monitor capacity, or more spikes in one step than the exchange buffer holds: HHD_XCHG_CAP)");
  {
    unsigned long long slots[CNT_SLOTS];
    CU(cudaMemcpy(slots, h->P.del_slots, sizeof slots, cudaMemcpyDeviceToHost));
    for (int q = 0; q < CNT_SLOTS; ++q) c[CNT_DELIVERIES] += slots[q];
  }
  if (!h->resident && h->nsyn) {
    /* the graph plan counts a delivery when it is pushed into the conductance buffer; Brian (and the oracle) count it when
     * the delayed pathway executes: take off what is still in flight after the last completed step */
    const long long last = h->step - 1, s0 = h->step - h->P.max_delay - 1;
    std::vector<unsigned long long> ends((size_t)h->P.W);
    unsigned long long head = 0, *d_n = nullptr, n_fly = 0;
    CU(cudaMemcpy(ends.data(), h->P.step_end, ends.size() * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&head, h->P.head, 8, cudaMemcpyDeviceToHost));
    const unsigned long long lo = s0 >= 0 ? ends[(size_t)(s0 % h->P.W)] : 0ull;      /* head at the end of step s0 */
    if (head > lo) {
      if ((rc = dev_alloc(h, &d_n, 1, true))) return rc;
      k_inflight<<<(unsigned)std::min<unsigned long long>(1024, (head - lo + 3) / 4), 128, 0, h->stream>>>(h->P, last, lo, head, d_n);
      CU(cudaMemcpyAsync(&n_fly, d_n, 8, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      dev_free(h, d_n);
    }
    c[CNT_DELIVERIES] -= std::min(n_fly, c[CNT_DELIVERIES]);
  }
  h->cnt.syn_events = (long long)c[CNT_SYN_EVENTS]; h->cnt.syn_deliveries = (long long)c[CNT_DELIVERIES];
  h->cnt.ext_events = (long long)c[CNT_EXT]; h->cnt.spikes = (long long)c[CNT_SPIKES];
  h->cnt.w_crossings = (long long)c[CNT_WCROSS]; h->cnt.nonfinite = (long long)c[CNT_NONFINITE];
  return HHD_OK;
}

extern "C" int hhd_profile_kernels(hhd_handle h, int64_t n_steps, double us[3]) {
  if (!h || n_steps <= 0 || !us) return fail(h, HHD_E_ARG, "bad argument");
  cudaSetDevice(h->cfg.device);
  int rc = 0;
  if (!h->started && (rc = finalize_setup(h))) return rc;
  if (n_steps > h->seg_steps) n_steps = h->seg_steps;
  if (h->ext_dirty && (rc = rebuild_ext(h))) return rc;
  if ((rc = grow_monitors(h, n_steps))) return rc;
  const int prev_n_mon = h->P.n_mon;
  if ((rc = push_monitor_descs(h))) return rc;
  if (prev_n_mon != h->P.n_mon) h->graph_dirty = true;
  for (Monitor& m : h->mons) if (m.active && m.first_step < 0) m.first_step = h->step;
  CU(cudaMemcpyAsync(h->d_base_step, &h->step, 8, cudaMemcpyHostToDevice, h->stream));
  if (h->resident) {
    CU(cudaEventRecord(h->ev0, h->stream));
    if ((rc = launch_resident(h, (int)n_steps))) return rc;
    CU(cudaEventRecord(h->ev1, h->stream));
    if ((rc = drain_spikes(h))) return rc;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    us[0] = ms * 1e3 / (double)n_steps; us[1] = 0; us[2] = 0;
    for (Monitor& m : h->mons) if (m.active) m.n_samples += n_steps;
    h->step += n_steps;
    h->cnt.steps += n_steps;
    h->cnt.neuron_updates += n_steps * (long long)h->N;
    return HHD_OK;
  }
  std::vector<cudaEvent_t> ev((size_t)n_steps * 3 + 1);
  for (auto& e : ev) CU(cudaEventCreate(&e));
  CU(cudaEventRecord(ev[0], h->stream));
  for (int64_t s = 0; s < n_steps; ++s) {
    launch_neuron<true>(h, (int)s);
    CU(cudaEventRecord(ev[3 * s + 1], h->stream));
#ifdef HHD_WITH_NCCL
    if (h->cfg.n_ranks > 1 && !h->p2p) {   /* slot 1: the NCCL fallback of the exchange (over peer memory it is inside the two kernels) */
      ncclAllGather(h->P.xchg_send, h->P.xchg_recv, (size_t)1 + h->P.xchg_cap, ncclInt32, h->comm, h->stream);
      k_append<<<1, 256, 0, h->stream>>>(h->P, (int)s);
    }
#endif
    CU(cudaEventRecord(ev[3 * s + 2], h->stream));
    launch_synapse(h, (int)s);
    CU(cudaEventRecord(ev[3 * s + 3], h->stream));
  }
  k_advance<<<1, 1, 0, h->stream>>>(h->d_base_step, (int)n_steps);
  launch_neuron<false>(h, 0);
  h->cnt.kernel_launches += h->kernels_per_step * n_steps + 2;
  CU(cudaGetLastError());
  if ((rc = drain_spikes(h))) return rc;
  double acc[3] = {0, 0, 0};
  for (int64_t s = 0; s < n_steps; ++s)
    for (int q = 0; q < 3; ++q) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[3 * s + q], ev[3 * s + q + 1]);
      acc[q] += ms;
    }
  for (auto& e : ev) cudaEventDestroy(e);
  for (int q = 0; q < 3; ++q) us[q] = acc[q] * 1e3 / (double)n_steps;
  for (Monitor& m : h->mons) if (m.active) m.n_samples += n_steps;
  h->step += n_steps;
  h->cnt.steps += n_steps;
  h->cnt.neuron_updates += n_steps * (long long)h->N;
  return HHD_OK;
}

/* ---- results ----------------------------------------------------------------------------------------------------- */
extern "C" int hhd_spike_count(hhd_handle h, int64_t* n) {
  if (!h || !n) return fail(h, HHD_E_ARG, "null argument");
  *n = (int64_t)h->rec_neuron.size();
  return HHD_OK;
}

extern "C" int hhd_read_spikes(hhd_handle h, int32_t* neuron, int64_t* step, int64_t cap, int64_t* n_out) {
  if (!h) return HHD_E_ARG;
  finish_spike_order(h);
  const int64_t n = (int64_t)h->rec_neuron.size();
  if (n > cap) return fail(h, HHD_E_CAPACITY, "output buffer too small: need %lld", (long long)n);
  if (n) { memcpy(neuron, h->rec_neuron.data(), (size_t)n * 4); memcpy(step, h->rec_step.data(), (size_t)n * 8); }
  if (n_out) *n_out = n;
  return HHD_OK;
}

extern "C" int hhd_read_state(hhd_handle h, int32_t var, double* out) {
  if (!h || !out || var < 0 || var >= HHD_N_VARS) return fail(h, HHD_E_ARG, "bad argument");
  cudaSetDevice(h->cfg.device);
  double* tmp = nullptr;
  int rc = dev_alloc(h, &tmp, (size_t)h->N, false);
  if (rc) return rc;
  k_read_var<<<(h->N + 127) / 128, 128, 0, h->stream>>>(h->P, var, tmp);
  cudaError_t e = cudaMemcpyAsync(out, tmp, (size_t)h->N * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dev_free(h, tmp);
  if (e != cudaSuccess) return fail(h, HHD_E_CUDA, "read_state failed: %s", cudaGetErrorString(e));
  return HHD_OK;
}

extern "C" int hhd_write_state(hhd_handle h, int32_t var, const double* in) {
  if (!h || !in) return fail(h, HHD_E_ARG, "null argument");
  cudaSetDevice(h->cfg.device);
  if (var >= 0 && var < 8) return scatter_field(h, var, in);
  if (var == HHD_LAST_SPIKE) {       /* given in ms, held in seconds */
    std::vector<double> ls((size_t)h->N);
    for (int i = 0; i < h->N; ++i) ls[i] = in[i] * 1e-3;
    return upload(h, h->P.last_spike + h->cfg.global_offset, ls.data(), (size_t)h->N);
  }
  return fail(h, HHD_E_ARG, "variable %d is not writable", var);
}

extern "C" int hhd_read_syn_state(hhd_handle h, int32_t which, double* out) {
  if (!h || !out) return fail(h, HHD_E_ARG, "null argument");
  if (which < HHD_SYN_U || which > HHD_SYN_DELAY_STEPS) return fail(h, HHD_E_ARG, "unknown synapse variable %d", which);
  cudaSetDevice(h->cfg.device);
  if (!h->nsyn) return HHD_OK;
  double* tmp = nullptr;
  int rc = dev_alloc(h, &tmp, (size_t)h->nsyn, false);
  if (rc) return rc;
  k_syn_field<<<(unsigned)((h->nsyn + 255) / 256), 256, 0, h->stream>>>(h->nsyn, h->P.hot, h->P.stp, which, tmp);
  cudaError_t e = cudaMemcpyAsync(out, tmp, (size_t)h->nsyn * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dev_free(h, tmp);
  if (e != cudaSuccess) return fail(h, HHD_E_CUDA, "read_syn_state failed: %s", cudaGetErrorString(e));
  return HHD_OK;
}

extern "C" int hhd_get_counters(hhd_handle h, hhd_counters* out) {
  if (!h || !out) return fail(h, HHD_E_ARG, "null argument");
  *out = h->cnt;
  return HHD_OK;
}

extern "C" int hhd_current_step(hhd_handle h, int64_t* step) {
  if (!h || !step) return fail(h, HHD_E_ARG, "null argument");
  *step = h->step;
  return HHD_OK;
}

extern "C" int hhd_comm_unique_id(uint8_t out_id[128]) {
#ifdef HHD_WITH_NCCL
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return HHD_E_COMM;
  memcpy(out_id, &id, 128);
  return HHD_OK;
#else
  memset(out_id, 0, 128);
  return HHD_E_COMM;
#endif
}

extern "C" int hhd_comm_init(hhd_handle h, const uint8_t id_bytes[128]) {
  if (!h || !id_bytes) return fail(h, HHD_E_ARG, "null argument");
  if (h->started) return fail(h, HHD_E_STATE, "hhd_comm_init after run");
#ifdef HHD_WITH_NCCL
  cudaSetDevice(h->cfg.device);
  ncclUniqueId id;
  memcpy(&id, id_bytes, 128);
  ncclResult_t r = ncclCommInitRank(&h->comm, h->cfg.n_ranks, id, h->cfg.rank);
  if (r != ncclSuccess) return fail(h, HHD_E_COMM, "ncclCommInitRank failed: %s", ncclGetErrorString(r));
  h->comm_ready = true;
  return HHD_OK;
#else
  return fail(h, HHD_E_COMM, "this libhhd.so was built without NCCL");
#endif
}
