/*
 * hhd_model.cuh — device-side arithmetic of the simplified-AdEx neuron and its synapses, shared by the grid-wide
 * (k_neuron / k_synapse) and cluster-resident (k_resident) kernels of hhd_engine.cu.
 *
 * Every expression keeps the operand order of the reference's equation strings (codes/CortexNetwork.py:233-302 for the
 * neuron, :360-372 for the synapse pathways) and is compiled with -fmad=false, so a neuron's trajectory is bit-identical
 * to the CPU oracle's (oracle/hhd_oracle.c, gcc -ffp-contract=off).  exp() is hhd_exp from hhd_math.h on both sides.
 */
#ifndef HHD_MODEL_CUH_
#define HHD_MODEL_CUH_

#include "hhd_math.h"

struct ChannelConst {
  double inv_tau_on[3], inv_tau_off[3];   /* 1/tau (codes/CortexNetwork.py:256-263 write -(1/tau)*g) */
  double E[3], gsyn[3];
  double mg_fac, mg_slope, mg_half;
  double window_s;                         /* `t - last_spike < 5*ms` (:251), in seconds like t and last_spike (hhd_time_s) */
};

struct NeuronPar {
  double C, gL, EL, dT, Vup, tauw, b, Vr, VT, Iref, Vdep, IDC;
};

struct SubExpr {
  double I_AMPA, I_GABA, I_NMDA, I_syn, I_inj, I_tot, I_exp, w_V, D0, dD0, ex, Mg, g[3];
};

/* state order: V, w, gA_on, gA_off, gG_on, gG_off, gN_on, gN_off (= HHD_V .. HHD_G_NMDA_OFF) */
__device__ __forceinline__ void hhd_subexpr(const ChannelConst& cc, const NeuronPar& p, const double x[8], SubExpr& s,
                                            const double iac = 0.0) {
  const double V = x[0];
  s.g[0] = x[3] - x[2];
  s.g[1] = x[5] - x[4];
  s.g[2] = x[7] - x[6];
  s.I_AMPA = s.g[0] * (cc.E[0] - V);
  const double Mg = hhd_div(1.0, 1.0 + cc.mg_fac * hhd_exp(cc.mg_slope * (cc.mg_half - V)));
  s.I_GABA = s.g[1] * (cc.E[1] - V);
  s.I_NMDA = (s.g[2] * (cc.E[2] - V)) * Mg;
  s.Mg = Mg;
  s.I_syn = (s.I_AMPA + s.I_NMDA) + s.I_GABA;
  s.I_inj = p.IDC + iac;                                  /* I_inj = I_DC + I_AC   (codes/CortexNetwork.py:236-237) */
  s.I_tot = s.I_syn + s.I_inj;
  s.ex = hhd_exp(hhd_div(V - p.VT, p.dT));
  s.I_exp = (p.gL * p.dT) * s.ex;
  s.w_V = (s.I_tot + s.I_exp) - p.gL * (V - p.EL);
  s.D0 = hhd_div(p.C, p.gL) * s.w_V;
  s.dD0 = p.C * (s.ex - 1.0);
}

/* The subexpressions again after ONLY the conductances changed (synaptic increments): Mg, exp((V-V_T)/delta_T), I_exp and
 * dD0 depend on V alone and are kept; everything else is recomputed with the operations and order of hhd_subexpr, so the
 * result is bit-identical to a full re-evaluation at a third of its cost (no exponential, one division). */
__device__ __forceinline__ void hhd_subexpr_new_g(const ChannelConst& cc, const NeuronPar& p, const double x[8], SubExpr& s) {
  const double V = x[0];
  s.g[0] = x[3] - x[2];
  s.g[1] = x[5] - x[4];
  s.g[2] = x[7] - x[6];
  s.I_AMPA = s.g[0] * (cc.E[0] - V);
  s.I_GABA = s.g[1] * (cc.E[1] - V);
  s.I_NMDA = (s.g[2] * (cc.E[2] - V)) * s.Mg;
  s.I_syn = (s.I_AMPA + s.I_NMDA) + s.I_GABA;
  s.I_tot = s.I_syn + s.I_inj;
  s.w_V = (s.I_tot + s.I_exp) - p.gL * (V - p.EL);
  s.D0 = hhd_div(p.C, p.gL) * s.w_V;
}

/* Right-hand sides at state x given the subexpressions `s` of that state. */
/* ts = stage time, ls = the neuron's last_spike, both in SECONDS (Brian's SI values, see hhd_time_s in hhd_math.h). */
__device__ __forceinline__ void hhd_rhs(const ChannelConst& cc, const NeuronPar& p, const double x[8], double ts, double ls,
                                        const SubExpr& s, double d[8]) {
  const double V = x[0], w = x[1];
  const double a = ((s.I_tot >= p.Iref) ? 1.0 : 0.0) * ((ts - ls < cc.window_s) ? 1.0 : 0.0);
  const double dV = (a * hhd_div(-p.gL, p.C)) * (V - p.Vdep) + hhd_div((1.0 - a) * (s.w_V - w), p.C);
  const double band = hhd_div(s.D0, p.tauw);
  const double gate = ((w > s.w_V - band) ? 1.0 : 0.0) * ((w < s.w_V + band) ? 1.0 : 0.0) *
                      ((V <= p.VT) ? 1.0 : 0.0) * ((s.I_tot < p.Iref) ? 1.0 : 0.0);
  d[0] = dV;
  d[1] = (gate * -(p.gL * (1.0 - s.ex) + hhd_div(s.dD0, p.tauw))) * dV;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    d[2 + 2 * c] = -(cc.inv_tau_on[c] * x[2 + 2 * c]);
    d[3 + 2 * c] = -(cc.inv_tau_off[c] * x[3 + 2 * c]);
  }
}

/* Subexpressions + right-hand sides (stages after the first). */
__device__ __forceinline__ void hhd_deriv(const ChannelConst& cc, const NeuronPar& p, const double x[8], double t,
                                          double ls, double d[8], SubExpr& s, const double iac = 0.0) {
  hhd_subexpr(cc, p, x, s, iac);
  hhd_rhs(cc, p, x, t, ls, s, d);
}

/* Brian 2's ExplicitStateUpdaters: METHOD 0 euler, 1 rk2 (midpoint), 2 rk4.  `s1` = subexpressions of the state x on entry
 * (the caller has them already: they also decide the deferred w_crossing test and feed the monitors). */
template <int METHOD>
__device__ __forceinline__ void hhd_integrate(const ChannelConst& cc, const NeuronPar& p, double x[8], double t /* s */,
                                              double ls /* s */, double dt /* ms */, const SubExpr& s1, const double iac0 = 0.0,
                                              const double iac1 = 0.0,     /* I_AC of this step / of the next (stage at t+dt) */
                                              double* hcar = nullptr) {    /* adaptive method only: the carried step size */
  double k1[8], y[8];
  SubExpr sk;
  const double dts = hhd_dt_s(dt);
  hhd_rhs(cc, p, x, t, ls, s1, k1);
#pragma unroll
  for (int v = 0; v < 8; ++v) k1[v] = dt * k1[v];
  if (METHOD == 0) {
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] = x[v] + k1[v];
    return;
  }
  double k2[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) y[v] = x[v] + k1[v] * 0.5;
  hhd_deriv(cc, p, y, t + dts * 0.5, ls, k2, sk, iac0);
#pragma unroll
  for (int v = 0; v < 8; ++v) k2[v] = dt * k2[v];
  if (METHOD == 1) {
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] = x[v] + k2[v];
    return;
  }
  double k3[8], k4[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) y[v] = x[v] + k2[v] * 0.5;
  hhd_deriv(cc, p, y, t + dts * 0.5, ls, k3, sk, iac0);
#pragma unroll
  for (int v = 0; v < 8; ++v) k3[v] = dt * k3[v];
#pragma unroll
  for (int v = 0; v < 8; ++v) y[v] = x[v] + k3[v];
  hhd_deriv(cc, p, y, t + dts, ls, k4, sk, iac1);
#pragma unroll
  for (int v = 0; v < 8; ++v) k4[v] = dt * k4[v];
#pragma unroll
  for (int v = 0; v < 8; ++v) x[v] = x[v] + hhd_div(((k1[v] + 2.0 * k2[v]) + 2.0 * k3[v]) + k4[v], 6.0);
}

/* METHOD 3: `gsl_rk2` — embedded RK2(3) (gsl_odeiv2_step_rk2) under GSL's standard step-size control with per-neuron
 * sub-stepping inside dt and the step size carried from one dt to the next; operation for operation the function
 * integrate_rk23 of oracle/hhd_oracle.c, see there for the algorithm and its sources. */
template <>
__device__ __forceinline__ void hhd_integrate<3>(const ChannelConst& cc, const NeuronPar& p, double x[8], double t0,
                                                 double ls, double dt, const SubExpr& s1, const double iac0,
                                                 const double iac1, double* hcar) {
  const double inv_tol[8] = {1e3, 1e-6, 1e-3, 1e-3, 1e-3, 1e-3, 1e-3, 1e-3};
  double t = 0.0, h0 = *hcar;
  bool first = true;
  for (int it = 0; it < 4096 && t < dt; ++it) {
    double hs = h0;
    bool final_step = false;
    if (hs > dt - t) { hs = dt - t; final_step = true; }
    double k1[8], k2[8], k3[8], y[8], ynew[8];
    SubExpr sk;
    if (first) hhd_rhs(cc, p, x, t0 + t * 1e-3, ls, s1, k1);
    else hhd_deriv(cc, p, x, t0 + t * 1e-3, ls, k1, sk, iac0);
    first = false;
#pragma unroll
    for (int v = 0; v < 8; ++v) y[v] = x[v] + (0.5 * hs) * k1[v];
    hhd_deriv(cc, p, y, t0 + (t + 0.5 * hs) * 1e-3, ls, k2, sk, iac0);
#pragma unroll
    for (int v = 0; v < 8; ++v) y[v] = x[v] + hs * (2.0 * k2[v] - k1[v]);
    hhd_deriv(cc, p, y, t0 + (t + hs) * 1e-3, ls, k3, sk, final_step ? iac1 : iac0);
    double rmax = 0.0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      ynew[v] = x[v] + hs * hhd_div((k1[v] + 4.0 * k2[v]) + k3[v], 6.0);
      const double err = hs * hhd_div((k1[v] - 2.0 * k2[v]) + k3[v], 6.0);
      const double r = fabs(err) * inv_tol[v];
      if (r > rmax) rmax = r;
    }
    if (rmax > 1.1 && hs > dt * 9.5367431640625e-07) {
      double f = hhd_div(0.9, sqrt(rmax));
      if (f < 0.2) f = 0.2;
      h0 = hs * f;
      continue;
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) x[v] = ynew[v];
    t = final_step ? dt : t + hs;
    double hn = hs;
    if (rmax < 0.5) {
      double f = 5.0;
      if (rmax > 9.765625e-4) {
        f = hhd_div(0.9, hhd_cbrt(rmax));
        if (f > 5.0) f = 5.0;
        if (f < 1.0) f = 1.0;
      }
      hn = hs * f;
    }
    if (!final_step) h0 = hn;
  }
  *hcar = h0;
}

/* Derived variable (id >= 8, != last_spike) from subexpressions already evaluated at the sampled state. */
__device__ __forceinline__ double hhd_var_from_subexpr(const SubExpr& s, int var) {
  switch (var) {
    case 8: return s.g[0];
    case 9: return s.g[1];
    case 10: return s.g[2];
    case 11: return s.I_AMPA;
    case 12: return s.I_GABA;
    case 13: return s.I_NMDA;
    case 14: return s.I_syn;
    case 15: return s.I_inj;
    case 16: return s.I_tot;
    case 17: return s.I_AMPA + s.I_NMDA;
    case 18: return (s.I_AMPA + s.I_NMDA) + s.I_inj;
    case 19: return s.I_exp;
  }
  return 0.0;
}
__device__ __forceinline__ bool hhd_var_is_derived(int var) { return var >= 8 && var != 20; }

/* A state variable (id < 8) or last_spike (id 20) — no subexpression is evaluated.  (hhd_var_value below is side-effect free,
 * so when both arms of `derived ? from_subexpr : var_value` are inlined the compiler evaluates its two exponentials and the
 * division for every neuron and selects afterwards: 130 of the 800 instructions of a monitored Euler step.) */
__device__ __forceinline__ double hhd_var_state(const double x[8], int var, double last_spike) {
  double v = last_spike * 1e3;               /* held in seconds, reported in ms */
#pragma unroll
  for (int q = 0; q < 8; ++q) v = (var == q) ? x[q] : v;
  return v;
}

/* Value of a monitored / derived variable (ids of include/hhd.h) from the state. */
__device__ __forceinline__ double hhd_var_value(const ChannelConst& cc, const NeuronPar& p, const double x[8], int var,
                                                double last_spike, const double iac = 0.0) {
  switch (var) {   /* constant indices keep x[] in registers */
    case 0: return x[0];
    case 1: return x[1];
    case 2: return x[2];
    case 3: return x[3];
    case 4: return x[4];
    case 5: return x[5];
    case 6: return x[6];
    case 7: return x[7];
    case 20: return last_spike * 1e3;      /* held in seconds, reported in ms */
  }
  SubExpr s;
  hhd_subexpr(cc, p, x, s, iac);
  switch (var) {
    case 8: return s.g[0];
    case 9: return s.g[1];
    case 10: return s.g[2];
    case 11: return s.I_AMPA;
    case 12: return s.I_GABA;
    case 13: return s.I_NMDA;
    case 14: return s.I_syn;
    case 15: return s.I_inj;
    case 16: return s.I_tot;
    case 17: return s.I_AMPA + s.I_NMDA;
    case 18: return (s.I_AMPA + s.I_NMDA) + s.I_inj;
    case 19: return s.I_exp;
  }
  return 0.0;
}

/* Pathways p0..p6 for one synapse at presynaptic spike time t (codes/CortexNetwork.py:360-366; t, last, tau_rec, tau_fac in
 * seconds); returns the amplitude
 * ((g_max*a_syn)*(1-failure))*Gsyn that the order-7 pathways add when the delay has elapsed (:367-372). */
__device__ __forceinline__ double hhd_stp_update(double& u, double& R, double U, double tau_rec, double tau_fac,
                                                 double gmax, double gsyn, double pfail, double t, double last,
                                                 unsigned long long seed, unsigned long long syn_index,
                                                 unsigned long long step) {
  const double failure = hhd_philox_u01(seed, syn_index, step, 0) < pfail ? 1.0 : 0.0;
  const double u1 = U + (u * (1.0 - U)) * hhd_exp(hhd_div(-(t - last), tau_fac));
  const double R1 = 1.0 + ((R - u * R) - 1.0) * hhd_exp(hhd_div(-(t - last), tau_rec));
  u = u1;
  R = R1;
  const double a_syn = u1 * R1;
  return ((gmax * a_syn) * (1.0 - failure)) * gsyn;
}

#endif /* HHD_MODEL_CUH_ */
