// ev2b_device.cuh -- device data layout and kernels of the batched EV2Gym step engine (sm_100a).
//
// What this replaces (paths relative to /root/reference): the Python per-object cascade
//   EV2Gym.step (ev2gym/models/ev2gym_env.py:333-447) -> EV_Charger.step (ev_charger.py:114-233)
//   -> EV.step/_charge/_discharge (ev.py:138-186, 240-355, 357-405) -> Transformer.reset/step/
//   get_how_overloaded (transformer.py:258-302), the spawn loop (ev2gym_env.py:399-417),
//   calculate_charge_power_potential (utils.py:760-791), the three stock rewards (reward.py) and the
//   three stock state functions (state.py), for E env replicas at once.
//
// Mapping: ONE THREAD PER (env, charger).  A CTA owns EPB whole envs (EPB*C threads) so that the
// per-transformer and per-env reductions stay inside the CTA (shared memory + warp shuffles, fixed
// order => bitwise reproducible).  All per-port arrays are [E,P] struct-of-arrays: consecutive
// threads touch consecutive addresses.  Arithmetic is IEEE float64 in the reference's operation
// order; this translation unit is compiled with -fmad=false so no multiply-add is contracted.
// No tensor cores: the path is elementwise + segmented reduce (HBM-bound).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ev2b.h"

namespace ev2b {

constexpr int   kNoArrival = 32767;   // "no (further) session on this port"
constexpr int   kMaxThreads = 1024;
constexpr int   kNRed = 8;            // float64 partials per charger, see Red* below
enum { RedP = 0, RedA, RedProfit, RedSatExp, RedPot, RedCharged, RedDischarged, RedSatSum };

// ---- static tables -------------------------------------------------------------------------
struct CsStatic {            // one per charger; ev_charger.py:41-75 + derived constants
    double imax, imin, imax_dis_abs, imin_dis;
    double veff[4];          // voltage*sqrt(phases) for phases = 1..3 (index 0 unused)  ev.py:279,365
    double max_power;        // sqrt(ph)*V*Imax/1000   utils.py:779-780
    double min_power;        // sqrt(ph)*V*Imin/1000   utils.py:781-782
    int    port_off, n_ports, tr, phases;
    int    cls, pad0, pad1, pad2;   // charger class (distinct imax/voltage/phases) for the potential table
};

struct EvSpec {              // de-duplicated EV model; ev.py:45-113
    double B, pmax_ac, pmin_ac, pmax_dis, pmin_dis, bmin, bmin_em, desired, mult;
    double ts, eta_c, eta_d; // used when the per-session milli encodings are 0xFFFF
    int    ev_phases, lut;   // lut < 0: scalar efficiencies
    double pad[3];
};
static_assert(sizeof(EvSpec) == 128, "EvSpec must be 128 B");

// "hot" words of the session currently (or last) connected to a port: 16 B, read every step.
//   w.x = t_arr (i16) | t_dep (i16) << 16
//   w.y = next_arr (i16) | cursor (u8) << 16        cursor = index of the NEXT session of this port
//   w.z = spec_id (u16) | ts_milli (u16) << 16      ts = k/1000.0 (np.round(.,3), utils.py:309) or 0xFFFF
//   w.w = eta_c_milli (u16) | eta_d_milli (u16) << 16
// A port is occupied at step t iff t_arr <= t <= t_dep (the EV is charged in step t_dep and then
// leaves, ev_charger.py:209-224).
struct SessRec { uint4 hot; double cap0; double pad; };   // cold table, read once per arrival
static_assert(sizeof(SessRec) == 32, "SessRec must be 32 B");

struct EnvT { double cp, dp, setpoint, pad; };            // per (scenario, t)
struct TrT  { double infl, solar, maxp, minp; };          // per (scenario, t, transformer)
struct DrEv { int16_t start, end; float value; };         // value = limit - limit*cap/100  transformer.py:158-163

struct Params {
    // sizes
    int E, C, P, Tr, T, D, EPB, n_dr, lut_len, Smax, S, n_cls;
    int reward_kind, state_kind, dr_steps_ahead;
    double c60;        // 60 / timescale          ev.py:296
    double p60;        // timescale / 60          ev.py:355
    double period;     // timescale
    double tr_voltage;
    unsigned c_magic;  // ceil(2^32 / C): tid / C == __umulhi(tid, c_magic)
    int cs_uniform;    // all chargers identical: read the layout from cs0 (constant bank) instead of global memory
    CsStatic cs0;
    // static
    const CsStatic *cs; const int *tr_cs_off; const int *tr_cs_idx; const int *obs_slot; const int *tr_obs_off;
    // scenario bank
    const EnvT *env_t; const TrT *tr_t; const SessRec *sess; const EvSpec *spec;
    const double *luts_c, *luts_d; const double *pot_kw;
    const float *trA, *trF, *tr_limit; const DrEv *dr; const uint8_t *dr_count;
    // state
    uint4 *hot; double *cap; float *exch; int *env_step; int *env_scn; double *env_pot; double *env_usage;
    double *env_kpi;
    // io
    const void *actions;
    ev2b_step_out out;
};

__device__ __forceinline__ int hot_t_arr(const uint4 &h)   { return (int)(int16_t)(h.x & 0xFFFFu); }
__device__ __forceinline__ int hot_t_dep(const uint4 &h)   { return (int)(int16_t)(h.x >> 16); }
__device__ __forceinline__ int hot_next_arr(const uint4 &h){ return (int)(int16_t)(h.y & 0xFFFFu); }
__device__ __forceinline__ int hot_cursor(const uint4 &h)  { return (int)((h.y >> 16) & 0xFFu); }
__device__ __forceinline__ int hot_spec(const uint4 &h)    { return (int)(h.z & 0xFFFFu); }

__device__ __forceinline__ double lut_get(const double *lut, int lut_len, double key) {
    // dict.get(np.round(amps), 1): integer keys 0..lut_len-1, default 1   ev.py:288,376
    if (key >= 0.0 && key < (double)lut_len) return __ldg(lut + (int)key);
    return 1.0;
}

// EV._charge  ev.py:240-355
__device__ __forceinline__ double ev_charge(const EvSpec &sp, double ts, double eta, double &cap, double amps,
                                            double veff, const Params &p, double &energy) {
    double pilot = eta * amps * veff / 1000.0 / sp.B / p.c60;      // :295-296
    const double maxd = eta * sp.pmax_ac / sp.B / p.c60;           // :297-298
    if (pilot > maxd) pilot = maxd;                                // :300-301
    const double soc = cap / sp.B;
    double curr;
    if (ts == 1.0) {                                               // :303-306
        curr = pilot + soc;
        if (curr > 1.0) curr = 1.0;
    } else {
        const double pts = ts + (pilot - maxd) / maxd * (ts - 1.0);    // :312-314
        double nsoc;
        if (soc < pts) {
            if (1.0 <= (pts - soc) / pilot) nsoc = pilot + soc;        // :323-324
            else nsoc = 1.0 + exp(sp.mult * (pilot + soc - pts) / (pts - 1.0)) * (pts - 1.0);  // :326-330
        } else {
            nsoc = 1.0 + exp(sp.mult * pilot / (pts - 1.0)) * (soc - 1.0);                     // :332-334
        }
        const double lim = (maxd > pilot) ? pilot : maxd;              // :336-339
        curr = (nsoc - soc > lim) ? lim + soc : nsoc;                  // :341-344
    }
    const double dsoc = curr - soc;
    cap = curr * sp.B;                                             // :348
    energy = dsoc * sp.B;                                          // :352
    return energy / p.p60 * 1000.0 / veff;                         // :355
}

// EV._discharge  ev.py:357-405
__device__ __forceinline__ double ev_discharge(const EvSpec &sp, double eta, double &cap, double amps,
                                               double veff, const Params &p, double &energy) {
    double given_power = amps * veff / 1000.0;                     // :367
    if (fabs(given_power) > fabs(sp.pmax_dis)) given_power = sp.pmax_dis;   // :370-371
    double given_energy = given_power * eta * p.period / 60.0;     // :381
    if (cap + given_energy < sp.bmin) {                            // :382-393
        if (cap > sp.bmin) { energy = -(cap - sp.bmin); given_energy = energy; }
        else { energy = 0.0; given_energy = 0.0; }
        cap = sp.bmin;
    } else {
        energy = given_energy;
        cap += given_energy;
    }
    return given_energy * 60.0 / p.period * 1000.0 / veff;         // :405
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- observation pieces shared by the step and reset kernels --------------------------------
// Header of the stock state functions at observation time tq (= current_step after the increment).
__device__ __forceinline__ void obs_header(const Params &p, float *row, int s, int tq, double prev_usage) {
    if (p.state_kind == EV2B_STATE_PUBLIC_PST) {                    // state.py:11-34
        row[0] = (float)((double)tq / (double)p.T);
        row[1] = (tq < p.T) ? (float)p.env_t[(size_t)s * p.T + tq].setpoint : 0.f;
        row[2] = (float)prev_usage;
    } else {                                                        // state.py:70-82, 113-125
        row[0] = (float)tq;
        row[1] = (float)prev_usage;
    }
}
// One value of the header price window / per-transformer forecast + limit blocks.  i indexes the
// flat list [20 prices][Tr * 40]; returns the obs offset through *off.
__device__ __forceinline__ float obs_series_value(const Params &p, int s, int tq, int i, int *off) {
    if (i < 20) {                                                   // abs(charge_prices[0, t:t+20]), zero padded
        *off = 2 + i;
        return (tq + i < p.T) ? (float)fabs(p.env_t[(size_t)s * p.T + tq + i].cp) : 0.f;
    }
    i -= 20;
    const int k = i / 40, j = i - k * 40;
    *off = p.tr_obs_off[k] + j;
    const size_t base = ((size_t)s * p.Tr + k) * p.T;
    if (j < 20) {                                                   // loads - pv   transformer.py:173-188
        const int idx = tq + j;
        if (j == 0 && tq < p.T) return p.trA[base + tq];            // write-through of the actual value
        if (idx < p.T) return p.trF[base + idx];
        return (tq >= p.T - 1) ? p.trA[base + p.T - 1] : p.trF[base + p.T - 1];
    }
    const int h = j - 20;                                           // get_power_limits  transformer.py:142-171
    float v = p.tr_limit[(size_t)s * p.Tr + k];
    const int nd = p.dr_count[(size_t)s * p.Tr + k];
    for (int ev = 0; ev < nd; ++ev) {
        const DrEv d = p.dr[((size_t)s * p.Tr + k) * p.n_dr + ev];
        if (tq + p.dr_steps_ahead >= d.start && d.end >= tq) {
            int a, b;
            if (tq > d.start) { a = 0; b = d.end - tq; }
            else { a = abs(d.start - tq); b = abs(d.end - tq); }
            if (h >= a && h < b) v = d.value;
        }
    }
    return v;
}

// ---- the fused step kernel --------------------------------------------------------------------
__device__ __forceinline__ const CsStatic &cs_of(const Params &p, int c) { return p.cs_uniform ? p.cs0 : p.cs[c]; }

// One port of one charger: EV.step + departure + arrival + potential + obs tuple.
// Accumulates into the charger's partial sums (acc) in the reference's port order.
struct ChargerAcc { double P, A, profit, satexp, pot, ch, dis, sat; int cnt; bool overflow; };

template <typename ActT>
__device__ __forceinline__ void process_port(const Params &p, const CsStatic &cs, const int t, const int s,
                                             const size_t ip, const int port, uint4 h, double a, const double sum,
                                             const double cp, const double dp, ChargerAcc &acc, float *obs_row,
                                             const bool want_obs) {
    const int tq = t + 1;
    bool occ = hot_t_arr(h) <= t && t <= hot_t_dep(h);
    if (!occ) a = 0.0;
    if (sum > 1.0) a = a / sum; else if (sum < -1.0) a = -a / sum;                    // ev_charger.py:143-149
    const double action = (a == 0.0) ? 0.0 : rint(a * 100000.0) / 100000.0;           // round(action, 5)  :157
    const EvSpec *sp = p.spec + hot_spec(h);
    double capv = 0.0, energy = 0.0;
    float exch_new = 0.f;
    bool exch_valid = false;
    if (occ) capv = p.cap[ip];
    if (occ && action != 0.0) {
        double amps, act_amps = 0.0;
        const double veff_cs = cs.veff[cs.phases];
        if (action > 0.0) {                                                            // :167-170
            amps = action * cs.imax;
            if (amps < cs.imin - 0.01) amps = 0.0;
            const double pmin = __ldg(&sp->pmin_ac);
            if (amps > 0.0 && pmin != 0.0 && amps < pmin * 1000.0 / veff_cs) amps = 0.0;   // ev.py:151-152
        } else {                                                                       // :183-186
            amps = action * cs.imax_dis_abs;
            if (amps > cs.imin_dis - 0.01) amps = cs.imin_dis;
            const double pmin = __ldg(&sp->pmin_dis);
            if (amps > 0.0) {                                                          // only if imin_dis > 0
                const double pm = __ldg(&sp->pmin_ac);
                if (amps < pm * 1000.0 / veff_cs) amps = 0.0;
            } else if (amps < 0.0 && pmin != 0.0 && amps > pmin * 1000.0 / veff_cs) amps = 0.0;  // ev.py:153-154
        }
        if (amps != 0.0) {
            const int evph = __ldg(&sp->ev_phases);
            const double veff = cs.veff[cs.phases < evph ? cs.phases : evph];            // ev.py:169
            const int lut = __ldg(&sp->lut);
            const double B = __ldg(&sp->B);
            if (amps > 0.0) {                                                          // EV._charge  ev.py:240-355
                double eta;
                const unsigned tsm = h.z >> 16, ecm = h.w & 0xFFFFu;
                if (lut >= 0) eta = lut_get(p.luts_c + (size_t)lut * p.lut_len, p.lut_len, rint(amps)) / 100.0;
                else eta = (ecm == 0xFFFFu) ? __ldg(&sp->eta_c) : (double)ecm / 1000.0;
                const double ts = (tsm == 0xFFFFu) ? __ldg(&sp->ts) : (double)tsm / 1000.0;
                const double pmax = __ldg(&sp->pmax_ac);
                double pilot = eta * amps * veff / 1000.0 / B / p.c60;                 // :295-296
                const double maxd = eta * pmax / B / p.c60;                            // :297-298
                if (pilot > maxd) pilot = maxd;                                        // :300-301
                const double soc = capv / B;
                double curr;
                if (ts == 1.0) {                                                       // :303-306
                    curr = pilot + soc;
                    if (curr > 1.0) curr = 1.0;
                } else {
                    const double pts = ts + (pilot - maxd) / maxd * (ts - 1.0);        // :312-314
                    double nsoc;
                    if (soc < pts) {
                        if (1.0 <= (pts - soc) / pilot) nsoc = pilot + soc;            // :323-324
                        else nsoc = 1.0 + exp(__ldg(&sp->mult) * (pilot + soc - pts) / (pts - 1.0)) * (pts - 1.0);
                    } else {
                        nsoc = 1.0 + exp(__ldg(&sp->mult) * pilot / (pts - 1.0)) * (soc - 1.0);   // :332-334
                    }
                    const double lim = (maxd > pilot) ? pilot : maxd;                  // :336-339
                    curr = (nsoc - soc > lim) ? lim + soc : nsoc;                      // :341-344
                }
                capv = curr * B;                                                       // :348
                energy = (curr - soc) * B;                                             // :346,352
                act_amps = energy / p.p60 * 1000.0 / veff;                             // :355
            } else {                                                                   // EV._discharge  ev.py:357-405
                double eta;
                const unsigned edm = h.w >> 16;
                if (lut >= 0) eta = lut_get(p.luts_d + (size_t)lut * p.lut_len, p.lut_len, fabs(rint(amps))) / 100.0;
                else eta = (edm == 0xFFFFu) ? __ldg(&sp->eta_d) : (double)edm / 1000.0;
                const double pmd = __ldg(&sp->pmax_dis), bmin = __ldg(&sp->bmin);
                double given_power = amps * veff / 1000.0;                             // :367
                if (fabs(given_power) > fabs(pmd)) given_power = pmd;                  // :370-371
                double given_energy = given_power * eta * p.period / 60.0;             // :381
                if (capv + given_energy < bmin) {                                      // :382-393
                    if (capv > bmin) { energy = -(capv - bmin); given_energy = energy; }
                    else { energy = 0.0; given_energy = 0.0; }
                    capv = bmin;
                } else {
                    energy = given_energy;
                    capv += given_energy;
                }
                act_amps = given_energy * 60.0 / p.period * 1000.0 / veff;             // :405
            }
            capv = ceil(capv * 100.0) / 100.0;                                         // my_ceil  ev.py:183
            p.cap[ip] = capv;
            exch_new = p.exch[ip] + (float)energy;                                     // total_energy_exchanged ev.py:178
            exch_valid = true;
            p.exch[ip] = exch_new;
        }
        const double ae = fabs(energy);
        if (action > 0.0) { acc.profit += ae * cp; acc.ch += ae; }                     // :178-179
        else              { acc.profit += ae * dp; acc.dis += ae; }                    // :194-195
        acc.P += energy * 60.0 / p.period;                                             // :180,196
        acc.A += act_amps;                                                             // :181,197
    }
    if (acc.A - 0.0001 > cs.imax) acc.overflow = true;                                 // :203-205
    if (p.out.port_energy) p.out.port_energy[ip] = (float)energy;

    // departure (charger step counter == t)        ev_charger.py:209-224, ev.py:199-214
    float dsat = __int_as_float(0x7fc00000);
    if (occ && t >= hot_t_dep(h)) {
        const double des = __ldg(&sp->desired);
        const double sat = (capv < des - 0.001) ? capv / des : 1.0;
        acc.satexp += 100.0 * exp(-10.0 * sat);                                        // reward.py:42,85
        acc.sat += sat;
        acc.cnt += 1 << 10;
        dsat = (float)sat;
    }
    if (p.out.dep_sat) p.out.dep_sat[ip] = dsat;

    // arrival of the next session at t+1            ev2gym_env.py:399-417, ev_charger.py:266-285
    if (hot_next_arr(h) == tq) {
        const SessRec r = p.sess[((size_t)s * p.P + port) * p.Smax + hot_cursor(h)];
        h = r.hot;
        capv = r.cap0;
        p.hot[ip] = h;
        p.cap[ip] = capv;
        p.exch[ip] = 0.f;
        exch_new = 0.f; exch_valid = true;
        sp = p.spec + hot_spec(h);
        acc.cnt += 1 << 20;
    }
    const bool occ_after = hot_t_arr(h) <= tq && tq <= hot_t_dep(h);
    if (p.out.action_mask) p.out.action_mask[ip] = occ_after ? 1 : 0;                 // ev2gym_env.py:452-457
    if (occ_after) {
        const double B = __ldg(&sp->B);
        // charge power potential for step t+1            utils.py:766-777
        if (capv < B && hot_t_dep(h) > tq) acc.pot += __ldg(&p.pot_kw[hot_spec(h) * p.n_cls + cs.cls]);
        if (want_obs) {       // observation tuple (transformer-major slot)   state.py:37-57, 85-102, 137-151
            float *o = obs_row + p.obs_slot[port];
            if (p.state_kind == EV2B_STATE_PUBLIC_PST) {
                o[0] = (capv == B) ? 1.f : 0.5f;
                o[1] = exch_valid ? exch_new : p.exch[ip];
                o[2] = (float)(tq - hot_t_arr(h));
            } else {
                o[0] = (float)(capv / B);
                o[1] = (float)(hot_t_dep(h) - tq);
            }
        }
    } else if (want_obs) {
        float *o = obs_row + p.obs_slot[port];
        o[0] = 0.f; o[1] = 0.f;
        if (p.state_kind == EV2B_STATE_PUBLIC_PST) o[2] = 0.f;
    }
}

// ActT: float or double actions.  NP: ports per charger when uniform (1, 2), 0 = read from CsStatic.
template <typename ActT, int NP, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) step_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x;
    double *red   = reinterpret_cast<double *>(smem_raw);                 // [kNRed][NT]
    double *trov  = red + (size_t)kNRed * NT;                             // [EPB*Tr] overload per transformer
    double *envs  = trov + (size_t)p.EPB * p.Tr;                          // [EPB][kNRed]
    int    *cnt   = reinterpret_cast<int *>(envs + (size_t)p.EPB * kNRed);// [NT] invalid | dep<<10 | arr<<20
    int    *envi  = cnt + NT;                                             // [EPB][4] t, scn, cnt, flags
    float  *obs_s = reinterpret_cast<float *>(envi + (size_t)p.EPB * 4);  // [EPB][D]

    const int tid = threadIdx.x;
    const int el = p.C == 1 ? tid : (int)__umulhi((unsigned)tid, p.c_magic);   // tid / C
    const int c = tid - el * p.C;
    const int e = blockIdx.x * p.EPB + el;
    const bool valid = (el < p.EPB) && (e < p.E);
    const ActT *actions = reinterpret_cast<const ActT *>(p.actions);
    const bool want_obs = (p.out.obs != nullptr) && (p.state_kind != EV2B_STATE_NONE);

    ChargerAcc acc;
    acc.P = acc.A = acc.profit = acc.satexp = acc.pot = acc.ch = acc.dis = acc.sat = 0.0;
    acc.cnt = 0; acc.overflow = false;
    int t = 0, s = 0;
    bool live = false;
    if (valid) {
        t = p.env_step[e];
        s = p.env_scn[e];
        live = t < p.T;
        if (c == 0) { envi[el * 4 + 0] = t; envi[el * 4 + 1] = s; envi[el * 4 + 3] = 0; }
    }

    if (valid && live) {
        const CsStatic &cs = cs_of(p, c);
        const int port0 = p.cs_uniform ? c * (NP > 0 ? NP : p.cs0.n_ports) : cs.port_off;
        const int n = NP > 0 ? NP : cs.n_ports;
        const size_t pbase = (size_t)e * p.P + port0;
        float *obs_row = obs_s + (size_t)el * p.D;
        if (NP > 0) {
            // issue every independent load of this charger up front
            uint4 h[NP > 0 ? NP : 1];
            double a[NP > 0 ? NP : 1];
#pragma unroll
            for (int j = 0; j < NP; ++j) { h[j] = p.hot[pbase + j]; a[j] = (double)actions[pbase + j]; }
            const EnvT et = p.env_t[(size_t)s * p.T + t];
            double sum = 0.0;                                       // python sum(): left to right  ev_charger.py:143
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const bool occ = hot_t_arr(h[j]) <= t && t <= hot_t_dep(h[j]);
                acc.cnt += occ ? 0 : 1;                             // invalid_action_punishment  :137-140
                sum = sum + (occ ? a[j] : 0.0);
            }
#pragma unroll
            for (int j = 0; j < NP; ++j)
                process_port<ActT>(p, cs, t, s, pbase + j, port0 + j, h[j], a[j], sum, et.cp, et.dp, acc, obs_row, want_obs);
        } else {
            const EnvT et = p.env_t[(size_t)s * p.T + t];
            double sum = 0.0;
            for (int j = 0; j < n; ++j) {
                const uint4 h = p.hot[pbase + j];
                const bool occ = hot_t_arr(h) <= t && t <= hot_t_dep(h);
                acc.cnt += occ ? 0 : 1;
                sum = sum + (occ ? (double)actions[pbase + j] : 0.0);
            }
            for (int j = 0; j < n; ++j)
                process_port<ActT>(p, cs, t, s, pbase + j, port0 + j, p.hot[pbase + j], (double)actions[pbase + j], sum,
                                   et.cp, et.dp, acc, obs_row, want_obs);
        }
        // clamp the charger's potential                      utils.py:779-789
        if (acc.pot > cs.max_power) acc.pot = cs.max_power;
        else if (acc.pot < cs.min_power) acc.pot = 0.0;
        if (acc.overflow) atomicOr(&envi[el * 4 + 3], (int)EV2B_ST_AMPS_OVERFLOW);
        if (p.out.cs_power)   p.out.cs_power[(size_t)e * p.C + c] = (float)acc.P;
        if (p.out.cs_current) p.out.cs_current[(size_t)e * p.C + c] = (float)acc.A;
    }
    red[RedP * NT + tid] = acc.P;             red[RedProfit * NT + tid] = acc.profit;
    red[RedSatExp * NT + tid] = acc.satexp;   red[RedPot * NT + tid] = acc.pot;
    red[RedCharged * NT + tid] = acc.ch;      red[RedDischarged * NT + tid] = acc.dis;
    red[RedSatSum * NT + tid] = acc.sat;
    cnt[tid] = acc.cnt;
    __syncthreads();

    // ---- phase B: fixed-order reductions (warp per job) + series part of the observation -------
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = NT >> 5;
        const int njobs = p.EPB * (p.Tr + 1);
        for (int job = warp; job < njobs; job += nwarps) {
            const int jel = job / (p.Tr + 1), k = job - jel * (p.Tr + 1);
            const int je = blockIdx.x * p.EPB + jel;
            if (je >= p.E) continue;
            const int jt = envi[jel * 4 + 0], js = envi[jel * 4 + 1];
            if (jt >= p.T) continue;
            if (k < p.Tr) {       // transformer k: Transformer.step accumulation   transformer.py:269-274
                double sp_ = 0;
                for (int i = p.tr_cs_off[k] + lane; i < p.tr_cs_off[k + 1]; i += 32)
                    sp_ += red[RedP * NT + jel * p.C + p.tr_cs_idx[i]];
                sp_ = warp_sum(sp_);
                if (lane == 0) {
                    const TrT tt = p.tr_t[((size_t)js * p.T + jt) * p.Tr + k];
                    const double base = tt.infl + tt.solar;                       // transformer.py:264-265
                    const double ptot = base + sp_;
                    double ov = 0.0;                                              // transformer.py:284-302
                    if (ptot > tt.maxp + 0.0001 || ptot < tt.minp - 0.0001) ov = fabs(ptot - tt.maxp);
                    trov[jel * p.Tr + k] = ov;
                    if (p.out.tr_power)    p.out.tr_power[(size_t)je * p.Tr + k] = ptot;
                    if (p.out.tr_overload) p.out.tr_overload[(size_t)je * p.Tr + k] = ov;
                }
            } else {              // env-level sums
                double v[kNRed];
#pragma unroll
                for (int q = 0; q < kNRed; ++q) v[q] = 0.0;
                int ci = 0, cd = 0, ca = 0;
                for (int i = lane; i < p.C; i += 32) {
                    const int cc = jel * p.C + i;
#pragma unroll
                    for (int q = 0; q < kNRed; ++q) if (q != RedA) v[q] += red[q * NT + cc];
                    const int w = cnt[cc];
                    ci += w & 1023; cd += (w >> 10) & 1023; ca += (w >> 20) & 1023;
                }
#pragma unroll
                for (int q = 0; q < kNRed; ++q) if (q != RedA) v[q] = warp_sum(v[q]);
                ci = warp_sum_i(ci); cd = warp_sum_i(cd); ca = warp_sum_i(ca);
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < kNRed; ++q) envs[jel * kNRed + q] = v[q];
                    envi[jel * 4 + 2] = ci | (cd << 10) | (ca << 20);
                }
            }
        }
        if (want_obs && p.state_kind != EV2B_STATE_PUBLIC_PST) {
            const int per_env = 20 + (p.state_kind == EV2B_STATE_V2G_PROFIT_MAX_LOADS ? p.Tr * 40 : 0);
            for (int jel = 0; jel < p.EPB; ++jel) {
                const int je = blockIdx.x * p.EPB + jel;
                if (je >= p.E) break;
                const int jt = envi[jel * 4 + 0];
                if (jt >= p.T) continue;
                const int js = envi[jel * 4 + 1];
                for (int i = tid; i < per_env; i += NT) {
                    int off;
                    const float v = obs_series_value(p, js, jt + 1, i, &off);
                    obs_s[(size_t)jel * p.D + off] = v;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase C: one thread per env: reward, KPIs, step counter -------------------------------
    if (tid < p.EPB) {
        const int je = blockIdx.x * p.EPB + tid;
        if (je < p.E) {
            const int jt = envi[tid * 4 + 0], js = envi[tid * 4 + 1];
            unsigned status = (unsigned)envi[tid * 4 + 3];
            double reward = 0.0, costs = 0.0;
            if (jt < p.T) {
                const double *v = envs + tid * kNRed;
                const int w = envi[tid * 4 + 2];
                const EnvT et = p.env_t[(size_t)js * p.T + jt];
                const double usage = v[RedP];                                  // current_power_usage[t]  ev2gym_env.py:375
                costs = v[RedProfit];
                double ovsum = 0.0;
                for (int k = 0; k < p.Tr; ++k) ovsum += trov[tid * p.Tr + k];
                if (p.reward_kind == EV2B_REWARD_SQ_TRACKING) {                // reward.py:11-12
                    const double pot = p.env_pot[je];
                    const double m = et.setpoint < pot ? et.setpoint : pot;
                    reward = -((m - usage) * (m - usage));
                } else if (p.reward_kind == EV2B_REWARD_PROFIT_TR_USER) {      // reward.py:36-44
                    reward = costs - 100.0 * ovsum - v[RedSatExp];
                } else if (p.reward_kind == EV2B_REWARD_PROFIT_MAX) {          // reward.py:81-87
                    reward = costs - v[RedSatExp];
                }
                double *kpi = p.env_kpi + (size_t)je * EV2B_KPI_COUNT;
                kpi[EV2B_KPI_TOTAL_REWARD] += reward;
                kpi[EV2B_KPI_TOTAL_PROFITS] += costs;
                kpi[EV2B_KPI_ENERGY_CHARGED] += v[RedCharged];
                kpi[EV2B_KPI_ENERGY_DISCHARGED] += v[RedDischarged];
                kpi[EV2B_KPI_TR_OVERLOAD] += ovsum;
                kpi[EV2B_KPI_EVS_SERVED] += (double)((w >> 10) & 1023);
                kpi[EV2B_KPI_SAT_SUM] += v[RedSatSum];
                const double d = et.setpoint - usage;                          // utils.py:37-44
                kpi[EV2B_KPI_TRACKING_ERROR] += d * d;
                kpi[EV2B_KPI_ENERGY_TRACKING_ERROR] += fabs(d);
                if (usage > et.setpoint) kpi[EV2B_KPI_TRACKER_VIOLATION] += usage - et.setpoint;
                kpi[EV2B_KPI_EVS_SPAWNED] += (double)((w >> 20) & 1023);
                kpi[EV2B_KPI_INVALID_ACTIONS] += (double)(w & 1023);
                kpi[EV2B_KPI_STEPS] += 1.0;
                p.env_pot[je] = (jt + 1 < p.T) ? v[RedPot] : 0.0;              // ev2gym_env.py:424-426
                p.env_usage[je] = usage;
                p.env_step[je] = jt + 1;
                if (jt + 1 >= p.T) status |= EV2B_ST_DONE;                     // ev2gym_env.py:460
                if (want_obs) obs_header(p, obs_s + (size_t)tid * p.D, js, jt + 1, usage);
            } else {
                status |= EV2B_ST_DONE | EV2B_ST_WAS_DONE;                     // ev2gym_env.py:343
            }
            if (p.out.reward) p.out.reward[je] = reward;
            if (p.out.total_costs) p.out.total_costs[je] = costs;
            if (p.out.status) p.out.status[je] = status;
        }
    }
    if (want_obs) {
        __syncthreads();
        for (int jel = 0; jel < p.EPB; ++jel) {
            const int je = blockIdx.x * p.EPB + jel;
            if (je >= p.E) break;
            if (envi[jel * 4 + 0] >= p.T) continue;
            float *dst = p.out.obs + (size_t)je * p.D;
            const float *src = obs_s + (size_t)jel * p.D;
            for (int i = tid; i < p.D; i += NT) dst[i] = src[i];
        }
    }
}

// ---- reset ------------------------------------------------------------------------------------
// Per-port part of reset(): every port empty, cursor at its first session.   ev2gym_env.py:298-306
// mode 0: envs [lo,hi) take scn_ids[e-lo] (or e mod S); mode 1: only finished envs, next scenario.
__global__ void reset_ports_kernel(const Params p, int lo, int hi, const int *scn_ids, int mode) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)(hi - lo) * p.P;
    if (i >= n) return;
    const int e = lo + (int)(i / p.P), port = (int)(i % p.P);
    int s;
    if (mode == 1) {
        if (p.env_step[e] < p.T) return;
        s = (p.env_scn[e] + p.E) % p.S;
    } else {
        s = scn_ids ? scn_ids[e - lo] : e % p.S;
    }
    const SessRec r0 = p.sess[((size_t)s * p.P + port) * p.Smax];
    uint4 h;
    h.x = ((unsigned)kNoArrival & 0xFFFFu) | (0xFFFFu << 16);        // t_arr = 32767, t_dep = -1
    h.y = (r0.hot.x & 0xFFFFu);                                      // next_arr = first session's t_arr, cursor 0
    h.z = 0; h.w = 0;
    p.hot[(size_t)e * p.P + port] = h;
    p.cap[(size_t)e * p.P + port] = 0.0;
    p.exch[(size_t)e * p.P + port] = 0.f;
}

// Per-env part of reset() + first observation.  Must run AFTER reset_ports_kernel (same stream).
__global__ void reset_envs_kernel(const Params p, int lo, int hi, const int *scn_ids, int mode, float *obs0) {
    const int e = lo + blockIdx.x;
    if (e >= hi) return;
    __shared__ int sh_s, sh_go;
    if (threadIdx.x == 0) {
        int go = 1, s;
        if (mode == 1) {
            go = p.env_step[e] >= p.T;
            s = (p.env_scn[e] + p.E) % p.S;
        } else {
            s = scn_ids ? scn_ids[e - lo] : e % p.S;
        }
        sh_s = s; sh_go = go;
    }
    __syncthreads();
    if (!sh_go) return;
    const int s = sh_s;
    if (obs0 && p.state_kind != EV2B_STATE_NONE) {
        float *row = obs0 + (size_t)e * p.D;
        for (int i = threadIdx.x; i < p.D; i += blockDim.x) row[i] = 0.f;   // no EV is connected at t = 0
        __syncthreads();
        if (threadIdx.x == 0) obs_header(p, row, s, 0, 0.0);
        if (p.state_kind != EV2B_STATE_PUBLIC_PST) {
            const int per_env = 20 + (p.state_kind == EV2B_STATE_V2G_PROFIT_MAX_LOADS ? p.Tr * 40 : 0);
            for (int i = threadIdx.x; i < per_env; i += blockDim.x) {
                int off;
                const float v = obs_series_value(p, s, 0, i, &off);
                row[off] = v;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.env_step[e] = 0; p.env_scn[e] = s; p.env_pot[e] = 0.0; p.env_usage[e] = 0.0;
        for (int k = 0; k < EV2B_KPI_COUNT; ++k) p.env_kpi[(size_t)e * EV2B_KPI_COUNT + k] = 0.0;
    }
}

}  // namespace ev2b
