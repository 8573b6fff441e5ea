// ev2b_device.cuh -- device data layout and kernels of the batched EV2Gym step engine (sm_100a).
//
// What this replaces (paths relative to /root/reference): the Python per-object cascade
//   EV2Gym.step (ev2gym/models/ev2gym_env.py:333-447) -> EV_Charger.step (ev_charger.py:114-233)
//   -> EV.step/_charge/_discharge (ev.py:138-186, 240-355, 357-405) -> Transformer.reset/step/
//   get_how_overloaded (transformer.py:258-302), the spawn loop (ev2gym_env.py:399-417),
//   calculate_charge_power_potential (utils.py:760-791), the three stock rewards (reward.py) and the
//   three stock state functions (state.py), for E env replicas at once.
//
// Kernel structure (one CTA owns EPB whole envs, so every reduction stays inside the CTA):
//   A1  thread per (env, charger): coalesced loads of the per-port state, empty-port masking,
//       action normalisation; ports that will actually move energy are COMPACTED into a shared
//       memory work list (mid-episode only a few % of ports are active: running the float64
//       battery model per lane would leave ~30 of 32 lanes idle).
//   A2  thread per work item (dense warps): EV.step -- the float64 battery model.
//   A3  thread per (env, charger): charger accounting in port order, departures, arrivals,
//       potential, observation tuples; coalesced stores of the updated state.
//   B   warp per (env, transformer) / per env: fixed-order shared-memory + shuffle reductions.
//   C   thread per env: reward, KPI sums, step counter.   D: coalesced observation rows.
// Arithmetic is IEEE float64 in the reference's operation order; this translation unit is
// compiled with -fmad=false so no multiply-add is contracted (the only FMAs are the explicit ones
// of the exact constant division, ev2b_math.h).  No tensor cores: elementwise + segmented reduce.
#pragma once
#ifdef EV2B_SIMT_EMU
// tests/simt_emu: the same sources compiled by g++ and run on a fiber-based SIMT emulator (test infrastructure only)
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#define EV2B_NOINLINE __noinline__
#define EV2B_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define EV2B_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#include <stdint.h>
#include "../../include/ev2b.h"
#include "ev2b_math.h"

namespace ev2b {

// optional per-step outputs (everything but reward / status / obs): the OPTOUT = false instantiation of the step kernel,
// launched when the caller asked for none of them, has no code for them at all
#define EV2B_OPT(ptr) (OPTOUT && (ptr))

constexpr int   kNoArrival = 32767;   // "no (further) session on this port"
constexpr int   kMaxThreads = 1024;
constexpr int   kRrAbsent = 0x7fffffff;
// Per-env records every step needs late (phase B / C) are fetched into shared memory with cp.async before the first
// barrier, so the single-thread tail of a CTA never waits on a global load:
//   [0,13) KPI sums   [13] charge_power_potential[t]   [14] setpoint[t]   [15] setpoint[t+1]   [16 + 4k, +4) TrT of transformer k
constexpr int   kPrePot = 13, kPreSet = 14, kPreSetNext = 15, kPreTr = 16;
constexpr int   kNRed = 7;            // float64 partials per charger, see Red* below
enum { RedP = 0, RedProfit, RedSatExp, RedPot, RedCharged, RedDischarged, RedSatSum };

// ---- static tables -------------------------------------------------------------------------
struct CsStatic {            // one per charger; ev_charger.py:41-75 + derived constants
    double imax, imin, imax_dis_abs, imin_dis;
    double veff[4];          // voltage*sqrt(phases) for phases = 1..3 (index 0 unused)  ev.py:279,365
    double rveff[4];         // RN(1/veff[k])
    double max_power;        // sqrt(ph)*V*Imax/1000   utils.py:779-780
    double min_power;        // sqrt(ph)*V*Imin/1000   utils.py:781-782
    int    port_off, n_ports, tr, phases;
    int    cls, pad0, pad1, pad2;   // charger class (distinct imax/voltage/phases) for the potential table
    double calap_kw;         // imax*V*sqrt(ph)/1000 in the operation order of heuristics.py:120-121
    double pad3;
};

struct alignas(16) EvSpec {  // de-duplicated EV model; ev.py:45-113
    // 16-byte pairs: EV.step loads a pair with one instruction, and picks the pair of the action's direction by address
    // ({pmin, pmax} and {ts, mult} when charging, {pmin_dis, pmax_dis} and {bmin, bmin_em} when discharging), so the loads
    // of a step are all in flight together instead of one dependent cache round trip per field.
    double B, rB;            // rB = RN(1/B)
    double pmin_ac, pmax_ac;
    double pmin_dis, pmax_dis;
    double ts, mult;         // ts, eta_c, eta_d: used when the per-session milli encodings are 0xFFFF
    double bmin, bmin_em;
    double eta_c, eta_d;
    double desired, pad0;
    int    ev_phases, lut;   // lut < 0: scalar efficiencies
    double pad1;
};
static_assert(sizeof(EvSpec) == 128, "EvSpec must be 128 B");
static_assert(offsetof(EvSpec, pmin_dis) == offsetof(EvSpec, pmin_ac) + 16 && offsetof(EvSpec, bmin) == offsetof(EvSpec, ts) + 16,
              "the discharging pair follows the charging pair");

// "hot" words of the session currently (or last) connected to a port: 16 B, read every step.
//   w.x = t_arr (i16) | t_dep (i16) << 16
//   w.y = next_arr (i16) | cursor (u8) << 16        cursor = index of the NEXT session of this port
//   w.z = spec_id (u16) | ts_milli (u16) << 16      ts = k/1000.0 (np.round(.,3), utils.py:309) or 0xFFFF
//   w.w = eta_c_milli (u16) | eta_d_milli (u16) << 16
// A port is occupied at step t iff t_arr <= t <= t_dep (the EV is charged in step t_dep and then
// leaves, ev_charger.py:209-224).
struct SessRec { uint4 hot; double cap0; double afap; };  // cold table, read once per arrival; afap: ev.py:407-440
static_assert(sizeof(SessRec) == 32, "SessRec must be 32 B");

struct EnvT { double cp, dp, setpoint; int arr0, n_arr; };  // per (scenario, t); the sessions arriving at step t+1 are
                                                            // arr_list[arr0 .. arr0 + n_arr)  (ev2b_evlist.cuh)
static_assert(sizeof(EnvT) == 32, "EnvT must be 32 B");
struct TrT  { double infl, solar, maxp, minp; };          // per (scenario, t, transformer)
struct DrEv { int16_t start, end; float value; };         // value = limit - limit*cap/100  transformer.py:158-163

struct Params {
    // sizes
    int E, C, P, Tr, T, D, EPB, n_dr, lut_len, Smax, S, n_cls, W;
    int reward_kind, state_kind, dr_steps_ahead;
    int agent_kind;          // enum ev2b_agent_kind; != 0: actions are generated, `actions` is not read
    unsigned agent_seed_lo, agent_seed_hi; double action_low;
    int env0, env_end;       // this launch advances envs [env0, env_end)  (ev2b_step_host pipelines chunks)
    int obs_full;            // 1: rewrite every observation entry; 0: the caller's obs buffer still holds last step's
                             //    rows, only entries that can change are written (occupied ports, header, series)
    double c60, rc60;        // 60 / timescale (ev.py:296) and its reciprocal
    double p60, rp60;        // timescale / 60 (ev.py:355)
    double period, rperiod;  // timescale
    unsigned c_magic;        // ceil(2^32 / C): tid / C == __umulhi(tid, c_magic)
    unsigned p_magic;        // ceil(2^32 / P)
    int cs_uniform;          // all chargers identical: read the layout from cs0 (constant bank)
    // byte offsets of the step kernel's shared-memory arrays (computed once on the host: ev2b_handle::layout_smem)
    int o_resE, o_resA, o_resC, o_trov, o_trp, o_lossv, o_pfv, o_envs, o_whot, o_cnt, o_envi, o_wl, o_wcnt, o_pflag, o_pre;
    int pre_stride;          // doubles per env in the prefetch area: kPreTr + 4 * Tr
    CsStatic cs0;
    // static
    const CsStatic *cs; const int *cs_tr; const int *port_cs; const int *tr_cs_off; const int *tr_cs_idx; const int *obs_slot;
    const int *tr_obs_off; const int *series_off;
    // scenario bank
    const EnvT *env_t; const TrT *tr_t; const SessRec *sess; const EvSpec *spec;
    const double *luts_c, *luts_d; const double *pot_kw;
    const float *trA, *trF, *tr_limit; const DrEv *dr; const uint8_t *dr_count;
    const float *obs_static;   // [S][T+1][W] precomputed price window + forecast/limit blocks, or null
    // distribution grid: Laurent power flow  grid.py:120-141, numbarize.py:268-325
    int n_bus; double s_base;
    const double2 *grid_Kt;    // [n_bus][n_bus] K transposed (Kt[c][r] = K[r][c]): lanes read consecutive rows
    const double2 *grid_L;     // [n_bus]
    const double *grid_act, *grid_rea;   // [S][T+1][n_bus]
    const double *date_feat;   // [S][T+1][3]
    // statistics mode (EV2B_F_STATS): per-EV histories for get_statistics  utils.py:12-123, ev.py:442-521
    int stats, L;              // L: trace slots per port (longest session + 2)
    double k_cal, k_exp, k_cyc; // 0.75/730^0.25, exp(-e2/theta), 0.5*(2.05/78)/sqrt(Q_acc)
    double *st_soc_sum, *st_abs_e, *st_act, *st_r;   // [E,P], [E,P], [E,P,L], [E,P,Smax]
    int *st_cnt, *st_nfin;     // [E,P] n_hist | n_act << 16 ; finalised sessions on the port
    double *cs_sat_sum, *cs_dcal, *cs_dcyc; int *cs_served, *cs_em;   // [E,C]
    // RoundRobin agent queue (heuristics.py:29,33-52): sort key per port (kRrAbsent = not queued) + front/back counters
    int *rr_key; int *rr_fb;   // [E,P], [E,2]; null until the agent is first used
    double rr_avg_power, rr_share;   // heuristics.py:19-23 ; 1 / number_of_ports_per_cs
    // event-driven step kernel (ev2b_evlist.cuh): connected-EV list per env, arrival schedule per scenario, smem map
    uint16_t *occ_list;        // [E][P] ports holding an EV (first occ_n[e] entries); null: kernel not in use
    int *occ_n;                // [E]
    const unsigned *arr_list;  // port | session index on that port << 16, arrival-sorted; bucket of a step: see EnvT
    int tr_lg;                 // 2^tr_lg lanes share one transformer in the CSR sums (largest with 2^tr_lg * Tr <= 32)
    int v_stride, v_amp, v_pot, v_csP, v_pre, v_wsum, v_stage, v_occ;   // byte offsets inside one env's block
    int v_trp, v_pfv, v_dsat, v_dcal, v_dcyc;   // HEAVY instantiation only (grid / statistics mode)
    int v_hdr;                 // 16 B: env_step / env_scn / occ_n published by warp 0 (whole-CTA groups)
    int mask_full;             // 1: out.action_mask does not hold last step's rows (evl_step_kernel rewrites them)
    int scn_stride;            // auto-reset: next scenario = (current + scn_stride) mod S, gcd(scn_stride, S) = 1
    int k_steps, auto_reset;   // evl_step_kernel<..., KSTEP = true>: steps per launch, device-side reset of finished envs
    int series_pairs;          // 1: W is even, series values 2j / 2j+1 land on neighbouring floats at an even offset and D is
                               //    even: the (scenario, time) observation values are copied as float2
    int obs_pairs;             // 1: an EV's observation tuple is two floats at an even offset of an 8-byte aligned row: one store
    int act_pairs;             // 1: two ports per charger, caller-supplied actions on a 2-element boundary: one load per charger
    // state
    uint4 *hot; double *cap; double *exch;   // exch: float64 like the reference's total_energy_exchanged (ev.py:178)
     int *env_step; int *env_scn; double *env_pot; double *env_usage;
    double *env_pot_prev;      // charge_power_potential[t-1], kept only for SquaredTrackingErrorRewardWithPenalty (reward.py:50)
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

// Counter-based RNG of the UNIFORM device agent (documented in include/ev2b.h; mirrored by tests).
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <typename ActT>
__device__ __forceinline__ double agent_action(const Params &p, const ActT *actions, size_t ip, int t) {
    if (p.agent_kind == EV2B_AGENT_EXTERNAL) return (double)actions[ip];
    if (p.agent_kind == EV2B_AGENT_AFAP) return 1.0;
    if (p.agent_kind == EV2B_AGENT_ZERO) return 0.0;
    const unsigned hsh = mix32(mix32((unsigned)ip ^ p.agent_seed_lo) + (unsigned)t * 0x9E3779B9u + p.agent_seed_hi);
    return p.action_low + (1.0 - p.action_low) * ((double)(hsh >> 8) * (1.0 / 16777216.0));
}

__device__ __forceinline__ double lut_get(const double *lut, int lut_len, double key) {
    // dict.get(np.round(amps), 1): integer keys 0..lut_len-1, default 1   ev.py:288,376
    if (key >= 0.0 && key < (double)lut_len) return __ldg(lut + (int)key);
    return 1.0;
}

#ifdef EV2B_SIMT_EMU
__device__ __forceinline__ void prefetch_l2(const void *) {}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) { simt::cp_async(smem_dst, gmem_src, 8); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) { simt::cp_async(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async_wait_all() { simt::cp_async_wait_all(); }
#else
// HBM -> L2 without a destination register: the later load of the same sector finds it in L2.
__device__ __forceinline__ void prefetch_l2(const void *gmem) { asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem)); }
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

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

// Hides a value from the optimiser (it must materialise x and may assume nothing about it); nothing is emitted.
#if defined(__CUDA_ARCH__)
#define EV2B_OPAQUE_F64(x) asm volatile("" : "+d"(x))
#else
#define EV2B_OPAQUE_F64(x) ((void)0)
#endif

template <bool UNI>
__device__ __forceinline__ const CsStatic &cs_of(const Params &p, int c) {
    if (UNI) return p.cs0;          // kernel-parameter constant bank: fields become immediate operands
    return p.cs[c];
}

// ---- observation pieces shared by the step and reset kernels --------------------------------
// Header of the stock state functions at observation time tq (= current_step after the increment).
__device__ __forceinline__ void obs_header(const Params &p, float *row, int s, int tq, double prev_usage, double setpoint_tq) {
    if (p.state_kind == EV2B_STATE_V2G_GRID) {                      // state.py:247 (the rest of the header is static)
        row[5] = (float)prev_usage;
    } else if (p.state_kind == EV2B_STATE_PUBLIC_PST) {             // state.py:11-34
        row[0] = (float)((double)tq / (double)p.T);
        row[1] = (tq < p.T) ? (float)setpoint_tq : 0.f;        // power_setpoints[current_step]
        row[2] = (float)prev_usage;
    } else {                                                        // state.py:70-82, 113-125
        row[0] = (float)tq;
        row[1] = (float)prev_usage;
    }
}
// One value of the (scenario, time)-only part of the observation: i indexes the flat list
// [20 prices][Tr * (20 load-pv forecast + 20 power limits)].
// (out of line: the step kernels only call it when the precomputed table `obs_static` does not exist -- a bank too large
//  for it -- and inlined it was ~700 instructions at every call site)
__device__ EV2B_NOINLINE float obs_series_value(const Params &p, int s, int tq, int i) {
    if (p.state_kind == EV2B_STATE_V2G_GRID) {                      // V2G_grid_state  state.py:221-255
        if (i < 3) return (float)p.date_feat[((size_t)s * (p.T + 1) + tq) * 3 + i];
        if (i == 3) return (tq < p.T) ? (float)p.env_t[(size_t)s * p.T + tq].cp : 0.f;        // charge_prices[0, t:t+1]
        if (i == 4) return (tq < p.T) ? (float)p.env_t[(size_t)s * p.T + tq].setpoint : 0.f;
        const int j = i - 5;   // node_active_power[1:, step-1] == base powers of step `step` (grid.step returns the next step's)
        const size_t base = ((size_t)s * (p.T + 1) + tq) * p.n_bus;
        return (float)(j < p.n_bus ? p.grid_act[base + j] : p.grid_rea[base + j - p.n_bus]);
    }
    if (i < 20)                                                     // abs(charge_prices[0, t:t+20]), zero padded
        return (tq + i < p.T) ? (float)fabs(p.env_t[(size_t)s * p.T + tq + i].cp) : 0.f;
    i -= 20;
    const int k = i / 40, j = i - k * 40;
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
__device__ __forceinline__ float obs_series_fetch(const Params &p, int s, int tq, int i) {
    if (p.obs_static) return __ldg(p.obs_static + ((size_t)s * (p.T + 1) + tq) * p.W + i);
    return obs_series_value(p, s, tq, i);
}
// Fills obs_static[s][tq][0..W) for every scenario and observation time (run once per bank).
__global__ void obs_static_kernel(const Params p, float *table) {
    const size_t n = (size_t)p.S * (p.T + 1) * p.W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % p.W);
        const size_t r = i / p.W;
        table[i] = obs_series_value(p, (int)(r / (p.T + 1)), (int)(r % (p.T + 1)), w);
    }
}

// Loads of the EV model's fields in EV.step: 1 = one batch up front, 0 = at their points of use (pairs), 2 = a batch only in
// the instantiations with registers to spare (statistics / grid); measured in profiles/r2_ab_spec_loads.jsonl.
#ifndef EV2B_SPEC_BATCH
#define EV2B_SPEC_BATCH 2
#endif

// ---- A2: EV.step for one work item ------------------------------------------------------------
// a = normalised action (non-zero), cap = battery level, hz/hw = hot words z/w of the session.
// Returns through (energy, act_amps, cap); result = the EV saw non-zero amps (ev.py:158).
template <bool STATS>
__device__ __forceinline__ bool ev_step_item(const Params &p, const CsStatic &cs, const unsigned hz, const unsigned hw,
                                             const double a, double &cap, double &energy, double &act_amps,
                                             bool &em_cross) {
    energy = 0.0; act_amps = 0.0; em_cross = false;
    const double action = ev2b_div_c(rint(a * 100000.0), 100000.0, 1e-5);          // round(action, 5)  ev_charger.py:157
    if (action == 0.0) return false;
    // Everything the step needs of the EV's model, requested here in one batch (four 16-byte / 8-byte loads in flight
    // together): the kernels are bound by the latency of dependent loads, and field-by-field loads at their points of
    // use cost one L1 / L2 round trip each (round 2, profiles/r2i_evl_lines_busiest.txt: 26 % of the long-scoreboard
    // stalls of evl_step_kernel sat on them).
    const EvSpec *sp = p.spec + (hz & 0xFFFFu);
    const bool pos = action > 0.0;
    constexpr bool BATCH = EV2B_SPEC_BATCH == 1 || (EV2B_SPEC_BATCH == 2 && STATS);
    double2 Bv, lim, par;
    int2 pl;
    if (BATCH) {
        Bv = __ldg(reinterpret_cast<const double2 *>(&sp->B));                                // B, 1/B
        pl = __ldg(reinterpret_cast<const int2 *>(&sp->ev_phases));                           // ev_phases, lut
        lim = __ldg(reinterpret_cast<const double2 *>(pos ? &sp->pmin_ac : &sp->pmin_dis));   // pmin, pmax of the direction
        par = __ldg(reinterpret_cast<const double2 *>(pos ? &sp->ts : &sp->bmin));            // (ts, mult) | (bmin, bmin_em)
    }
    const double veff_cs = cs.veff[cs.phases];
    double amps;
    if (pos) {                                                                     // ev_charger.py:167-170
        amps = action * cs.imax;
        if (amps < cs.imin - 0.01) amps = 0.0;
    } else {                                                                       // ev_charger.py:183-186
        amps = action * cs.imax_dis_abs;
        if (amps > cs.imin_dis - 0.01) amps = cs.imin_dis;
    }
    if (!BATCH) lim = __ldg(reinterpret_cast<const double2 *>(pos ? &sp->pmin_ac : &sp->pmin_dis));
    if (lim.x != 0.0) {                                                            // ev.py:151-154
        const double thr = lim.x * 1000.0 / veff_cs;
        if (amps > 0.0 ? amps < thr : (amps < 0.0 && amps > thr)) amps = 0.0;
    }
    if (amps == 0.0) return false;
    if (!BATCH) {
        Bv = __ldg(reinterpret_cast<const double2 *>(&sp->B));
        pl = __ldg(reinterpret_cast<const int2 *>(&sp->ev_phases));
        par = __ldg(reinterpret_cast<const double2 *>(pos ? &sp->ts : &sp->bmin));
    }
    const int evph = pl.x;
    const int ph = cs.phases < evph ? cs.phases : evph;                            // ev.py:169
    const double veff = cs.veff[ph], rveff = cs.rveff[ph];
    const int lut = pl.y;
    const double B = Bv.x, rB = Bv.y;
    // charge / discharge efficiency: a scalar, or dict.get(np.round(amps), 1) / 100 resp. dict.get(abs(np.round(amps)), 1) / 100
    // (ev.py:287-290, 375-378).  np.round(amps) >= 0 when charging, so one lookup with |round(amps)| serves both directions:
    // every lane of the warp does it once instead of the charging and the discharging lanes one after the other.
    const bool charging = amps > 0.0;
    double eta;
    {
        const unsigned em = charging ? (hw & 0xFFFFu) : (hw >> 16);
        if (lut >= 0) eta = ev2b_div_c(lut_get((charging ? p.luts_c : p.luts_d) + (size_t)lut * p.lut_len, p.lut_len, fabs(rint(amps))), 100.0, 0.01);
        else eta = (em == 0xFFFFu) ? __ldg(charging ? &sp->eta_c : &sp->eta_d) : ev2b_div_c((double)em, 1000.0, 0.001);
    }
    // (Round 2 also measured a formulation with the two directions as warp-uniform blocks of straight-line code -- selects
    //  instead of per-lane branches, a warp of one direction skipping the other block: bit-identical, 10 % SLOWER on the
    //  B200 (c3 26.2 vs 23.7 us, c4 46.7 vs 40.7): the extra live results of both blocks cost more registers than the
    //  removed reconvergence points saved.  profiles/r2_ab_model_formulation.jsonl)
    if (charging) {                                                                // EV._charge  ev.py:240-355
        const unsigned tsm = hz >> 16;
        const double ts = (tsm == 0xFFFFu) ? par.x : ev2b_div_c((double)tsm, 1000.0, 0.001);
        const double pmax = lim.y;
        // pilot_dsoc = eta*amps*voltage/1000/B/(60/period)                            :295-296
        double pilot = ev2b_div_c(ev2b_div_c(ev2b_div_c(eta * amps * veff, 1000.0, 0.001), B, rB), p.c60, p.rc60);
        const double maxd = ev2b_div_c(ev2b_div_c(eta * pmax, B, rB), p.c60, p.rc60);  // :297-298
        if (pilot > maxd) pilot = maxd;                                            // :300-301
        const double soc = ev2b_div_c(cap, B, rB);
        double curr;
        if (ts == 1.0) {                                                           // :303-306
            curr = pilot + soc;
            if (curr > 1.0) curr = 1.0;
        } else {
            // pts = ts + (pilot - maxd) / maxd * (ts - 1)   :312-314.  A saturated request has pilot == maxd exactly, and a
            // zero numerator sends the compiler's fp64 division into its out-of-line slow path (ncu: ~270 warp instructions
            // per env-step); 0 / maxd * (ts - 1) is +-0 and ts + (+-0) == ts, so those lanes skip the division.
            // (round 2: ptxas turned the round-1 `if (dz != 0) ratio = dz / maxd` into a SELECT, so the saturated lanes --
            //  half of the charging lanes of the bench workload -- still ran the division and 98 % of the warps called its
            //  slow path, 5.5 % of the kernel's instructions (profiles/r2c_evl_g1_ncu_busy_lines.txt).  Those lanes now
            //  divide maxd / maxd, which stays on the fast path, and discard the quotient.  The numerator goes through an
            //  empty asm: without it the optimiser sees that the saturated lanes never use the quotient, drops the select
            //  and divides dz / maxd in every lane again -- 18 197 of 18 604 warp-iterations of the busiest c3 step still
            //  called the slow path, 5.8 % of the instructions: profiles/r2g_evl_lines_busiest.txt.)
            const double dz = pilot - maxd;
            const bool sat = dz == 0.0 && maxd != 0.0;
            double num = sat ? maxd : dz;
            EV2B_OPAQUE_F64(num);
            const double quot = num / maxd;
            const double ratio = sat ? 0.0 : quot;
            const double pts = ts + ratio * (ts - 1.0);
            double nsoc;
            // `1 <= (pts - soc) / pilot`  (:323)  without the division: for 0 < pilot and x = pts - soc > 0, x < pilot
            // implies x / pilot < 1 - 2^-53 exactly, which rounds below 1, and x >= pilot implies a quotient >= 1 -- so the
            // float64 comparison of the rounded quotient with 1 is the comparison of x with pilot.
            if (soc < pts && pilot <= pts - soc) {
                nsoc = pilot + soc;                                                // constant-current stage  :323-324
            } else {
                // the two constant-voltage expressions differ only in their operands; selecting the operands first lets
                // the lanes of a warp share one exp / one division instead of running both branches in turn
                //   soc <  pts: 1 + exp(mult * (pilot + soc - pts) / (pts - 1)) * (pts - 1)      :326-330
                //   soc >= pts: 1 + exp(mult * pilot / (pts - 1)) * (soc - 1)                    :332-334
                const bool below = soc < pts;
                const double x = below ? pilot + soc - pts : pilot;
                const double y = below ? pts - 1.0 : soc - 1.0;
                nsoc = 1.0 + exp(par.y * x / (pts - 1.0)) * y;
            }
            const double lim = (maxd > pilot) ? pilot : maxd;                      // :336-339
            curr = (nsoc - soc > lim) ? lim + soc : nsoc;                          // :341-344
        }
        cap = curr * B;                                                            // :348
        energy = (curr - soc) * B;                                                 // :346,352
        act_amps = ev2b_div_c(ev2b_div_c(energy, p.p60, p.rp60) * 1000.0, veff, rveff);   // :355
    } else {                                                                       // EV._discharge  ev.py:357-405
        const double pmd = lim.y, bmin = par.x;
        double given_power = ev2b_div_c(amps * veff, 1000.0, 0.001);               // :367
        if (fabs(given_power) > fabs(pmd)) given_power = pmd;                      // :370-371
        double given_energy = ev2b_div_c(given_power * eta * p.period, 60.0, 1.0 / 60.0);   // :381
        const double prev_cap = cap;
        if (cap + given_energy < bmin) {                                           // :382-393
            if (cap > bmin) { energy = -(cap - bmin); given_energy = energy; }
            else { energy = 0.0; given_energy = 0.0; }
            cap = bmin;
        } else {
            energy = given_energy;
            cap += given_energy;
        }
        if (STATS && p.stats) { const double be = par.y; em_cross = prev_cap > be && cap < be; }   // :401-402
        act_amps = ev2b_div_c(ev2b_div_c(given_energy * 60.0, p.period, p.rperiod) * 1000.0, veff, rveff);   // :405
    }
    cap = ev2b_div_c(ceil(cap * 100.0), 100.0, 0.01);                              // my_ceil  ev.py:183,188-189
    return true;
}

// ---- statistics mode: what EV.get_battery_degradation / get_statistics need of one EV ----------
// Called when the EV leaves (or at the last step for EVs still connected).  ev.py:442-521, utils.py:49-63
// Returns the EV's calendar / cyclic degradation (the caller adds them to its charger's totals, in port order).
__device__ EV2B_NOINLINE void finalize_ev(const Params &p, size_t ip, const EvSpec *sp, double afap,
                                         int t_arr, int t_dep, double cap_final, double &d_cal_out, double &d_cyc_out) {
    const double e0 = 7.543e6, e1 = 23.75e6, z0 = 7.348e-3, z1 = 3.667, z2 = 7.6e-4, z3 = 4.081e-3;
    const double k = 0.8263, v_min = 3.3324, b_cap_kwh = 78;
    const double B = __ldg(&sp->B);
    const double final_soc = cap_final / B;
    const int cnt = p.st_cnt[ip], n_hist = cnt & 0xFFFF, n_act = cnt >> 16;
    const double T_sim = (double)(t_dep - t_arr + 1) * p.period / (60.0 * 24.0);
    const double avg_soc = (p.st_soc_sum[ip] + final_soc) / (double)(n_hist + 1);
    const double alpha = (e0 * (v_min + k * avg_soc) - e1) * p.k_exp;
    const double d_cal = alpha * T_sim * p.k_cal;
    const double *f = p.st_act + ip * (size_t)p.L;
    double fsum = final_soc;
    for (int j = 0; j < n_act; ++j) fsum += f[j];
    const double avg_f = fsum / (double)(n_act + 1);
    double dev = fabs(avg_f - final_soc);
    for (int j = 0; j < n_act; ++j) dev += fabs(avg_f - f[j]);
    const double delta_dod = 2.0 * (dev / (double)(n_act + 1));
    const double v_half = v_min + k * 0.5;
    const double beta = z0 * (v_half - z1) * (v_half - z1) + z2 + z3 * delta_dod;
    const double d_cyc = beta * (p.st_abs_e[ip] / b_cap_kwh) * p.k_cyc;
    d_cal_out = d_cal;
    d_cyc_out = d_cyc;
    const int k_fin = p.st_nfin[ip];
    p.st_r[ip * (size_t)p.Smax + k_fin] = (cap_final / afap) * 100.0;      // utils.py:59-62
    p.st_nfin[ip] = k_fin + 1;
}

// ---- B2: Laurent power flow of one env by one warp  (grid.py:120-141, numbarize.py:268-325) ----------
// S/V/Lm: shared scratch [n_bus] each; trp_env: transformer (= bus) EV power of this env.  Returns the
// voltage-band loss sum(min(0, 0.05 - |1 - vm|)) (reward.py:107-110) in every lane.
__device__ __forceinline__ double power_flow_env(const Params &p, double2 *S, const double *trp_env, int js, int jt, int je,
                                              int lane) {
    const int n = p.n_bus;
    double2 *V = S + n, *Lm = V + n;
            const size_t gb = ((size_t)js * (p.T + 1) + jt) * n;
    for (int r = lane; r < n; r += 32) {                     // S = (P + jQ) / s_base, flat start  grid_tensor.py:594-630
        S[r] = make_double2((p.grid_act[gb + r] + trp_env[r]) / p.s_base, p.grid_rea[gb + r] / p.s_base);
        V[r] = make_double2(1.0, 0.0);
    }
    __syncwarp();
    int it = 0; double tol = 1e300;
    while (it < 100 && tol >= 1e-6) {
        for (int r = lane; r < n; r += 32) {                 // lambda = conj(S * (1 / v0))   numbarize.py:304
            const double a = V[r].x, b = V[r].y;
            double rr, ri;                                   // numpy complex reciprocal (Smith)
            if (fabs(a) >= fabs(b)) { const double rat = b / a, scl = 1.0 / (a + b * rat); rr = scl; ri = (0.0 - rat) * scl; }
            else { const double rat = a / b, scl = 1.0 / (b + a * rat); rr = rat * scl; ri = (0.0 - 1.0) * scl; }
            Lm[r] = make_double2(S[r].x * rr - S[r].y * ri, -(S[r].x * ri + S[r].y * rr));
        }
        __syncwarp();
        double dmax = 0.0;
        double2 nv[4];                                       // rows lane, lane+32, ... (n_bus <= 128)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = lane + 32 * q;
            nv[q] = make_double2(0.0, 0.0);
            if (r < n) {
                double zr = 0.0, zi = 0.0;
                for (int c2 = 0; c2 < n; ++c2) {             // Z = K @ lambda   numbarize.py:305
                    const double2 kk = __ldg(&p.grid_Kt[(size_t)c2 * n + r]);
                    const double2 lm = Lm[c2];
                    zr = fma(kk.x, lm.x, fma(-kk.y, lm.y, zr));     // voltages carry a 1e-6 solver tolerance: FMA is fine here
                    zi = fma(kk.x, lm.y, fma(kk.y, lm.x, zi));
                }
                const double2 ll = __ldg(&p.grid_L[r]);
                nv[q] = make_double2(zr + ll.x, zi + ll.y);  // voltage_k = Z + L
                dmax = fmax(dmax, fabs(sqrt(nv[q].x * nv[q].x + nv[q].y * nv[q].y) - sqrt(V[r].x * V[r].x + V[r].y * V[r].y)));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        tol = dmax;                                          // numbarize.py:308
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int r = lane + 32 * q; if (r < n) V[r] = nv[q]; }
        __syncwarp();
        ++it;
    }
    double lv = 0.0;                                         // sum(min(0, 0.05 - |1 - vm|)) incl. the slack (vm = 1)
    for (int r = lane; r < n; r += 32) {
        const double vm = sqrt(V[r].x * V[r].x + V[r].y * V[r].y);
        const double x = 0.05 - fabs(1.0 - vm);
        lv += x < 0.0 ? x : 0.0;
        if (p.out.node_voltage) p.out.node_voltage[(size_t)je * (n + 1) + r + 1] = vm;
    }
    lv = warp_sum(lv);
    return lv;
}

// ---- the fused step kernel --------------------------------------------------------------------
// ActT: float or double actions.  NP: ports per charger when uniform (1, 2), 0 = ragged (CsStatic).
template <typename ActT, int NP, bool UNI, int MAXT, int MINB, bool HEAVY, bool OPTOUT>   // HEAVY: statistics mode and/or distribution grid compiled in
__global__ void __launch_bounds__(MAXT, MINB) step_kernel(const __grid_constant__ Params p) {   // __grid_constant__: &p may be passed to finalize_ev without a per-thread copy
    EV2B_DYNAMIC_SMEM(smem_raw);
    const int NT = blockDim.x;
    const int PP = p.EPB * p.P;
    double *red   = reinterpret_cast<double *>(smem_raw);                 // [kNRed][NT]
    double *resE  = reinterpret_cast<double *>(smem_raw + p.o_resE);      // [PP] action in, energy out
    double *resA  = reinterpret_cast<double *>(smem_raw + p.o_resA);      // [PP] actual amps out
    double *resC  = reinterpret_cast<double *>(smem_raw + p.o_resC);      // [PP] battery level in/out (-1: EV inactive)
    double *trov  = reinterpret_cast<double *>(smem_raw + p.o_trov);      // [EPB*Tr] overload per transformer
    double *trp   = reinterpret_cast<double *>(smem_raw + p.o_trp);       // [EPB*Tr] transformer power (grid: bus EV power)
    double *lossv = reinterpret_cast<double *>(smem_raw + p.o_lossv);     // [EPB] voltage-band loss  reward.py:107-110
    double2 *pfv  = reinterpret_cast<double2 *>(smem_raw + p.o_pfv);      // [EPB][3][n_bus] S, V, lambda
    double *envs  = reinterpret_cast<double *>(smem_raw + p.o_envs);      // [EPB][kNRed]
    double *pre   = reinterpret_cast<double *>(smem_raw + p.o_pre);       // [EPB][pre_stride] prefetched per-env records
    uint2  *whot  = reinterpret_cast<uint2 *>(smem_raw + p.o_whot);       // [PP] hot words z, w of work items
    int    *cnt   = reinterpret_cast<int *>(smem_raw + p.o_cnt);          // [NT] invalid | dep<<10 | arr<<20
    int    *envi  = reinterpret_cast<int *>(smem_raw + p.o_envi);         // [EPB][8] t, scn, -, flags, invalid, departures, arrivals, -
    int    *wl    = reinterpret_cast<int *>(smem_raw + p.o_wl);           // [PP] work list (port_local)
    int    *wcnt  = reinterpret_cast<int *>(smem_raw + p.o_wcnt);         // [1] (+3 pad)
    signed char *pflag = reinterpret_cast<signed char *>(smem_raw + p.o_pflag);   // [PP] ragged path only

    const int tid = threadIdx.x;
    const int el = p.C == 1 ? tid : (int)__umulhi((unsigned)tid, p.c_magic);   // tid / C
    const int c = tid - el * p.C;
    const int e = (p.env0 + blockIdx.x * p.EPB) + el;
    const bool valid = (el < p.EPB) && (e < p.env_end);
    const ActT *actions = reinterpret_cast<const ActT *>(p.actions);
    const bool want_obs = (p.out.obs != nullptr) && (p.state_kind != EV2B_STATE_NONE);
    constexpr int NPR = NP > 0 ? NP : 1;

    if (tid == 0) { wcnt[0] = 0; wcnt[1] = 0; }   // charge items fill wl from the front, discharge items from the back
    int t = 0, s = 0;
    bool live = false;
    // ---- A1: per charger: loads, empty-port masking, normalisation, work-list compaction -------
    // Every independent global load of this thread is issued BEFORE the first barrier.
    uint4 h[NPR];
    double araw[NPR];
    double capv[NPR];
    unsigned pushed = 0, asign = 0;      // bit j: port j is a work item / its action is > 0
    int invalid = 0;
    int port0 = 0, n = 0;
    double pre_cp = 0.0, pre_dp = 0.0;     // prices of this step, fetched before the first barrier
    double exch0[NPR];
    if (valid) {
        t = p.env_step[e];
        s = p.env_scn[e];
        double *pe = pre + el * p.pre_stride;
        for (int i = c; i <= kPrePot; i += p.C)                 // KPI sums and the potential of this step: need only e
            cp_async8(pe + i, i < kPrePot ? p.env_kpi + (size_t)e * EV2B_KPI_COUNT + i : p.env_pot + e);
        if (NP > 0) {
            port0 = UNI ? c * NP : p.cs[c].port_off;
            const size_t pb = (size_t)e * p.P + port0;
#pragma unroll
            for (int j = 0; j < NP; ++j) { h[j] = p.hot[pb + j]; araw[j] = agent_action<ActT>(p, actions, pb + j, t); }
        }
        live = t < p.T;
        if (live) {
            const EnvT et0 = p.env_t[(size_t)s * p.T + t]; pre_cp = et0.cp; pre_dp = et0.dp;
            for (int i = c; i < 2 + 2 * p.Tr; i += p.C) {        // setpoint[t], setpoint[t+1], TrT rows (two 16 B halves each)
                if (i < 2) { if (t + i < p.T) cp_async8(pe + kPreSet + i, &p.env_t[(size_t)s * p.T + t + i].setpoint); }
                else cp_async16(pe + kPreTr + 2 * (i - 2),
                                reinterpret_cast<const double *>(p.tr_t + ((size_t)s * p.T + t) * p.Tr) + 2 * (i - 2));
            }
        }
        if (c == 0) { envi[el * 8 + 0] = t; envi[el * 8 + 1] = s; envi[el * 8 + 2] = 0; envi[el * 8 + 3] = 0; }   // [4..6] are written in phase B
    }
    __syncthreads();

    if (valid && live) {
        const CsStatic &cs = cs_of<UNI>(p, c);
        if (NP == 0) port0 = cs.port_off;
        n = NP > 0 ? NP : cs.n_ports;
        const size_t pbase = (size_t)e * p.P + port0;
        double sum = 0.0;
        if (NP > 0) {
            double a[NPR];
#pragma unroll
            for (int j = 0; j < NP; ++j) a[j] = araw[j];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const bool occ = hot_t_arr(h[j]) <= t && t <= hot_t_dep(h[j]);
                capv[j] = occ ? p.cap[pbase + j] : 0.0;
                exch0[j] = occ ? p.exch[pbase + j] : 0.0;
                if (!occ) { a[j] = 0.0; ++invalid; }                     // ev_charger.py:137-140
                sum = sum + a[j];                                        // python sum(), left to right  :143
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                double an = a[j];
                if (sum > 1.0) an = an / sum; else if (sum < -1.0) an = -an / sum;     // :143-149
                if (an != 0.0) {                                         // occupied and a non-zero request
                    const int pl = el * p.P + port0 + j;
                    if (an > 0.0) wl[atomicAdd(&wcnt[0], 1)] = pl;
                    else          wl[PP - 1 - atomicAdd(&wcnt[1], 1)] = pl;
                    resE[pl] = an; resC[pl] = capv[j]; whot[pl] = make_uint2(h[j].z, h[j].w);
                    pushed |= 1u << j;
                    if (an > 0.0) asign |= 1u << j;
                }
            }
        } else {
            for (int j = 0; j < n; ++j) {
                const uint4 hj = p.hot[pbase + j];
                const bool occ = hot_t_arr(hj) <= t && t <= hot_t_dep(hj);
                if (!occ) ++invalid;
                sum = sum + (occ ? agent_action<ActT>(p, actions, pbase + j, t) : 0.0);
            }
            for (int j = 0; j < n; ++j) {
                const uint4 hj = p.hot[pbase + j];
                const bool occ = hot_t_arr(hj) <= t && t <= hot_t_dep(hj);
                double an = occ ? agent_action<ActT>(p, actions, pbase + j, t) : 0.0;
                if (sum > 1.0) an = an / sum; else if (sum < -1.0) an = -an / sum;
                const int pl = el * p.P + port0 + j;
                signed char f = 0;
                if (an != 0.0) {
                    if (an > 0.0) wl[atomicAdd(&wcnt[0], 1)] = pl;
                    else          wl[PP - 1 - atomicAdd(&wcnt[1], 1)] = pl;
                    resE[pl] = an; resC[pl] = p.cap[pbase + j]; whot[pl] = make_uint2(hj.z, hj.w);
                    f = an > 0.0 ? 1 : -1;
                }
                pflag[pl] = f;
            }
        }
    }
    // (scenario, time)-only part of the observation: price window + forecast / limit blocks
    if (want_obs && p.W > 0) {
        for (int jel = 0; jel < p.EPB; ++jel) {
            const int je = (p.env0 + blockIdx.x * p.EPB) + jel;
            if (je >= p.env_end) break;
            const int jt = envi[jel * 8 + 0];
            if (jt >= p.T) continue;
            const int js = envi[jel * 8 + 1];
            for (int i = tid; i < p.W; i += NT)
                p.out.obs[(size_t)je * p.D + p.series_off[i]] = obs_series_fetch(p, js, jt + 1, i);
        }
    }
    __syncthreads();

    // ---- A2: the float64 battery model on the compacted work list (dense warps) ----------------
    {
        // two passes so that a warp runs either the charge or the discharge model, not both
        const int n_ch = wcnt[0], n_dis = wcnt[1];
        const int n_ch_pad = (n_ch + 31) & ~31;                        // discharge items start on a warp boundary
        for (int i = tid; i < n_ch_pad + n_dis; i += NT) {
            if (i >= n_ch && i < n_ch_pad) continue;
            const int pl = i < n_ch ? wl[i] : wl[PP - 1 - (i - n_ch_pad)];
            const int port = p.EPB == 1 ? pl : pl - (int)__umulhi((unsigned)pl, p.p_magic) * p.P;
            const CsStatic &cs = cs_of<UNI>(p, UNI ? 0 : p.port_cs[port]);
            const uint2 hw = whot[pl];
            double cap = resC[pl], energy, amps;
            bool em_cross;
            const bool active = ev_step_item<HEAVY>(p, cs, hw.x, hw.y, resE[pl], cap, energy, amps, em_cross);
            if (HEAVY && p.stats && em_cross) {   // min_emergency_battery_capacity_metric  ev.py:401-402 (integer atomics: order-free)
                const int jel = p.EPB == 1 ? 0 : (int)__umulhi((unsigned)pl, p.p_magic);
                atomicAdd(&p.cs_em[(size_t)((p.env0 + blockIdx.x * p.EPB) + jel) * p.C + p.port_cs[port]], 1);
            }
            resE[pl] = energy; resA[pl] = amps;
            resC[pl] = active ? cap : -1.0;                              // EV saw amps == 0: nothing changes (ev.py:158-163)
        }
    }
    __syncthreads();

    // ---- A3: charger accounting in port order, departures, arrivals, potential, obs tuples -----
    double rP = 0, rA = 0, rProfit = 0, rSatExp = 0, rPot = 0, rCh = 0, rDis = 0, rSat = 0;
    int rCnt = invalid;
    if (valid && live) {
        const CsStatic &cs = cs_of<UNI>(p, c);
        EnvT et; et.cp = pre_cp; et.dp = pre_dp;
        const size_t pbase = (size_t)e * p.P + port0;
        const int tq = t + 1;
        float *obs_row = p.out.obs + (size_t)e * p.D;      // observation rows live in the caller's buffer across steps
        bool overflow = false;
#pragma unroll
        for (int j = 0; j < (NP > 0 ? NP : n); ++j) {
            const size_t ip = pbase + j;
            const int pl = el * p.P + port0 + j;
            uint4 hj; double cv; bool was_item, apos;
            if (NP > 0) { hj = h[j]; cv = capv[j]; was_item = (pushed >> j) & 1u; apos = (asign >> j) & 1u; }
            else {
                hj = p.hot[ip];
                const signed char f = pflag[pl];
                was_item = f != 0; apos = f > 0;
                cv = (hot_t_arr(hj) <= t && t <= hot_t_dep(hj)) ? p.cap[ip] : 0.0;
            }
            const bool occ = hot_t_arr(hj) <= t && t <= hot_t_dep(hj);
            double energy = 0.0, act_amps = 0.0;
            const double cv_old = cv;
            double exch_new = 0.0; bool exch_valid = false;
            if (was_item) {
                energy = resE[pl];
                const double cnew = resC[pl];
                if (cnew >= 0.0) {                                        // the EV was active this step
                    cv = cnew;
                    p.cap[ip] = cv;
                    exch_new = (NP > 0 ? exch0[j] : p.exch[ip]) + energy;   // total_energy_exchanged  ev.py:178
                    exch_valid = true;
                    p.exch[ip] = exch_new;
                }
                const double ae = fabs(energy);
                if (apos) { rProfit += ae * et.cp; rCh += ae; }           // ev_charger.py:178-179
                else      { rProfit += ae * et.dp; rDis += ae; }          // ev_charger.py:194-195
                rP += ev2b_div_c(energy * 60.0, p.period, p.rperiod);     // :180,196
                act_amps = resA[pl];
                rA += act_amps;                                           // :181,197
            }
            if (rA - 0.0001 > cs.imax) overflow = true;                   // :203-205
            if (HEAVY && p.stats && occ) {      // EV.step bookkeeping: historic_soc / active_steps / |energy|  ev.py:156,178-185
                const EvSpec *sq = p.spec + hot_spec(hj);
                const double soc0 = ev2b_div_c(cv_old, __ldg(&sq->B), __ldg(&sq->rB));
                int cn = p.st_cnt[ip];
                p.st_soc_sum[ip] += soc0;
                if (was_item && act_amps != 0.0) {
                    p.st_act[ip * (size_t)p.L + (cn >> 16)] = soc0;
                    cn += 1 << 16;
                }
                if (was_item) p.st_abs_e[ip] += fabs(energy);
                p.st_cnt[ip] = cn + 1;
            }
            if (EV2B_OPT(p.out.port_energy)) p.out.port_energy[ip] = (float)energy;

            // departure (charger step counter == t)        ev_charger.py:209-224, ev.py:199-214
            double dsat = __longlong_as_double(0x7ff8000000000000LL), dcap = dsat;
            if (occ && t >= hot_t_dep(hj)) {
                const double des = __ldg(&p.spec[hot_spec(hj)].desired);
                const double sat = (cv < des - 0.001) ? cv / des : 1.0;
                if (p.reward_kind == EV2B_REWARD_GRID_FULL || p.reward_kind == EV2B_REWARD_GRID_SIMPLE)
                    rSatExp += (cv - des) * (cv - des);                   // -user_costs  reward.py:99-102
                else if (HEAVY && p.reward_kind >= EV2B_REWARD_SQTR_TR_USER) {   // per-departure penalty of the other stock rewards
                    if (p.reward_kind == EV2B_REWARD_SQTR_TR_USER) rSatExp += 1000.0 * (1.0 - sat);                // reward.py:29-30
                    else if (p.reward_kind == EV2B_REWARD_V2G_PROFITMAX) { if (des > cv) rSatExp += 100.0 * (des - cv); }   // :136-138
                    else if (p.reward_kind >= EV2B_REWARD_V2G_PROFITMAX_V2 && p.reward_kind <= EV2B_REWARD_PST_PROFITMAX_V2) { if (des > cv) rSatExp += 0.05 * ((des - cv) * (des - cv)); }   // :199-207
                } else
                    rSatExp += 100.0 * exp(-10.0 * sat);                  // reward.py:42,85
                rSat += sat;
                rCnt += 1 << 10;
                dsat = sat; dcap = cv;
                if (HEAVY && p.stats) {
                    const size_t ec = (size_t)e * p.C + c;
                    p.cs_sat_sum[ec] += sat; p.cs_served[ec] += 1;          // ev_charger.py:218-220
                    const SessRec r0 = p.sess[((size_t)s * p.P + port0 + j) * p.Smax + hot_cursor(hj) - 1];
                    double d1, d2;
                    finalize_ev(p, ip, p.spec + hot_spec(hj), r0.afap, hot_t_arr(hj), hot_t_dep(hj), cv, d1, d2);
                    p.cs_dcal[ec] += d1; p.cs_dcyc[ec] += d2;
                }
            }
            if (EV2B_OPT(p.out.dep_sat)) p.out.dep_sat[ip] = dsat;
            if (EV2B_OPT(p.out.dep_cap)) p.out.dep_cap[ip] = dcap;

            // arrival of the next session at t+1            ev2gym_env.py:399-417, ev_charger.py:266-285
            if (hot_next_arr(hj) == tq) {
                const SessRec r = p.sess[((size_t)s * p.P + port0 + j) * p.Smax + hot_cursor(hj)];
                hj = r.hot;
                cv = r.cap0;
                p.hot[ip] = hj;
                p.cap[ip] = cv;
                p.exch[ip] = 0.0;
                exch_new = 0.0; exch_valid = true;
                rCnt += 1 << 20;
                if (HEAVY && p.stats) { p.st_soc_sum[ip] = 0.0; p.st_abs_e[ip] = 0.0; p.st_cnt[ip] = 0; }
            }
            const bool occ_after = hot_t_arr(hj) <= tq && tq <= hot_t_dep(hj);
            if (EV2B_OPT(p.out.action_mask)) p.out.action_mask[ip] = occ_after ? 1 : 0;          // ev2gym_env.py:452-457
            if (HEAVY && p.stats && occ_after && tq >= p.T) {   // episode over: EVs still connected count too (env.EVs)
                const SessRec r0 = p.sess[((size_t)s * p.P + port0 + j) * p.Smax + hot_cursor(hj) - 1];
                const size_t ec = (size_t)e * p.C + c;
                double d1, d2;
                finalize_ev(p, ip, p.spec + hot_spec(hj), r0.afap, hot_t_arr(hj), hot_t_dep(hj), cv, d1, d2);
                p.cs_dcal[ec] += d1; p.cs_dcyc[ec] += d2;
            }
            if (occ_after) {
                const EvSpec *sp = p.spec + hot_spec(hj);
                const double B = __ldg(&sp->B);
                if (HEAVY && p.reward_kind >= EV2B_REWARD_V2G_PROFITMAX_V2 && p.reward_kind <= EV2B_REWARD_PST_PROFITMAX_V2) {   // V2G_profitmaxV2 family: EVs that can no longer
                    const double des = __ldg(&sp->desired), pmax = __ldg(&sp->pmax_ac);   // reach their desired level  reward.py:172-190
                    const double min_steps = (des - cv) / (pmax / p.c60);
                    const int dstep = hot_t_dep(hj) - tq;
                    if (min_steps > (double)dstep) {
                        const double gap = (des - ((double)(dstep + 1) * pmax / p.c60)) - cv;
                        rSatExp += 0.05 * (gap * gap);
                    }
                }
                // charge power potential for step t+1            utils.py:766-777
                if (cv < B && hot_t_dep(hj) > tq) rPot += __ldg(&p.pot_kw[hot_spec(hj) * p.n_cls + cs.cls]);
                if (want_obs) {       // observation tuple (transformer-major slot)   state.py:37-57, 85-102, 137-151
                    float *o = obs_row + p.obs_slot[port0 + j];
                    if (p.state_kind == EV2B_STATE_V2G_GRID) {          // state.py:262-270
                        o[0] = (float)cv;
                        o[1] = (float)(hot_t_dep(hj) - tq + 1);
                        o[2] = (float)__ldg(&p.cs_tr[c]);     // cs.connected_bus (cs0 is shared in the uniform layout)
                    } else if (p.state_kind == EV2B_STATE_PUBLIC_PST) {
                        o[0] = (cv == B) ? 1.f : 0.5f;
                        o[1] = (float)(exch_valid ? exch_new : (NP > 0 ? exch0[j] : p.exch[ip]));
                        o[2] = (float)(tq - hot_t_arr(hj));
                    } else {
                        o[0] = (float)ev2b_div_c(cv, B, __ldg(&sp->rB));
                        o[1] = (float)(hot_t_dep(hj) - tq);
                    }
                }
            } else if (want_obs && (occ || p.obs_full)) {     // the port just emptied (or a full rewrite was asked for)
                float *o = obs_row + p.obs_slot[port0 + j];
                o[0] = 0.f; o[1] = 0.f;
                if (p.state_kind == EV2B_STATE_PUBLIC_PST || p.state_kind == EV2B_STATE_V2G_GRID) o[2] = 0.f;
            }
        }
        // clamp the charger's potential                      utils.py:779-789
        if (rPot > cs.max_power) rPot = cs.max_power;
        else if (rPot < cs.min_power) rPot = 0.0;
        if (overflow) atomicOr(&envi[el * 8 + 3], (int)EV2B_ST_AMPS_OVERFLOW);
        if (EV2B_OPT(p.out.cs_power))   p.out.cs_power[(size_t)e * p.C + c] = (float)rP;
        if (EV2B_OPT(p.out.cs_current)) p.out.cs_current[(size_t)e * p.C + c] = (float)rA;
        if (EV2B_OPT(p.out.hist_cs_power))   p.out.hist_cs_power[((size_t)e * p.T + t) * p.C + c] = (float)rP;
        if (EV2B_OPT(p.out.hist_cs_current)) p.out.hist_cs_current[((size_t)e * p.T + t) * p.C + c] = (float)rA;
    }
    red[RedP * NT + tid] = rP;             red[RedProfit * NT + tid] = rProfit;
    red[RedSatExp * NT + tid] = rSatExp;   red[RedPot * NT + tid] = rPot;
    red[RedCharged * NT + tid] = rCh;      red[RedDischarged * NT + tid] = rDis;
    red[RedSatSum * NT + tid] = rSat;
    cnt[tid] = rCnt;
    cp_async_wait_all();
    __syncthreads();

    // ---- B: fixed-order reductions.  Three small warp jobs per env, every lane busy: lanes are split
    //      into (quantity, segment) pairs, each lane sums its segment serially, then a short xor tree.
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = NT >> 5;
        for (int job = warp; job < 3 * p.EPB; job += nwarps) {
            const int jel = job / 3, kind = job - jel * 3;
            const int je = (p.env0 + blockIdx.x * p.EPB) + jel;
            if (je >= p.env_end) continue;
            const int jt = envi[jel * 8 + 0];
            if (jt >= p.T) continue;
            if (kind == 0) {          // Transformer.step accumulation + overload   transformer.py:264-302
                int nseg = 1;
                while (nseg * 2 * p.Tr <= 32) nseg *= 2;
                const int per = 32 / nseg;                              // transformers per pass
                for (int k0 = 0; k0 < p.Tr; k0 += per) {
                    const int k = k0 + lane / nseg, seg = lane & (nseg - 1);
                    double sp_ = 0.0;
                    if (k < p.Tr && lane / nseg < per) {
                        const int i0 = p.tr_cs_off[k], n_k = p.tr_cs_off[k + 1] - i0;
                        const int chunk = (n_k + nseg - 1) / nseg;
                        const int lo = seg * chunk, hi = min(n_k, lo + chunk);
                        for (int i = lo; i < hi; ++i) sp_ += red[RedP * NT + jel * p.C + p.tr_cs_idx[i0 + i]];
                    }
                    for (int o = nseg >> 1; o > 0; o >>= 1) sp_ += __shfl_xor_sync(0xffffffffu, sp_, o);
                    if (seg == 0 && k < p.Tr && lane / nseg < per) {
                        const double *tq4 = pre + jel * p.pre_stride + kPreTr + 4 * k;
                        TrT tt; tt.infl = tq4[0]; tt.solar = tq4[1]; tt.maxp = tq4[2]; tt.minp = tq4[3];
                        const double ptot = (tt.infl + tt.solar) + sp_;
                        double ov = 0.0;
                        if (ptot > tt.maxp + 0.0001 || ptot < tt.minp - 0.0001) ov = fabs(ptot - tt.maxp);
                        trov[jel * p.Tr + k] = ov;
                        trp[jel * p.Tr + k] = ptot;
                        if (EV2B_OPT(p.out.tr_power))    p.out.tr_power[(size_t)je * p.Tr + k] = ptot;
                        if (EV2B_OPT(p.out.tr_overload)) p.out.tr_overload[(size_t)je * p.Tr + k] = ov;
                        if (EV2B_OPT(p.out.hist_tr_overload)) p.out.hist_tr_overload[((size_t)je * p.T + jt) * p.Tr + k] = ov;
                    }
                }
            } else if (kind == 1) {   // env-level float64 sums: 7 quantities x 4 segments = 28 lanes
                const int q = lane >> 2, seg = lane & 3;
                double v = 0.0;
                if (q < kNRed) {
                    const int chunk = (p.C + 3) >> 2;
                    const int lo = seg * chunk, hi = min(p.C, lo + chunk);
                    const double *r = red + q * NT + jel * p.C;
                    for (int i = lo; i < hi; ++i) v += r[i];
                }
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (q < kNRed && seg == 0) envs[jel * kNRed + q] = v;
            } else {                  // env-level integer counts: 3 fields x 8 segments = 24 lanes
                const int q = lane >> 3, seg = lane & 7;
                int v = 0;
                if (q < 3) {
                    const int chunk = (p.C + 7) >> 3;
                    const int lo = seg * chunk, hi = min(p.C, lo + chunk);
                    for (int i = lo; i < hi; ++i) v += (cnt[jel * p.C + i] >> (10 * q)) & 1023;
                }
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (q < 3 && seg == 0) envi[jel * 8 + 4 + q] = v;     // (whole ints: the env totals exceed 10 bits for P > 1023)
            }
        }
    }
    __syncthreads();

    // ---- B2: distribution grid: Laurent power flow, one warp per env ------------------------------------
    if (HEAVY && p.n_bus > 0) {
        const int warp = tid >> 5, lane = tid & 31, nwarps = NT >> 5, n = p.n_bus;
        for (int jel = warp; jel < p.EPB; jel += nwarps) {
            const int je = (p.env0 + blockIdx.x * p.EPB) + jel;
            if (je >= p.env_end) continue;
            const int jt = envi[jel * 8 + 0];
            if (jt >= p.T) continue;
            const double lv = power_flow_env(p, pfv + (size_t)jel * 3 * n, trp + jel * p.Tr, envi[jel * 8 + 1], jt, je, lane);
            if (lane == 0) { lossv[jel] = lv; if (p.out.node_voltage) p.out.node_voltage[(size_t)je * (n + 1)] = 1.0; }
        }
        __syncthreads();
    }

    // ---- C: one thread per env: reward, KPIs, step counter ---------------------------------------
    if (tid < p.EPB) {
        const int je = (p.env0 + blockIdx.x * p.EPB) + tid;
        if (je < p.env_end) {
            const int jt = envi[tid * 8 + 0], js = envi[tid * 8 + 1];
            unsigned status = (unsigned)envi[tid * 8 + 3];
            double reward = 0.0, costs = 0.0;
            if (jt < p.T) {
                const double *v = envs + tid * kNRed;
                const int n_inv = envi[tid * 8 + 4], n_dep = envi[tid * 8 + 5], n_arr = envi[tid * 8 + 6];
                const double *pe = pre + tid * p.pre_stride;
                EnvT et; et.setpoint = pe[kPreSet];
                const double usage = v[RedP];                                  // current_power_usage[t]  ev2gym_env.py:375
                costs = v[RedProfit];
                double ovsum = 0.0;
                for (int k = 0; k < p.Tr; ++k) ovsum += trov[tid * p.Tr + k];
                if (p.reward_kind == EV2B_REWARD_SQ_TRACKING) {                // reward.py:11-12
                    const double pot = pe[kPrePot];
                    const double m = et.setpoint < pot ? et.setpoint : pot;
                    reward = -((m - usage) * (m - usage));
                } else if (p.reward_kind == EV2B_REWARD_PROFIT_TR_USER) {      // reward.py:36-44
                    reward = costs - 100.0 * ovsum - v[RedSatExp];
                } else if (p.reward_kind == EV2B_REWARD_PROFIT_MAX) {          // reward.py:81-87
                    reward = costs - v[RedSatExp];
                } else if (p.reward_kind == EV2B_REWARD_GRID_FULL) {           // reward.py:89-111
                    reward = costs + 1000.0 * lossv[tid] - v[RedSatExp];
                } else if (p.reward_kind == EV2B_REWARD_GRID_SIMPLE) {         // reward.py:114-121
                    reward = 1000.0 * lossv[tid];
                } else if (HEAVY && p.reward_kind == EV2B_REWARD_SQTR_TR_USER) {   // reward.py:16-32
                    double m = et.setpoint < pe[kPrePot] ? et.setpoint : pe[kPrePot];
                    const double lim = pe[kPreTr + 2];                         // transformers[0].max_power[t]
                    if (lim < m) m = lim;
                    reward = -((m - usage) * (m - usage)) - 100.0 * ovsum - v[RedSatExp];
                } else if (HEAVY && p.reward_kind == EV2B_REWARD_SQ_TRACKING_PENALTY) {   // reward.py:46-58
                    const double pot = pe[kPrePot];
                    const double m = et.setpoint < pot ? et.setpoint : pot;
                    reward = -((m - usage) * (m - usage));
                    if (usage == 0.0 && p.env_pot_prev[je] != 0.0) reward = reward - 100.0;   // potential[current_step-2]; 0 at t = 0
                    p.env_pot_prev[je] = pot;
                } else if (HEAVY && p.reward_kind == EV2B_REWARD_SIMPLE) {         // reward.py:60-65
                    reward = -((et.setpoint - usage) * (et.setpoint - usage));
                } else if (HEAVY && p.reward_kind == EV2B_REWARD_MIN_TRACKER_SURPLUS) {   // reward.py:67-76
                    if (et.setpoint < usage) reward -= (usage - et.setpoint) * (usage - et.setpoint);
                    reward += usage;
                } else if (HEAVY && p.reward_kind == EV2B_REWARD_V2G_COSTS_SIMPLE) {   // reward.py:150-153
                    reward = costs;
                } else if (HEAVY && p.reward_kind >= EV2B_REWARD_V2G_PROFITMAX && p.reward_kind <= EV2B_REWARD_PST_PROFITMAX_V2) {   // V2G_profitmax, V2G_profitmaxV2 (+ grid / pst)
                    reward = costs - v[RedSatExp];
                    if (p.reward_kind == EV2B_REWARD_GRID_PROFITMAX_V2) reward += 50000.0 * lossv[tid];       // reward.py:274-279
                    if (p.reward_kind == EV2B_REWARD_PST_PROFITMAX_V2 && et.setpoint < usage) reward += 1000.0 * (et.setpoint - usage);   // :333-339
                }
                double *kpi = p.env_kpi + (size_t)je * EV2B_KPI_COUNT;      // old sums come from the prefetch area
                kpi[EV2B_KPI_TOTAL_REWARD] = pe[EV2B_KPI_TOTAL_REWARD] + reward;
                kpi[EV2B_KPI_TOTAL_PROFITS] = pe[EV2B_KPI_TOTAL_PROFITS] + costs;
                kpi[EV2B_KPI_ENERGY_CHARGED] = pe[EV2B_KPI_ENERGY_CHARGED] + v[RedCharged];
                kpi[EV2B_KPI_ENERGY_DISCHARGED] = pe[EV2B_KPI_ENERGY_DISCHARGED] + v[RedDischarged];
                kpi[EV2B_KPI_TR_OVERLOAD] = pe[EV2B_KPI_TR_OVERLOAD] + ovsum;
                kpi[EV2B_KPI_EVS_SERVED] = pe[EV2B_KPI_EVS_SERVED] + (double)n_dep;
                kpi[EV2B_KPI_SAT_SUM] = pe[EV2B_KPI_SAT_SUM] + v[RedSatSum];
                const double d = et.setpoint - usage;                          // utils.py:37-44
                kpi[EV2B_KPI_TRACKING_ERROR] = pe[EV2B_KPI_TRACKING_ERROR] + d * d;
                kpi[EV2B_KPI_ENERGY_TRACKING_ERROR] = pe[EV2B_KPI_ENERGY_TRACKING_ERROR] + fabs(d);
                if (usage > et.setpoint) kpi[EV2B_KPI_TRACKER_VIOLATION] = pe[EV2B_KPI_TRACKER_VIOLATION] + (usage - et.setpoint);
                kpi[EV2B_KPI_EVS_SPAWNED] = pe[EV2B_KPI_EVS_SPAWNED] + (double)n_arr;
                kpi[EV2B_KPI_INVALID_ACTIONS] = pe[EV2B_KPI_INVALID_ACTIONS] + (double)n_inv;
                kpi[EV2B_KPI_STEPS] = pe[EV2B_KPI_STEPS] + 1.0;
                p.env_pot[je] = (jt + 1 < p.T) ? v[RedPot] : 0.0;              // ev2gym_env.py:424-426
                p.env_usage[je] = usage;
                p.env_step[je] = jt + 1;
                if (EV2B_OPT(p.out.hist_usage)) p.out.hist_usage[(size_t)je * p.T + jt] = usage;
                if (jt + 1 >= p.T) status |= EV2B_ST_DONE;                     // ev2gym_env.py:460
                if (want_obs) obs_header(p, p.out.obs + (size_t)je * p.D, js, jt + 1, usage, pe[kPreSetNext]);
            } else {
                status |= EV2B_ST_DONE | EV2B_ST_WAS_DONE;                     // ev2gym_env.py:343
            }
            if (p.out.reward) p.out.reward[je] = reward;
            if (EV2B_OPT(p.out.total_costs)) p.out.total_costs[je] = costs;
            if (p.out.status) p.out.status[je] = status;
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
        s = (p.env_scn[e] + p.scn_stride) % p.S;
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
    p.exch[(size_t)e * p.P + port] = 0.0;
    if (p.rr_key) {                       // a new episode starts with a fresh agent: empty queue
        p.rr_key[(size_t)e * p.P + port] = kRrAbsent;
        if (port == 0) { p.rr_fb[2 * e] = 0; p.rr_fb[2 * e + 1] = 1; }
    }
    if (p.stats) {
        const size_t ip = (size_t)e * p.P + port;
        p.st_soc_sum[ip] = 0.0; p.st_abs_e[ip] = 0.0; p.st_cnt[ip] = 0; p.st_nfin[ip] = 0;
    }
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
            s = (p.env_scn[e] + p.scn_stride) % p.S;
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
        if (threadIdx.x == 0) obs_header(p, row, s, 0, 0.0, p.env_t[(size_t)s * p.T].setpoint);
        for (int i = threadIdx.x; i < p.W; i += blockDim.x) row[p.series_off[i]] = obs_series_fetch(p, s, 0, i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.env_step[e] = 0; p.env_scn[e] = s; p.env_pot[e] = 0.0; p.env_usage[e] = 0.0; p.env_pot_prev[e] = 0.0;
        if (p.occ_n) p.occ_n[e] = 0;          // no EV is connected at t = 0 (ev2b_evlist.cuh)
        for (int k = 0; k < EV2B_KPI_COUNT; ++k) p.env_kpi[(size_t)e * EV2B_KPI_COUNT + k] = 0.0;
    }
    if (p.stats)
        for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
            const size_t ec = (size_t)e * p.C + c;
            p.cs_sat_sum[ec] = 0.0; p.cs_dcal[ec] = 0.0; p.cs_dcyc[ec] = 0.0; p.cs_served[ec] = 0; p.cs_em[ec] = 0;
        }
}

// ---- stock heuristic agents: agent.get_action(env) for every env, one CTA per env -------------------------
// ROUNDROBIN (heuristics.py:7-95).  The reference keeps an ordered list of waiting ports: newly waiting EVs are
// inserted at the FRONT while the ports are scanned in ascending order, EVs that are full or gone are removed, the first
// k = min(ceil(setpoint / average port power), len) entries are served and moved, in order, to the BACK.  Restated with
// one integer sort key per port (queue order == ascending key): a new port takes `front - 1 - (#new ports before it)`
// (so later ports end up nearer the front, as repeated insert(0, .) does), served ports take `back + rank`.  A port's
// position in the queue is the number of queued ports with a smaller key -- an O(P) shared-memory scan per port instead of
// the reference's O(P) list searches per port.
// CALAP (heuristics.py:98-150): per-port, stateless.
__global__ void agent_kernel(const Params p, int kind, double *act) {
    EV2B_DYNAMIC_SMEM(a_raw);                     // [P] keys, then [P] bytes: bit0 queued, bit1 new
    int *a_sm = reinterpret_cast<int *>(a_raw);
    int *skey = a_sm;
    unsigned char *sflag = reinterpret_cast<unsigned char *>(a_sm + p.P);
    __shared__ int s_new, s_len;
    const int e = blockIdx.x;
    const int t = p.env_step[e], s = p.env_scn[e];
    double *arow = act + (size_t)e * p.P;
    if (t >= p.T || kind == EV2B_AGENT_ZERO || kind == EV2B_AGENT_AFAP) {
        const double v = (t < p.T && kind == EV2B_AGENT_AFAP) ? 1.0 : 0.0;     // heuristics.py:161-166
        for (int i = threadIdx.x; i < p.P; i += blockDim.x) arow[i] = v;
        return;
    }
    if (kind == EV2B_AGENT_CALAP) {
        for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
            const size_t ip = (size_t)e * p.P + i;
            const uint4 h = p.hot[ip];
            double a = 0.0;
            if (hot_t_arr(h) <= t && t <= hot_t_dep(h)) {
                const CsStatic &cs = p.cs[p.port_cs[i]];
                const EvSpec *sp = p.spec + hot_spec(h);
                const double pmax = sp->pmax_ac;
                const double kw = pmax < cs.calap_kw ? pmax : cs.calap_kw;                            // min(cs, ev)  :123-124
                const double soc = p.cap[ip] / sp->B;
                const double steps = ceil((1.0 - soc) / (kw * p.period / 60.0 / sp->B));                          // :127-128
                if (soc < 1.0 && (double)hot_t_dep(h) - steps <= (double)t) a = 1.0;                               // :130-132
            }
            arow[i] = a;
        }
        return;
    }
    // ---- ROUNDROBIN
    if (threadIdx.x == 0) { s_new = 0; s_len = 0; }
    __syncthreads();
    int *gkey = p.rr_key + (size_t)e * p.P;
    const int front = p.rr_fb[2 * e], back = p.rr_fb[2 * e + 1];
    for (int i = threadIdx.x; i < p.P; i += blockDim.x) {                          // update_ev_buffer :33-52
        const size_t ip = (size_t)e * p.P + i;
        const uint4 h = p.hot[ip];
        bool waiting = false;
        if (hot_t_arr(h) <= t && t <= hot_t_dep(h)) waiting = p.cap[ip] / p.spec[hot_spec(h)].B < 1.0;   // get_soc() < 1
        const int key = gkey[i];
        const bool fresh = waiting && key == kRrAbsent;
        skey[i] = waiting ? key : kRrAbsent;
        sflag[i] = (unsigned char)((waiting ? 1 : 0) | (fresh ? 2 : 0));
        if (fresh) atomicAdd(&s_new, 1);
        if (waiting) atomicAdd(&s_len, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
        if (!(sflag[i] & 2)) continue;
        int before = 0;
        for (int j = 0; j < i; ++j) before += (sflag[j] >> 1) & 1;
        skey[i] = front - 1 - before;
    }
    __syncthreads();
    const double want = p.env_t[(size_t)s * p.T + t].setpoint * 1000.0 / p.rr_avg_power;   // :58-60
    const double cw = ceil(want);
    const int len = s_len;
    const int k = cw < (double)len ? (int)cw : len;                                // min(int(ceil(.)), len(buffer))
    for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
        double a = 0.0;
        int key = skey[i];
        if (sflag[i] & 1) {
            int rank = 0;
            for (int j = 0; j < p.P; ++j) rank += ((sflag[j] & 1) && skey[j] < key) ? 1 : 0;
            if (rank < k) {
                a = p.rr_share;                                                    // :84
                if (rank == k - 1 && want < (double)k) a = want - (double)rank;    // :85-86
                key = back + rank;                                                 // served EVs go to the back, in order
            }
        }
        arow[i] = a;
        gkey[i] = key;
    }
    if (threadIdx.x == 0) { p.rr_fb[2 * e] = front - s_new; p.rr_fb[2 * e + 1] = back + (k > 0 ? k : 0); }
}

// ---- get_statistics(env)  utils.py:12-123: one CTA per env, fixed-order sums by thread 0 ----------
__global__ void episode_stats_kernel(const Params p, double *out) {
    const int e = blockIdx.x;
    if (e >= p.E || threadIdx.x != 0) return;
    const double *kpi = p.env_kpi + (size_t)e * EV2B_KPI_COUNT;
    double *o = out + (size_t)e * EV2B_STAT_COUNT;
    double avg_sum = 0, dcal = 0, dcyc = 0; int n_cs = 0, em = 0;
    for (int c = 0; c < p.C; ++c) {
        const size_t ec = (size_t)e * p.C + c;
        if (p.cs_served[ec] > 0) { avg_sum += p.cs_sat_sum[ec] / (double)p.cs_served[ec]; ++n_cs; }   // utils.py:21-23
        dcal += p.cs_dcal[ec]; dcyc += p.cs_dcyc[ec]; em += p.cs_em[ec];
    }
    double rsum = 0, rmin = __longlong_as_double(0x7ff0000000000000LL); int n = 0;
    for (int port = 0; port < p.P; ++port) {
        const size_t ip = (size_t)e * p.P + port;
        for (int k = 0; k < p.st_nfin[ip]; ++k) { const double r = p.st_r[ip * p.Smax + k]; rsum += r; rmin = fmin(rmin, r); ++n; }
    }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double rmean = n ? rsum / n : nan;
    double var = 0;
    for (int port = 0; port < p.P; ++port) {
        const size_t ip = (size_t)e * p.P + port;
        for (int k = 0; k < p.st_nfin[ip]; ++k) { const double d = p.st_r[ip * p.Smax + k] - rmean; var += d * d; }
    }
    o[EV2B_STAT_EV_SERVED] = kpi[EV2B_KPI_EVS_SERVED];           o[EV2B_STAT_PROFITS] = kpi[EV2B_KPI_TOTAL_PROFITS];
    o[EV2B_STAT_ENERGY_CHARGED] = kpi[EV2B_KPI_ENERGY_CHARGED];  o[EV2B_STAT_ENERGY_DISCHARGED] = kpi[EV2B_KPI_ENERGY_DISCHARGED];
    o[EV2B_STAT_AVG_USER_SAT] = n_cs ? avg_sum / n_cs : nan;
    o[EV2B_STAT_TRACKER_VIOLATION] = kpi[EV2B_KPI_TRACKER_VIOLATION];
    o[EV2B_STAT_TRACKING_ERROR] = kpi[EV2B_KPI_TRACKING_ERROR];
    o[EV2B_STAT_ENERGY_TRACKING_ERROR] = kpi[EV2B_KPI_ENERGY_TRACKING_ERROR] * p.period / 60.0;   // utils.py:46
    o[EV2B_STAT_ENERGY_USER_SAT] = rmean;
    o[EV2B_STAT_STD_ENERGY_USER_SAT] = n ? sqrt(var / n) : nan;
    o[EV2B_STAT_MIN_ENERGY_USER_SAT] = n ? rmin : nan;
    o[EV2B_STAT_EMERGENCY_STEPS] = (double)em;
    o[EV2B_STAT_TR_OVERLOAD] = kpi[EV2B_KPI_TR_OVERLOAD];
    o[EV2B_STAT_DEGRADATION] = dcal + dcyc; o[EV2B_STAT_DEGRADATION_CAL] = dcal; o[EV2B_STAT_DEGRADATION_CYC] = dcyc;
    o[EV2B_STAT_TOTAL_REWARD] = kpi[EV2B_KPI_TOTAL_REWARD];
}

}  // namespace ev2b
