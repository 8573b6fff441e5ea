// ev2b_evlist.cuh -- the event-driven step kernel: visits CONNECTED EVs instead of ports.
//
// step_kernel (ev2b_device.cuh) runs one thread per (env, charger) and touches every port every step although, over an
// episode of the stock scenarios, 80 % of the ports are empty (30-40 % at the busy part, all of them at night); it is
// instruction-issue bound, not HBM bound.
// This kernel keeps, per env, the list of ports that currently hold an EV (`occ_list`, rewritten in place every step) and
// a per-scenario arrival schedule (`arr_list` bucketed by step), and does the reference's work in that order:
//
//   --  an env with nobody connected and nobody arriving takes evl_idle_step: base load, reward, KPI sums, observation
//   P0  prefetch of the per-env records (cp.async), (scenario, time)-only observation values
//   EV  one thread per CONNECTED EV (dense lanes): loads, Sigma-normalisation with the charger's other ports,
//       EV.step (the float64 battery model, ev_step_item), charger accounting, departure, observation tuple,
//       potential; per-port results go to shared memory                        ev_charger.py:114-233, ev.py:138-405
//   AR  one thread per ARRIVAL of step t+1 (from the schedule)                  ev2gym_env.py:399-417
//   CS  one thread per charger: power / amps / potential in port order, clamp   transformer.py:264-274, utils.py:779-789
//       (not with one port per charger: there the EV's own thread does it)
//   TR  warp 0: transformer sums (CSR) + overload; reward, KPI sums, step counter (same code path as step_kernel C)
//   LS  last warp: stable compaction of the kept EVs + arrivals back into the list (staged in shared memory meanwhile)
//
// An env is owned by a GROUP of G warps (G = 1, 2, 4; a 128-thread CTA holds 4 / G envs), so every barrier is a
// warp barrier (G = 1), a named barrier (G = 2) or __syncthreads (G = 4), and no phase leaves more than one warp of
// a group running alone for long.  The state arrays (hot / cap / exch) are exactly step_kernel's: the two kernels are
// interchangeable launch by launch (evl_rebuild_kernel re-derives the list from the hot words after step_kernel ran).
// Handles: every stock reward and state function that needs no distribution grid, without statistics mode and without
// the per-port optional outputs dep_sat, dep_cap, port_energy (action_mask is covered); everything else takes step_kernel.
// Sums are formed in a different (still fixed) order than step_kernel's, so float64 outputs agree to ~1e-15
// relative, not bitwise; battery levels, indices, counts and flags are identical.
#pragma once
#include "ev2b_device.cuh"

namespace ev2b {

constexpr int kEvlThreads = 128;
constexpr unsigned kEvlGone = 0xFFFFu;       // staging mark: this EV left during the step
enum { EvlProfit = 0, EvlSatExp, EvlCharged, EvlDischarged, EvlSatSum, EvlUsage, EvlPot, EvlCounts, EvlNSum };
static_assert(EvlNSum == 8, "warp_sum8 reduces exactly 8 quantities");

// Warp totals of 8 per-lane values in 9 exchanges instead of 40 (a reduce-scatter butterfly: at distance 16 every lane
// keeps 4 of the 8 quantities and hands the other 4 to its partner, at 8 it keeps 2, at 4 one; distances 2 and 1 are
// plain).  Returns, in lane L, the total of quantity 4*bit4(L) + 2*bit3(L) + bit2(L); the order of additions is fixed.
__device__ __forceinline__ double warp_sum8(const double (&q)[8], int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    double a[4], b[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double send = h16 ? q[j] : q[j + 4], keep = h16 ? q[j + 4] : q[j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double send = h8 ? a[j] : a[j + 2], keep = h8 ? a[j + 2] : a[j];
        b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const double send = h4 ? b[0] : b[1], keep = h4 ? b[1] : b[0];
    double c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}

__device__ __forceinline__ void evl_prefetch_l2(const void *ptr) {
#ifndef EV2B_SIMT_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
#ifdef EV2B_SIMT_EMU
    simt::cp_async(smem_dst, gmem_src, 4);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}

// Named barrier 1 + g of the CTA for the 64 threads of group g.  The id is a literal: with the id in a register ptxas
// reserves all 16 barriers for the CTA, which capped the SM at 3 resident CTAs (ncu: 21 % warps active).
__device__ __forceinline__ void evl_bar_sync64(int g) {
#ifdef EV2B_SIMT_EMU
    simt::bar_sync(1 + g, 64);
#else
    if (g == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else        asm volatile("bar.sync 2, 64;" ::: "memory");
#endif
}
template <int G>
__device__ __forceinline__ void evl_group_sync(int g) {
    static_assert(G == 1 || G == 2 || G * 32 == kEvlThreads, "group sizes: one warp, two warps, or the whole CTA");
    if (G == 1) __syncwarp();
    else if (G * 32 == kEvlThreads) __syncthreads();
    else evl_bar_sync64(g);
}

// What the env's reward function adds per EV that leaves this step (cv = its final battery level, des = the level it
// asked for, sat = its user satisfaction); summed into EvlSatExp and subtracted by the reward.
__device__ __forceinline__ double evl_departure_penalty(const Params &p, double cv, double des, double sat) {
    const int k = p.reward_kind;
    if (k == EV2B_REWARD_PROFIT_TR_USER || k == EV2B_REWARD_PROFIT_MAX) return 100.0 * exp(-10.0 * sat);   // reward.py:42,85
    if (k == EV2B_REWARD_SQTR_TR_USER) return 1000.0 * (1.0 - sat);                                         // reward.py:29-30
    if (k == EV2B_REWARD_V2G_PROFITMAX) return des > cv ? 100.0 * (des - cv) : 0.0;                         // reward.py:136-138
    if (k == EV2B_REWARD_V2G_PROFITMAX_V2 || k == EV2B_REWARD_PST_PROFITMAX_V2)
        return des > cv ? 0.05 * ((des - cv) * (des - cv)) : 0.0;                                           // reward.py:199-207
    return 0.0;
}
// V2G_profitmaxV2 family: an EV that stays connected but can no longer reach its desired level   reward.py:172-190
__device__ __forceinline__ double evl_unreachable_penalty(const Params &p, const EvSpec *sp, double cv, int steps_left) {
    if (p.reward_kind != EV2B_REWARD_V2G_PROFITMAX_V2 && p.reward_kind != EV2B_REWARD_PST_PROFITMAX_V2) return 0.0;
    const double des = __ldg(&sp->desired), pmax = __ldg(&sp->pmax_ac);
    const double min_steps = (des - cv) / (pmax / p.c60);
    if (!(min_steps > (double)steps_left)) return 0.0;
    const double gap = (des - ((double)(steps_left + 1) * pmax / p.c60)) - cv;
    return 0.05 * (gap * gap);
}

// Reward, KPI sums, step counter, done flag and observation header of env e (one thread; the same statements as
// step_kernel's phase C).  old = the env's KPI sums before the step (shared-memory prefetch, or env_kpi itself),
// q = the step's totals (Evl*), ovsum = sum of the transformers' overloads, pot_now = charge_power_potential[t].
struct EvlTotals {            // the step's totals of one env, indexed by Evl*
    double v[EvlNSum];
    __device__ __forceinline__ double operator[](int k) const { return v[k]; }
};
__device__ EV2B_NOINLINE void evl_finish_env(const Params &p, int e, int s, int tq, const double *old, double pot_now,
                                             double setpoint, double setpoint_next, double tr0_max_power,
                                             const EvlTotals q, double ovsum, int n_arr, int n_connected, bool want_obs) {
    const int cnts = (int)q[EvlCounts];
    unsigned status = (cnts >> 20) ? EV2B_ST_AMPS_OVERFLOW : 0u;
    const int n_dep = cnts & 0xFFFFF;
    const double usage = q[EvlUsage];                                     // current_power_usage[t]  ev2gym_env.py:375
    const double costs = q[EvlProfit];
    double reward = 0.0;
    if (p.reward_kind == EV2B_REWARD_SQ_TRACKING) {                       // reward.py:11-12
        const double m = setpoint < pot_now ? setpoint : pot_now;
        reward = -((m - usage) * (m - usage));
    } else if (p.reward_kind == EV2B_REWARD_PROFIT_TR_USER) {             // reward.py:36-44
        reward = costs - 100.0 * ovsum - q[EvlSatExp];
    } else if (p.reward_kind == EV2B_REWARD_PROFIT_MAX) {                 // reward.py:81-87
        reward = costs - q[EvlSatExp];
    } else if (p.reward_kind == EV2B_REWARD_SQTR_TR_USER) {               // reward.py:16-32
        double m = setpoint < pot_now ? setpoint : pot_now;
        if (tr0_max_power < m) m = tr0_max_power;                         // transformers[0].max_power[t]
        reward = -((m - usage) * (m - usage)) - 100.0 * ovsum - q[EvlSatExp];
    } else if (p.reward_kind == EV2B_REWARD_SQ_TRACKING_PENALTY) {        // reward.py:46-58
        const double m = setpoint < pot_now ? setpoint : pot_now;
        reward = -((m - usage) * (m - usage));
        if (usage == 0.0 && p.env_pot_prev[e] != 0.0) reward = reward - 100.0;   // potential[current_step-2]; 0 at t = 0
        p.env_pot_prev[e] = pot_now;
    } else if (p.reward_kind == EV2B_REWARD_SIMPLE) {                     // reward.py:60-65
        reward = -((setpoint - usage) * (setpoint - usage));
    } else if (p.reward_kind == EV2B_REWARD_MIN_TRACKER_SURPLUS) {        // reward.py:67-76
        if (setpoint < usage) reward -= (usage - setpoint) * (usage - setpoint);
        reward += usage;
    } else if (p.reward_kind == EV2B_REWARD_V2G_COSTS_SIMPLE) {           // reward.py:150-153
        reward = costs;
    } else if (p.reward_kind == EV2B_REWARD_V2G_PROFITMAX || p.reward_kind == EV2B_REWARD_V2G_PROFITMAX_V2 ||
               p.reward_kind == EV2B_REWARD_PST_PROFITMAX_V2) {           // reward.py:123-148, 155-213, 281-339
        reward = costs - q[EvlSatExp];
        if (p.reward_kind == EV2B_REWARD_PST_PROFITMAX_V2 && setpoint < usage) reward += 1000.0 * (setpoint - usage);
    }
    double *kpi = p.env_kpi + (size_t)e * EV2B_KPI_COUNT;
    kpi[EV2B_KPI_TOTAL_REWARD] = old[EV2B_KPI_TOTAL_REWARD] + reward;
    kpi[EV2B_KPI_TOTAL_PROFITS] = old[EV2B_KPI_TOTAL_PROFITS] + costs;
    kpi[EV2B_KPI_ENERGY_CHARGED] = old[EV2B_KPI_ENERGY_CHARGED] + q[EvlCharged];
    kpi[EV2B_KPI_ENERGY_DISCHARGED] = old[EV2B_KPI_ENERGY_DISCHARGED] + q[EvlDischarged];
    kpi[EV2B_KPI_TR_OVERLOAD] = old[EV2B_KPI_TR_OVERLOAD] + ovsum;
    kpi[EV2B_KPI_EVS_SERVED] = old[EV2B_KPI_EVS_SERVED] + (double)n_dep;
    kpi[EV2B_KPI_SAT_SUM] = old[EV2B_KPI_SAT_SUM] + q[EvlSatSum];
    const double d = setpoint - usage;                                    // utils.py:37-44
    kpi[EV2B_KPI_TRACKING_ERROR] = old[EV2B_KPI_TRACKING_ERROR] + d * d;
    kpi[EV2B_KPI_ENERGY_TRACKING_ERROR] = old[EV2B_KPI_ENERGY_TRACKING_ERROR] + fabs(d);
    if (usage > setpoint) kpi[EV2B_KPI_TRACKER_VIOLATION] = old[EV2B_KPI_TRACKER_VIOLATION] + (usage - setpoint);
    kpi[EV2B_KPI_EVS_SPAWNED] = old[EV2B_KPI_EVS_SPAWNED] + (double)n_arr;
    kpi[EV2B_KPI_INVALID_ACTIONS] = old[EV2B_KPI_INVALID_ACTIONS] + (double)(p.P - n_connected);   // every empty port  ev_charger.py:137-140
    kpi[EV2B_KPI_STEPS] = old[EV2B_KPI_STEPS] + 1.0;
    p.env_pot[e] = (tq < p.T) ? q[EvlPot] : 0.0;                          // ev2gym_env.py:424-426
    p.env_usage[e] = usage;
    p.env_step[e] = tq;
    if (tq >= p.T) status |= EV2B_ST_DONE;                                // ev2gym_env.py:460
    if (want_obs) obs_header(p, p.out.obs + (size_t)e * p.D, s, tq, usage, setpoint_next);
    if (p.out.reward) p.out.reward[e] = reward;
    if (p.out.total_costs) p.out.total_costs[e] = costs;
    if (p.out.status) p.out.status[e] = status;
}

// A step of an env with no EV connected and none arriving (half of the steps of the stock workplace scenarios: the site
// is empty at night).  Every per-EV and per-charger quantity is zero; what remains is the transformers' base load, the
// reward, the KPI sums and the observation.  No shared memory; the group's threads share the copies, its first thread
// does the per-env part straight from global memory -- after a barrier, because it advances env_step, which every thread
// of the group has just read (the SIMT emulator caught exactly that: a lane that ran ahead saw the next step).
template <int G>
__device__ __forceinline__ void evl_idle_step(const Params &p, int e, int t, int s, int g, int gtid, bool want_obs) {
    constexpr int GT = 32 * G;
    const int tq = t + 1;
    evl_group_sync<G>(g);
    if (want_obs) {
        float *obs_row = p.out.obs + (size_t)e * p.D;
        for (int i = gtid; i < p.W; i += GT) obs_row[p.series_off[i]] = obs_series_fetch(p, s, tq, i);
        if (p.obs_full) {
            const bool three = p.state_kind == EV2B_STATE_PUBLIC_PST;
#pragma unroll 1
            for (int i = gtid; i < p.P; i += GT) {
                float *o = obs_row + p.obs_slot[i];
                o[0] = 0.f; o[1] = 0.f;
                if (three) o[2] = 0.f;
            }
        }
    }
    if (p.out.action_mask && (p.mask_full || t == 0)) {
#pragma unroll 1
        for (int i = gtid; i < p.P; i += GT) p.out.action_mask[(size_t)e * p.P + i] = 0;
    }
    if (p.out.cs_power || p.out.cs_current) {
#pragma unroll 1
        for (int c = gtid; c < p.C; c += GT) {
            if (p.out.cs_power)   p.out.cs_power[(size_t)e * p.C + c] = 0.f;
            if (p.out.cs_current) p.out.cs_current[(size_t)e * p.C + c] = 0.f;
        }
    }
    if (gtid != 0) return;
    double ovsum = 0.0;
    for (int k = 0; k < p.Tr; ++k) {                                      // transformer.py:264-302 with no charger load
        const TrT tt = p.tr_t[((size_t)s * p.T + t) * p.Tr + k];
        const double ptot = (tt.infl + tt.solar) + 0.0;
        double ov = 0.0;
        if (ptot > tt.maxp + 0.0001 || ptot < tt.minp - 0.0001) ov = fabs(ptot - tt.maxp);
        if (p.out.tr_power)    p.out.tr_power[(size_t)e * p.Tr + k] = ptot;
        if (p.out.tr_overload) p.out.tr_overload[(size_t)e * p.Tr + k] = ov;
        ovsum += ov;
    }
    const EvlTotals q = {{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}};
    const EnvT *et = p.env_t + (size_t)s * p.T + t;
    evl_finish_env(p, e, s, tq, p.env_kpi + (size_t)e * EV2B_KPI_COUNT, p.env_pot[e], et->setpoint,
                   tq < p.T ? et[1].setpoint : 0.0, p.tr_t[((size_t)s * p.T + t) * p.Tr].maxp, q, ovsum, 0, 0, want_obs);
}

// Rebuilds occ_list / occ_n of envs [lo, hi) from the hot words (one warp per env): ports in ascending order.
__global__ void evl_rebuild_kernel(const Params p, int lo, int hi) {
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31u);
    const int e = lo + warp;
    if (e >= hi) return;
    const int t = p.env_step[e];
    uint16_t *lst = p.occ_list + (size_t)e * p.P;
    int base = 0;
    for (int p0 = 0; p0 < p.P; p0 += 32) {
        const int port = p0 + lane;
        bool occ = false;
        if (port < p.P && t < p.T) {
            const unsigned hx = p.hot[(size_t)e * p.P + port].x;
            occ = (int)(int16_t)(hx & 0xFFFFu) <= t && t <= (int)(int16_t)(hx >> 16);
        }
        const unsigned m = __ballot_sync(0xffffffffu, occ);
        if (occ) lst[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)port;
        base += __popc(m);
    }
    if (lane == 0) p.occ_n[e] = base;
}

// STG: every EV record of the env (hot words, battery level, exchanged energy, action) is copied to shared memory with
// cp.async before the EV loop instead of being loaded inside it (all of an env's DRAM requests in flight at once).
template <typename ActT, int NP, bool UNI, int G, bool STG>
__global__ void __launch_bounds__(kEvlThreads, 8) evl_step_kernel(const __grid_constant__ Params p) {
    EV2B_DYNAMIC_SMEM(smem_raw);
    constexpr int GT = 32 * G, EPB = kEvlThreads / GT;
    const int tid = threadIdx.x;
    const int g = tid / GT, gtid = tid - g * GT, lane = tid & 31, gw = gtid >> 5;
    const int e = p.env0 + (int)blockIdx.x * EPB + g;
    if (e >= p.env_end) return;                       // whole group: its barriers are its own

    unsigned char *sm = smem_raw + (size_t)g * p.v_stride;
    double *pw    = reinterpret_cast<double *>(sm);                  // [P] charger power contribution of the port's EV (kW)
    double *amp   = reinterpret_cast<double *>(sm + p.v_amp);        // [P] its actual current (A)
    double *pot   = reinterpret_cast<double *>(sm + p.v_pot);        // [P] its charge-power potential for step t+1
    double *csP   = reinterpret_cast<double *>(sm + p.v_csP);        // [C] charger power, for the transformer sums
    double *pre   = reinterpret_cast<double *>(sm + p.v_pre);        // [pre_stride] prefetched per-env records (kPre*)
    double *wsum  = reinterpret_cast<double *>(sm + p.v_wsum);       // [G][EvlNSum] per-warp partial sums (the last one holds counts)
    double *trov  = reinterpret_cast<double *>(sm + p.v_trov);       // [Tr] overload per transformer
    uint16_t *stage = reinterpret_cast<uint16_t *>(sm + p.v_stage);  // [P] by list position: port, or kEvlGone
    unsigned char *occ = sm + p.v_occ;                               // [P] the port holds per-port results this step
    uint4  *s_hot  = reinterpret_cast<uint4 *>(sm + p.v_shot);       // STG only, by list position: [P] hot words,
    double *s_cap  = reinterpret_cast<double *>(sm + p.v_scap);      //   [P] battery level,
    float  *s_exch = reinterpret_cast<float *>(sm + p.v_sexch);      //   [P] exchanged energy,
    ActT   *s_act  = reinterpret_cast<ActT *>(sm + p.v_sact);        //   [P] action (caller-supplied actions only)

    const ActT *actions = reinterpret_cast<const ActT *>(p.actions);
    const bool want_obs = (p.out.obs != nullptr) && (p.state_kind != EV2B_STATE_NONE);
    // The EV loop starts with a chain of dependent global loads (occ_n -> list -> hot words -> spec): the list is a
    // single buffer rewritten in place (phase LS), so its address needs nothing but e and this thread's first entry is
    // read together with the per-env scalars (entries at or beyond occ_n are stale and never used).
    const uint16_t *lst = p.occ_list + (size_t)e * p.P;
    const unsigned first = gtid < p.P ? (unsigned)lst[gtid] : 0u;
    const int t = p.env_step[e];
    if (t >= p.T) {                                   // step() on a finished env   ev2gym_env.py:343
        if (gtid == 0) {
            if (p.out.reward) p.out.reward[e] = 0.0;
            if (p.out.total_costs) p.out.total_costs[e] = 0.0;
            if (p.out.status) p.out.status[e] = EV2B_ST_DONE | EV2B_ST_WAS_DONE;
        }
        return;
    }
    const int s = p.env_scn[e];
    const int n_old = p.occ_n[e];
    const int tq = t + 1;
    const int a0 = p.arr_off[(size_t)s * (p.T + 2) + tq], a1 = p.arr_off[(size_t)s * (p.T + 2) + tq + 1];
    const int nArr = a1 - a0;
    if (n_old == 0 && nArr == 0) {                    // nobody connected, nobody arriving: the short path
        evl_idle_step<G>(p, e, t, s, g, gtid, want_obs);
        return;
    }
    float *obs_row = p.out.obs + (size_t)e * p.D;

    // ---- P0: prefetch, zero the per-port flags, (scenario, time)-only observation values ----------------------
    const bool ext_actions = p.agent_kind == EV2B_AGENT_EXTERNAL;
#pragma unroll 1
    for (int i = gtid; i < n_old; i += GT) {               // the list is staged in shared memory: phase LS rewrites it in place
        const int port = i == gtid ? (int)first : (int)lst[i];
        stage[i] = (uint16_t)port;
        const size_t ip = (size_t)e * p.P + port;
        if (STG) {                                         // this thread consumes exactly the records it requests here
            cp_async16(s_hot + i, p.hot + ip);
            cp_async8(s_cap + i, p.cap + ip);
            cp_async4(s_exch + i, p.exch + ip);
            if (ext_actions) {
                if (sizeof(ActT) == 8) cp_async8(s_act + i, actions + ip); else cp_async4(s_act + i, actions + ip);
            }
        } else if ((p.evl_pf & 1) && i != gtid) {          // later EVs of this thread: pull their lines into L2 now
            evl_prefetch_l2(p.hot + ip); evl_prefetch_l2(p.cap + ip); evl_prefetch_l2(p.exch + ip);
            if (ext_actions) evl_prefetch_l2(actions + ip);
        }
    }
    if (p.evl_pf & 2) {                                    // the env a later CTA of this launch will own: its rows into L2
        const int e2 = e + p.evl_pf_dist;
        if (e2 < p.env_end) {
            const size_t r0 = (size_t)e2 * p.P;
#pragma unroll 1
            for (int o = gtid * 128; o < p.P * 16; o += GT * 128) evl_prefetch_l2(reinterpret_cast<const char *>(p.hot + r0) + o);
#pragma unroll 1
            for (int o = gtid * 128; o < p.P * 8; o += GT * 128) evl_prefetch_l2(reinterpret_cast<const char *>(p.cap + r0) + o);
#pragma unroll 1
            for (int o = gtid * 128; o < p.P * 4; o += GT * 128) evl_prefetch_l2(reinterpret_cast<const char *>(p.exch + r0) + o);
            if (ext_actions)
    #pragma unroll 1
            for (int o = gtid * 128; o < p.P * (int)sizeof(ActT); o += GT * 128) evl_prefetch_l2(reinterpret_cast<const char *>(actions + r0) + o);
#pragma unroll 1
            for (int o = gtid * 128; o < p.P * 2; o += GT * 128) evl_prefetch_l2(reinterpret_cast<const char *>(p.occ_list + r0) + o);
        }
    }
#pragma unroll 1
    for (int i = gtid; i <= kPrePot; i += GT)
        cp_async8(pre + i, i < kPrePot ? p.env_kpi + (size_t)e * EV2B_KPI_COUNT + i : p.env_pot + e);
#pragma unroll 1
    for (int i = gtid; i < 2 + 2 * p.Tr; i += GT) {
        if (i < 2) { if (t + i < p.T) cp_async8(pre + kPreSet + i, &p.env_t[(size_t)s * p.T + t + i].setpoint); }
        else cp_async16(pre + kPreTr + 2 * (i - 2), reinterpret_cast<const double *>(p.tr_t + ((size_t)s * p.T + t) * p.Tr) + 2 * (i - 2));
    }
    const EnvT et0 = p.env_t[(size_t)s * p.T + t];
    if (NP == 1) {                                    // one port per charger: the EV's thread is the charger's thread (no CS phase)
#pragma unroll 1
        for (int i = gtid; i < p.C; i += GT) {
            csP[i] = 0.0;
            if (p.out.cs_power)   p.out.cs_power[(size_t)e * p.C + i] = 0.f;
            if (p.out.cs_current) p.out.cs_current[(size_t)e * p.C + i] = 0.f;
        }
    } else {
#pragma unroll 1
        for (int i = gtid; i < (p.P + 3) >> 2; i += GT) reinterpret_cast<unsigned *>(occ)[i] = 0u;
    }
    if (want_obs) {
        for (int i = gtid; i < p.W; i += GT) obs_row[p.series_off[i]] = obs_series_fetch(p, s, tq, i);
        if (p.obs_full) {                             // the caller's buffer does not hold last step's rows: clear every tuple
            const bool three = p.state_kind == EV2B_STATE_PUBLIC_PST;
#pragma unroll 1
            for (int i = gtid; i < p.P; i += GT) {
                float *o = obs_row + p.obs_slot[i];
                o[0] = 0.f; o[1] = 0.f;
                if (three) o[2] = 0.f;
            }
        }
    }
    // action_mask is maintained incrementally like the observation tuples: the caller's buffer still holds last step's
    // row, only ports whose EV arrives or leaves change.  A new buffer (mask_full) or a new episode (t == 0: the row may
    // hold the previous episode's terminal mask) rewrites the row.                                  ev2gym_env.py:452-457
    uint8_t *mask_row = p.out.action_mask ? p.out.action_mask + (size_t)e * p.P : nullptr;
    if (mask_row && (p.mask_full || t == 0)) {
#pragma unroll 1
        for (int i = gtid; i < p.P; i += GT) mask_row[i] = 0;
    }
    if (STG) cp_async_wait_all();                          // (also completes the per-env records requested above)
    evl_group_sync<G>(g);

    // ---- EV: one thread per connected EV ------------------------------------------------------------------------
    double aProfit = 0, aSatExp = 0, aCh = 0, aDis = 0, aSat = 0, aUsage = 0, aPot = 0;
    int nDep = 0;
    bool overflow = false;
#pragma unroll 1
    for (int i = gtid; i < n_old; i += GT) {
        const int port = stage[i];
        const size_t ip = (size_t)e * p.P + port;
        const uint4 h = STG ? s_hot[i] : p.hot[ip];
        double cv = STG ? s_cap[i] : p.cap[ip];
        float exch_new = STG ? s_exch[i] : p.exch[ip];
        const double a = (STG && ext_actions) ? (double)s_act[i] : agent_action<ActT>(p, actions, ip, t);
        const unsigned hx = NP == 2 ? p.hot[ip ^ 1].x : 0u;
        const double am_raw = NP == 2 ? agent_action<ActT>(p, actions, ip ^ 1, t) : 0.0;   // same 32 B sector as `a`
        const int c = NP == 1 ? port : (NP == 2 ? port >> 1 : p.port_cs[port]);
        const CsStatic &cs = cs_of<UNI>(p, c);
        // Sigma over the charger's occupied ports, in port order (python sum())   ev_charger.py:137-149
        double sum = 0.0;
        if (NP == 1) {
            sum = sum + a;
        } else if (NP == 2) {
            // the other port of this charger is ip ^ 1 (P is even and port offsets are 2c); its action was loaded with ours
            const bool occ_m = (int)(int16_t)(hx & 0xFFFFu) <= t && t <= (int)(int16_t)(hx >> 16);
            const double am = occ_m ? am_raw : 0.0;
            sum = sum + ((port & 1) ? am : a);
            sum = sum + ((port & 1) ? a : am);
        } else {
            const int p0 = cs.port_off, n = cs.n_ports;
            for (int j = 0; j < n; ++j) {
                double aj = a;
                if (p0 + j != port) {
                    const size_t ij = (size_t)e * p.P + p0 + j;
                    const unsigned hx = p.hot[ij].x;
                    const bool occ_j = (int)(int16_t)(hx & 0xFFFFu) <= t && t <= (int)(int16_t)(hx >> 16);
                    aj = occ_j ? agent_action<ActT>(p, actions, ij, t) : 0.0;
                }
                sum = sum + aj;
            }
        }
        double an = a;
        {   // if sum > 1: a / sum; if sum < -1: -a / sum  (:143-149) -- one division for both branches
            const bool over = sum > 1.0, under = sum < -1.0;
            if (over || under) an = (over ? an : -an) / sum;
        }
        double energy = 0.0, act_amps = 0.0, pwv = 0.0;
        if (an != 0.0) {
            double cap = cv;
            bool em_cross;
            const bool active = ev_step_item<false>(p, cs, h.z, h.w, an, cap, energy, act_amps, em_cross);
            if (active) {                                                 // amps == 0: nothing changes  ev.py:158-163
                cv = cap;
                p.cap[ip] = cv;
                exch_new = exch_new + (float)energy;                     // total_energy_exchanged  ev.py:178
                p.exch[ip] = exch_new;
            }
            const double ae = fabs(energy);
            if (an > 0.0) { aProfit += ae * et0.cp; aCh += ae; }          // ev_charger.py:178-179
            else          { aProfit += ae * et0.dp; aDis += ae; }         // ev_charger.py:194-195
            pwv = ev2b_div_c(energy * 60.0, p.period, p.rperiod);         // :180,196
        }
        if (NP != 1) { pw[port] = pwv; amp[port] = act_amps; }
        double potv = 0.0;
        if (t >= hot_t_dep(h)) {                                          // departure  ev_charger.py:209-224, ev.py:199-214
            const double des = __ldg(&p.spec[hot_spec(h)].desired);
            const double sat = (cv < des - 0.001) ? cv / des : 1.0;
            aSatExp += evl_departure_penalty(p, cv, des, sat);
            aSat += sat;
            ++nDep;
            stage[i] = (uint16_t)kEvlGone;                                // (stage[i] already holds the port of an EV that stays)
            if (mask_row) mask_row[port] = 0;
            if (want_obs) {
                float *o = obs_row + p.obs_slot[port];
                o[0] = 0.f; o[1] = 0.f;
                if (p.state_kind == EV2B_STATE_PUBLIC_PST) o[2] = 0.f;
            }
        } else {
            if (mask_row) mask_row[port] = 1;
            const EvSpec *sp = p.spec + hot_spec(h);
            const double B = __ldg(&sp->B);
            aSatExp += evl_unreachable_penalty(p, sp, cv, hot_t_dep(h) - tq);
            if (cv < B && hot_t_dep(h) > tq) potv = __ldg(&p.pot_kw[hot_spec(h) * p.n_cls + cs.cls]);   // utils.py:766-777
            if (want_obs) {                                               // state.py:37-57, 85-102, 137-151
                float *o = obs_row + p.obs_slot[port];
                if (p.state_kind == EV2B_STATE_PUBLIC_PST) {
                    o[0] = (cv == B) ? 1.f : 0.5f;
                    o[1] = exch_new;
                    o[2] = (float)(tq - hot_t_arr(h));
                } else {
                    o[0] = (float)ev2b_div_c(cv, B, __ldg(&sp->rB));
                    o[1] = (float)(hot_t_dep(h) - tq);
                }
            }
        }
        if (NP == 1) {                                                    // charger == port: accounting in place
            const double rP = 0.0 + pwv, rA = 0.0 + act_amps;
            if (rA - 0.0001 > cs.imax) overflow = true;                   // ev_charger.py:203-205
            double rPot = potv;
            if (rPot > cs.max_power) rPot = cs.max_power;                 // utils.py:779-789
            else if (rPot < cs.min_power) rPot = 0.0;
            csP[port] = rP;
            aUsage += rP; aPot += rPot;
            if (p.out.cs_power)   p.out.cs_power[(size_t)e * p.C + port] = (float)rP;
            if (p.out.cs_current) p.out.cs_current[(size_t)e * p.C + port] = (float)rA;
        } else {
            pot[port] = potv;
            occ[port] = 1;
        }
    }
    evl_group_sync<G>(g);

    // ---- AR: arrivals of step t+1, one thread each (the highest threads: they had the least EV work) ---------
#pragma unroll 1
    for (int k = GT - 1 - gtid; k < nArr; k += GT) {                      // ev2gym_env.py:399-417, ev_charger.py:266-285
        const unsigned u = p.arr_list[a0 + k];
        const int port = (int)(u & 0xFFFFu), cur = (int)(u >> 16);
        const size_t ip = (size_t)e * p.P + port;
        const SessRec r = p.sess[((size_t)s * p.P + port) * p.Smax + cur];
        p.hot[ip] = r.hot;
        p.cap[ip] = r.cap0;
        p.exch[ip] = 0.f;
        const int c = NP == 1 ? port : (NP == 2 ? port >> 1 : p.port_cs[port]);
        const CsStatic &cs = cs_of<UNI>(p, c);
        const EvSpec *sp = p.spec + hot_spec(r.hot);
        const double B = __ldg(&sp->B);
        double potv = 0.0;
        aSatExp += evl_unreachable_penalty(p, sp, r.cap0, hot_t_dep(r.hot) - tq);
        if (r.cap0 < B && hot_t_dep(r.hot) > tq) potv = __ldg(&p.pot_kw[hot_spec(r.hot) * p.n_cls + cs.cls]);
        if (NP == 1) {                                                    // (an EV that left this very port in step t added 0)
            double rPot = potv;
            if (rPot > cs.max_power) rPot = cs.max_power;
            else if (rPot < cs.min_power) rPot = 0.0;
            aPot += rPot;
        } else {
            if (!occ[port]) { pw[port] = 0.0; amp[port] = 0.0; }         // (an EV may have left this very port in step t)
            pot[port] = potv;
            occ[port] = 1;
        }
        if (mask_row) mask_row[port] = 1;
        if (want_obs) {
            float *o = obs_row + p.obs_slot[port];
            if (p.state_kind == EV2B_STATE_PUBLIC_PST) {
                o[0] = (r.cap0 == B) ? 1.f : 0.5f;
                o[1] = 0.f;
                o[2] = (float)(tq - hot_t_arr(r.hot));
            } else {
                o[0] = (float)ev2b_div_c(r.cap0, B, __ldg(&sp->rB));
                o[1] = (float)(hot_t_dep(r.hot) - tq);
            }
        }
    }
    evl_group_sync<G>(g);

    // ---- CS: one thread per charger, ports in order -------------------------------------------------------------
#pragma unroll 1
    for (int c = NP == 1 ? p.C : gtid; c < p.C; c += GT) {
        const CsStatic &cs = cs_of<UNI>(p, c);
        const int p0 = NP > 0 ? c * NP : cs.port_off, n = NP > 0 ? NP : cs.n_ports;
        double rP = 0, rA = 0, rPot = 0;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            if (occ[p0 + j]) { rP += pw[p0 + j]; rA += amp[p0 + j]; rPot += pot[p0 + j]; }
            if (rA - 0.0001 > cs.imax) overflow = true;                   // ev_charger.py:203-205
        }
        if (rPot > cs.max_power) rPot = cs.max_power;                     // utils.py:779-789
        else if (rPot < cs.min_power) rPot = 0.0;
        csP[c] = rP;
        aUsage += rP; aPot += rPot;
        if (p.out.cs_power)   p.out.cs_power[(size_t)e * p.C + c] = (float)rP;
        if (p.out.cs_current) p.out.cs_current[(size_t)e * p.C + c] = (float)rA;
    }
    // per-warp partial sums (fixed butterfly); the group total is formed warp by warp in the reward phase
    {
        const double q[EvlNSum] = {aProfit, aSatExp, aCh, aDis, aSat, aUsage, aPot,
                                   (double)nDep + (overflow ? 1048576.0 : 0.0)};   // departures < 2^20; overflow votes above
        const double tot = warp_sum8(q, lane);
        if ((lane & 3) == 0) wsum[gw * EvlNSum + (lane >> 2)] = tot;
    }
    cp_async_wait_all();
    evl_group_sync<G>(g);

    // ---- LS: the group's last warp writes next step's list: kept EVs in list order, then the arrivals ----------
    if (gw == G - 1) {
        uint16_t *nxt = p.occ_list + (size_t)e * p.P;      // in place: every thread staged the old list before the first barrier
        int base = 0;
#pragma unroll 1
        for (int i0 = 0; i0 < n_old; i0 += 32) {
            const int i = i0 + lane;
            const unsigned v = i < n_old ? (unsigned)stage[i] : kEvlGone;
            const unsigned m = __ballot_sync(0xffffffffu, v != kEvlGone);
            if (v != kEvlGone) nxt[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)v;
            base += __popc(m);
        }
#pragma unroll 1
        for (int k = lane; k < nArr; k += 32) nxt[base + k] = (uint16_t)(p.arr_list[a0 + k] & 0xFFFFu);
        if (lane == 0) p.occ_n[e] = base + nArr;
    }
    if (gw != 0) return;

    // ---- TR: transformer sums + overload (warp 0), same lane split as step_kernel's phase B ---------------------
    {
        const int lg = p.tr_lg, nseg = 1 << lg, per = 32 >> lg;   // 2^lg lanes share one transformer (host: largest with 2^lg * Tr <= 32)
        for (int k0 = 0; k0 < p.Tr; k0 += per) {
            const int k = k0 + (lane >> lg), seg = lane & (nseg - 1);
            double sp_ = 0.0;
            if (k < p.Tr) {
                const int i0 = p.tr_cs_off[k], n_k = p.tr_cs_off[k + 1] - i0;
                const int chunk = (n_k + nseg - 1) >> lg;
                const int lo = seg * chunk, hi = min(n_k, lo + chunk);
                for (int i = lo; i < hi; ++i) sp_ += csP[p.tr_cs_idx[i0 + i]];
            }
            for (int o = nseg >> 1; o > 0; o >>= 1) sp_ += __shfl_xor_sync(0xffffffffu, sp_, o);
            if (seg == 0 && k < p.Tr) {                                   // transformer.py:264-302
                const double *tq4 = pre + kPreTr + 4 * k;
                const double ptot = (tq4[0] + tq4[1]) + sp_;
                double ov = 0.0;
                if (ptot > tq4[2] + 0.0001 || ptot < tq4[3] - 0.0001) ov = fabs(ptot - tq4[2]);
                trov[k] = ov;
                if (p.out.tr_power)    p.out.tr_power[(size_t)e * p.Tr + k] = ptot;
                if (p.out.tr_overload) p.out.tr_overload[(size_t)e * p.Tr + k] = ov;
            }
        }
        __syncwarp();
    }
    // ---- reward, KPI sums, step counter: one lane ------------------------------------------------------------------
    if (lane == 0) {
        EvlTotals q;
#pragma unroll
        for (int k = 0; k < EvlNSum; ++k) {
            double v = wsum[k];
            for (int w = 1; w < G; ++w) v += wsum[w * EvlNSum + k];
            q.v[k] = v;
        }
        double ovsum = 0.0;
        for (int k = 0; k < p.Tr; ++k) ovsum += trov[k];
        evl_finish_env(p, e, s, tq, pre, pre[kPrePot], pre[kPreSet], pre[kPreSetNext], pre[kPreTr + 2], q, ovsum, nArr,
                       n_old, want_obs);
    }
}

}  // namespace ev2b
