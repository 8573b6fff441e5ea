// ev2b_evlist.cuh -- the event-driven step kernel: visits CONNECTED EVs instead of ports.
//
// step_kernel (ev2b_device.cuh) runs one thread per (env, charger) and touches every port every step although, over an
// episode of the stock scenarios, 80 % of the ports are empty (30-40 % at the busy part, all of them at night).
// This kernel keeps, per env, the list of ports that currently hold an EV (`occ_list`, rewritten in place every step) and
// a per-scenario arrival schedule (`arr_list` bucketed by step; the bucket of step t+1 is named by the (scenario, t)
// record EnvT itself), and does the reference's work in that order:
//
//   P0  per-env scalars, cp.async prefetch of the per-env records (KPI sums, potential, transformer rows), fills of the
//       per-step outputs, (scenario, time)-only observation values
//   EV  one thread per CONNECTED EV (dense lanes): loads, Sigma-normalisation with the charger's other ports,
//       EV.step (the float64 battery model, ev_step_item), charger accounting, departure, observation tuple,
//       potential; per-port results go to shared memory                        ev_charger.py:114-233, ev.py:138-405
//   AR  one thread per ARRIVAL of step t+1 (from the schedule)                  ev2gym_env.py:399-417
//   CS  one thread per charger: power / amps / potential in port order, clamp   transformer.py:264-274, utils.py:779-789
//       (not with one port per charger: there the EV's own thread does it)
//   LS  last warp: the list of step t+1, IN PORT ORDER (ballot compaction of per-port "connected at t+1" flags), so that
//       neighbouring threads of the EV phase work on neighbouring ports and their loads / stores share memory sectors
//   TR  warp 0: transformer sums (CSR) + overload; distribution-grid power flow; then reward, the 13 KPI sums, step
//       counter, done flag and observation header, ONE LANE PER QUANTITY (lane k owns KPI k: a load, an add, a store --
//       not thirteen dependent read-modify-writes by one thread)
//
// An env with nobody connected and nobody arriving (half of the steps of the stock workplace scenarios: the site is
// empty at night) runs the same code with the EV / AR / CS / LS phases and their barriers skipped, on warp 0 of the
// group alone: base load, reward, KPI sums, observation.
//
// An env is owned by a GROUP of G warps (G = 1, 2, 4; a 128-thread CTA holds 4 / G envs), so every barrier is a
// warp barrier (G = 1), a named barrier (G = 2) or a CTA barrier (G = 4).  The state arrays (hot / cap / exch) are
// exactly step_kernel's: the two kernels are interchangeable launch by launch (evl_rebuild_kernel re-derives the list from
// the hot words after step_kernel ran).  The kernel covers EVERY feature of the C ABI: all stock rewards and state
// functions, statistics mode, the distribution grid and the per-port optional outputs live in the HEAVY instantiation
// (so the lean one does not pay for them).  Sums are formed in a different (still fixed) order than step_kernel's, so
// float64 outputs agree to ~1e-15 relative, not bitwise; battery levels, indices, counts and flags are identical.
#pragma once
#include "ev2b_device.cuh"

namespace ev2b {

constexpr int kEvlThreads = 128;             // threads per CTA (a second instantiation with 32 serves multi-wave launches)
// prefetch area of this kernel (doubles): [0,13) KPI sums, [13] charge_power_potential[t] (kPrePot), then
constexpr int kEvlPotPrev = 14;              // charge_power_potential[t-1]
constexpr int kEvlSet = 15, kEvlSetNext = 16;   // power_setpoints[t], [t+1]
constexpr int kEvlTr = 18;                   // [18 + 4k, +4) TrT of transformer k (16 B aligned)
enum { EvlProfit = 0, EvlSatExp, EvlCharged, EvlDischarged, EvlSatSum, EvlUsage, EvlPot, EvlCounts, EvlNSum };
#ifndef EV2B_EVL_PF
#define EV2B_EVL_PF 1         // L2 prefetch of the next iteration's EV state: 0 none, 1 hot + cap + exch + action, 2 hot only, 3 no action
                              // (us per launch c3 / c4 / c5 with the port-ordered list: 0: 18.61 / 32.38 / 55.54, 1: 18.53 / 31.65 / 53.74,
                              //  profiles/r2_ab_build_modes.jsonl; before the list was ordered the prefetch was worth 9 % on c3)
#endif
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

// Barriers of a group.  G = 2: named barrier 1 + g, 64 threads (the ids are literals: with the id in a register ptxas
// reserves all 16 barriers for the CTA, which capped the SM at 3 resident CTAs -- ncu: 21 % warps active).  G = 4: the
// group is the CTA: __syncthreads.  `arrive` does not wait: the idle path of G = 2 uses it so that the warp with nothing
// to do leaves at once while warp 0 still learns that it has read the env's step counter before advancing it.  (With
// G = 4 warp 0 reads the env scalars and publishes them through shared memory instead, see evl_env_step: `bar.arrive 1, 128`
// by three warps that then exit, paired with a `bar.sync 1, 128` of warp 0, passed the emulator and HUNG on the B200.)
template <int G>
__device__ __forceinline__ void evl_group_sync(int g) {
    static_assert(G == 1 || G == 2 || G * 32 == kEvlThreads, "group sizes: one warp, two warps, or the whole 128-thread CTA");
    if (G == 1) { __syncwarp(); return; }
    if (G * 32 == kEvlThreads) { __syncthreads(); return; }
#ifdef EV2B_SIMT_EMU
    simt::bar_sync(1 + g, 32 * G);
#else
    if (g == 0) asm volatile("bar.sync 1, 64;" ::: "memory"); else asm volatile("bar.sync 2, 64;" ::: "memory");
#endif
}
// The idle path's "I have read the step counter" of a warp other than warp 0.
template <int G>
__device__ __forceinline__ void evl_group_arrive(int g) {
    if (G == 1 || G * 32 == kEvlThreads) return;      // (whole-CTA groups publish the env scalars through shared memory instead)
#ifdef EV2B_SIMT_EMU
    simt::bar_arrive(1 + g, 32 * G);
#else
    if (g == 0) asm volatile("bar.arrive 1, 64;" ::: "memory"); else asm volatile("bar.arrive 2, 64;" ::: "memory");
#endif
}

// What the env's reward function adds per EV that leaves this step (cv = its final battery level, des = the level it
// asked for, sat = its user satisfaction); summed into EvlSatExp and subtracted by the reward.
template <bool HEAVY>
__device__ __forceinline__ double evl_departure_penalty(const Params &p, double cv, double des, double sat) {
    const int k = p.reward_kind;
    if (k == EV2B_REWARD_PROFIT_TR_USER || k == EV2B_REWARD_PROFIT_MAX) return 100.0 * exp(-10.0 * sat);   // reward.py:42,85
    if (k == EV2B_REWARD_SQTR_TR_USER) return 1000.0 * (1.0 - sat);                                         // reward.py:29-30
    if (k == EV2B_REWARD_V2G_PROFITMAX) return des > cv ? 100.0 * (des - cv) : 0.0;                         // reward.py:136-138
    if (k == EV2B_REWARD_V2G_PROFITMAX_V2 || k == EV2B_REWARD_PST_PROFITMAX_V2 || (HEAVY && k == EV2B_REWARD_GRID_PROFITMAX_V2))
        return des > cv ? 0.05 * ((des - cv) * (des - cv)) : 0.0;                                           // reward.py:199-207
    if (HEAVY && (k == EV2B_REWARD_GRID_FULL || k == EV2B_REWARD_GRID_SIMPLE)) return (cv - des) * (cv - des);   // -user_costs  reward.py:99-102
    return 0.0;
}
// V2G_profitmaxV2 family: an EV that stays connected but can no longer reach its desired level   reward.py:172-190
template <bool HEAVY>
__device__ __forceinline__ double evl_unreachable_penalty(const Params &p, const EvSpec *sp, double cv, int steps_left) {
    if (p.reward_kind != EV2B_REWARD_V2G_PROFITMAX_V2 && p.reward_kind != EV2B_REWARD_PST_PROFITMAX_V2 &&
        !(HEAVY && p.reward_kind == EV2B_REWARD_GRID_PROFITMAX_V2)) return 0.0;
    const double des = __ldg(&sp->desired), pmax = __ldg(&sp->pmax_ac);
    const double min_steps = (des - cv) / (pmax / p.c60);
    if (!(min_steps > (double)steps_left)) return 0.0;
    const double gap = (des - ((double)(steps_left + 1) * pmax / p.c60)) - cv;
    return 0.05 * (gap * gap);
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

template <typename T> struct alignas(2 * sizeof(T)) Pair { T x, y; };      // two neighbouring actions

// Observation tuple of the EV on `port` after the step (cv = battery level, h = its hot words)   state.py:37-57, 85-102, 137-151, 262-270
template <bool HEAVY>
__device__ __forceinline__ void evl_obs_tuple(const Params &p, float *obs_row, int port, int c, const uint4 &h, double cv,
                                              double B, double rB, double exch, int tq) {
    float *o = obs_row + p.obs_slot[port];
    if (HEAVY && p.state_kind == EV2B_STATE_V2G_GRID) {
        o[0] = (float)cv;
        o[1] = (float)(hot_t_dep(h) - tq + 1);
        o[2] = (float)__ldg(&p.cs_tr[c]);                     // cs.connected_bus
    } else if (p.state_kind == EV2B_STATE_PUBLIC_PST) {
        o[0] = (cv == B) ? 1.f : 0.5f;
        o[1] = (float)exch;
        o[2] = (float)(tq - hot_t_arr(h));
    } else {
        const float soc = (float)ev2b_div_c(cv, B, rB), left = (float)(hot_t_dep(h) - tq);
        if (p.obs_pairs) *reinterpret_cast<float2 *>(o) = make_float2(soc, left);
        else { o[0] = soc; o[1] = left; }
    }
}
__device__ __forceinline__ void evl_obs_clear(const Params &p, float *obs_row, int port) {
    float *o = obs_row + p.obs_slot[port];
    if (p.obs_pairs) { *reinterpret_cast<float2 *>(o) = make_float2(0.f, 0.f); return; }
    o[0] = 0.f; o[1] = 0.f;
    if (p.state_kind == EV2B_STATE_PUBLIC_PST || p.state_kind == EV2B_STATE_V2G_GRID) o[2] = 0.f;
}

// Starts the HBM -> L2 transfer of the state one EV's thread will load (hot words, battery level, energy exchanged, the
// action): issued one loop iteration ahead (and for the first iteration from the prologue), so that the dependent gather
// list entry -> port state, the longest wait of the EV loop, finds its sectors in L2.
template <typename ActT>
__device__ __forceinline__ void evl_prefetch_ev(const Params &p, const ActT *actions, size_t ip) {
#if EV2B_EVL_PF >= 1
    prefetch_l2(p.hot + ip);
#endif
#if EV2B_EVL_PF == 1 || EV2B_EVL_PF == 3
    prefetch_l2(p.cap + ip);
    prefetch_l2(p.exch + ip);
#endif
#if EV2B_EVL_PF == 1
    if (p.agent_kind == EV2B_AGENT_EXTERNAL) prefetch_l2(actions + ip);
#endif
}

// One step of env e by its group (g = group in the CTA, sm = the group's shared memory).  KSTEP: called from the k-step
// loop of evl_step_kernel (every warp of the group meets at a barrier after each call, so the idle path's arrive / wait
// pair is not needed); `actions` = the action tensor of this step, `first_it` = first step of the launch (only then may
// the caller's obs / mask buffers need a full rewrite).  Returns the env's step counter after the call (T + 1: the env
// was already finished).
template <typename ActT, int NP, bool UNI, int G, bool HEAVY, bool KSTEP>
__device__ __forceinline__ int evl_env_step(const Params &p, unsigned char *sm, const int e, const int g, const int gtid,
                                            const ActT *actions, const bool first_it) {
    constexpr int GT = 32 * G;
    const int lane = gtid & 31, gw = gtid >> 5;
    double *pw    = reinterpret_cast<double *>(sm);                  // [P] charger power contribution of the port's EV (kW)
    double *amp   = reinterpret_cast<double *>(sm + p.v_amp);        // [P] its actual current (A)
    double *pot   = reinterpret_cast<double *>(sm + p.v_pot);        // [P] its charge-power potential for step t+1
    double *csP   = reinterpret_cast<double *>(sm + p.v_csP);        // [C] charger power, for the transformer sums
    double *pre   = reinterpret_cast<double *>(sm + p.v_pre);        // [pre_stride] prefetched per-env records (kPre*)
    double *wsum  = reinterpret_cast<double *>(sm + p.v_wsum);       // [G][EvlNSum] per-warp partial sums (the last one holds counts)
    unsigned char *occ = sm + p.v_occ;                               // [P] bit 0: the port holds per-port results this step;
                                                                     //     bit 3: an EV is connected to it at step t+1;
                                                                     //     bit 1 / 2 (statistics): an EV left it / was finalised on it
    double *trp   = reinterpret_cast<double *>(sm + p.v_trp);        // HEAVY: [Tr] transformer power (grid: bus EV power)
    double2 *pfv  = reinterpret_cast<double2 *>(sm + p.v_pfv);       // HEAVY: [3][n_bus] power-flow scratch S, V, lambda
    double *dsat  = reinterpret_cast<double *>(sm + p.v_dsat);       // HEAVY statistics, several ports per charger: [P] satisfaction,
    double *dcal  = reinterpret_cast<double *>(sm + p.v_dcal);       //   calendar and
    double *dcyc  = reinterpret_cast<double *>(sm + p.v_dcyc);       //   cyclic degradation of the EVs finalised this step

    const bool want_obs = (p.out.obs != nullptr) && (p.state_kind != EV2B_STATE_NONE);
    const bool obs_full = p.obs_full && first_it, mask_full = p.mask_full && first_it;
    // The EV loop starts with a chain of dependent global loads (list -> hot words -> spec): the list is a single buffer
    // rewritten in place (phase LS), so its address needs nothing but e and this thread's first entry is read together
    // with the per-env scalars (entries at or beyond occ_n are stale and never used).
    const uint16_t *lst = p.occ_list + (size_t)e * p.P;
    const unsigned first = gtid < p.P ? (unsigned)lst[gtid] : 0u;
    // volatile: read exactly once.  Warp 0 uses n_old in the KPI update while the group's last warp is storing the new
    // occ_n; a plain load may be re-issued by the compiler at that later use (it was, under register pressure, with
    // G = 4: invalid_actions went wrong on the B200 while every other quantity was right -- round 2, test_gpu_evlist).
    int t, s, n_old;
    if (G * 32 == kEvlThreads) {
        // The group is the whole CTA: warp 0 reads the scalars and publishes them through shared memory, so the other
        // warps never look at env_step themselves and the idle path needs no handshake before warp 0 advances it (every
        // thread meets at THIS __syncthreads; compute-sanitizer's synccheck rejected the earlier variant, in which the
        // other warps synchronised at a different call site and exited).
        int *hdr = reinterpret_cast<int *>(sm + p.v_hdr);
        if (gtid == 0) {
            hdr[0] = *reinterpret_cast<const volatile int *>(p.env_step + e);
            hdr[1] = *reinterpret_cast<const volatile int *>(p.env_scn + e);
            hdr[2] = *reinterpret_cast<const volatile int *>(p.occ_n + e);
        }
        __syncthreads();
        t = hdr[0]; s = hdr[1]; n_old = hdr[2];
    } else {
        t = *reinterpret_cast<const volatile int *>(p.env_step + e);
        s = *reinterpret_cast<const volatile int *>(p.env_scn + e);
        n_old = *reinterpret_cast<const volatile int *>(p.occ_n + e);
    }
    if (gtid < n_old) evl_prefetch_ev<ActT>(p, actions, (size_t)e * p.P + first);
    if (gw == 0) {                                     // KPI sums, potential[t], potential[t-1]: need nothing but e
        if (lane <= kPrePot) cp_async8(pre + lane, lane < kPrePot ? p.env_kpi + (size_t)e * EV2B_KPI_COUNT + lane : p.env_pot + e);
        else if (lane == kEvlPotPrev) cp_async8(pre + lane, p.env_pot_prev + e);
    }
    if (t >= p.T) {                                   // step() on a finished env   ev2gym_env.py:343
        if (gtid == 0) {
            if (p.out.reward) p.out.reward[e] = 0.0;
            if (p.out.total_costs) p.out.total_costs[e] = 0.0;
            if (p.out.status) p.out.status[e] = EV2B_ST_DONE | EV2B_ST_WAS_DONE;
        }
        cp_async_wait_all();
        return p.T + 1;
    }
    const int tq = t + 1;
    const EnvT *etp = p.env_t + (size_t)s * p.T + t;
    const double2 price = *reinterpret_cast<const double2 *>(etp);           // cp, dp
    const int a0 = etp->arr0, nArr = etp->n_arr;                             // the sessions arriving at step t+1
    if (gw == 0) {                                    // setpoints and the transformers' rows of this step (two 16 B halves each)
        if (lane == kEvlSet) cp_async8(pre + lane, &etp->setpoint);
        else if (lane == kEvlSetNext) { if (tq < p.T) cp_async8(pre + lane, &etp[1].setpoint); else pre[lane] = 0.0; }
#pragma unroll 1
        for (int i = lane; i < 2 * p.Tr; i += 32)
            cp_async16(pre + kEvlTr + 2 * i, reinterpret_cast<const double *>(p.tr_t + ((size_t)s * p.T + t) * p.Tr) + 2 * i);
    }
    float *obs_row = p.out.obs + (size_t)e * p.D;
    if (want_obs) {                                   // (before the idle decision: these loads travel with the ones above)
        // batches of 4 values per thread: the loads of a batch are in flight together (an idle env's warp would otherwise
        // wait out seven dependent L2 round trips, one per value)
        if (p.series_pairs && p.obs_static) {         // (the host checked: values 2j, 2j+1 are neighbours at an even offset)
            const float2 *src = reinterpret_cast<const float2 *>(p.obs_static + ((size_t)s * (p.T + 1) + tq) * p.W);
            const int W2 = p.W >> 1;
#pragma unroll 1
            for (int i0 = gtid; i0 < W2; i0 += 4 * GT) {
                float2 v[4]; int o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + j * GT;
                    if (i < W2) { o[j] = __ldg(&p.series_off[2 * i]); v[j] = __ldg(src + i); }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) if (i0 + j * GT < W2) *reinterpret_cast<float2 *>(obs_row + o[j]) = v[j];
            }
        } else {
#pragma unroll 1
        for (int i0 = gtid; i0 < p.W; i0 += 4 * GT) {
            float v[4]; int o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * GT;
                if (i < p.W) { o[j] = __ldg(&p.series_off[i]); v[j] = obs_series_fetch(p, s, tq, i); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) if (i0 + j * GT < p.W) obs_row[o[j]] = v[j];
        }
        }
    }
    const bool idle = n_old == 0 && nArr == 0;        // nobody connected, nobody arriving: warp 0 alone, no barriers
    if (idle && gw != 0) { if (!KSTEP && G * 32 != kEvlThreads) evl_group_arrive<G>(g); return tq; }
    const int NT = idle ? 32 : GT;                    // threads that share the fills below
    uint8_t *mask_row = p.out.action_mask ? p.out.action_mask + (size_t)e * p.P : nullptr;
    float *h_csP = p.out.hist_cs_power ? p.out.hist_cs_power + ((size_t)e * p.T + t) * p.C : nullptr;     // row t of the histories
    float *h_csA = p.out.hist_cs_current ? p.out.hist_cs_current + ((size_t)e * p.T + t) * p.C : nullptr;

    // ---- P0: zero the per-port flags, fill the per-step outputs, (scenario, time)-only observation values ---------
    if (!idle) {
        if (NP == 1) {                                // one port per charger: the EV's thread is the charger's thread (no CS phase)
#pragma unroll 1
            for (int i = gtid; i < p.C; i += GT) csP[i] = 0.0;
        }
#pragma unroll 1
        for (int i = gtid; i < (p.P + 3) >> 2; i += GT) reinterpret_cast<unsigned *>(occ)[i] = 0u;
    }
    if (NP == 1 || idle) {                            // (with a CS phase the charger's thread writes these)
        if (p.out.cs_power || p.out.cs_current || h_csP || h_csA) {
#pragma unroll 1
            for (int c = gtid; c < p.C; c += NT) {
                if (p.out.cs_power)   p.out.cs_power[(size_t)e * p.C + c] = 0.f;
                if (p.out.cs_current) p.out.cs_current[(size_t)e * p.C + c] = 0.f;
                if (h_csP) h_csP[c] = 0.f;
                if (h_csA) h_csA[c] = 0.f;
            }
        }
    }
    if (want_obs) {
        if (obs_full) {                               // the caller's buffer does not hold last step's rows: clear every tuple
#pragma unroll 1
            for (int i = gtid; i < p.P; i += NT) evl_obs_clear(p, obs_row, i);
        }
    }
    // action_mask is maintained incrementally like the observation tuples: the caller's buffer still holds last step's
    // row, only ports whose EV arrives or leaves change.  A new buffer (mask_full) or a new episode (t == 0: the row may
    // hold the previous episode's terminal mask) rewrites the row.                                  ev2gym_env.py:452-457
    if (mask_row && (mask_full || t == 0)) {
#pragma unroll 1
        for (int i = gtid; i < p.P; i += NT) mask_row[i] = 0;
    }
    if (HEAVY && (p.out.port_energy || p.out.dep_sat || p.out.dep_cap)) {   // dense per-port outputs: defaults first
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll 1
        for (int i = gtid; i < p.P; i += NT) {
            const size_t ip = (size_t)e * p.P + i;
            if (p.out.port_energy) p.out.port_energy[ip] = 0.f;
            if (p.out.dep_sat) p.out.dep_sat[ip] = nan;
            if (p.out.dep_cap) p.out.dep_cap[ip] = nan;
        }
    }

    double aProfit = 0, aSatExp = 0, aCh = 0, aDis = 0, aSat = 0, aUsage = 0, aPot = 0;
    int nDep = 0;
    bool overflow = false;
    if (!idle) {
    evl_group_sync<G>(g);

    // ---- EV: one thread per connected EV ------------------------------------------------------------------------
    // (Round 2 also measured ordering a warp's entries by the direction of the action first -- charging EVs to the front,
    //  ballot + popc compaction, so that an iteration runs one branch of the battery model: c3 -1.5 %, c4 +3.4 %, removed.
    //  The warps wait on the loads of the common part of an iteration, not on the two branches.
    //  profiles/r2_ab_direction_sort.jsonl)
    int port = (int)first;
#pragma unroll 1
    for (int i = gtid; i < n_old; i += GT) {
        const int port_next = i + GT < n_old ? (int)lst[i + GT] : -1;    // next iteration's entry, and what to prefetch for it
        const size_t ip = (size_t)e * p.P + port;
        const uint4 h = p.hot[ip];
        double cv = p.cap[ip];
        double exch_new = p.exch[ip];
        double a, am_raw = 0.0;
        if (NP == 2 && p.act_pairs) {            // both ports of the charger with one load
            const Pair<ActT> v = *reinterpret_cast<const Pair<ActT> *>(actions + (ip & ~(size_t)1));
            a = (double)((ip & 1) ? v.y : v.x); am_raw = (double)((ip & 1) ? v.x : v.y);
        } else {
            a = agent_action<ActT>(p, actions, ip, t);
            if (NP == 2) am_raw = agent_action<ActT>(p, actions, ip ^ 1, t);                 // same 32 B sector as `a`
        }
        const unsigned hx = NP == 2 ? p.hot[ip ^ 1].x : 0u;
        if (port_next >= 0) evl_prefetch_ev<ActT>(p, actions, (size_t)e * p.P + port_next);
        const EvSpec *sp = p.spec + hot_spec(h);
        constexpr bool BATCH = EV2B_SPEC_BATCH == 1 || (EV2B_SPEC_BATCH == 2 && HEAVY);
        double2 Bv;
        if (BATCH) Bv = __ldg(reinterpret_cast<const double2 *>(&sp->B));      // B, 1 / B: with the model's own batch of loads
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
                    const unsigned hj = p.hot[ij].x;
                    const bool occ_j = (int)(int16_t)(hj & 0xFFFFu) <= t && t <= (int)(int16_t)(hj >> 16);
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
        const double cv_old = cv;
        if (an != 0.0) {
            double cap = cv;
            bool em_cross;
            const bool active = ev_step_item<HEAVY>(p, cs, h.z, h.w, an, cap, energy, act_amps, em_cross);
            if (active) {                                                 // amps == 0: nothing changes  ev.py:158-163
                cv = cap;
                p.cap[ip] = cv;
                exch_new = exch_new + energy;                            // total_energy_exchanged  ev.py:178
                p.exch[ip] = exch_new;
            }
            const double ae = fabs(energy);
            if (an > 0.0) { aProfit += ae * price.x; aCh += ae; }         // ev_charger.py:178-179
            else          { aProfit += ae * price.y; aDis += ae; }        // ev_charger.py:194-195
            pwv = ev2b_div_c(energy * 60.0, p.period, p.rperiod);         // :180,196
            if (HEAVY && p.stats && em_cross) atomicAdd(&p.cs_em[(size_t)e * p.C + c], 1);   // ev.py:401-402 (integer: order-free)
        }
        if (!BATCH) Bv = __ldg(reinterpret_cast<const double2 *>(&sp->B));
        if (HEAVY && p.stats) {             // EV.step bookkeeping: historic_soc / active_steps / |energy|  ev.py:156,178-185
            const double soc0 = ev2b_div_c(cv_old, Bv.x, Bv.y);
            int cn = p.st_cnt[ip];
            p.st_soc_sum[ip] += soc0;
            if (an != 0.0 && act_amps != 0.0) { p.st_act[ip * (size_t)p.L + (cn >> 16)] = soc0; cn += 1 << 16; }
            if (an != 0.0) p.st_abs_e[ip] += fabs(energy);
            p.st_cnt[ip] = cn + 1;
        }
        if (HEAVY && p.out.port_energy) p.out.port_energy[ip] = (float)energy;
        if (NP != 1) { pw[port] = pwv; amp[port] = act_amps; }
        double potv = 0.0;
        unsigned flags = 1u;
        if (t >= hot_t_dep(h)) {                                          // departure  ev_charger.py:209-224, ev.py:199-214
            const double des = __ldg(&sp->desired);
            const double sat = (cv < des - 0.001) ? cv / des : 1.0;
            aSatExp += evl_departure_penalty<HEAVY>(p, cv, des, sat);
            aSat += sat;
            ++nDep;
            if (mask_row) mask_row[port] = 0;
            if (want_obs) evl_obs_clear(p, obs_row, port);
            if (HEAVY) {
                if (p.out.dep_sat) p.out.dep_sat[ip] = sat;
                if (p.out.dep_cap) p.out.dep_cap[ip] = cv;
                if (p.stats) {                                            // ev_charger.py:218-220, utils.py:49-63
                    const SessRec r0 = p.sess[((size_t)s * p.P + port) * p.Smax + hot_cursor(h) - 1];
                    double d1, d2;
                    finalize_ev(p, ip, sp, r0.afap, hot_t_arr(h), hot_t_dep(h), cv, d1, d2);
                    if (NP == 1) {
                        const size_t ec = (size_t)e * p.C + c;
                        p.cs_sat_sum[ec] += sat; p.cs_served[ec] += 1; p.cs_dcal[ec] += d1; p.cs_dcyc[ec] += d2;
                    } else { dsat[port] = sat; dcal[port] = d1; dcyc[port] = d2; flags |= 6u; }
                }
            }
        } else {
            flags |= 8u;                                                  // stays connected: in next step's list
            if (mask_row) mask_row[port] = 1;
            const double B = Bv.x;
            aSatExp += evl_unreachable_penalty<HEAVY>(p, sp, cv, hot_t_dep(h) - tq);
            if (cv < B && hot_t_dep(h) > tq) potv = __ldg(&p.pot_kw[hot_spec(h) * p.n_cls + cs.cls]);   // utils.py:766-777
            if (want_obs) evl_obs_tuple<HEAVY>(p, obs_row, port, c, h, cv, B, Bv.y, exch_new, tq);
            if (HEAVY && p.stats && tq >= p.T) {      // episode over: EVs still connected count too (env.EVs)
                const SessRec r0 = p.sess[((size_t)s * p.P + port) * p.Smax + hot_cursor(h) - 1];
                double d1, d2;
                finalize_ev(p, ip, sp, r0.afap, hot_t_arr(h), hot_t_dep(h), cv, d1, d2);
                if (NP == 1) { const size_t ec = (size_t)e * p.C + c; p.cs_dcal[ec] += d1; p.cs_dcyc[ec] += d2; }
                else { dcal[port] = d1; dcyc[port] = d2; flags |= 4u; }
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
            if (h_csP) h_csP[port] = (float)rP;
            if (h_csA) h_csA[port] = (float)rA;
        } else {
            pot[port] = potv;
        }
        occ[port] = (unsigned char)flags;
        port = port_next;
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
        p.exch[ip] = 0.0;
        const int c = NP == 1 ? port : (NP == 2 ? port >> 1 : p.port_cs[port]);
        const CsStatic &cs = cs_of<UNI>(p, c);
        const EvSpec *sp = p.spec + hot_spec(r.hot);
        const double2 Bv = __ldg(reinterpret_cast<const double2 *>(&sp->B));
        const double B = Bv.x;
        double potv = 0.0;
        aSatExp += evl_unreachable_penalty<HEAVY>(p, sp, r.cap0, hot_t_dep(r.hot) - tq);
        if (r.cap0 < B && hot_t_dep(r.hot) > tq) potv = __ldg(&p.pot_kw[hot_spec(r.hot) * p.n_cls + cs.cls]);
        unsigned flags = (unsigned)occ[port];
        if (HEAVY && p.stats) {
            p.st_soc_sum[ip] = 0.0; p.st_abs_e[ip] = 0.0; p.st_cnt[ip] = 0;
            if (tq >= p.T) {                          // arrives as the episode ends: finalised at once (env.EVs)
                double d1, d2;
                finalize_ev(p, ip, sp, r.afap, hot_t_arr(r.hot), hot_t_dep(r.hot), r.cap0, d1, d2);
                if (NP == 1) { const size_t ec = (size_t)e * p.C + c; p.cs_dcal[ec] += d1; p.cs_dcyc[ec] += d2; }
                else {
                    if (flags & 4u) { dcal[port] += d1; dcyc[port] += d2; } else { dcal[port] = d1; dcyc[port] = d2; }
                    flags |= 4u;
                }
            }
        }
        if (NP == 1) {                                                    // (an EV that left this very port in step t added 0)
            double rPot = potv;
            if (rPot > cs.max_power) rPot = cs.max_power;
            else if (rPot < cs.min_power) rPot = 0.0;
            aPot += rPot;
        } else {
            if (!(flags & 1u)) { pw[port] = 0.0; amp[port] = 0.0; }      // (an EV may have left this very port in step t)
            pot[port] = potv;
        }
        occ[port] = (unsigned char)(flags | 9u);                          // per-port results + connected at t+1
        if (mask_row) mask_row[port] = 1;
        if (want_obs) evl_obs_tuple<HEAVY>(p, obs_row, port, c, r.hot, r.cap0, B, Bv.y, 0.0, tq);
    }
    evl_group_sync<G>(g);

    // ---- CS: one thread per charger, ports in order -------------------------------------------------------------
#pragma unroll 1
    for (int c = NP == 1 ? p.C : gtid; c < p.C; c += GT) {
        const CsStatic &cs = cs_of<UNI>(p, c);
        const int p0 = NP > 0 ? c * NP : cs.port_off, n = NP > 0 ? NP : cs.n_ports;
        double rP = 0, rA = 0, rPot = 0;
        if (NP == 2 && !(HEAVY && p.stats)) {
            // both ports of the charger at once: one 2-byte load of the flags, three 16-byte loads of the staged values
            // (entries of empty ports are stale and never enter a sum); same additions in the same order as below
            const unsigned f2 = reinterpret_cast<const uint16_t *>(occ)[c];
            const double2 w = reinterpret_cast<const double2 *>(pw)[c], a2 = reinterpret_cast<const double2 *>(amp)[c],
                          o2 = reinterpret_cast<const double2 *>(pot)[c];
            if (f2 & 0x001u) { rP += w.x; rA += a2.x; rPot += o2.x; }
            if (rA - 0.0001 > cs.imax) overflow = true;
            if (f2 & 0x100u) { rP += w.y; rA += a2.y; rPot += o2.y; }
            if (rA - 0.0001 > cs.imax) overflow = true;
        } else
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const unsigned f = occ[p0 + j];
            if (f & 1u) { rP += pw[p0 + j]; rA += amp[p0 + j]; rPot += pot[p0 + j]; }
            if (rA - 0.0001 > cs.imax) overflow = true;                   // ev_charger.py:203-205
            if (HEAVY && p.stats && (f & 6u)) {                           // the charger's totals, in port order
                const size_t ec = (size_t)e * p.C + c;
                if (f & 2u) { p.cs_sat_sum[ec] += dsat[p0 + j]; p.cs_served[ec] += 1; }
                p.cs_dcal[ec] += dcal[p0 + j]; p.cs_dcyc[ec] += dcyc[p0 + j];
            }
        }
        if (rPot > cs.max_power) rPot = cs.max_power;                     // utils.py:779-789
        else if (rPot < cs.min_power) rPot = 0.0;
        csP[c] = rP;
        aUsage += rP; aPot += rPot;
        if (p.out.cs_power)   p.out.cs_power[(size_t)e * p.C + c] = (float)rP;
        if (p.out.cs_current) p.out.cs_current[(size_t)e * p.C + c] = (float)rA;
        if (h_csP) h_csP[c] = (float)rP;
        if (h_csA) h_csA[c] = (float)rA;
    }
    // per-warp partial sums (fixed butterfly); the group total is formed warp by warp in the reward phase
    {
        const double q[EvlNSum] = {aProfit, aSatExp, aCh, aDis, aSat, aUsage, aPot,
                                   (double)nDep + (overflow ? 1048576.0 : 0.0)};   // departures < 2^20; overflow votes above
        const double tot = warp_sum8(q, lane);
        if ((lane & 3) == 0) wsum[gw * EvlNSum + (lane >> 2)] = tot;
    }
    evl_group_sync<G>(g);

    // ---- LS: the group's last warp writes next step's list IN PORT ORDER (a ballot compaction of the "connected at t+1"
    // flags).  Neighbouring lanes of the EV loop then work on neighbouring ports: their battery levels / energies (four
    // per 32-byte sector), hot words (two), actions (eight) and observation tuples share sectors -- the gathers and
    // scatters of an iteration touch 40 % fewer sectors than with the list in arrival order (c3 at its busiest step: 14.6
    // instead of 25.5 sectors per 32 battery levels, 8.2 instead of 19.6 per 32 actions).
    if (gw == G - 1) {
        uint16_t *nxt = p.occ_list + (size_t)e * p.P;      // in place: every thread is past its last read of the old list
        int base = 0;
#pragma unroll 1
        for (int i0 = 0; i0 < p.P; i0 += 32) {
            const int pt = i0 + lane;
            const bool on = pt < p.P && (occ[pt] & 8u);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) nxt[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)pt;
            base += __popc(m);
        }
        if (lane == 0) p.occ_n[e] = base;
    }
    if (gw != 0) return tq;
    }  // !idle

    // ---- TR (warp 0): transformer sums + overload, same lane split as step_kernel's phase B ---------------------
    cp_async_wait_all();
    __syncwarp();
    double ovsum = 0.0;
    {
        const int lg = p.tr_lg, nseg = 1 << lg, per = 32 >> lg;   // 2^lg lanes share one transformer (host: largest with 2^lg * Tr <= 32)
        for (int k0 = 0; k0 < p.Tr; k0 += per) {
            const int k = k0 + (lane >> lg), seg = lane & (nseg - 1);
            double sp_ = 0.0;
            if (!idle && k < p.Tr) {
                const int i0 = p.tr_cs_off[k], n_k = p.tr_cs_off[k + 1] - i0;
                const int chunk = (n_k + nseg - 1) >> lg;
                const int lo = seg * chunk, hi = min(n_k, lo + chunk);
                for (int i = lo; i < hi; ++i) sp_ += csP[p.tr_cs_idx[i0 + i]];
            }
            for (int o = nseg >> 1; o > 0; o >>= 1) sp_ += __shfl_xor_sync(0xffffffffu, sp_, o);
            if (seg == 0 && k < p.Tr) {                                   // transformer.py:264-302
                const double *tq4 = pre + kEvlTr + 4 * k;
                const double ptot = (tq4[0] + tq4[1]) + sp_;
                double ov = 0.0;
                if (ptot > tq4[2] + 0.0001 || ptot < tq4[3] - 0.0001) ov = fabs(ptot - tq4[2]);
                ovsum += ov;
                if (HEAVY && p.n_bus > 0) trp[k] = ptot;
                if (p.out.tr_power)    p.out.tr_power[(size_t)e * p.Tr + k] = ptot;
                if (p.out.tr_overload) p.out.tr_overload[(size_t)e * p.Tr + k] = ov;
                if (p.out.hist_tr_overload) p.out.hist_tr_overload[((size_t)e * p.T + t) * p.Tr + k] = ov;
            }
        }
        ovsum = warp_sum(ovsum);
    }
    // ---- distribution grid: Laurent power flow of this env (the warp)   grid.py:120-141 -------------------------
    double lossv = 0.0;
    if (HEAVY && p.n_bus > 0) {
        __syncwarp();
        lossv = power_flow_env(p, pfv, trp, s, t, e, lane);
        if (lane == 0 && p.out.node_voltage) p.out.node_voltage[(size_t)e * (p.n_bus + 1)] = 1.0;
    }
    // ---- reward, KPI sums, step counter: every lane computes the (uniform) reward, lane k stores quantity k ------
    if (idle) {                                       // every per-EV total is zero (stored so that the reads below are
        if (lane < EvlNSum) wsum[lane] = 0.0;         //  unconditional: one LDS each instead of eight guarded blocks)
        __syncwarp();
    }
    double q[EvlNSum];
#pragma unroll
    for (int k = 0; k < EvlNSum; ++k) {
        double v = wsum[k];
        if (!idle) for (int w = 1; w < G; ++w) v += wsum[w * EvlNSum + k];
        q[k] = v;
    }
    const int cnts = (int)q[EvlCounts];
    const int n_dep = cnts & 0xFFFFF;
    const double usage = q[EvlUsage];                                     // current_power_usage[t]  ev2gym_env.py:375
    const double costs = q[EvlProfit];
    const double pot_now = pre[kPrePot];                                  // charge_power_potential[t]
    const double setpoint = pre[kEvlSet], setpoint_next = pre[kEvlSetNext];
    double reward = 0.0;
    {
        const int rk = p.reward_kind;
        if (rk == EV2B_REWARD_SQ_TRACKING) {                              // reward.py:11-12
            const double m = setpoint < pot_now ? setpoint : pot_now;
            reward = -((m - usage) * (m - usage));
        } else if (rk == EV2B_REWARD_PROFIT_TR_USER) {                    // reward.py:36-44
            reward = costs - 100.0 * ovsum - q[EvlSatExp];
        } else if (rk == EV2B_REWARD_PROFIT_MAX) {                        // reward.py:81-87
            reward = costs - q[EvlSatExp];
        } else if (rk == EV2B_REWARD_SQTR_TR_USER) {                      // reward.py:16-32
            double m = setpoint < pot_now ? setpoint : pot_now;
            const double lim = pre[kEvlTr + 2];                           // transformers[0].max_power[t]
            if (lim < m) m = lim;
            reward = -((m - usage) * (m - usage)) - 100.0 * ovsum - q[EvlSatExp];
        } else if (rk == EV2B_REWARD_SQ_TRACKING_PENALTY) {               // reward.py:46-58
            const double m = setpoint < pot_now ? setpoint : pot_now;
            reward = -((m - usage) * (m - usage));
            if (usage == 0.0 && pre[kEvlPotPrev] != 0.0) reward = reward - 100.0;   // potential[current_step-2]; 0 at t = 0
        } else if (rk == EV2B_REWARD_SIMPLE) {                            // reward.py:60-65
            reward = -((setpoint - usage) * (setpoint - usage));
        } else if (rk == EV2B_REWARD_MIN_TRACKER_SURPLUS) {               // reward.py:67-76
            if (setpoint < usage) reward -= (usage - setpoint) * (usage - setpoint);
            reward += usage;
        } else if (rk == EV2B_REWARD_V2G_COSTS_SIMPLE) {                  // reward.py:150-153
            reward = costs;
        } else if (rk == EV2B_REWARD_V2G_PROFITMAX || rk == EV2B_REWARD_V2G_PROFITMAX_V2 ||
                   rk == EV2B_REWARD_PST_PROFITMAX_V2) {                  // reward.py:123-148, 155-213, 281-339
            reward = costs - q[EvlSatExp];
            if (rk == EV2B_REWARD_PST_PROFITMAX_V2 && setpoint < usage) reward += 1000.0 * (setpoint - usage);
        } else if (HEAVY && rk == EV2B_REWARD_GRID_FULL) {                // reward.py:89-111
            reward = costs + 1000.0 * lossv - q[EvlSatExp];
        } else if (HEAVY && rk == EV2B_REWARD_GRID_SIMPLE) {              // reward.py:114-121
            reward = 1000.0 * lossv;
        } else if (HEAVY && rk == EV2B_REWARD_GRID_PROFITMAX_V2) {        // reward.py:215-279
            reward = (costs - q[EvlSatExp]) + 50000.0 * lossv;
        }
    }
    const double dsp = setpoint - usage;                                  // utils.py:37-44
    // lane k keeps quantity k.  A chain of selects on purpose: a switch (lane) compiles to thirteen one-lane paths that
    // the warp walks one after the other (75 instructions per env-step on the B200, 11 % of an idle step:
    // profiles/r2g_evl_lines_idle.txt); every value below is warp-uniform and already computed.
    double delta = 1.0;                                                              // EV2B_KPI_STEPS
    delta = lane == EV2B_KPI_TOTAL_REWARD ? reward : delta;
    delta = lane == EV2B_KPI_TOTAL_PROFITS ? costs : delta;
    delta = lane == EV2B_KPI_ENERGY_CHARGED ? q[EvlCharged] : delta;
    delta = lane == EV2B_KPI_ENERGY_DISCHARGED ? q[EvlDischarged] : delta;
    delta = lane == EV2B_KPI_TR_OVERLOAD ? ovsum : delta;
    delta = lane == EV2B_KPI_EVS_SERVED ? (double)n_dep : delta;
    delta = lane == EV2B_KPI_SAT_SUM ? q[EvlSatSum] : delta;
    delta = lane == EV2B_KPI_TRACKING_ERROR ? dsp * dsp : delta;
    delta = lane == EV2B_KPI_ENERGY_TRACKING_ERROR ? fabs(dsp) : delta;
    delta = lane == EV2B_KPI_TRACKER_VIOLATION ? (usage > setpoint ? usage - setpoint : 0.0) : delta;
    delta = lane == EV2B_KPI_EVS_SPAWNED ? (double)nArr : delta;
    delta = lane == EV2B_KPI_INVALID_ACTIONS ? (double)(p.P - n_old) : delta;        // every empty port  ev_charger.py:137-140
    if (idle && !KSTEP && G * 32 != kEvlThreads) evl_group_sync<G>(g);   // the group's other warp has read env_step (it only arrives)
    if (lane < EV2B_KPI_COUNT) {
        p.env_kpi[(size_t)e * EV2B_KPI_COUNT + lane] = pre[lane] + delta;
    } else if (lane == 13) {
        p.env_pot[e] = (tq < p.T) ? q[EvlPot] : 0.0;                      // ev2gym_env.py:424-426
    } else if (lane == 14) {
        p.env_usage[e] = usage;
        p.env_step[e] = tq;
        if (p.out.hist_usage) p.out.hist_usage[(size_t)e * p.T + t] = usage;
    } else if (lane == 15) {
        unsigned status = (cnts >> 20) ? EV2B_ST_AMPS_OVERFLOW : 0u;
        if (tq >= p.T) status |= EV2B_ST_DONE;                            // ev2gym_env.py:460
        if (p.out.status) p.out.status[e] = status;
        if (p.out.reward) p.out.reward[e] = reward;
    } else if (lane == 16) {
        if (p.out.total_costs) p.out.total_costs[e] = costs;
        if (p.reward_kind == EV2B_REWARD_SQ_TRACKING_PENALTY) p.env_pot_prev[e] = pot_now;
    } else if (lane == 17) {
        if (want_obs) obs_header(p, obs_row, s, tq, usage, setpoint_next);
    }
    return tq;
}

// Device-side reset of env e by its group (the per-env work of reset_ports_kernel + reset_envs_kernel, mode 1): next
// scenario of the bank, every port empty, counters and KPI sums zero, first observation.      ev2gym_env.py:298-331
template <int G>
__device__ __forceinline__ void evl_reset_env(const Params &p, const int e, const int g, const int gtid) {
    constexpr int GT = 32 * G;
    const int s = (p.env_scn[e] + p.scn_stride) % p.S;
    evl_group_sync<G>(g);                             // (every thread has read env_scn before thread 0 replaces it)
#pragma unroll 1
    for (int port = gtid; port < p.P; port += GT) {
        const size_t ip = (size_t)e * p.P + port;
        uint4 h;
        h.x = ((unsigned)kNoArrival & 0xFFFFu) | (0xFFFFu << 16);        // t_arr = 32767, t_dep = -1
        h.y = p.sess[((size_t)s * p.P + port) * p.Smax].hot.x & 0xFFFFu;  // next_arr = first session's t_arr, cursor 0
        h.z = 0; h.w = 0;
        p.hot[ip] = h; p.cap[ip] = 0.0; p.exch[ip] = 0.0;
        if (p.rr_key) p.rr_key[ip] = kRrAbsent;
    }
    const bool want_obs = (p.out.obs != nullptr) && (p.state_kind != EV2B_STATE_NONE);
    if (want_obs) {
        float *row = p.out.obs + (size_t)e * p.D;
#pragma unroll 1
        for (int i = gtid; i < p.D; i += GT) row[i] = 0.f;               // no EV is connected at t = 0
        evl_group_sync<G>(g);
        if (gtid == 0) obs_header(p, row, s, 0, 0.0, p.env_t[(size_t)s * p.T].setpoint);
#pragma unroll 1
        for (int i = gtid; i < p.W; i += GT) row[p.series_off[i]] = obs_series_fetch(p, s, 0, i);
    }
    if (gtid == 0) {
        p.env_step[e] = 0; p.env_scn[e] = s; p.env_pot[e] = 0.0; p.env_usage[e] = 0.0; p.env_pot_prev[e] = 0.0;
        p.occ_n[e] = 0;
        if (p.rr_fb) { p.rr_fb[2 * e] = 0; p.rr_fb[2 * e + 1] = 1; }
        for (int k = 0; k < EV2B_KPI_COUNT; ++k) p.env_kpi[(size_t)e * EV2B_KPI_COUNT + k] = 0.0;
    }
}

// The step kernel.  KSTEP = false: one step per launch (ev2b_step).  KSTEP = true (ev2b_step_k with an agent that needs no
// other env's state): the group advances ITS env p.k_steps times in one launch -- envs are independent, so no grid-wide
// synchronisation exists; from the second step on the env's rows are found in L1 / L2 instead of HBM and there is no
// launch gap between steps.  Finished envs restart on their next scenario when p.auto_reset is set (= ev2b_reset_done).
#ifndef EV2B_EVL_MINB
#define EV2B_EVL_MINB 7       // resident 128-thread CTAs per SM the lean instantiation is compiled for: 7 -> 72 registers per thread.
                              // B200, c3 whole episodes, one warp per env: 8 (64 regs) 25.8 us, 7 (72) 23.3, 6 (80) 31.4 --
                              // 4096 envs are 1024 CTAs = 6.9 per SM, so 7 slots still hold the launch in one wave
#endif
// TPB = threads per CTA: 128 (4 / G envs per CTA), or 32 with one warp per env -- a launch of more CTAs than the machine
// holds at once (c4: 8192 envs) ends with a shorter tail when the CTAs are single envs (B200, c4: 44.2 -> 40.9 us per
// launch; c3, a one-wave launch: 22.6 -> 23.5, so it keeps 128; profiles/r2_ab_cta_size.jsonl).
template <typename ActT, int NP, bool UNI, int G, bool HEAVY, bool KSTEP, int TPB>
__global__ void __launch_bounds__(TPB, (HEAVY ? 4 : EV2B_EVL_MINB) * kEvlThreads / TPB) evl_step_kernel(const __grid_constant__ Params p) {
    static_assert(TPB == kEvlThreads || (TPB == 32 && G == 1), "CTA shapes: 128 threads, or one warp = one env");
    EV2B_DYNAMIC_SMEM(smem_raw);
    constexpr int GT = 32 * G, EPB = TPB / GT;
    const int tid = threadIdx.x;
    const int g = tid / GT, gtid = tid - g * GT;
    const int e = p.env0 + (int)blockIdx.x * EPB + g;
    if (e >= p.env_end) return;                       // whole group: its barriers are its own
    unsigned char *sm = smem_raw + (size_t)g * p.v_stride;
    const ActT *actions = reinterpret_cast<const ActT *>(p.actions);
    if (!KSTEP) {
        evl_env_step<ActT, NP, UNI, G, HEAVY, false>(p, sm, e, g, gtid, actions, true);
        return;
    }
    const size_t act_stride = p.agent_kind == EV2B_AGENT_EXTERNAL ? (size_t)p.E * p.P : 0;
#pragma unroll 1
    for (int it = 0; it < p.k_steps; ++it) {
        const int tq = evl_env_step<ActT, NP, UNI, G, HEAVY, true>(p, sm, e, g, gtid, actions + act_stride * it, it == 0);
        evl_group_sync<G>(g);                         // the step's writes are visible to the whole group before the next one
        if (p.auto_reset && tq >= p.T) {              // (tq = T + 1: the env was finished before this launch and never reset)
            evl_reset_env<G>(p, e, g, gtid);
            evl_group_sync<G>(g);
        }
    }
}

}  // namespace ev2b
