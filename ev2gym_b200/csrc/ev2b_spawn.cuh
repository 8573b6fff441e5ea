// ev2b_spawn.cuh -- device-side scenario sampling: the EV sessions of every scenario of the bank are (re)drawn on the GPU.
//
// What this replaces (paths relative to /root/reference): EV_spawner (ev2gym/utilities/utils.py:477-557) and
// spawn_single_EV (utils.py:177-345), which reset() runs through load_ev_profiles (loaders.py:368-389) once per episode --
// 0.3-0.65 s of single-core Python per scenario, five orders of magnitude below the step rate.  The reference consumes
// numpy's global Mersenne-Twister stream, so draws cannot be reproduced bit for bit; parity is DISTRIBUTIONAL (arrival,
// stay, energy and model histograms against reference-sampled scenarios, tests/test_spawn.py) plus the exact structural
// rules, which are restated literally:
//   * a port spawns at loop step t only if it was free at t, t-1 and t-2                               utils.py:534-536
//   * arrival probability per port and step: rand * 100 < tau(t) * multiplier * timescale / 60 * spawn_multiplier   :538
//   * workplace sites are closed before 6 h, after 18 h and at weekends                                :509-520
//   * the clock of loop step t is sim_date + (t - 2) * timescale (the loop starts at t = 2 with `time = sim_date`)  :492,504
//   * time_of_arrival = t + 1, time_of_departure = int(stay + t + 3), occupied on [t + 1, t_dep)       :316-318, 551-552
//   * sessions that would end after the simulation are dropped (stay + t + 4 >= T)                     :254-256
//   * required energy ~ N(m, m/2) (< 5: randint(5, 10)), battery level at arrival from it              :207-229
//   * stay ~ N(m, m/5) hours -> steps + 1, at least min_time_of_stay                                   :236-252
//   * EV model ~ registrations; efficiency round(1 - (u + 1e-5) / 20, 3), transition_soc round(0.9 - (u + 1e-5) / 5, 3)
//   * the arriving EV takes the FIRST FREE PORT of its charger, not the port the spawner drew          ev_charger.py:273
// The scenario's time series (prices, loads, PV, limits, forecasts) stay those of the bank entry.  power_setpoints are
// derived from the sessions by the reference (generate_power_setpoints, utils.py:664-757: every EV's required energy is
// spread over its stay with price-weighted normal noise, pushed inside the power limits, summed, median-smoothed), so with
// power_setpoint_enabled they are regenerated here too (spawn_setpoints_kernel), with the same distributional parity.
//
// Kernels: spawn_sessions_kernel (thread per scenario x spawner port: the sequential loop over t with a counter-based
// RNG) -> spawn_assign_kernel (thread per scenario x charger: first-free-port replay, packs the SessRec table)
// -> spawn_schedule_kernel (CTA per scenario: arrival buckets of the event-driven kernel, EnvT.arr0 / n_arr)
// -> spawn_setpoints_kernel (CTA per scenario, only with power_setpoint_enabled).
#pragma once
#include "ev2b_device.cuh"

namespace ev2b {

struct SpawnParams {
    // tables (device pointers)
    const double *arrival_week, *arrival_weekend;   // [96]
    const double *req_energy_mean, *stay_mean;      // [48]
    const double *model_cdf;                        // [M] cumulative registrations
    const double *model_B;                          // [M]
    const int *model_lut;                           // [M] (-1: scalar efficiencies are drawn)
    const int *start;                               // [S][3] weekday, hour, minute of sim_date
    int M, workplace, heterogeneous, empty_ports_at_end, min_stay_steps, timescale;
    double spawn_multiplier, desired_frac, min_battery_capacity;
    unsigned homog_ts_milli, homog_eta_c_milli, homog_eta_d_milli;   // homogeneous config: the fixed encodings
    unsigned seed_lo, seed_hi;
    int cap_per_scn;                                // arr_list entries reserved per scenario (P * Smax)
    double setpoint_mult;                           // 100 + power_setpoint_flexiblity                  utils.py:680-681
    double min_cs_power, max_cs_power;              // charging_stations[0].get_min_charge_power() / get_max_power()  :683-684
    int setpoint_threads, median_window;            // CTA size of spawn_setpoints_kernel; 5 * max(1, 15 // timescale)  :754-757
    // scratch: what the spawner drew, per (scenario, spawner port)
    SessRec *raw; int *raw_n;                       // [S][P][Smax], [S][P]
    // outputs
    SessRec *sess; EnvT *env_t; unsigned *arr_list; int *n_sess;   // n_sess: [S] sessions per scenario
};

// splitmix64 of (seed, counter): the uniform generator.  counter = ((scenario * P + port) * T + t) * 16 + draw
__device__ __forceinline__ double spawn_uniform(const SpawnParams &sp, unsigned long long counter) {
    unsigned long long z = (((unsigned long long)sp.seed_hi << 32) | sp.seed_lo) + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
}
__device__ __forceinline__ double spawn_normal(const SpawnParams &sp, unsigned long long c, double mean, double sd) {
    const double u1 = 1.0 - spawn_uniform(sp, c), u2 = spawn_uniform(sp, c + 1);     // (0, 1], [0, 1)
    return mean + sd * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);            // Box-Muller
}
__device__ __forceinline__ double spawn_randint(const SpawnParams &sp, unsigned long long c, int lo, int hi) {   // np.random.randint(lo, hi)
    if (hi <= lo) return (double)lo;
    int v = lo + (int)(spawn_uniform(sp, c) * (double)(hi - lo));
    return (double)(v < hi ? v : hi - 1);
}

// The sequential loop of EV_spawner for one (scenario, spawner port).
__global__ void spawn_sessions_kernel(const Params p, const SpawnParams sp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.S * p.P) return;
    const int s = (int)(idx / p.P);                                        // (idx = s * P + spawner port)
    SessRec *out = sp.raw + (size_t)idx * p.Smax;
    const int wd0 = sp.start[3 * s], m0 = sp.start[3 * s + 1] * 60 + sp.start[3 * s + 2];
    int n = 0, next_free = 0;                                             // ports are free at t = 0, 1, 2
    for (int t = 2; t < p.T - sp.min_stay_steps - 1; ++t) {               // utils.py:504
        if (t < next_free) continue;
        const int mins = m0 + (t - 2) * sp.timescale;
        const int wd = (wd0 + mins / 1440) % 7, hour = (mins / 60) % 24, minute = mins % 60;
        double tau;
        if (wd < 5) {
            if (sp.workplace && (hour < 6 || hour > 18)) continue;        // :509-512
            tau = __ldg(&sp.arrival_week[hour * 4 + minute / 15]);
        } else {
            if (sp.workplace) continue;                                   // :518-520
            tau = __ldg(&sp.arrival_weekend[hour * 4 + minute / 15]);
        }
        const unsigned long long c = ((unsigned long long)idx * (unsigned long long)p.T + (unsigned long long)t) * 16ull;
        if (!(spawn_uniform(sp, c) * 100.0 < tau * 1.0 * ((double)sp.timescale / 60.0) * sp.spawn_multiplier)) continue;   // :538
        // ---- spawn_single_EV (utils.py:177-345)
        const int hh = hour * 2 + (minute >= 30 ? 1 : 0);                 // arrival time rounded down to the half hour  :194-201
        const double em = __ldg(&sp.req_energy_mean[hh]);
        double required = spawn_normal(sp, c + 1, em, 0.5 * em);          // :207-208
        if (required < 5.0) required = spawn_randint(sp, c + 3, 5, 10);   // :210-211
        int model = 0;
        if (sp.heterogeneous) {                                           // np.random.choice(models, p=registrations)  :213-216
            const double u = spawn_uniform(sp, c + 4);
            while (model < sp.M - 1 && u >= __ldg(&sp.model_cdf[model])) ++model;
        }
        const double B = __ldg(&sp.model_B[model]);
        double cap0 = B < required ? spawn_randint(sp, c + 5, 1, (int)B) : B - required;       // :220-223
        if (cap0 > sp.desired_frac * B) cap0 = spawn_randint(sp, c + 6, 1, (int)B);            // :225-226
        if (cap0 < sp.min_battery_capacity && B > 2.0 * sp.min_battery_capacity) cap0 = sp.min_battery_capacity;   // :228-229
        const double sm = __ldg(&sp.stay_mean[hh]);
        double stay = spawn_normal(sp, c + 7, sm, 0.2 * sm);              // hours  :236-237
        stay = stay * 60.0 / (double)sp.timescale + 1.0;                  // :240
        if (stay < (double)sp.min_stay_steps) stay = (double)sp.min_stay_steps;       // :251-252
        if (sp.empty_ports_at_end && stay + (double)t + 4.0 >= (double)p.T) continue; // :254-256 (no occupancy is marked)
        const int t_arr = t + 1, t_dep = (int)(stay + (double)t + 3.0);   // :316-318
        unsigned tsm, ecm, edm;
        if (sp.heterogeneous) {
            tsm = (unsigned)rint(1000.0 * (0.9 - (spawn_uniform(sp, c + 9) + 0.00001) / 5.0));             // :309-310
            if (__ldg(&sp.model_lut[model]) >= 0) { ecm = 0u; edm = 0u; }                                   // efficiency curve
            else {
                ecm = (unsigned)rint(1000.0 * (1.0 - (spawn_uniform(sp, c + 10) + 0.00001) / 20.0));       // :293-296
                edm = (unsigned)rint(1000.0 * (1.0 - (spawn_uniform(sp, c + 11) + 0.00001) / 20.0));
            }
        } else { tsm = sp.homog_ts_milli; ecm = sp.homog_eta_c_milli; edm = sp.homog_eta_d_milli; }
        if (n < p.Smax) {
            SessRec r;
            r.hot.x = ((unsigned)t_arr & 0xFFFFu) | (((unsigned)min(t_dep, 32766) & 0xFFFFu) << 16);
            r.hot.y = 0u;
            r.hot.z = (unsigned)model | (tsm << 16);
            r.hot.w = ecm | (edm << 16);
            r.cap0 = cap0; r.afap = 0.0;
            out[n++] = r;
        }
        next_free = t_dep + 2;                                            // occupied on [t + 1, t_dep); free at t', t'-1, t'-2
    }
    sp.raw_n[idx] = n;
}

// First-free-port replay for one (scenario, charger): the charger's sessions in arrival order (ties: spawner port order,
// which is the order EV_spawner appends them in) take evs_connected.index(None).          ev_charger.py:266-285
__global__ void spawn_assign_kernel(const Params p, const SpawnParams sp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.S * p.C) return;
    const int s = (int)(idx / p.C), c = (int)(idx - (long long)s * p.C);
    const CsStatic &cs = p.cs[c];
    const int p0 = cs.port_off, n = cs.n_ports;
    const size_t base = ((size_t)s * p.P + p0) * p.Smax;                  // records of port p0 + j start at base + j * Smax
    SessRec empty;
    empty.hot.x = ((unsigned)kNoArrival & 0xFFFFu) | (0xFFFFu << 16);
    empty.hot.y = (unsigned)kNoArrival & 0xFFFFu; empty.hot.z = 0u; empty.hot.w = 0u; empty.cap0 = 0.0; empty.afap = 0.0;
    // per final port: sessions placed so far and the departure step of the last one (state lives in the output table itself)
    for (int j = 0; j < n; ++j) sp.sess[base + (size_t)j * p.Smax] = empty;
    int placed_total = 0;
    // cursors into the raw lists are kept in the raw records' hot.y of slot 0 ... simpler: rescan (lists are <= Smax long)
    for (;;) {
        int best_j = -1, best_k = 0, best_t = 0x7fffffff;
        for (int j = 0; j < n; ++j) {                                     // next unplaced session of every spawner port
            const int cnt = sp.raw_n[(size_t)s * p.P + p0 + j];
            const SessRec *rj = sp.raw + base + (size_t)j * p.Smax;
            for (int k = 0; k < cnt; ++k) {
                if (rj[k].hot.y) continue;                                // already placed
                const int ta = (int)(rj[k].hot.x & 0xFFFFu);
                if (ta < best_t) { best_t = ta; best_j = j; best_k = k; }
                break;                                                    // a port's sessions are in arrival order
            }
        }
        if (best_j < 0) break;
        SessRec *src = sp.raw + base + (size_t)best_j * p.Smax + best_k;
        SessRec r = *src;
        src->hot.y = 1u;
        const int ta = best_t, td = (int)(int16_t)(r.hot.x >> 16);
        int dst = -1, dst_k = 0;
        for (int j = 0; j < n && dst < 0; ++j) {                          // first free port: its last occupant left by step ta - 1
            const SessRec *fj = sp.sess + base + (size_t)j * p.Smax;
            int k = 0;
            while (k < p.Smax && (int)(fj[k].hot.x & 0xFFFFu) != kNoArrival) ++k;
            if (k == 0 || (int)(int16_t)(fj[k - 1].hot.x >> 16) <= ta - 1) { if (k < p.Smax) { dst = j; dst_k = k; } }
        }
        if (dst < 0) continue;                                            // (cannot happen: the spawner port itself is free)
        SessRec *fj = sp.sess + base + (size_t)dst * p.Smax;
        r.hot.y = ((unsigned)kNoArrival & 0xFFFFu) | ((unsigned)(dst_k + 1) << 16);
        {   // EV.calculate_max_energy_with_AFAP(cs.get_max_power())   ev.py:407-440, ev_charger.py:251-252, 279
            const EvSpec &es = p.spec[r.hot.z & 0xFFFFu];
            const double max_cs_power = cs.imax * cs.veff[1] * sqrt((double)cs.phases) / 1000.0;
            const double max_power = fabs(max_cs_power) > fabs(es.pmax_ac) ? es.pmax_ac : max_cs_power;
            double eff;
            if (es.lut >= 0) { eff = 0.0; for (int q = 0; q < p.lut_len; ++q) eff = fmax(eff, p.luts_c[(size_t)es.lut * p.lut_len + q]); eff = eff / 100.0; }
            else eff = (double)(r.hot.w & 0xFFFFu) / 1000.0;
            double afap = r.cap0;
            for (int q = ta; q < td + 1; ++q) {
                afap += max_power * eff * p.period / 60.0;
                afap = ceil(afap * 100.0) / 100.0;
                if (afap > es.B) { afap = es.B; break; }
            }
            r.afap = afap;
        }
        fj[dst_k] = r;
        if (dst_k + 1 < p.Smax) fj[dst_k + 1] = empty;
        if (dst_k > 0) fj[dst_k - 1].hot.y = (fj[dst_k - 1].hot.y & 0xFFFF0000u) | ((unsigned)ta & 0xFFFFu);   // next_arr of the previous one
        ++placed_total;
    }
    atomicAdd(&sp.n_sess[s], placed_total);
}

// Arrival buckets of one scenario (one CTA): sessions arriving at step q are arr_list[arr0 .. arr0 + n_arr) of the record
// EnvT[s][q - 1], in port order.  A bitmap (step x port) gives both the bucket sizes and every session's rank in its bucket.
__global__ void spawn_schedule_kernel(const Params p, const SpawnParams sp) {
    EV2B_DYNAMIC_SMEM(sm_raw);
    const int s = blockIdx.x, W = (p.P + 31) >> 5, R = p.T + 2;
    unsigned *bits = reinterpret_cast<unsigned *>(sm_raw);                // [R][W]
    int *off = reinterpret_cast<int *>(bits + (size_t)R * W);             // [R + 1]
    for (int i = threadIdx.x; i < R * W; i += blockDim.x) bits[i] = 0u;
    __syncthreads();
    for (int port = threadIdx.x; port < p.P; port += blockDim.x) {
        const SessRec *f = sp.sess + ((size_t)s * p.P + port) * p.Smax;
        for (int k = 0; k < p.Smax; ++k) {
            const int ta = (int)(f[k].hot.x & 0xFFFFu);
            if (ta == kNoArrival) break;
            if (ta <= p.T) atomicOr(&bits[(size_t)ta * W + (port >> 5)], 1u << (port & 31));
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < R; q += blockDim.x) {
        int n = 0;
        for (int w = 0; w < W; ++w) n += __popc(bits[(size_t)q * W + w]);
        off[q + 1] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) { off[0] = 0; for (int q = 1; q <= R; ++q) off[q] += off[q - 1]; }   // off[q] = first entry of step q
    __syncthreads();
    const int base = s * sp.cap_per_scn;
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) {                // the step that starts at t spawns the arrivals of t + 1
        EnvT *e = sp.env_t + (size_t)s * p.T + t;
        e->arr0 = base + off[t + 1]; e->n_arr = off[t + 2] - off[t + 1];
    }
    for (int port = threadIdx.x; port < p.P; port += blockDim.x) {
        const SessRec *f = sp.sess + ((size_t)s * p.P + port) * p.Smax;
        for (int k = 0; k < p.Smax; ++k) {
            const int ta = (int)(f[k].hot.x & 0xFFFFu);
            if (ta == kNoArrival) break;
            if (ta > p.T) continue;
            int rank = __popc(bits[(size_t)ta * W + (port >> 5)] & ((1u << (port & 31)) - 1u));
            for (int w = 0; w < (port >> 5); ++w) rank += __popc(bits[(size_t)ta * W + w]);
            sp.arr_list[base + off[ta] + rank] = (unsigned)port | ((unsigned)k << 16);
        }
    }
}

// generate_power_setpoints (utils.py:664-757) for one scenario (one CTA).  Every thread takes the sessions of some ports:
// the EV's required energy is spread over [t_arr + 1, t_dep) with |N(1 - price, min price)| weights, loads below the
// minimum / above the maximum power are pushed to the next slot (at most 11 sweeps), and the result is added to the
// thread's OWN row of partial setpoints; the rows are then summed in thread order (deterministic: the same seed gives the
// same setpoints) and median-smoothed.  Shared memory: [NT][T] partial setpoints, [NT][T] the session being spread,
// [T] normalised prices, [T] the sum.
__global__ void spawn_setpoints_kernel(const Params p, const SpawnParams sp) {
    EV2B_DYNAMIC_SMEM(sm_raw);
    const int s = blockIdx.x, NT = blockDim.x, T = p.T, tid = threadIdx.x;
    double *acc = reinterpret_cast<double *>(sm_raw);                     // [NT][T]
    double *vec = acc + (size_t)NT * T;                                   // [NT][T]
    double *price = vec + (size_t)NT * T;                                 // [T]
    double *total = price + T;                                            // [T]
    for (int i = tid; i < NT * T; i += NT) acc[i] = 0.0;
    if (tid == 0) {                                                       // prices = |charge_prices[0]| / max  :676-677
        double mx = 0.0;
        for (int t = 0; t < T; ++t) { price[t] = fabs(sp.env_t[(size_t)s * T + t].cp); mx = fmax(mx, price[t]); }
        for (int t = 0; t < T; ++t) price[t] = price[t] / mx;
    }
    __syncthreads();
    double *mine = acc + (size_t)tid * T, *sh = vec + (size_t)tid * T;
    for (int port = tid; port < p.P; port += NT) {
        const SessRec *f = sp.sess + ((size_t)s * p.P + port) * p.Smax;
        for (int k = 0; k < p.Smax; ++k) {
            const int ta = (int)(f[k].hot.x & 0xFFFFu);
            if (ta == kNoArrival) break;
            const int td = min((int)(int16_t)(f[k].hot.x >> 16), T), w0 = ta + 1, L = td - ta - 1;   // window [t + 2, t_dep), t = ta - 1
            if (L <= 0 || w0 >= T) continue;
            const EvSpec &es = p.spec[f[k].hot.z & 0xFFFFu];
            const double required = (es.B - f[k].cap0) * sp.setpoint_mult / 100.0;          // :692-693
            const double minp = fmax(es.pmin_ac, sp.min_cs_power), maxp = fmin(es.pmax_ac, sp.max_cs_power);   // :694-695
            double scale = price[w0];
            for (int i = 1; i < L; ++i) scale = fmin(scale, price[w0 + i]);
            const unsigned long long c = (1ull << 62) + (((unsigned long long)s * p.P + port) * (unsigned long long)T + ta) * 2ull * T;
            double sum = 0.0;
            for (int i = 0; i < L; ++i) {                                 // :698-703
                sh[i] = fabs(spawn_normal(sp, c + 2ull * i, 1.0 - price[w0 + i], scale));
                sum += sh[i];
            }
            for (int i = 0; i < L; ++i) sh[i] = sh[i] / sum * required * 60.0 / (double)sp.timescale;   // :704-705
            for (int step = 0; step <= 10; ++step) {                      // :708-733
                double mn = 1e300, mx = -1e300;
                for (int i = 0; i < L; ++i) { if (sh[i] != 0.0) mn = fmin(mn, sh[i]); mx = fmax(mx, sh[i]); }
                if (!(mn < minp || mx > maxp)) break;
                for (int i = 0; i < L; ++i) {
                    const int nxt = i == L - 1 ? 0 : i + 1;
                    if (sh[i] < minp && sh[i] > 0.0) { const double mv = sh[i]; sh[i] = 0.0; sh[nxt] += mv; }
                    else if (sh[i] > maxp) { const double mv = sh[i] - maxp; sh[i] = maxp; sh[nxt] += mv; }
                }
            }
            for (int i = 0; i < L; ++i) mine[w0 + i] += sh[i];            // :735
        }
    }
    __syncthreads();
    for (int t = tid; t < T; t += NT) {
        double v = 0.0;
        for (int r = 0; r < NT; ++r) v += acc[(size_t)r * T + t];
        total[t] = v;
    }
    __syncthreads();
    const int half = sp.median_window / 2;                                // median_smoothing  :652-661
    for (int t = tid; t < T; t += NT) {
        const int lo = max(0, t - half), hi = min(T, t + half + 1), n = hi - lo;
        // the median of n <= window + 1 values by rank counting (no sorting buffer): the k-th smallest is the value with
        // k smaller-or-equal-and-earlier values before it
        double m1 = 0.0, m2 = 0.0;
        const int k1 = (n - 1) / 2, k2 = n / 2;                           // numpy: mean of the two middle values for even n
        for (int i = lo; i < hi; ++i) {
            int rank = 0;
            for (int j = lo; j < hi; ++j) rank += (total[j] < total[i] || (total[j] == total[i] && j < i)) ? 1 : 0;
            if (rank == k1) m1 = total[i];
            if (rank == k2) m2 = total[i];
        }
        sp.env_t[(size_t)s * T + t].setpoint = 0.5 * (m1 + m2);
    }
}

// Every env reads as finished until it is reset (a new bank invalidates every running episode).
__global__ void spawn_invalidate_envs_kernel(const Params p) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < p.E) p.env_step[e] = p.T;
}

}  // namespace ev2b
