/*
 * ev2o.c -- ORACLE (test infrastructure, NOT product code).  See ev2o.h.
 *
 * A literal float64 restatement of the reference's step() cascade: same loops,
 * same operation order, same quirks.  Build with -ffp-contract=off -O2 (no FMA
 * contraction, no -ffast-math) so every expression rounds exactly as CPython /
 * numpy float64 arithmetic does.
 *
 * Nothing here is shared with the CUDA product path; the two are written
 * independently against the reference and compared by tests/.
 */
#include "ev2o.h"

#include <math.h>
#include <stddef.h>
#include <string.h>

/* The reference squares Python/numpy float64 SCALARS with `x**2` (reward.py:12,102; utils.py:38; ev.py:508), which
 * calls libm pow(x, 2.0).  glibc's pow is not correctly rounded: about 0.1 % of arguments come out 1 ulp away from
 * x*x.  To restate the reference literally the oracle calls the same libm function; the Makefile passes
 * -fno-builtin-pow because gcc otherwise folds pow(x, 2.0) into x*x.  (The CUDA path uses the correctly rounded
 * x*x, see DESIGN.md "known deviations".) */
static double py_sq(double x) { return pow(x, 2.0); }

/* ------------------------------------------------------------------------- */
/* helpers                                                                   */

/* numpy float64 round(x, 5): multiply, rint (half-even), divide.
 * Reference call site: ev_charger.py:157 `action = round(action, 5)` on an np.float64. */
static double np_round5(double a) { return rint(a * 100000.0) / 100000.0; }

/* EV.my_ceil(a, 2) = np.true_divide(np.ceil(a * 10**2), 10**2)   ev.py:188-189 */
static double my_ceil2(double a) { return ceil(a * 100.0) / 100.0; }

/* dict.get(np.round(amps), 1) / 100 on the efficiency table   ev.py:287-288, 375-376
 * The dict has integer keys 0..lut_len-1 (utils.py:280-288); any other key yields the default 1. */
static double lut_get(const double *lut, int lut_len, double key) {
    if (key >= 0.0 && key < (double)lut_len && key == floor(key)) return lut[(int)key];
    return 1.0;
}

typedef struct {
    double B, pmax_ac, pmin_ac, pmax_dis, pmin_dis, bmin, bmin_em, ts, mult, eta_c, eta_d;
    int ev_phases;
    const double *lut_c, *lut_d; /* NULL => scalar efficiency */
    int lut_len;
} ev_params;

/* EV._charge  ev.py:240-355.  Updates *cap, returns actual current; *energy = current_energy. */
static double ev_charge(const ev_params *p, double *cap, double amps, double voltage, int phases,
                        int timescale, double *energy) {
    double pilot = amps;
    voltage = voltage * sqrt((double)phases);                       /* :279 */
    double period = (double)timescale;
    double eta;
    if (p->lut_c) eta = lut_get(p->lut_c, p->lut_len, rint(amps)) / 100.0;   /* :287-288 */
    else          eta = p->eta_c;
    double pilot_dsoc = eta * pilot * voltage / 1000.0 / p->B / (60.0 / period); /* :295-296 */
    double max_dsoc   = eta * p->pmax_ac / p->B / (60.0 / period);               /* :297-298 */
    if (pilot_dsoc > max_dsoc) pilot_dsoc = max_dsoc;               /* :300-301 */
    double soc = *cap / p->B;                                       /* get_soc() :229 */
    double curr_soc;
    if (p->ts == 1.0) {                                             /* :303-306 */
        curr_soc = pilot_dsoc + soc;
        if (curr_soc > 1.0) curr_soc = 1.0;
    } else {
        double pts = p->ts + (pilot_dsoc - max_dsoc) / max_dsoc * (p->ts - 1.0);  /* :312-314 */
        double new_soc;
        if (soc < pts) {                                            /* :318 */
            if (1.0 <= (pts - soc) / pilot_dsoc)                    /* :323 */
                new_soc = pilot_dsoc + soc;
            else
                new_soc = 1.0 + exp(p->mult * (pilot_dsoc + soc - pts) / (pts - 1.0)) * (pts - 1.0); /* :326-330 */
        } else {
            new_soc = 1.0 + exp(p->mult * pilot_dsoc / (pts - 1.0)) * (soc - 1.0);   /* :332-334 */
        }
        double dsoc_limit = (max_dsoc > pilot_dsoc) ? pilot_dsoc : max_dsoc;          /* :336-339 */
        if (new_soc - soc > dsoc_limit) curr_soc = dsoc_limit + soc;                  /* :341-344 */
        else                            curr_soc = new_soc;
    }
    double dsoc = curr_soc - soc;                                   /* :346 */
    *cap = curr_soc * p->B;                                         /* :348 */
    *energy = dsoc * p->B;                                          /* :352 */
    return *energy / (period / 60.0) * 1000.0 / voltage;            /* :355 */
}

/* EV._discharge  ev.py:357-405 */
static double ev_discharge(const ev_params *p, double *cap, double amps, double voltage, int phases,
                           int timescale, double *energy, int *em_cross) {
    voltage = voltage * sqrt((double)phases);                       /* :365 */
    double given_power = (amps * voltage / 1000.0);                 /* :367 */
    double prev_capacity = *cap;
    if (fabs(given_power) > fabs(p->pmax_dis)) given_power = p->pmax_dis;   /* :370-371 */
    double eta;
    if (p->lut_c) eta = lut_get(p->lut_d, p->lut_len, fabs(rint(amps))) / 100.0; /* :375-377 */
    else          eta = p->eta_d;
    double given_energy = given_power * eta * (double)timescale / 60.0;     /* :381 */
    if (*cap + given_energy < p->bmin) {                            /* :382 */
        if (*cap > p->bmin) {
            *energy = -(*cap - p->bmin);
            given_energy = *energy;
            *cap = p->bmin;
        } else {
            *energy = 0.0;
            given_energy = 0.0;
            *cap = p->bmin;                                         /* may RAISE cap :393 */
        }
    } else {
        *energy = given_energy;
        *cap += given_energy;
    }
    *em_cross = (prev_capacity > p->bmin_em && *cap < p->bmin_em);  /* :401-402 */
    return given_energy * 60.0 / (double)timescale * 1000.0 / voltage;      /* :405 */
}

/* EV.step  ev.py:138-186.  Returns 1 if the EV was "active" (amps != 0 after gating). */
static int ev_step(const ev_params *p, double *cap, double amps, double voltage, int phases,
                   int timescale, double *energy, double *actual, double prev_power,
                   int *cycle_inc, int *em_cross) {
    *cycle_inc = 0; *em_cross = 0;
    if (amps > 0 && amps < p->pmin_ac * 1000.0 / (voltage * sqrt((double)phases))) amps = 0;     /* :151-152 */
    else if (amps < 0 && amps > p->pmin_dis * 1000.0 / (voltage * sqrt((double)phases))) amps = 0; /* :153-154 */
    if (amps == 0) { *energy = 0; *actual = 0; return 0; }           /* :158-163 */
    if (prev_power == 0 || (prev_power / amps) < 0) *cycle_inc = 1;  /* :166-167 */
    if (p->ev_phases < phases) phases = p->ev_phases;                /* :169 */
    if (amps > 0) *actual = ev_charge(p, cap, amps, voltage, phases, timescale, energy);
    else          *actual = ev_discharge(p, cap, amps, voltage, phases, timescale, energy, em_cross);
    *cap = my_ceil2(*cap);                                           /* :183 */
    return 1;
}

static void session_params(const ev2o_scenario *sc, int s, ev_params *p) {
    p->B = sc->s_B[s]; p->pmax_ac = sc->s_pmax_ac[s]; p->pmin_ac = sc->s_pmin_ac[s];
    p->pmax_dis = sc->s_pmax_dis[s]; p->pmin_dis = sc->s_pmin_dis[s];
    p->bmin = sc->s_bmin[s]; p->bmin_em = sc->s_bmin_em[s];
    p->ts = sc->s_ts[s]; p->mult = sc->s_mult[s];
    p->eta_c = sc->s_eta_c[s]; p->eta_d = sc->s_eta_d[s];
    p->ev_phases = sc->s_ev_phases[s];
    p->lut_len = sc->lut_len;
    if (sc->s_lut[s] >= 0) {
        p->lut_c = sc->luts_c + (size_t)sc->s_lut[s] * sc->lut_len;
        p->lut_d = sc->luts_d + (size_t)sc->s_lut[s] * sc->lut_len;
    } else { p->lut_c = NULL; p->lut_d = NULL; }
}

double ev2o_ev_step(const double *q, int ev_phases, double amps, double voltage, int phases,
                    int timescale, double *energy, double *actual_amps) {
    ev_params p;
    double cap = q[0];
    p.B = q[1]; p.pmax_ac = q[2]; p.pmin_ac = q[3]; p.pmax_dis = q[4]; p.pmin_dis = q[5];
    p.bmin = q[6]; p.bmin_em = 0; p.ts = q[7]; p.mult = q[8]; p.eta_c = q[9]; p.eta_d = q[10];
    p.ev_phases = ev_phases; p.lut_c = p.lut_d = NULL; p.lut_len = 0;
    int ci, ec;
    ev_step(&p, &cap, amps, voltage, phases, timescale, energy, actual_amps, 0.0, &ci, &ec);
    return cap;
}

/* ------------------------------------------------------------------------- */
/* state functions  (state.py)                                               */

int ev2o_obs_dim(const ev2o_topology *tp, int kind) {
    switch (kind) {
    case EV2O_STATE_PUBLIC_PST:            return 3 + 3 * tp->P;               /* state.py:11-57 */
    case EV2O_STATE_V2G_PROFIT_MAX:        return 2 + 20 + 2 * tp->P;          /* state.py:70-102 */
    case EV2O_STATE_V2G_PROFIT_MAX_LOADS:  return 2 + 20 + 40 * tp->Tr + 2 * tp->P; /* state.py:113-151 */
    case EV2O_STATE_V2G_GRID:              return 6 + 2 * tp->n_bus + 3 * tp->P;    /* state.py:216-278 */
    default: return 0;
    }
}

/* Transformer.get_power_limits  transformer.py:142-171 (horizon fixed to 20 by state.py:131) */
static void power_limits(const ev2o_topology *tp, const ev2o_scenario *sc, int tr, int step, double *out) {
    const int H = 20;
    const double *mp = sc->tr_max_power + (size_t)tr * tp->T;
    double limit = mp[0];
    for (int t = 1; t < tp->T; ++t) if (mp[t] > limit) limit = mp[t];   /* max(self.max_power) */
    for (int k = 0; k < H; ++k) out[k] = limit;
    for (int e = 0; e < sc->dr_count[tr]; ++e) {
        int s0 = sc->dr_start[tr * sc->n_dr + e], s1 = sc->dr_end[tr * sc->n_dr + e];
        double cap = sc->dr_cap[tr * sc->n_dr + e];
        if (step + tp->dr_steps_ahead >= s0 && s1 >= step) {
            double v = limit - limit * cap / 100.0;
            int a, b;
            if (step > s0) { a = 0; b = s1 - step; }
            else { a = s0 - step; if (a < 0) a = -a; b = s1 - step; if (b < 0) b = -b; }
            if (b > H) b = H;
            for (int k = a; k < b; ++k) out[k] = v;
        }
    }
}

/* Transformer.get_load_pv_forecast  transformer.py:173-188, including the in-place
 * write-through of the actual value into the (sliced view of the) forecast arrays. */
static void load_pv_forecast(const ev2o_topology *tp, const ev2o_scenario *sc, ev2o_state *st,
                             int tr, int step, double *loads_minus_pv) {
    const int H = 20, T = tp->T;
    double *lf = st->load_fc_live + (size_t)tr * T, *pf = st->pv_fc_live + (size_t)tr * T;
    if (step < T) {
        lf[step] = sc->tr_infl[(size_t)tr * T + step];
        pf[step] = sc->tr_solar[(size_t)tr * T + step];
    }
    for (int k = 0; k < H; ++k) {
        int i = step + k;
        double l = (i < T) ? lf[i] : lf[T - 1];
        double p = (i < T) ? pf[i] : pf[T - 1];
        loads_minus_pv[k] = l - p;
    }
}

static void write_obs(const ev2o_topology *tp, const ev2o_scenario *sc, ev2o_state *st, int kind, double *obs) {
    if (!obs || kind == EV2O_STATE_NONE) return;
    const int T = tp->T, t = st->current_step;
    int o = 0;
    double prev_usage = st->usage[(t - 1 + T) % T];  /* python negative index at t == 0 */
    if (kind == EV2O_STATE_V2G_GRID) {                               /* V2G_grid_state  state.py:216-278 */
        const int n = tp->n_bus;
        obs[o++] = sc->date_feat[t * 3 + 0]; obs[o++] = sc->date_feat[t * 3 + 1]; obs[o++] = sc->date_feat[t * 3 + 2];
        obs[o++] = (t < T) ? sc->charge_price[t] : 0.0;              /* charge_prices[0, t:t+1], zero padded */
        obs[o++] = (t < T) ? sc->setpoint[t] : 0.0;
        obs[o++] = prev_usage;
        /* node_active_power[1:, step-1] holds the base powers of step `step` (grid.step returns the NEXT step's) */
        for (int i = 0; i < n; ++i) obs[o++] = sc->grid_active[(size_t)t * n + i];
        for (int i = 0; i < n; ++i) obs[o++] = sc->grid_reactive[(size_t)t * n + i];
        for (int c = 0; c < tp->C; ++c)
            for (int p = tp->cs_port_off[c]; p < tp->cs_port_off[c + 1]; ++p) {
                int s = st->port_session[p];
                if (s >= 0) { obs[o++] = st->port_cap[p]; obs[o++] = (double)(sc->s_t_dep[s] - t + 1); obs[o++] = (double)tp->cs_tr[c]; }
                else { obs[o++] = 0; obs[o++] = 0; obs[o++] = 0; }
            }
        return;
    }
    if (kind == EV2O_STATE_PUBLIC_PST) {
        obs[o++] = (double)t / (double)T;
        obs[o++] = (t < T) ? sc->setpoint[t] : 0.0;
        obs[o++] = prev_usage;
    } else {
        obs[o++] = (double)t;
        obs[o++] = prev_usage;
        for (int k = 0; k < 20; ++k) obs[o++] = (t + k < T) ? fabs(sc->charge_price[t + k]) : 0.0;
    }
    for (int tr = 0; tr < tp->Tr; ++tr) {
        if (kind == EV2O_STATE_V2G_PROFIT_MAX_LOADS) {
            load_pv_forecast(tp, sc, st, tr, t, obs + o); o += 20;
            power_limits(tp, sc, tr, t, obs + o);         o += 20;
        }
        for (int c = 0; c < tp->C; ++c) {
            if (tp->cs_tr[c] != tr) continue;
            for (int p = tp->cs_port_off[c]; p < tp->cs_port_off[c + 1]; ++p) {
                int s = st->port_session[p];
                if (kind == EV2O_STATE_PUBLIC_PST) {
                    if (s >= 0) {
                        double soc = st->port_cap[p] / sc->s_B[s];
                        obs[o++] = (soc == 1.0) ? 1.0 : 0.5;
                        obs[o++] = st->port_energy_exch[p];
                        obs[o++] = (double)(t - sc->s_t_arr[s]);
                    } else { obs[o++] = 0; obs[o++] = 0; obs[o++] = 0; }
                } else {
                    if (s >= 0) {
                        obs[o++] = st->port_cap[p] / sc->s_B[s];
                        obs[o++] = (double)(sc->s_t_dep[s] - t);
                    } else { obs[o++] = 0; obs[o++] = 0; }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* reset / step                                                              */

void ev2o_reset(const ev2o_topology *tp, const ev2o_scenario *sc, ev2o_state *st, int state_kind, double *obs0) {
    const int P = tp->P, C = tp->C, T = tp->T, Tr = tp->Tr;
    st->current_step = 0; st->total_evs_spawned = 0; st->current_evs_parked = 0; st->done = 0;
    st->total_reward = 0;
    for (int p = 0; p < P; ++p) {
        st->port_session[p] = -1; st->port_cap[p] = 0; st->port_energy_exch[p] = 0; st->port_abs_energy[p] = 0;
        st->port_prev_power[p] = 0; st->port_required[p] = 0; st->port_cur_energy[p] = 0; st->port_cur_amps[p] = 0;
        st->port_cycles[p] = 0; st->port_em_metric[p] = 0;
    }
    for (int c = 0; c < C; ++c) {
        st->cs_total_charged[c] = st->cs_total_discharged[c] = st->cs_total_profits[c] = st->cs_total_sat[c] = 0;
        st->cs_total_served[c] = 0;
    }
    memset(st->usage, 0, sizeof(double) * T);
    memset(st->potential, 0, sizeof(double) * T);
    memset(st->tr_overload_hist, 0, sizeof(double) * Tr * T);
    memset(st->cs_power_hist, 0, sizeof(double) * C * T);
    memset(st->cs_current_hist, 0, sizeof(double) * C * T);
    for (int i = 0; i < sc->n_sessions; ++i) {
        st->ev_spawned[i] = 0; st->ev_final_cap[i] = 0; st->ev_afap[i] = 0; st->ev_soc_sum[i] = 0; st->ev_n_hist[i] = 0;
        st->ev_abs_energy[i] = 0; st->ev_em_metric[i] = 0; st->ev_n_act[i] = 0;
    }
    if (tp->n_bus > 0) memset(st->node_voltage, 0, sizeof(double) * (tp->n_bus + 1) * T);
    memcpy(st->load_fc_live, sc->tr_load_fc, sizeof(double) * Tr * T);
    memcpy(st->pv_fc_live, sc->tr_pv_fc, sizeof(double) * Tr * T);
    write_obs(tp, sc, st, state_kind, obs0);
}

/* calculate_charge_power_potential  utils.py:760-791 (called after current_step += 1) */
static double charge_power_potential(const ev2o_topology *tp, const ev2o_scenario *sc, const ev2o_state *st) {
    double power_potential = 0;
    for (int c = 0; c < tp->C; ++c) {
        double cs_pot = 0;
        for (int p = tp->cs_port_off[c]; p < tp->cs_port_off[c + 1]; ++p) {
            int s = st->port_session[p];
            if (s < 0) continue;
            if (st->port_cap[p] / sc->s_B[s] < 1.0 && sc->s_t_dep[s] > st->current_step) {
                int phases = tp->cs_phases[c] < sc->s_ev_phases[s] ? tp->cs_phases[c] : sc->s_ev_phases[s];
                double ev_current = sc->s_pmax_ac[s] * 1000.0 / (sqrt((double)phases) * tp->cs_voltage[c]);
                double current = tp->cs_imax[c] < ev_current ? tp->cs_imax[c] : ev_current;
                cs_pot += sqrt((double)phases) * tp->cs_voltage[c] * current / 1000.0;
            }
        }
        double max_cs = sqrt((double)tp->cs_phases[c]) * tp->cs_voltage[c] * tp->cs_imax[c] / 1000.0;
        double min_cs = sqrt((double)tp->cs_phases[c]) * tp->cs_voltage[c] * tp->cs_imin[c] / 1000.0;
        if (cs_pot > max_cs) power_potential += max_cs;
        else if (cs_pot < min_cs) power_potential += 0;
        else power_potential += cs_pot;
    }
    return power_potential;
}

int ev2o_step(const ev2o_topology *tp, const ev2o_scenario *sc, ev2o_state *st,
              const double *actions_in, int reward_kind, int state_kind, ev2o_out *out) {
    const int P = tp->P, C = tp->C, T = tp->T, Tr = tp->Tr, t = st->current_step;
    if (st->done) return 2;                                          /* ev2gym_env.py:343 */
    out->error = 0;

    double total_costs = 0;
    int total_invalid = 0, n_departed = 0, n_arrived = 0;
    double sat_exp_sum = 0;      /* sum over departing EVs of exp(-10*score)  reward.py:41-42,83-85 */
    double user_costs = 0, loss_v = 0;
    double dep_score[P > 0 ? P : 1], dep_capv[P > 0 ? P : 1], dep_des[P > 0 ? P : 1];   /* env.departing_evs of this step, in order */
    int n_dep_rec = 0;
    double tr_power[Tr > 0 ? Tr : 1], tr_amps[Tr > 0 ? Tr : 1];
    double act[P > 0 ? P : 1];
    memcpy(act, actions_in, sizeof(double) * P);
    if (out->dep_sat) for (int p = 0; p < P; ++p) out->dep_sat[p] = NAN;

    /* Transformer.reset(step)  transformer.py:258-267 */
    for (int k = 0; k < Tr; ++k) {
        tr_power[k] = sc->tr_infl[(size_t)k * T + t] + sc->tr_solar[(size_t)k * T + t];
        tr_amps[k] = (tr_power[k] * 1000.0) / tp->tr_voltage;
    }

    /* charger loop  ev2gym_env.py:363-385 -> EV_Charger.step ev_charger.py:114-233 */
    for (int c = 0; c < C; ++c) {
        const int lo = tp->cs_port_off[c], n = tp->cs_n_ports[c];
        const double cp = sc->charge_price[t], dp = sc->discharge_price[t];
        double profit = 0, power_out = 0, total_amps = 0;
        int invalid = 0;
        for (int j = 0; j < n; ++j)                                  /* :137-140 */
            if (st->port_session[lo + j] < 0) { act[lo + j] = 0; invalid++; }
        double sum = 0;                                              /* python sum(): 0 + a0 + a1 ... */
        for (int j = 0; j < n; ++j) sum = sum + act[lo + j];
        double norm[n > 0 ? n : 1];
        for (int j = 0; j < n; ++j) {                                /* :143-149 */
            if (sum > 1)       norm[j] = act[lo + j] / sum;
            else if (sum < -1) norm[j] = -act[lo + j] / sum;
            else               norm[j] = act[lo + j];
        }
        for (int j = 0; j < n; ++j) {                                /* :155-205 */
            const int p = lo + j, s = st->port_session[p];
            double action = np_round5(norm[j]);                      /* :157 */
            double amps = 0, energy = 0, actual = 0;
            int stepped = 0, active = 0, ci = 0, ec = 0;
            ev_params prm;
            double soc_before = 0;
            if (s >= 0) { session_params(sc, s, &prm); soc_before = st->port_cap[p] / prm.B; }
            if (action == 0 && s >= 0) {                             /* :162-165  ev.step(0, V) */
                stepped = 1;
                active = ev_step(&prm, &st->port_cap[p], 0.0, tp->cs_voltage[c], 1, tp->timescale,
                                 &energy, &actual, st->port_prev_power[p], &ci, &ec);
            } else if (action > 0) {                                 /* :167-181 */
                amps = action * tp->cs_imax[c];
                if (amps < tp->cs_imin[c] - 0.01) amps = 0;
                stepped = 1;
                active = ev_step(&prm, &st->port_cap[p], amps, tp->cs_voltage[c], tp->cs_phases[c], tp->timescale,
                                 &energy, &actual, st->port_prev_power[p], &ci, &ec);
                profit += fabs(energy) * cp;
                st->cs_total_charged[c] += fabs(energy);
                power_out += energy * 60.0 / (double)tp->timescale;
                total_amps += actual;
            } else if (action < 0) {                                 /* :183-197 */
                amps = action * fabs(tp->cs_imax_dis[c]);
                if (amps > tp->cs_imin_dis[c] - 0.01) amps = tp->cs_imin_dis[c];
                stepped = 1;
                active = ev_step(&prm, &st->port_cap[p], amps, tp->cs_voltage[c], tp->cs_phases[c], tp->timescale,
                                 &energy, &actual, st->port_prev_power[p], &ci, &ec);
                profit += fabs(energy) * dp;
                st->cs_total_discharged[c] += fabs(energy);
                power_out += energy * 60.0 / (double)tp->timescale;
                total_amps += actual;
            }
            if (stepped) {
                /* EV.step bookkeeping for get_battery_degradation: historic_soc / active_steps  ev.py:156,162,185 */
                st->ev_soc_sum[s] += soc_before; st->ev_n_hist[s] += 1;
                if (active && actual != 0) st->ev_act_soc[(size_t)s * T + st->ev_n_act[s]++] = soc_before;
                st->port_cur_energy[p] = energy; st->port_cur_amps[p] = actual;
                if (active) {                                        /* ev.py:166-183 bookkeeping */
                    st->port_cycles[p] += ci;
                    st->port_prev_power[p] = energy;
                    st->port_energy_exch[p] += energy;
                    st->port_abs_energy[p] += fabs(energy);
                    st->port_em_metric[p] += ec;
                    st->ev_abs_energy[s] += fabs(energy); st->ev_em_metric[s] += ec;
                    if (amps > 0) st->port_required[p] -= energy;    /* ev.py:353 (after gating amps keeps its sign) */
                    else          st->port_required[p] += energy;    /* ev.py:399 */
                }
            }
            if (total_amps - 0.0001 > tp->cs_imax[c]) out->error = 1;   /* :203-205 raise Exception */
        }
        for (int j = 0; j < n; ++j) if (st->port_session[lo + j] >= 0) st->ev_final_cap[st->port_session[lo + j]] = st->port_cap[lo + j];
        st->cs_total_profits[c] += profit;                          /* :207 */
        /* departures, with the charger's own step counter == t   :209-231 */
        for (int j = 0; j < n; ++j) {
            const int p = lo + j, s = st->port_session[p];
            if (s < 0) continue;
            if (t >= sc->s_t_dep[s]) {                               /* ev.py:199 */
                double capn = st->port_cap[p], des = sc->s_desired[s];
                double sat = (capn < des - 0.001) ? capn / des : 1.0;    /* ev.py:211-214 */
                st->port_session[p] = -1;
                st->cs_total_served[c] += 1;
                st->cs_total_sat[c] += sat;
                sat_exp_sum += 100.0 * exp(-10.0 * sat);
                user_costs += -py_sq(capn - des);                        /* reward.py:99-102 */
                dep_score[n_dep_rec] = sat; dep_capv[n_dep_rec] = capn; dep_des[n_dep_rec] = des; n_dep_rec++;   /* env.departing_evs order */
                if (out->dep_sat) out->dep_sat[p] = sat;
                n_departed++;
            }
        }
        st->usage[t] += power_out;                                   /* ev2gym_env.py:375 */
        tr_amps[tp->cs_tr[c]] += total_amps;                         /* transformer.py:273-274 */
        tr_power[tp->cs_tr[c]] += power_out;
        total_costs += profit;
        total_invalid += invalid;
        if (out->cs_power) out->cs_power[c] = power_out;
        if (out->cs_current) out->cs_current[c] = total_amps;
        st->cs_power_hist[(size_t)c * T + t] = power_out;            /* ev2gym_env.py:534-535 */
        st->cs_current_hist[(size_t)c * T + t] = total_amps;
    }

    /* distribution grid: node_ev_power <- tr.current_power; PowerGrid.step  ev2gym_env.py:387-397, grid.py:120-141 */
    out->pf_iterations = 0;
    if (tp->n_bus > 0) {
        const int n = tp->n_bus;
        double Sr[n], Si[n], vr[n], vi[n], lr[n], li[n], nr[n], ni[n];
        for (int i = 0; i < n; ++i) {
            double act = sc->grid_active[(size_t)t * n + i] + tr_power[i];       /* grid.py:121 */
            Sr[i] = act / tp->s_base; Si[i] = sc->grid_reactive[(size_t)t * n + i] / tp->s_base;   /* grid_tensor.py:594-598 */
            vr[i] = 1.0; vi[i] = 0.0;                                            /* flat start */
        }
        int it = 0; double tol = INFINITY;
        while (it < 100 && tol >= 1e-6) {                                        /* numbarize.py:301-315 */
            for (int i = 0; i < n; ++i) {
                /* 1 / v0 (numpy complex division, Smith), then S * that, then conj */
                double a = vr[i], b = vi[i], rr, ri;
                if (fabs(a) >= fabs(b)) { double rat = b / a, scl = 1.0 / (a + b * rat); rr = (1.0 + 0.0 * rat) * scl; ri = (0.0 - 1.0 * rat) * scl; }
                else { double rat = a / b, scl = 1.0 / (b + a * rat); rr = (1.0 * rat + 0.0) * scl; ri = (0.0 * rat - 1.0) * scl; }
                lr[i] = Sr[i] * rr - Si[i] * ri;
                li[i] = -(Sr[i] * ri + Si[i] * rr);
            }
            tol = 0;
            for (int r = 0; r < n; ++r) {
                double zr = 0, zi = 0;
                for (int c2 = 0; c2 < n; ++c2) {
                    const double kr = tp->grid_K[2 * ((size_t)r * n + c2)], ki = tp->grid_K[2 * ((size_t)r * n + c2) + 1];
                    zr += kr * lr[c2] - ki * li[c2];
                    zi += kr * li[c2] + ki * lr[c2];
                }
                nr[r] = zr + tp->grid_L[2 * r]; ni[r] = zi + tp->grid_L[2 * r + 1];
                double d = fabs(hypot(nr[r], ni[r]) - hypot(vr[r], vi[r]));
                if (d > tol) tol = d;
            }
            for (int r = 0; r < n; ++r) { vr[r] = nr[r]; vi[r] = ni[r]; }
            it++;
        }
        out->pf_iterations = it;
        st->node_voltage[(size_t)0 * T + t] = 1.0;                               /* grid.py:127-129: slack = 1 */
        for (int i = 0; i < n; ++i) st->node_voltage[(size_t)(i + 1) * T + t] = hypot(vr[i], vi[i]);
        for (int i = 0; i <= n; ++i) {                                           /* reward.py:107-110 */
            double vm = st->node_voltage[(size_t)i * T + t];
            double x = 0.05 - fabs(1.0 - vm);
            loss_v += x < 0 ? x : 0.0;
            if (out->node_vm) out->node_vm[i] = vm;
        }
    }

    /* spawn EVs arriving at t+1  ev2gym_env.py:399-417, ev_charger.py:266-285 */
    for (int i = st->total_evs_spawned; i < sc->n_sessions; ++i) {
        if (sc->s_t_arr[i] == t + 1) {
            int c = sc->s_loc[i], idx = -1;
            for (int p = tp->cs_port_off[c]; p < tp->cs_port_off[c + 1]; ++p)
                if (st->port_session[p] < 0) { idx = p; break; }     /* evs_connected.index(None) */
            if (idx < 0) { out->error = 3; break; }                  /* assert n_evs_connected < n_ports */
            st->port_session[idx] = i;
            st->port_cap[idx] = sc->s_cap0[i];                       /* EV.reset() ev.py:115-136 */
            st->port_energy_exch[idx] = 0; st->port_abs_energy[idx] = 0; st->port_prev_power[idx] = 0;
            st->port_required[idx] = sc->s_B[i] - sc->s_cap0[i];
            st->port_cur_energy[idx] = 0; st->port_cur_amps[idx] = 0;
            st->port_cycles[idx] = 0; st->port_em_metric[idx] = 0;
            st->ev_spawned[i] = 1; st->ev_final_cap[i] = sc->s_cap0[i];
            {   /* EV.calculate_max_energy_with_AFAP(cs.get_max_power())  ev.py:407-440, ev_charger.py:251-252,279 */
                double max_cs_power = tp->cs_imax[c] * tp->cs_voltage[c] * sqrt((double)tp->cs_phases[c]) / 1000.0;
                double max_power = fabs(max_cs_power) > fabs(sc->s_pmax_ac[i]) ? sc->s_pmax_ac[i] : max_cs_power;
                double eff = sc->s_eta_c[i];
                if (sc->s_lut[i] >= 0) {
                    const double *l = sc->luts_c + (size_t)sc->s_lut[i] * sc->lut_len;
                    double m = 0; for (int k = 0; k < sc->lut_len; ++k) if (l[k] > m) m = l[k];
                    eff = m / 100.0;
                }
                double afap = sc->s_cap0[i];
                for (int k = sc->s_t_arr[i]; k < sc->s_t_dep[i] + 1; ++k) {
                    afap += max_power * eff * (double)tp->timescale / 60.0;
                    afap = my_ceil2(afap);
                    if (afap > sc->s_B[i]) { afap = sc->s_B[i]; break; }
                }
                st->ev_afap[i] = afap;
            }
            st->total_evs_spawned++;
            n_arrived++;
        } else if (sc->s_t_arr[i] > t + 1) break;
    }

    /* _update_power_statistics  ev2gym_env.py:520-531; overload uses tr.current_step == t */
    double overload_sum = 0;
    for (int k = 0; k < Tr; ++k) {
        double mx = sc->tr_max_power[(size_t)k * T + t], mn = sc->tr_min_power[(size_t)k * T + t];
        double ov = 0;
        if (tr_power[k] > mx + 0.0001 || tr_power[k] < mn - 0.0001) ov = fabs(tr_power[k] - mx); /* transformer.py:284-300 */
        st->tr_overload_hist[(size_t)k * T + t] = ov;
        overload_sum += 100.0 * ov;                                  /* reward.py:38-39 */
        if (out->tr_power) out->tr_power[k] = tr_power[k];
        if (out->tr_amps) out->tr_amps[k] = tr_amps[k];
        if (out->tr_overload) out->tr_overload[k] = ov;
    }

    st->current_step += 1;                                           /* :421 */
    if (st->current_step < T)                                        /* :424-426 */
        st->potential[st->current_step] = charge_power_potential(tp, sc, st);
    st->current_evs_parked += n_arrived - n_departed;                /* :428 */

    /* reward  ev2gym_env.py:579-586 */
    double reward = 0;
    const int tm1 = st->current_step - 1;
    switch (reward_kind) {
    case EV2O_REWARD_SQ_TRACKING: {                                  /* reward.py:11-12 */
        double m = sc->setpoint[tm1] < st->potential[tm1] ? sc->setpoint[tm1] : st->potential[tm1];
        double d = m - st->usage[tm1];
        reward = -py_sq(d);
    } break;
    case EV2O_REWARD_PROFIT_TR_USER: {                               /* reward.py:36-44: costs, then -= per tr, then -= per EV */
        reward = total_costs;
        for (int k = 0; k < Tr; ++k) reward -= 100.0 * st->tr_overload_hist[(size_t)k * T + tm1];
        for (int p = 0; p < P; ++p)
            if (out->dep_sat ? !isnan(out->dep_sat[p]) : 0) reward -= 100.0 * exp(-10.0 * out->dep_sat[p]);
        if (!out->dep_sat) reward -= sat_exp_sum;
    } break;
    case EV2O_REWARD_PROFIT_MAX: {                                   /* reward.py:81-87 */
        reward = total_costs;
        for (int p = 0; p < P; ++p)
            if (out->dep_sat ? !isnan(out->dep_sat[p]) : 0) reward -= 100.0 * exp(-10.0 * out->dep_sat[p]);
        if (!out->dep_sat) reward -= sat_exp_sum;
    } break;
    case EV2O_REWARD_GRID_FULL:   reward = total_costs + 1000.0 * loss_v + user_costs; break;   /* reward.py:89-111 */
    case EV2O_REWARD_GRID_SIMPLE: reward = 1000.0 * loss_v; break;                              /* reward.py:114-121 */
    case EV2O_REWARD_SQTR_TR_USER: {                                 /* reward.py:16-32 */
        double m = sc->setpoint[tm1];
        if (st->potential[tm1] < m) m = st->potential[tm1];
        if (sc->tr_max_power[tm1] < m) m = sc->tr_max_power[tm1];   /* transformers[0].max_power[t] */
        reward = -py_sq(m - st->usage[tm1]);
        for (int k = 0; k < Tr; ++k) reward -= 100.0 * st->tr_overload_hist[(size_t)k * T + tm1];
        for (int i = 0; i < n_dep_rec; ++i) reward -= 1000.0 * (1.0 - dep_score[i]);
    } break;
    case EV2O_REWARD_SQ_TRACKING_PENALTY: {                          /* reward.py:46-58 */
        double m = sc->setpoint[tm1] < st->potential[tm1] ? sc->setpoint[tm1] : st->potential[tm1];
        int prev = st->current_step - 2;                             /* python index: -1 wraps to the last entry */
        if (prev < 0) prev += T;
        reward = -py_sq(m - st->usage[tm1]);
        if (st->usage[tm1] == 0 && st->potential[prev] != 0) reward = reward - 100;
    } break;
    case EV2O_REWARD_SIMPLE:                                         /* reward.py:60-65 */
        reward = -py_sq(sc->setpoint[tm1] - st->usage[tm1]);
        break;
    case EV2O_REWARD_MIN_TRACKER_SURPLUS:                            /* reward.py:67-76 */
        reward = 0;
        if (sc->setpoint[tm1] < st->usage[tm1]) reward -= py_sq(st->usage[tm1] - sc->setpoint[tm1]);
        reward += st->usage[tm1];
        break;
    case EV2O_REWARD_V2G_PROFITMAX: {                                /* reward.py:123-148 */
        double uc = 0;
        for (int i = 0; i < n_dep_rec; ++i) if (dep_des[i] > dep_capv[i]) uc += -100.0 * (dep_des[i] - dep_capv[i]);
        reward = total_costs + uc;
    } break;
    case EV2O_REWARD_V2G_COSTS_SIMPLE: reward = total_costs; break;  /* reward.py:150-153 */
    case EV2O_REWARD_V2G_PROFITMAX_V2: case EV2O_REWARD_GRID_PROFITMAX_V2: case EV2O_REWARD_PST_PROFITMAX_V2: {
        /* reward.py:155-213 (and the grid / pst variants :215-339): EVs that can no longer reach their desired level */
        double uc = 0;
        const double mult = 0.05, per_hour = 60.0 / (double)tp->timescale;
        for (int p = 0; p < P; ++p) {                                /* connected EVs, incl. the ones that just arrived */
            const int s = st->port_session[p];
            if (s < 0) continue;
            const double des = sc->s_desired[s], capn = st->port_cap[p], pmax = sc->s_pmax_ac[s];
            const double min_steps = (des - capn) / (pmax / per_hour);
            const int departing_step = sc->s_t_dep[s] - st->current_step;
            if (min_steps > (double)departing_step) {
                const double min_cap = des - ((double)(departing_step + 1) * pmax / per_hour);
                uc += -(mult * py_sq(min_cap - capn));
            }
        }
        for (int i = 0; i < n_dep_rec; ++i)
            if (dep_des[i] > dep_capv[i]) uc += -mult * py_sq(dep_des[i] - dep_capv[i]);
        if (reward_kind == EV2O_REWARD_GRID_PROFITMAX_V2) reward = total_costs + uc + 50000.0 * loss_v;
        else if (reward_kind == EV2O_REWARD_PST_PROFITMAX_V2) {
            double viol = 0;
            if (sc->setpoint[tm1] < st->usage[tm1]) viol += sc->setpoint[tm1] - st->usage[tm1];
            reward = total_costs + uc + 1000.0 * viol;
        } else reward = total_costs + uc;
    } break;
    default: reward = 0;
    }
    (void)overload_sum;
    st->total_reward += reward;

    /* _check_termination  ev2gym_env.py:449-496 */
    if (out->action_mask)
        for (int p = 0; p < P; ++p) out->action_mask[p] = st->port_session[p] >= 0 ? 1.0 : 0.0;
    if (st->current_step >= T) st->done = 1;
    write_obs(tp, sc, st, state_kind, out->obs);
    if (out->actions_eff) memcpy(out->actions_eff, act, sizeof(double) * P);

    out->reward = reward; out->total_costs = total_costs; out->done = st->done;
    out->invalid_actions = total_invalid; out->n_departed = n_departed; out->n_arrived = n_arrived;
    return out->error ? 1 : 0;
}


/* ------------------------------------------------------------------------- */
/* get_statistics  utils.py:12-123                                           */

static double mean_of(const double *v, int n) { double s = 0; for (int i = 0; i < n; ++i) s += v[i]; return n ? s / n : NAN; }

void ev2o_statistics(const ev2o_topology *tp, const ev2o_scenario *sc, const ev2o_state *st, double *out) {
    const int C = tp->C, T = tp->T, Tr = tp->Tr;
    double served = 0, profits = 0, ch = 0, dis = 0, avg_sat_sum = 0; int n_served_cs = 0;
    for (int c = 0; c < C; ++c) {
        served += st->cs_total_served[c]; profits += st->cs_total_profits[c];
        ch += st->cs_total_charged[c]; dis += st->cs_total_discharged[c];
        if (st->cs_total_served[c] > 0) { avg_sat_sum += st->cs_total_sat[c] / st->cs_total_served[c]; n_served_cs++; }
    }
    double tr_ov = 0;
    for (int i = 0; i < Tr * T; ++i) tr_ov += st->tr_overload_hist[i];
    double te = 0, ete = 0, viol = 0;                                  /* utils.py:31-46 */
    for (int t = 0; t < T; ++t) {
        double d = sc->setpoint[t] - st->usage[t];
        te += py_sq(d); ete += fabs(d);                                /* utils.py:37-38 */
        if (st->usage[t] > sc->setpoint[t]) viol += st->usage[t] - sc->setpoint[t];
    }
    ete *= (double)tp->timescale / 60.0;
    /* per EV: energy user satisfaction + battery degradation (ev.py:442-521) */
    const double e0 = 7.543e6, e1 = 23.75e6, e2 = 6976, z0 = 7.348e-3, z1 = 3.667, z2 = 7.6e-4, z3 = 4.081e-3;
    const double b_cap_ah = 2.05, b_cap_kwh = 78, d_dist = 15000, b_age = 2 * 365, G = 0.186;
    const double theta = 298.15, k = 0.8263, v_min = 3.3324;
    double dcal = 0, dcyc = 0, eus_sum = 0, eus_sq = 0, eus_min = INFINITY; int n_ev = 0, em = 0;
    for (int i = 0; i < sc->n_sessions; ++i) {
        if (!st->ev_spawned[i]) continue;
        const double B = sc->s_B[i], final_soc = st->ev_final_cap[i] / B;
        double T_sim = (double)(sc->s_t_dep[i] - sc->s_t_arr[i] + 1) * (double)tp->timescale / (60.0 * 24.0);
        double avg_soc = (st->ev_soc_sum[i] + final_soc) / (double)(st->ev_n_hist[i] + 1);
        double v_avg = v_min + k * avg_soc;
        double alpha = (e0 * v_avg - e1) * exp(-e2 / theta);
        double d_cal = alpha * 0.75 * T_sim / pow(b_age, 0.25);
        const double *f = st->ev_act_soc + (size_t)i * T;
        int nf = st->ev_n_act[i];
        double fsum = final_soc; for (int j = 0; j < nf; ++j) fsum += f[j];
        double avg_f = fsum / (double)(nf + 1);
        double dev = fabs(avg_f - final_soc); for (int j = 0; j < nf; ++j) dev += fabs(avg_f - f[j]);
        double delta_DoD = 2.0 * (dev / (double)(nf + 1));
        double v_half = v_min + k * 0.5;
        double beta = z0 * py_sq(v_half - z1) + z2 + z3 * delta_DoD;          /* ev.py:508 */
        double Q_sim = (st->ev_abs_energy[i] / b_cap_kwh) * b_cap_ah;
        double Q_acc = 2 * (b_age * (d_dist / 365) * G * b_cap_ah) / b_cap_kwh;
        double d_cyc = beta * 0.5 * Q_sim / pow(Q_acc, 0.5);
        dcal += d_cal; dcyc += d_cyc;
        double r = (st->ev_final_cap[i] / st->ev_afap[i]) * 100.0;       /* utils.py:59-62 */
        eus_sum += r; eus_sq += r * r; if (r < eus_min) eus_min = r; n_ev++;
        em += st->ev_em_metric[i];
    }
    double eus_mean = n_ev ? eus_sum / n_ev : NAN;
    double eus_var = 0;                                                 /* np.std: two-pass population variance */
    for (int i = 0; i < sc->n_sessions; ++i) {
        if (!st->ev_spawned[i]) continue;
        double r = (st->ev_final_cap[i] / st->ev_afap[i]) * 100.0;
        eus_var += (r - eus_mean) * (r - eus_mean);
    }
    eus_var = n_ev ? eus_var / n_ev : NAN;
    (void)eus_sq;
    out[EV2O_STAT_EV_SERVED] = served; out[EV2O_STAT_PROFITS] = profits; out[EV2O_STAT_ENERGY_CHARGED] = ch;
    out[EV2O_STAT_ENERGY_DISCHARGED] = dis; out[EV2O_STAT_AVG_USER_SAT] = n_served_cs ? avg_sat_sum / n_served_cs : NAN;
    out[EV2O_STAT_TRACKER_VIOLATION] = viol; out[EV2O_STAT_TRACKING_ERROR] = te; out[EV2O_STAT_ENERGY_TRACKING_ERROR] = ete;
    out[EV2O_STAT_ENERGY_USER_SAT] = eus_mean; out[EV2O_STAT_STD_ENERGY_USER_SAT] = eus_var > 0 ? sqrt(eus_var) : 0.0;
    out[EV2O_STAT_MIN_ENERGY_USER_SAT] = eus_min; out[EV2O_STAT_EMERGENCY_STEPS] = em; out[EV2O_STAT_TR_OVERLOAD] = tr_ov;
    out[EV2O_STAT_DEGRADATION] = dcal + dcyc; out[EV2O_STAT_DEGRADATION_CAL] = dcal; out[EV2O_STAT_DEGRADATION_CYC] = dcyc;
    out[EV2O_STAT_TOTAL_REWARD] = st->total_reward;
    (void)mean_of;
}
