/*
 * ev2o.h -- ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C, single-threaded-per-env, float64 restatement of the reference's
 * EV2Gym.step() cascade.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (ev2gym_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this code against
 * traces recorded from the unmodified Python reference executed in the build
 * container (tools/make_golden.py -> tests/golden/ npz files) and against the
 * known-answer table of SURVEY.md section 8c.
 *
 * Reference files restated (paths relative to /root/reference):
 *   ev2gym/models/ev2gym_env.py:333-447   EV2Gym.step
 *   ev2gym/models/ev2gym_env.py:449-496   _check_termination
 *   ev2gym/models/ev2gym_env.py:520-556   _update_power_statistics
 *   ev2gym/models/ev_charger.py:114-233   EV_Charger.step
 *   ev2gym/models/ev_charger.py:266-285   EV_Charger.spawn_ev
 *   ev2gym/models/ev.py:138-186           EV.step
 *   ev2gym/models/ev.py:240-355           EV._charge
 *   ev2gym/models/ev.py:357-405           EV._discharge
 *   ev2gym/models/transformer.py:258-302  Transformer.reset/step/is_overloaded/get_how_overloaded
 *   ev2gym/models/transformer.py:142-188  get_power_limits / get_load_pv_forecast
 *   ev2gym/utilities/utils.py:760-791     calculate_charge_power_potential
 *   ev2gym/rl_agent/reward.py:7-14,34-44,78-87   three stock rewards
 *   ev2gym/rl_agent/state.py:6-63,65-106,108-155 three stock state functions
 */
#ifndef EV2O_H
#define EV2O_H

#ifdef __cplusplus
extern "C" {
#endif

enum { EV2O_REWARD_NONE = 0, EV2O_REWARD_SQ_TRACKING = 1, EV2O_REWARD_PROFIT_TR_USER = 2,
       EV2O_REWARD_PROFIT_MAX = 3, EV2O_REWARD_GRID_FULL = 4, EV2O_REWARD_GRID_SIMPLE = 5,
       EV2O_REWARD_SQTR_TR_USER = 6,       /* SqTrError_TrPenalty_UserIncentives       reward.py:16-32   */
       EV2O_REWARD_SIMPLE = 7,             /* SimpleReward                             reward.py:60-65   */
       EV2O_REWARD_MIN_TRACKER_SURPLUS = 8,/* MinimizeTrackerSurplusWithChargeRewards  reward.py:67-76   */
       EV2O_REWARD_V2G_PROFITMAX = 9,      /* V2G_profitmax                            reward.py:123-148 */
       EV2O_REWARD_V2G_COSTS_SIMPLE = 10,  /* V2G_costs_simple                         reward.py:150-153 */
       EV2O_REWARD_V2G_PROFITMAX_V2 = 11,  /* V2G_profitmaxV2                          reward.py:155-213 */
       EV2O_REWARD_GRID_PROFITMAX_V2 = 12, /* Grid_V2G_profitmaxV2                     reward.py:215-279 */
       EV2O_REWARD_PST_PROFITMAX_V2 = 13,  /* pst_V2G_profitmaxV2                      reward.py:281-339 */
       EV2O_REWARD_SQ_TRACKING_PENALTY = 14 /* SquaredTrackingErrorRewardWithPenalty   reward.py:46-58   */ };
enum { EV2O_STATE_NONE = 0, EV2O_STATE_PUBLIC_PST = 1, EV2O_STATE_V2G_PROFIT_MAX = 2,
       EV2O_STATE_V2G_PROFIT_MAX_LOADS = 3, EV2O_STATE_V2G_GRID = 4 };

/* Static layout shared by all envs. */
typedef struct {
    int C, P, Tr, T, timescale, dr_steps_ahead;
    const int    *cs_n_ports;    /* [C]   */
    const int    *cs_port_off;   /* [C+1] */
    const int    *cs_tr;         /* [C]   connected_transformer */
    const int    *cs_phases;     /* [C]   */
    const double *cs_imax, *cs_imin, *cs_imax_dis, *cs_imin_dis, *cs_voltage; /* [C] */
    double tr_voltage;           /* voltage*sqrt(phases), transformer.py:39-40 */
    /* distribution grid (simulate_grid): Laurent power flow  grid.py:120-141, numbarize.py:268-325 */
    int n_bus;                   /* buses without the slack (== Tr), 0 = no grid */
    const double *grid_K;        /* [n_bus*n_bus] complex128 row-major as (re,im) pairs */
    const double *grid_L;        /* [n_bus] complex128 */
    double s_base;
} ev2o_topology;

/* One env's pre-sampled episode. */
typedef struct {
    const double *charge_price, *discharge_price, *setpoint;           /* [T] */
    const double *tr_infl, *tr_solar, *tr_max_power, *tr_min_power;    /* [Tr*T] */
    const double *tr_load_fc, *tr_pv_fc;                               /* [Tr*T] */
    int n_dr;                                                          /* events per transformer (padded) */
    const int    *dr_start, *dr_end, *dr_count;                        /* [Tr*n_dr], [Tr*n_dr], [Tr] */
    const double *dr_cap;                                              /* [Tr*n_dr] */
    int n_sessions;
    const int    *s_loc, *s_t_arr, *s_t_dep, *s_ev_phases, *s_lut;     /* [S] */
    const double *s_cap0, *s_B, *s_pmax_ac, *s_pmin_ac, *s_pmax_dis, *s_pmin_dis,
                 *s_bmin, *s_bmin_em, *s_desired, *s_ts, *s_mult, *s_eta_c, *s_eta_d; /* [S] */
    int n_luts, lut_len;
    const double *luts_c, *luts_d;                                     /* [n_luts*lut_len] percent */
    const double *grid_active, *grid_reactive;                         /* [(T+1)*n_bus] base bus powers of steps 0..T */
    const double *date_feat;                                           /* [(T+1)*3] weekday/7, sin, cos at obs time */
} ev2o_scenario;

/* Mutable env state; all arrays are caller-allocated. */
typedef struct {
    int current_step, total_evs_spawned, current_evs_parked, done;
    double total_reward;
    int    *port_session;        /* [P] session index of the connected EV, -1 = empty (evs_connected) */
    double *port_cap;            /* [P] current_capacity */
    double *port_energy_exch;    /* [P] total_energy_exchanged */
    double *port_abs_energy;     /* [P] abs_total_energy_exchanged */
    double *port_prev_power;     /* [P] previous_power */
    double *port_required;       /* [P] required_energy */
    double *port_cur_energy;     /* [P] current_energy of the last EV.step */
    double *port_cur_amps;       /* [P] actual_current of the last EV.step */
    int    *port_cycles;         /* [P] charging_cycles */
    int    *port_em_metric;      /* [P] min_emergency_battery_capacity_metric */
    double *cs_total_charged, *cs_total_discharged, *cs_total_profits, *cs_total_sat; /* [C] */
    int    *cs_total_served;     /* [C] */
    double *usage;               /* [T] current_power_usage */
    double *potential;           /* [T] charge_power_potential */
    double *tr_overload_hist;    /* [Tr*T] */
    double *cs_power_hist, *cs_current_hist; /* [C*T] */
    double *node_voltage;        /* [(n_bus+1)*T] |V| per node and step, slack first  ev2gym_env.py:397 */
    double *load_fc_live, *pv_fc_live;       /* [Tr*T] forecasts incl. the write-through of transformer.py:178-180 */
    /* per spawned EV (index = session index; env.EVs keeps departed EVs), for get_statistics  utils.py:12-123 */
    int    *ev_spawned;          /* [S] 1 once spawned */
    double *ev_final_cap;        /* [S] current_capacity as of now / at departure */
    double *ev_afap;             /* [S] max_energy_AFAP  ev.py:407-440 */
    double *ev_soc_sum;          /* [S] sum(historic_soc) */
    int    *ev_n_hist;           /* [S] len(historic_soc) */
    double *ev_abs_energy;       /* [S] abs_total_energy_exchanged */
    int    *ev_em_metric;        /* [S] min_emergency_battery_capacity_metric */
    int    *ev_n_act;            /* [S] number of steps with actual_current != 0 */
    double *ev_act_soc;          /* [S*T] historic_soc at those steps */
} ev2o_state;

enum { EV2O_STAT_EV_SERVED = 0, EV2O_STAT_PROFITS, EV2O_STAT_ENERGY_CHARGED, EV2O_STAT_ENERGY_DISCHARGED,
       EV2O_STAT_AVG_USER_SAT, EV2O_STAT_TRACKER_VIOLATION, EV2O_STAT_TRACKING_ERROR, EV2O_STAT_ENERGY_TRACKING_ERROR,
       EV2O_STAT_ENERGY_USER_SAT, EV2O_STAT_STD_ENERGY_USER_SAT, EV2O_STAT_MIN_ENERGY_USER_SAT,
       EV2O_STAT_EMERGENCY_STEPS, EV2O_STAT_TR_OVERLOAD, EV2O_STAT_DEGRADATION, EV2O_STAT_DEGRADATION_CAL,
       EV2O_STAT_DEGRADATION_CYC, EV2O_STAT_TOTAL_REWARD, EV2O_STAT_COUNT };

/* Per-step outputs; arrays caller-allocated, any may be NULL. */
typedef struct {
    double reward, total_costs;
    int done, invalid_actions, n_departed, n_arrived, error;
    double *cs_power, *cs_current;             /* [C]  */
    double *tr_power, *tr_amps, *tr_overload;  /* [Tr] */
    double *dep_sat;                           /* [P] user satisfaction of the EV that left port p this step, NaN otherwise */
    double *action_mask;                       /* [P] */
    double *obs;                               /* [obs_dim] */
    double *actions_eff;                       /* [P] actions after the empty-port zeroing (ev_charger.py:137-140) */
    double *node_vm;                           /* [n_bus+1] |V| of this step */
    int pf_iterations;
} ev2o_out;

int  ev2o_obs_dim(const ev2o_topology *tp, int state_kind);
void ev2o_reset(const ev2o_topology *tp, const ev2o_scenario *sc, ev2o_state *st,
                int state_kind, double *obs0);
/* returns 0 ok, 1 = amps overflow exception (ev_charger.py:203-205), 2 = stepped a done env */
int  ev2o_step(const ev2o_topology *tp, const ev2o_scenario *sc, ev2o_state *st,
               const double *actions, int reward_kind, int state_kind, ev2o_out *out);

/* Micro entry point for the known-answer table: one EV.step (ev.py:138-186) on explicit params.
 * p = {cap, B, pmax_ac, pmin_ac, pmax_dis, pmin_dis, bmin, ts, mult, eta_c, eta_d}; returns new cap,
 * writes energy and amps. */
/* get_statistics(env)  ev2gym/utilities/utils.py:12-123 (+ EV.get_battery_degradation ev.py:442-521) */
void ev2o_statistics(const ev2o_topology *tp, const ev2o_scenario *sc, const ev2o_state *st, double *out);

double ev2o_ev_step(const double *p, int ev_phases, double amps, double voltage, int phases,
                    int timescale, double *energy, double *actual_amps);

#ifdef __cplusplus
}
#endif
#endif
