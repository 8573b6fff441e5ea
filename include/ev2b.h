/*
 * ev2b.h -- C ABI of the B200-native batched EV2Gym environment-step engine (libev2b.so).
 *
 * The reference (StavrosOrf/EV2Gym) has NO FFI on this path: its boundary is the Python call
 * convention of `EV2Gym.reset()/step()` (ev2gym/models/ev2gym_env.py:243-331, 333-447).  This
 * header is the C surface a binding for that boundary binds instead;
 * each entry point names the reference code it replaces.  Plain pointers and sizes only: no
 * torch / C++ types cross this boundary.  The Python facade (ev2gym_b200/env.py) and the
 * ctypes stub shown in INTEGRATION.md sit directly on top of it.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; ev2b_last_error() gives the text
 *   - one handle per GPU, not thread-safe (the reference env is not re-entrant either)
 *   - `stream` is a cudaStream_t passed as void*; all device work is stream-ordered,
 *     nothing synchronises the host unless documented (ev2b_step_host, ev2b_read_*)
 *   - "env" = one replica of the reference's EV2Gym object; "port" index is charger-major,
 *     i.e. the reference's action order (ev2gym_env.py:363-385)
 *   - arithmetic: float64, same operation order as the reference, FMA contraction off
 *     (battery level carries a ceil(x*100)/100 every active step, ev2gym/models/ev.py:183)
 */
#ifndef EV2B_H
#define EV2B_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EV2B_ABI_VERSION 2

/* reward_function / state_function fused on the device (ev2gym/rl_agent/reward.py, state.py). */
enum ev2b_reward_kind {
    EV2B_REWARD_NONE = 0,                 /* reward output = 0 (caller computes its own)            */
    EV2B_REWARD_SQ_TRACKING = 1,          /* SquaredTrackingErrorReward            reward.py:7-14   */
    EV2B_REWARD_PROFIT_TR_USER = 2,       /* ProfitMax_TrPenalty_UserIncentives    reward.py:34-44  */
    EV2B_REWARD_PROFIT_MAX = 3,           /* profit_maximization                   reward.py:78-87  */
    EV2B_REWARD_GRID_FULL = 4,            /* V2G_grid_full_reward   (needs a grid) reward.py:89-111 */
    EV2B_REWARD_GRID_SIMPLE = 5,          /* V2G_grid_simple_reward (needs a grid) reward.py:114-121 */
    /* the remaining stock rewards run in the kernel's full-featured instantiation (same as statistics mode) */
    EV2B_REWARD_SQTR_TR_USER = 6,         /* SqTrError_TrPenalty_UserIncentives    reward.py:16-32   */
    EV2B_REWARD_SIMPLE = 7,               /* SimpleReward                          reward.py:60-65   */
    EV2B_REWARD_MIN_TRACKER_SURPLUS = 8,  /* MinimizeTrackerSurplusWithChargeRewards reward.py:67-76 */
    EV2B_REWARD_V2G_PROFITMAX = 9,        /* V2G_profitmax                         reward.py:123-148 */
    EV2B_REWARD_V2G_COSTS_SIMPLE = 10,    /* V2G_costs_simple                      reward.py:150-153 */
    EV2B_REWARD_V2G_PROFITMAX_V2 = 11,    /* V2G_profitmaxV2                       reward.py:155-213 */
    EV2B_REWARD_GRID_PROFITMAX_V2 = 12,   /* Grid_V2G_profitmaxV2 (needs a grid)   reward.py:215-279 */
    EV2B_REWARD_PST_PROFITMAX_V2 = 13,    /* pst_V2G_profitmaxV2                   reward.py:281-339 */
    EV2B_REWARD_SQ_TRACKING_PENALTY = 14  /* SquaredTrackingErrorRewardWithPenalty reward.py:46-58   */
};
enum ev2b_state_kind {
    EV2B_STATE_NONE = 0,
    EV2B_STATE_PUBLIC_PST = 1,            /* PublicPST            state.py:6-63    D = 3 + 3P        */
    EV2B_STATE_V2G_PROFIT_MAX = 2,        /* V2G_profit_max       state.py:65-106  D = 22 + 2P       */
    EV2B_STATE_V2G_PROFIT_MAX_LOADS = 3,  /* V2G_profit_max_loads state.py:108-155 D = 22 + 40Tr + 2P */
    EV2B_STATE_V2G_GRID = 4               /* V2G_grid_state       state.py:216-278 D = 6 + 2(nb-1) + 3P */
};
enum ev2b_action_dtype { EV2B_F32 = 0, EV2B_F64 = 1 };
/* On-device agents for ev2b_step_k (no action tensor is read). */
enum ev2b_agent_kind {
    EV2B_AGENT_EXTERNAL = 0,      /* actions_k[k,E,P] supplied by the caller                                   */
    EV2B_AGENT_AFAP = 1,          /* ChargeAsFastAsPossible: all ones        ev2gym/baselines/heuristics.py:161-166 */
    EV2B_AGENT_ZERO = 2,          /* DoNothing: all zeros                    heuristics.py (DoNothing)          */
    EV2B_AGENT_UNIFORM = 3,       /* RandomAgent-like: uniform in the action space, counter-based hash RNG      */
    EV2B_AGENT_ROUNDROBIN = 4,    /* RoundRobin: setpoint-sized rotating queue of waiting EVs   heuristics.py:7-95    */
    EV2B_AGENT_CALAP = 5          /* ChargeAsLateAsPossible                                     heuristics.py:98-150  */
};

/* error codes */
enum {
    EV2B_OK = 0, EV2B_E_ARG = -1, EV2B_E_CUDA = -2, EV2B_E_STATE = -3, EV2B_E_SCENARIO = -4,
    EV2B_E_LIMIT = -5
};

/* per-env status bits written by the step kernel (ev2b_step_out.status) */
#define EV2B_ST_DONE          1u   /* current_step >= simulation_length       ev2gym_env.py:460        */
#define EV2B_ST_AMPS_OVERFLOW 2u   /* the reference would `raise Exception`   ev_charger.py:203-205    */
#define EV2B_ST_WAS_DONE      4u   /* step() on a finished env (reference: AssertionError, :343)       */

typedef struct ev2b_handle ev2b_handle;

/* Problem sizes; replaces the size-bearing part of the YAML config (ev2gym_env.py:78-128). */
typedef struct {
    int32_t n_envs;            /* E  env replicas resident on this GPU                        */
    int32_t n_chargers;        /* C  number_of_charging_stations                              */
    int32_t n_transformers;    /* Tr                                                          */
    int32_t sim_length;        /* T  simulation_length                                        */
    int32_t timescale;         /* minutes per step                                            */
    int32_t dr_steps_ahead;    /* notification_of_event_minutes // timescale (transformer.py:69) */
    int32_t reward_kind;       /* enum ev2b_reward_kind                                       */
    int32_t state_kind;        /* enum ev2b_state_kind                                        */
    double  tr_voltage;        /* voltage*sqrt(phases) of the config   (transformer.py:39-40) */
    int32_t flags;             /* EV2B_F_* */
    int32_t reserved;
} ev2b_dims;

#define EV2B_F_STATS 1         /* keep the per-EV histories get_statistics() needs (utils.py:12-123, ev.py:442-521) */

/* Static charger layout, HOST pointers; replaces load_ev_charger_profiles / cs_transformers
 * (ev2gym/utilities/loaders.py:299-365, 495-498).  Arrays have n_chargers entries. */
typedef struct {
    const int32_t *cs_n_ports;     /* n_ports                      */
    const int32_t *cs_tr;          /* connected_transformer        */
    const int32_t *cs_phases;      /* phases (1..3)                */
    const double  *cs_imax;        /* max_charge_current           */
    const double  *cs_imin;        /* min_charge_current           */
    const double  *cs_imax_dis;    /* max_discharge_current (<=0)  */
    const double  *cs_imin_dis;    /* min_discharge_current        */
    const double  *cs_voltage;     /* voltage                      */
    /* distribution grid (simulate_grid: True): Laurent power flow V <- K conj(S/V) + L per env and step
     * (ev2gym/models/grid.py:120-141, grid_utility/numbarize.py:268-325, K/L: grid_tensor.py:110-118).
     * n_bus = buses without the slack; must equal n_transformers (loaders.py:481).  0 = no grid. */
    int32_t        n_bus;
    const double  *grid_K;         /* [n_bus*n_bus] complex128 row-major as (re,im) pairs */
    const double  *grid_L;         /* [n_bus] complex128 */
    double         grid_s_base;    /* kVA */
} ev2b_topology;

/* A bank of n pre-sampled episodes ("scenarios"), HOST pointers, plain float64/int32 exactly as
 * the reference objects hold them after reset() (ev2gym_env.py:293-296).  The library replays
 * the first-free-port rule (ev_charger.py:273), de-duplicates EV specs / efficiency tables and
 * packs everything for the device.  Sessions of scenario i are rows sess_off[i]..sess_off[i+1]
 * (arrival-sorted, = env.EVs_profiles order); its efficiency tables are rows
 * lut_off[i]..lut_off[i+1] of luts_c/luts_d and s_lut indexes them locally (-1 = scalar). */
typedef struct {
    int32_t n;                                     /* scenarios in this call                       */
    int32_t n_dr;                                  /* DR events stored per transformer (padded)    */
    int32_t lut_len;                               /* entries per efficiency table (101)           */
    const double *charge_price, *discharge_price;  /* [n*T]      loaders.py:439-460 (row 0)        */
    const double *setpoint;                        /* [n*T]      power_setpoints                   */
    const double *tr_infl, *tr_solar;              /* [n*Tr*T]   inflexible_load, solar_power (<=0)*/
    const double *tr_max_power, *tr_min_power;     /* [n*Tr*T]   after DR events                   */
    const double *tr_load_fc, *tr_pv_fc;           /* [n*Tr*T]   forecasts                         */
    const int32_t *dr_start, *dr_end;              /* [n*Tr*n_dr]                                  */
    const double  *dr_cap;                         /* [n*Tr*n_dr] capacity_percentage              */
    const int32_t *dr_count;                       /* [n*Tr]                                       */
    const int64_t *sess_off;                       /* [n+1]                                        */
    const int32_t *s_loc, *s_t_arr, *s_t_dep, *s_ev_phases, *s_lut;
    const double  *s_cap0, *s_B, *s_pmax_ac, *s_pmin_ac, *s_pmax_dis, *s_pmin_dis, *s_bmin, *s_bmin_em,
                  *s_desired, *s_ts, *s_mult, *s_eta_c, *s_eta_d;
    const int64_t *lut_off;                        /* [n+1]                                        */
    const double  *luts_c, *luts_d;                /* [lut_off[n]*lut_len] percent                 */
    const double  *grid_active, *grid_reactive;    /* [n*(T+1)*n_bus] base bus powers of steps 0..T (grid.py:109-118,131-139) or NULL */
    const double  *date_feat;                      /* [n*(T+1)*3] weekday/7, sin, cos of the observation times (state.py:221-225) or NULL */
} ev2b_scenarios;

/* Per-step outputs, DEVICE pointers owned by the caller; any pointer may be NULL (= not wanted).
 * Replaces the 5-tuple of EV2Gym.step plus the attributes rewards/states/agents read afterwards.
 * `obs` and `action_mask` are updated IN PLACE: when a step is given the same buffer as the previous step (or as the
 * ev2b_reset before it), only the entries that can have changed are written (ports whose EV stayed, arrived or left,
 * the header, the price / forecast windows); a different pointer gets every entry rewritten.  So a caller that
 * overwrites such a buffer between two steps must hand in another buffer (or reset) -- reading it is always fine.
 * Rows of envs that were ALREADY finished when the step was called (status EV2B_ST_WAS_DONE: the reference raises an
 * AssertionError there, ev2gym_env.py:343) are not touched: `obs` / `action_mask` keep whatever the buffer held. */
typedef struct {
    double   *reward;        /* [E]     reward of this step                  ev2gym_env.py:430-432 */
    uint32_t *status;        /* [E]     EV2B_ST_* bits (done etc.)           ev2gym_env.py:460     */
    float    *obs;           /* [E,D]   state_function(env), float32         ev2gym_env.py:563-565 */
    float    *cs_power;      /* [E,C]   cs.current_power_output (kW)         ev_charger.py:180,196 */
    float    *cs_current;    /* [E,C]   cs.current_total_amps (A)            ev_charger.py:181,197 */
    double   *tr_power;      /* [E,Tr]  tr.current_power                     transformer.py:264-274 */
    double   *tr_overload;   /* [E,Tr]  tr.get_how_overloaded()              transformer.py:292-302 */
    double   *total_costs;   /* [E]     sum of charger profits (`total_costs`) ev2gym_env.py:381   */
    uint8_t  *action_mask;   /* [E,P]   1 if an EV is connected after the step ev2gym_env.py:452-457 */
    double   *dep_sat;       /* [E,P]   user satisfaction of the EV that left this port this step, NaN otherwise */
    double   *dep_cap;       /* [E,P]   its final battery level (kWh), NaN otherwise  (env.departing_evs)         */
    float    *port_energy;   /* [E,P]   ev.current_energy of this step (kWh)                       */
    double   *node_voltage;  /* [E,n_bus+1] |V| per node, slack first   env.node_voltage[:, t]  ev2gym_env.py:397 */
    /* Per-episode histories (init_statistic_variables, utils.py:794-861; written by _update_power_statistics,
     * ev2gym_env.py:520-556): the step that starts at t writes row t of every env, so after an episode the buffers hold
     * what the reference keeps in env.cs_power[C,T], env.cs_current[C,T], env.tr_overload[Tr,T] and
     * env.current_power_usage[T] -- time-major here (coalesced rows).  Rows of steps not yet run keep their old content. */
    float    *hist_cs_power;     /* [E,T,C]  */
    float    *hist_cs_current;   /* [E,T,C]  */
    double   *hist_tr_overload;  /* [E,T,Tr] */
    double   *hist_usage;        /* [E,T]    */
} ev2b_step_out;

/* Raw DEVICE pointers into the struct-of-arrays state, for zero-copy tensor views. */
typedef struct {
    int32_t n_envs, n_ports, n_chargers, n_transformers, obs_dim, n_kpi;
    double   *port_cap;        /* [E,P] current_capacity (kWh), float64                             */
    double   *port_exch;       /* [E,P] total_energy_exchanged (kWh), float64 like the reference's (ev.py:178)  */
    uint32_t *port_hot;        /* [E,P,4] packed session words, see DESIGN.md                      */
    int32_t  *env_step;        /* [E]   current_step                                               */
    int32_t  *env_scn;         /* [E]   scenario id in the bank                                    */
    double   *env_potential;   /* [E]   charge_power_potential[current_step]                       */
    double   *env_usage;       /* [E]   current_power_usage[current_step-1]                        */
    double   *env_kpi;         /* [E,n_kpi] running KPI sums, see enum ev2b_kpi                    */
} ev2b_state_view;

enum ev2b_kpi {
    EV2B_KPI_TOTAL_REWARD = 0, EV2B_KPI_TOTAL_PROFITS, EV2B_KPI_ENERGY_CHARGED, EV2B_KPI_ENERGY_DISCHARGED,
    EV2B_KPI_TR_OVERLOAD, EV2B_KPI_EVS_SERVED, EV2B_KPI_SAT_SUM, EV2B_KPI_TRACKING_ERROR,
    EV2B_KPI_ENERGY_TRACKING_ERROR, EV2B_KPI_TRACKER_VIOLATION, EV2B_KPI_EVS_SPAWNED, EV2B_KPI_INVALID_ACTIONS,
    EV2B_KPI_STEPS, EV2B_KPI_COUNT
};

enum ev2b_stat {
    EV2B_STAT_EV_SERVED = 0, EV2B_STAT_PROFITS, EV2B_STAT_ENERGY_CHARGED, EV2B_STAT_ENERGY_DISCHARGED,
    EV2B_STAT_AVG_USER_SAT, EV2B_STAT_TRACKER_VIOLATION, EV2B_STAT_TRACKING_ERROR, EV2B_STAT_ENERGY_TRACKING_ERROR,
    EV2B_STAT_ENERGY_USER_SAT, EV2B_STAT_STD_ENERGY_USER_SAT, EV2B_STAT_MIN_ENERGY_USER_SAT,
    EV2B_STAT_EMERGENCY_STEPS, EV2B_STAT_TR_OVERLOAD, EV2B_STAT_DEGRADATION, EV2B_STAT_DEGRADATION_CAL,
    EV2B_STAT_DEGRADATION_CYC, EV2B_STAT_TOTAL_REWARD, EV2B_STAT_COUNT
};

int         ev2b_abi_version(void);
const char *ev2b_last_error(const ev2b_handle *h);   /* h may be NULL: error of the last failed create */

/* EV2Gym.__init__ (sizes + topology only)            ev2gym_env.py:38-241 */
int  ev2b_create(const ev2b_dims *dims, const ev2b_topology *topo, int device, ev2b_handle **out);
void ev2b_destroy(ev2b_handle *h);
int  ev2b_obs_dim(const ev2b_handle *h);
int  ev2b_n_ports(const ev2b_handle *h);

/* The scenario part of reset(): load_transformers / load_ev_profiles / load_electricity_prices /
 * load_power_setpoints outputs (ev2gym_env.py:293-296).  Replaces the whole bank. Synchronous. */
int  ev2b_load_scenarios(ev2b_handle *h, const ev2b_scenarios *bank);
int  ev2b_n_scenarios(const ev2b_handle *h);

/* The state part of reset() for envs [env_lo, env_hi): init_statistic_variables + charger /
 * transformer resets + first observation (ev2gym_env.py:298-331).  scn_ids: HOST array of
 * env_hi-env_lo bank indices, or NULL for (env index mod bank size).  obs0: DEVICE [E,D] or NULL
 * (rows env_lo..env_hi are written). */
int  ev2b_reset(ev2b_handle *h, int env_lo, int env_hi, const int32_t *scn_ids, float *obs0, void *stream);

/* EV2Gym.step for all E envs: one fused kernel launch.   ev2gym_env.py:333-447
 * actions: DEVICE [E,P] float32 or float64, in [-1,1]; never written. */
int  ev2b_step(ev2b_handle *h, const void *actions, int action_dtype, const ev2b_step_out *out, void *stream);

/* Same, with HOST buffers (pinned for full speed): H2D(actions) -> kernel -> D2H(reward,status[,obs]),
 * then waits for completion.  This is the call a host-resident agent makes (end-to-end path). */
int  ev2b_step_host(ev2b_handle *h, const void *actions_host, int action_dtype,
                    double *reward_host, uint32_t *status_host, float *obs_host, void *stream);

/* k consecutive steps without returning to the host: `agent_kind` picks an on-device agent (or EXTERNAL with
 * actions_k = DEVICE [k,E,P]).  Outputs in `out` hold the LAST step; KPI sums cover all k.  With auto_reset != 0
 * finished envs restart on their next scenario between steps (ev2b_reset_done).  UNIFORM draws
 * a = low + (1-low) * u, u = (mix32(seed, env*P+port, step) >> 8) * 2^-24, low = -1 if v2g (action_low) else 0. */
int  ev2b_step_k(ev2b_handle *h, int k, int agent_kind, const void *actions_k, int action_dtype, uint64_t seed,
                 double action_low, int auto_reset, const ev2b_step_out *out, void *stream);

/* agent.get_action(env) of a stock heuristic for the CURRENT state of every env  (ev2gym/baselines/heuristics.py):
 * actions_out = DEVICE [E,P] float64 (the reference's agents return float64 arrays), ready for ev2b_step(..., EV2B_F64).
 * AFAP / ZERO / ROUNDROBIN / CALAP.  ROUNDROBIN keeps the reference's per-agent queue (`ev_buffer`) per env inside the
 * handle: every call is one get_action (it rotates the queue), and ev2b_reset / ev2b_reset_done empty the queue of the
 * envs they reset (the reference's scripts build a fresh agent per episode).  Finished envs get zeros.
 * ev2b_step_k accepts the same kinds and calls this before every step. */
int  ev2b_agent_actions(ev2b_handle *h, int agent_kind, double *actions_out, void *stream);

/* Device-side auto-reset of every env whose episode is over: next scenario id = (current id + stride) mod bank size,
 * stride = n_envs mod bank size (1 if that is 0), increased until it is coprime to the bank size -- so consecutive
 * episodes of an env walk the whole bank even when n_envs is a multiple of its size.  For vectorised RL rollouts. */
int  ev2b_reset_done(ev2b_handle *h, float *obs0, void *stream);

int  ev2b_state_view_get(ev2b_handle *h, ev2b_state_view *out);

/* ---- device-side scenario sampling (the EV part of reset(): EV_spawner + spawn_single_EV, utils.py:477-557, 177-345) ----
 * What the reference's spawner reads besides its random draws, HOST pointers (exported once from a reference env by
 * ev2gym_b200/reference_export.spawn_tables_from_env; shipped next to the scenario banks as ev2gym_b200/data/spawn_*.npz). */
typedef struct {
    int32_t workplace;                 /* scenario == "workplace": closed before 6 h, after 18 h, at weekends  :509-520 */
    int32_t heterogeneous;             /* heterogeneous_ev_specs                                                      */
    int32_t empty_ports_at_end;        /* empty_ports_at_end_of_simulation                                 :254-256   */
    int32_t min_stay_steps;            /* config["ev"]["min_time_of_stay"] // timescale                    :495-496   */
    int32_t n_models, n_luts;          /* EV models; efficiency curves (lut_len entries each, percent)                */
    int32_t power_setpoint_enabled;    /* config["power_setpoint_enabled"]: regenerate power_setpoints from the sessions
                                          (generate_power_setpoints, utils.py:664-757); 0: the bank's setpoints stay      */
    int32_t reserved;
    double  spawn_multiplier, desired_frac, min_battery_capacity, min_emergency_battery_capacity, ts_multiplier;
    double  homog_ts, homog_eta_c, homog_eta_d;   /* homogeneous config: transition_soc, efficiencies                 */
    double  power_setpoint_flexibility;           /* config["power_setpoint_flexiblity"] (percent)        utils.py:680 */
    const double *arrival_week, *arrival_weekend;  /* [96] percent per quarter of an hour                  :515, 522  */
    const double *req_energy_mean, *stay_mean;     /* [48] by half hour of arrival (kWh, hours)            :203, 231  */
    const double *model_prob, *model_B, *model_pmax_ac, *model_pmax_dis, *model_pmin_ac, *model_pmin_dis;   /* [n_models] */
    const int32_t *model_phases, *model_lut;       /* [n_models]; model_lut: row of luts or -1                        */
    const double *luts;                            /* [n_luts * 101] the reference uses one curve for both directions  */
} ev2b_spawn_tables;

/* Registers the tables.  Call BEFORE ev2b_load_scenarios (the session table is sized for what the sampler can draw). */
int  ev2b_set_spawn_tables(ev2b_handle *h, const ev2b_spawn_tables *tables);

/* Re-draws the EV sessions of EVERY scenario of the loaded bank on the device (counter-based RNG keyed by `seed`,
 * scenario, port and step: the same seed gives the same sessions on any GPU), keeping each scenario's time series.
 * start: HOST [n_scenarios][3] weekday (0 = Monday), hour, minute of sim_date at reset() for every scenario.
 * Afterwards every env reads as finished until it is reset.  Stream-ordered; a handful of small kernels.
 * With tables->power_setpoint_enabled the power setpoints are regenerated from the new sessions as well. */
int  ev2b_resample_sessions(ev2b_handle *h, uint64_t seed, const int32_t *start, void *stream);

/* The sessions of scenario `scn` as the bank currently holds them (arrival order), HOST arrays of capacity `cap`:
 * returns the number of sessions (or < 0).  port = flat port index the EV will occupy; model = index into the spawn
 * tables' models after ev2b_resample_sessions (the bank's own de-duplicated spec index otherwise); ts / eta_c / eta_d are
 * NaN where the value comes from an efficiency curve or from the spec.  Synchronous. */
int  ev2b_read_sessions(ev2b_handle *h, int scn, int cap, int32_t *port, int32_t *t_arr, int32_t *t_dep, int32_t *model,
                        double *cap0, double *ts, double *eta_c, double *eta_d);

/* env.power_setpoints of scenario `scn` as the bank currently holds them: HOST out[T] (kW).  Synchronous. */
int  ev2b_read_setpoints(ev2b_handle *h, int scn, double *out);

/* get_statistics(env) for every env (ev2gym/utilities/utils.py:12-123, incl. EV.get_battery_degradation
 * ev.py:442-521 and the AFAP bound ev.py:407-440).  Needs EV2B_F_STATS.  out: DEVICE [E, EV2B_STAT_COUNT]
 * float64, columns in enum ev2b_stat order.  Meaningful once an env is done (or at any time for the
 * EVs that already left). */
int  ev2b_episode_stats(ev2b_handle *h, double *out, void *stream);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t ev2b_launch_count(const ev2b_handle *h);
/* step launches by kernel: which = 0 step_kernel (thread per charger), 1 evl_step_kernel (thread per connected EV,
 * handles created under EV2B_KERNEL=evlist, see ev2b_evlist.cuh), 2 evl_rebuild_kernel (list re-derivation) */
int64_t ev2b_kernel_launches(const ev2b_handle *h, int which);

#ifdef __cplusplus
}
#endif
#endif /* EV2B_H */
