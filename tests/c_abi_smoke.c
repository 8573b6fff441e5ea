/* c_abi_smoke.c -- a C consumer of include/ev2b.h (compiled and run by tests/test_gpu_c_abi.py on the GPU box).
 *
 * Every other caller in this repository is a ctypes mirror of the header (ev2gym_b200/_lib.py), which could drift from
 * it silently: only symbol NAMES are compared on the CPU.  This program is compiled against the header itself, links
 * libev2b.so, builds a 2-charger env with one hand-written scenario, steps a whole episode through ev2b_step_host (host
 * buffers only: no CUDA API is needed on this side of the boundary) and checks the observations against the battery
 * model evaluated right here in plain C for the simplest case (transition_soc = 1, scalar efficiency: ev.py:295-306,
 * 346-355, 183).  Exit code 0 = pass.
 *
 *   gcc -O1 -ffp-contract=off -Iinclude tests/c_abi_smoke.c -Lev2gym_b200/csrc -lev2b -Wl,-rpath,$PWD/ev2gym_b200/csrc -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ev2b.h"

#define T 12
#define C 2
#define TR 1
#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "c_abi_smoke: " __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

int main(void) {
    CHECK(ev2b_abi_version() == EV2B_ABI_VERSION, "header says ABI %d, library says %d", EV2B_ABI_VERSION, ev2b_abi_version());
    CHECK(sizeof(ev2b_dims) == 48, "ev2b_dims is %zu bytes, expected 48 (the ctypes mirrors assume it)", sizeof(ev2b_dims));

    ev2b_dims d;
    memset(&d, 0, sizeof d);
    d.n_envs = 3; d.n_chargers = C; d.n_transformers = TR; d.sim_length = T; d.timescale = 15; d.dr_steps_ahead = 4;
    d.reward_kind = EV2B_REWARD_PROFIT_MAX; d.state_kind = EV2B_STATE_V2G_PROFIT_MAX; d.tr_voltage = 400.0 * sqrt(3.0);

    int32_t n_ports[C] = {1, 1}, cs_tr[C] = {0, 0}, phases[C] = {3, 3};
    double imax[C] = {32, 32}, imin[C] = {0, 0}, imax_dis[C] = {-32, -32}, imin_dis[C] = {0, 0}, volt[C] = {230, 230};
    ev2b_topology tp;
    memset(&tp, 0, sizeof tp);
    tp.cs_n_ports = n_ports; tp.cs_tr = cs_tr; tp.cs_phases = phases; tp.cs_imax = imax; tp.cs_imin = imin;
    tp.cs_imax_dis = imax_dis; tp.cs_imin_dis = imin_dis; tp.cs_voltage = volt;

    ev2b_handle *h = NULL;
    int rc = ev2b_create(&d, &tp, 0, &h);
    CHECK(rc == EV2B_OK, "ev2b_create: %d (%s)", rc, ev2b_last_error(NULL));
    CHECK(ev2b_n_ports(h) == 2 && ev2b_obs_dim(h) == 22 + 2 * 2, "ports %d, obs dim %d", ev2b_n_ports(h), ev2b_obs_dim(h));

    /* one scenario: an EV on charger 1 from step 2 to step 7, 50 kWh battery at 20 kWh, 11 kW, eta = 0.95, transition_soc = 1 */
    double cp[T], dp[T], sp[T], infl[T], solar[T], maxp[T], minp[T], lfc[T], pvfc[T];
    for (int t = 0; t < T; ++t) { cp[t] = -0.10 - 0.01 * t; dp[t] = 0.08; sp[t] = 0; infl[t] = 5; solar[t] = -1; maxp[t] = 100; minp[t] = -100; lfc[t] = 5; pvfc[t] = 1; }
    int32_t dr_start[1] = {0}, dr_end[1] = {0}, dr_count[1] = {0};
    double dr_cap[1] = {0};
    int64_t sess_off[2] = {0, 1}, lut_off[2] = {0, 0};
    int32_t s_loc[1] = {1}, s_ta[1] = {2}, s_td[1] = {7}, s_ph[1] = {3}, s_lut[1] = {-1};
    double s_cap0[1] = {20}, s_B[1] = {50}, s_pmax[1] = {11}, s_pmin[1] = {0}, s_pmd[1] = {-11}, s_pmind[1] = {0}, s_bmin[1] = {5},
           s_bem[1] = {10}, s_des[1] = {50}, s_ts[1] = {1.0}, s_mult[1] = {5}, s_ec[1] = {0.95}, s_ed[1] = {0.95};
    double lut_dummy[1] = {0};
    ev2b_scenarios b;
    memset(&b, 0, sizeof b);
    b.n = 1; b.n_dr = 1; b.lut_len = 101;
    b.charge_price = cp; b.discharge_price = dp; b.setpoint = sp; b.tr_infl = infl; b.tr_solar = solar;
    b.tr_max_power = maxp; b.tr_min_power = minp; b.tr_load_fc = lfc; b.tr_pv_fc = pvfc;
    b.dr_start = dr_start; b.dr_end = dr_end; b.dr_cap = dr_cap; b.dr_count = dr_count; b.sess_off = sess_off;
    b.s_loc = s_loc; b.s_t_arr = s_ta; b.s_t_dep = s_td; b.s_ev_phases = s_ph; b.s_lut = s_lut;
    b.s_cap0 = s_cap0; b.s_B = s_B; b.s_pmax_ac = s_pmax; b.s_pmin_ac = s_pmin; b.s_pmax_dis = s_pmd; b.s_pmin_dis = s_pmind;
    b.s_bmin = s_bmin; b.s_bmin_em = s_bem; b.s_desired = s_des; b.s_ts = s_ts; b.s_mult = s_mult; b.s_eta_c = s_ec; b.s_eta_d = s_ed;
    b.lut_off = lut_off; b.luts_c = lut_dummy; b.luts_d = lut_dummy;
    rc = ev2b_load_scenarios(h, &b);
    CHECK(rc == EV2B_OK, "ev2b_load_scenarios: %d (%s)", rc, ev2b_last_error(h));
    CHECK(ev2b_n_scenarios(h) == 1, "n_scenarios");
    rc = ev2b_reset(h, 0, d.n_envs, NULL, NULL, NULL);
    CHECK(rc == EV2B_OK, "ev2b_reset: %d (%s)", rc, ev2b_last_error(h));

    const int E = d.n_envs, P = 2, D = ev2b_obs_dim(h);
    float *act = calloc((size_t)E * P, sizeof(float)), *obs = calloc((size_t)E * D, sizeof(float));
    double *rew = calloc(E, sizeof(double));
    uint32_t *st = calloc(E, sizeof(uint32_t));
    for (int i = 0; i < E * P; ++i) act[i] = 1.0f;                 /* ChargeAsFastAsPossible */

    /* the model, right here: charger 32 A x 230 V x sqrt(3) = 12.7 kW > the EV's 11 kW -> saturated at pmax */
    double cap = 20.0;
    const double veff = 230.0 * sqrt(3.0), B = 50.0, eta = 0.95, c60 = 60.0 / 15.0;
    double total_reward = 0.0;
    for (int t = 0; t < T; ++t) {
        rc = ev2b_step_host(h, act, EV2B_F32, rew, st, obs, NULL);
        CHECK(rc == EV2B_OK, "ev2b_step_host: %d (%s)", rc, ev2b_last_error(h));
        double expect_reward = 0.0;
        if (t >= 2 && t <= 7) {                                   /* connected during steps t_arr .. t_dep */
            double pilot = eta * (1.0 * 32.0) * veff / 1000.0 / B / c60;
            const double maxd = eta * 11.0 / B / c60;
            if (pilot > maxd) pilot = maxd;
            const double soc = cap / B;
            double nsoc = pilot + soc;
            if (nsoc > 1.0) nsoc = 1.0;
            const double energy = (nsoc - soc) * B;
            cap = ceil(nsoc * B * 100.0) / 100.0;
            expect_reward = fabs(energy) * cp[t];
            if (t == 7) { const double sat = cap < 50.0 - 0.001 ? cap / 50.0 : 1.0; expect_reward -= 100.0 * exp(-10.0 * sat); }
        }
        for (int e = 0; e < E; ++e) {
            const float *o = obs + (size_t)e * D;
            CHECK(o[0] == (float)(t + 1), "env %d step %d: obs[0] = %g", e, t, o[0]);
            const int connected_after = (t + 1 >= 2 && t + 1 <= 7);
            const float soc_obs = o[22 + 2 * 1], left = o[22 + 2 * 1 + 1];       /* port 1 = charger 1 */
            if (connected_after) {
                CHECK(fabsf(soc_obs - (float)(cap / B)) < 1e-6f, "env %d step %d: soc %g, expected %g", e, t, soc_obs, cap / B);
                CHECK(left == (float)(7 - (t + 1)), "env %d step %d: steps to departure %g", e, t, left);
            } else {
                CHECK(soc_obs == 0.f && left == 0.f, "env %d step %d: empty port shows %g %g", e, t, soc_obs, left);
            }
            CHECK(o[22] == 0.f && o[23] == 0.f, "charger 0 is never used");
            CHECK(fabs(rew[e] - expect_reward) <= 1e-9 * fmax(1.0, fabs(expect_reward)), "env %d step %d: reward %.12g, expected %.12g", e, t, rew[e], expect_reward);
            CHECK(((st[e] & EV2B_ST_DONE) != 0) == (t == T - 1), "env %d step %d: status %u", e, t, st[e]);
        }
        total_reward += expect_reward;
    }
    rc = ev2b_step_host(h, act, EV2B_F32, rew, st, obs, NULL);     /* stepping a finished env  (ev2gym_env.py:343) */
    CHECK(rc == EV2B_OK && (st[0] & EV2B_ST_WAS_DONE), "WAS_DONE expected, status %u", st[0]);
    CHECK(ev2b_launch_count(h) > 0 && ev2b_kernel_launches(h, 0) + ev2b_kernel_launches(h, 1) >= T + 1, "launch counters");
    ev2b_destroy(h);
    printf("c_abi_smoke ok: %d envs x %d steps, final battery level %.2f kWh, episode reward %.6f\n", E, T, cap, total_reward);
    free(act); free(obs); free(rew); free(st);
    return 0;
}
