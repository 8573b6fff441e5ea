/*
 * ev2o_batch.c -- ORACLE (test infrastructure).  Runs many independent oracle envs on a pool
 * of POSIX threads (static partition of the env range).  The reference itself has no batching
 * or threading: this is its best case restated in C, used as the CPU
 * baseline of bench.py (`cpu_baseline.kind = "port"`) and by `bench.py --impl reference`.
 */
#include "ev2o.h"
#include <pthread.h>
#include <unistd.h>

int ev2o_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const ev2o_topology *tp; const ev2o_scenario *const *scs; ev2o_state *sts;
    const double *actions; int reward_kind, state_kind; ev2o_out *outs; double *reward; int *done;
    int lo, hi, err;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (int e = j->lo; e < j->hi; ++e) {
        int rc = ev2o_step(j->tp, j->scs[e], &j->sts[e], j->actions + (long)e * j->tp->P,
                           j->reward_kind, j->state_kind, &j->outs[e]);
        if (j->reward) j->reward[e] = j->outs[e].reward;
        if (j->done) j->done[e] = j->outs[e].done;
        j->err |= rc;
    }
    return 0;
}

/* Step envs [0,E) once.  actions: [E*P] float64; reward/done: [E] or NULL.
 * `outs` are E caller-prepared ev2o_out blocks (their array pointers may be NULL). */
int ev2o_step_batch(const ev2o_topology *tp, const ev2o_scenario *const *scs, ev2o_state *sts,
                    int E, const double *actions, int reward_kind, int state_kind,
                    ev2o_out *outs, double *reward, int *done, int n_threads) {
    if (n_threads <= 0) n_threads = ev2o_max_threads();
    if (n_threads > E) n_threads = E > 0 ? E : 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; job_t jobs[256];
    int err = 0;
    for (int i = 0; i < n_threads; ++i) {
        job_t j = { tp, scs, sts, actions, reward_kind, state_kind, outs, reward, done,
                    (int)((long)E * i / n_threads), (int)((long)E * (i + 1) / n_threads), 0 };
        jobs[i] = j;
        if (i > 0) pthread_create(&th[i], 0, worker, &jobs[i]);
    }
    worker(&jobs[0]);
    for (int i = 1; i < n_threads; ++i) pthread_join(th[i], 0);
    for (int i = 0; i < n_threads; ++i) err |= jobs[i].err;
    return err;
}
