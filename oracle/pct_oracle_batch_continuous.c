/* TEST INFRASTRUCTURE — multi-threaded batch driver over the single-env CONTINUOUS oracle: the role of the reference's
 * ShmemVecEnv (wrapper/shmem_vec_env.py:20-156: N independent workers, auto-reset on done) on the host cores for BASELINE
 * config 4 (PackingContinuous, sample_from_distribution U(0.1, 0.5), unit container), with the synthetic policy of SURVEY.md 8(d).
 * Used by tests (full-size parity of the final observations) and by bench.py --continuous (cpu_baseline / --impl reference).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct pctc_env pctc_env;
pctc_env *pctc_create(int setting, double W, double L, double H, int nb_holder, int nl_holder, double low_bound);
void pctc_destroy(pctc_env *e);
void pctc_set_random_sample(pctc_env *e, uint64_t seed, uint64_t gid, double lo, double hi);
void pctc_reset(pctc_env *e, double *obs);
int pctc_step(pctc_env *e, const double *action, int action_len, double *obs, double *reward, int *done, double *info);
uint64_t pcto_rnd_u64(uint64_t seed, uint64_t a, uint64_t b);

typedef struct {
    int setting, nb, nl;
    double W, L, H, lo, hi;
    uint64_t item_seed, policy_seed; int64_t gid_base;
    int n_envs, steps, first, last, phase;
    double *rew_sum; int *n_done; pctc_env **envs; double *obs;
    int64_t t0;
} cjob_t;

static void *cworker(void *arg) {
    cjob_t *j = arg;
    const int ol = (j->nb + j->nl + 1) * 9;
    double act[9], info[3];
    for (int e = j->first; e < j->last; e++) {
        if (j->phase == 0) {
            j->envs[e] = pctc_create(j->setting, j->W, j->L, j->H, j->nb, j->nl, j->lo);
            pctc_set_random_sample(j->envs[e], j->item_seed, (uint64_t)(j->gid_base + e), j->lo, j->hi);
            pctc_reset(j->envs[e], j->obs + (size_t)e * ol);
            j->rew_sum[e] = 0; j->n_done[e] = 0;
            continue;
        }
        double *obs = j->obs + (size_t)e * ol;
        for (int t = 0; t < j->steps; t++) {
            const double *leaf = obs + 9 * j->nb;
            int nvalid = 0;
            for (int k = 0; k < j->nl; k++) nvalid += leaf[9 * k + 8] == 1.0;
            if (nvalid == 0) memset(act, 0, sizeof act);
            else memcpy(act, leaf + 9 * (pcto_rnd_u64(j->policy_seed, (uint64_t)(j->gid_base + e), (uint64_t)(j->t0 + t)) % (uint64_t)nvalid), sizeof act);
            double r; int d;
            pctc_step(j->envs[e], act, 9, obs, &r, &d, info);
            j->rew_sum[e] += r;
            if (d) { j->n_done[e]++; pctc_reset(j->envs[e], obs); }
        }
    }
    return NULL;
}

typedef struct { cjob_t proto; int threads; int64_t t; } pctc_batch;

static void crun_phase(pctc_batch *b, int phase, int steps) {
    int T = b->threads < 1 ? 1 : b->threads;
    if (T > b->proto.n_envs) T = b->proto.n_envs;
    pthread_t *th = malloc(sizeof(pthread_t) * T);
    cjob_t *jobs = malloc(sizeof(cjob_t) * T);
    for (int i = 0; i < T; i++) {
        jobs[i] = b->proto;
        jobs[i].phase = phase; jobs[i].steps = steps; jobs[i].t0 = b->t;
        jobs[i].first = (int)((int64_t)b->proto.n_envs * i / T);
        jobs[i].last = (int)((int64_t)b->proto.n_envs * (i + 1) / T);
        pthread_create(&th[i], NULL, cworker, &jobs[i]);
    }
    for (int i = 0; i < T; i++) pthread_join(th[i], NULL);
    free(th); free(jobs);
}

pctc_batch *pctc_batch_create(int setting, double W, double L, double H, int nb, int nl, double lo, double hi, uint64_t item_seed,
                              uint64_t policy_seed, int64_t gid_base, int n_envs, int threads) {
    pctc_batch *b = calloc(1, sizeof *b);
    cjob_t *p = &b->proto;
    p->setting = setting; p->W = W; p->L = L; p->H = H; p->nb = nb; p->nl = nl; p->lo = lo; p->hi = hi;
    p->item_seed = item_seed; p->policy_seed = policy_seed; p->gid_base = gid_base; p->n_envs = n_envs;
    const int ol = (nb + nl + 1) * 9;
    p->envs = calloc(n_envs, sizeof(pctc_env *));
    p->obs = malloc(sizeof(double) * (size_t)n_envs * ol);
    p->rew_sum = calloc(n_envs, sizeof(double));
    p->n_done = calloc(n_envs, sizeof(int));
    b->threads = threads;
    crun_phase(b, 0, 0);
    return b;
}
/* runs `steps` vector steps; returns elapsed seconds */
double pctc_batch_run(pctc_batch *b, int steps) {
    struct timespec a, c;
    clock_gettime(CLOCK_MONOTONIC, &a);
    crun_phase(b, 1, steps);
    clock_gettime(CLOCK_MONOTONIC, &c);
    b->t += steps;
    return (c.tv_sec - a.tv_sec) + 1e-9 * (c.tv_nsec - a.tv_nsec);
}
void pctc_batch_get(pctc_batch *b, double *obs, double *rew_sum, int *n_done) {
    const int ol = (b->proto.nb + b->proto.nl + 1) * 9;
    if (obs) memcpy(obs, b->proto.obs, sizeof(double) * (size_t)b->proto.n_envs * ol);
    if (rew_sum) memcpy(rew_sum, b->proto.rew_sum, sizeof(double) * b->proto.n_envs);
    if (n_done) memcpy(n_done, b->proto.n_done, sizeof(int) * b->proto.n_envs);
}
void pctc_batch_destroy(pctc_batch *b) {
    for (int e = 0; e < b->proto.n_envs; e++) if (b->proto.envs[e]) pctc_destroy(b->proto.envs[e]);
    free(b->proto.envs); free(b->proto.obs); free(b->proto.rew_sum); free(b->proto.n_done); free(b);
}
