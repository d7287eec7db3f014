/* TEST INFRASTRUCTURE — multi-threaded batch driver over the single-env oracle (discrete domain).
 * Plays the role of the reference's ShmemVecEnv (wrapper/shmem_vec_env.py:20-156: N independent workers,
 * auto-reset on done) on the host cores, with the synthetic policy of SURVEY.md §8(d).  Used by
 *   - tests: full-size parity (final observations of thousands of envs after K steps must equal the GPU's),
 *   - bench.py: cpu_baseline and `--impl reference` timing.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct pcto_env pcto_env;
pcto_env *pcto_create(int setting, int W, int L, int H, int nb_holder, int nl_holder, double low_bound);
void pcto_destroy(pcto_env *e);
void pcto_set_random_items(pcto_env *e, const double *item_set3, int n, uint64_t seed, uint64_t gid);
int pcto_obs_len(pcto_env *e);
void pcto_reset(pcto_env *e, double *obs);
int pcto_step(pcto_env *e, const double *action, int action_len, double *obs, double *reward, int *done, double *info);
uint64_t pcto_rnd_u64(uint64_t seed, uint64_t a, uint64_t b);

typedef struct {
    int setting, W, L, H, nb, nl;
    double low_bound;
    const double *item_set; int n_items;
    uint64_t item_seed, policy_seed; int64_t gid_base;
    int n_envs, steps, first, last;
    double *final_obs;   /* n_envs x obs_len (may be NULL) */
    double *rew_sum;     /* n_envs */
    int *n_done;         /* n_envs */
    pcto_env **envs;     /* persistent envs (created by thread) */
    double *obs;         /* n_envs x obs_len working observations */
    int phase;           /* 0: create+reset, 1: run steps */
    int64_t t0;
} job_t;

static void *worker(void *arg) {
    job_t *j = arg;
    const int ol = (j->nb + j->nl + 1) * 9;
    double act[9], info[3];
    for (int e = j->first; e < j->last; e++) {
        if (j->phase == 0) {
            j->envs[e] = pcto_create(j->setting, j->W, j->L, j->H, j->nb, j->nl, j->low_bound);
            pcto_set_random_items(j->envs[e], j->item_set, j->n_items, j->item_seed, (uint64_t)(j->gid_base + e));
            pcto_reset(j->envs[e], j->obs + (size_t)e * ol);
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
            pcto_step(j->envs[e], act, 9, obs, &r, &d, info);
            j->rew_sum[e] += r;
            if (d) { j->n_done[e]++; pcto_reset(j->envs[e], obs); }
        }
    }
    return NULL;
}

typedef struct { job_t proto; int threads; int64_t t; } pcto_batch;

static void run_phase(pcto_batch *b, int phase, int steps) {
    int T = b->threads < 1 ? 1 : b->threads;
    if (T > b->proto.n_envs) T = b->proto.n_envs;
    pthread_t *th = malloc(sizeof(pthread_t) * T);
    job_t *jobs = malloc(sizeof(job_t) * T);
    for (int i = 0; i < T; i++) {
        jobs[i] = b->proto;
        jobs[i].phase = phase; jobs[i].steps = steps; jobs[i].t0 = b->t;
        jobs[i].first = (int)((int64_t)b->proto.n_envs * i / T);
        jobs[i].last = (int)((int64_t)b->proto.n_envs * (i + 1) / T);
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    for (int i = 0; i < T; i++) pthread_join(th[i], NULL);
    free(th); free(jobs);
}

pcto_batch *pcto_batch_create(int setting, int W, int L, int H, int nb, int nl, double low_bound, const double *item_set3, int n_items,
                              uint64_t item_seed, uint64_t policy_seed, int64_t gid_base, int n_envs, int threads) {
    pcto_batch *b = calloc(1, sizeof *b);
    job_t *p = &b->proto;
    p->setting = setting; p->W = W; p->L = L; p->H = H; p->nb = nb; p->nl = nl; p->low_bound = low_bound;
    double *is = malloc(sizeof(double) * 3 * n_items);
    memcpy(is, item_set3, sizeof(double) * 3 * n_items);
    p->item_set = is; p->n_items = n_items; p->item_seed = item_seed; p->policy_seed = policy_seed; p->gid_base = gid_base;
    p->n_envs = n_envs;
    const int ol = (nb + nl + 1) * 9;
    p->envs = calloc(n_envs, sizeof(pcto_env *));
    p->obs = malloc(sizeof(double) * (size_t)n_envs * ol);
    p->rew_sum = calloc(n_envs, sizeof(double));
    p->n_done = calloc(n_envs, sizeof(int));
    b->threads = threads;
    run_phase(b, 0, 0);
    return b;
}
/* runs `steps` vector steps; returns elapsed seconds */
double pcto_batch_run(pcto_batch *b, int steps) {
    struct timespec a, c;
    clock_gettime(CLOCK_MONOTONIC, &a);
    run_phase(b, 1, steps);
    clock_gettime(CLOCK_MONOTONIC, &c);
    b->t += steps;
    return (c.tv_sec - a.tv_sec) + 1e-9 * (c.tv_nsec - a.tv_nsec);
}
void pcto_batch_get(pcto_batch *b, double *obs, double *rew_sum, int *n_done) {
    const int ol = (b->proto.nb + b->proto.nl + 1) * 9;
    if (obs) memcpy(obs, b->proto.obs, sizeof(double) * (size_t)b->proto.n_envs * ol);
    if (rew_sum) memcpy(rew_sum, b->proto.rew_sum, sizeof(double) * b->proto.n_envs);
    if (n_done) memcpy(n_done, b->proto.n_done, sizeof(int) * b->proto.n_envs);
}
void pcto_batch_destroy(pcto_batch *b) {
    for (int e = 0; e < b->proto.n_envs; e++) if (b->proto.envs[e]) pcto_destroy(b->proto.envs[e]);
    free((void *)b->proto.item_set); free(b->proto.envs); free(b->proto.obs); free(b->proto.rew_sum); free(b->proto.n_done); free(b);
}
