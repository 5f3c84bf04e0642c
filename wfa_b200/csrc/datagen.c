/*
 * datagen.c -- deterministic synthetic pair generator.
 *
 * Shared by the tests, bench.py and the CPU baseline so that every arm sees
 * the same bytes.  PRNG = splitmix64; pair i of a config draws from the stream
 * seeded with splitmix64(base_seed ^ i), base_seed = 0x57464100 + config.
 * Mirrors WFA2-lib's generate_dataset shape (reference README.md:298-306):
 * query = L uniform ACGT bases, target = query with round(err*L) edits, each
 * uniformly a mismatch, a 1-base insertion or a 1-base deletion at a uniform
 * position.  Window mode (config 4): target = W uniform bases, query = the
 * L-base substring at a uniform start in [0, max_start] with the same edits.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};

typedef struct { uint32_t pos; uint8_t kind; uint8_t base; uint32_t ord; } edit_t;

static int edit_cmp(const void *a, const void *b)
{
    const edit_t *x = (const edit_t *)a, *y = (const edit_t *)b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return x->ord < y->ord ? -1 : x->ord > y->ord;
}

/* writes src with edits applied into dst, returns the new length (<= L + nedits) */
static uint32_t apply_edits(uint64_t *rng, const uint8_t *src, uint32_t L, uint32_t nedits, uint8_t *dst, edit_t *ed)
{
    for (uint32_t i = 0; i < nedits; i++) {
        uint64_t r = splitmix64(rng);
        ed[i].kind = (uint8_t)(r % 3);
        ed[i].pos = (uint32_t)((r >> 8) % L);
        ed[i].base = (uint8_t)((r >> 48) & 3);
        ed[i].ord = i;
    }
    qsort(ed, nedits, sizeof(edit_t), edit_cmp);
    uint32_t out = 0, e = 0;
    for (uint32_t p = 0; p < L; p++) {
        int consumed = 0;
        while (e < nedits && ed[e].pos == p) {
            if (ed[e].kind == 1) dst[out++] = (uint8_t)BASES[ed[e].base];          /* insertion before p */
            else if (!consumed) {
                if (ed[e].kind == 0) {                                              /* mismatch */
                    uint8_t b = (uint8_t)BASES[ed[e].base];
                    if (b == src[p]) b = (uint8_t)BASES[(ed[e].base + 1) & 3];
                    dst[out++] = b;
                }                                                                   /* kind 2: deletion */
                consumed = 1;
            }
            e++;
        }
        if (!consumed) dst[out++] = src[p];
    }
    if (out == 0) dst[out++] = src[0];
    return out;
}

typedef struct {
    uint64_t base_seed, first, n; uint32_t L, nedits, window, max_start;
    uint8_t *out; uint64_t stride; uint64_t *q_off, *t_off; uint32_t *q_len, *t_len;
    uint64_t *next;
} gen_job;

static void *gen_worker(void *arg)
{
    gen_job *j = (gen_job *)arg;
    edit_t *ed = (edit_t *)malloc(sizeof(edit_t) * (j->nedits ? j->nedits : 1));
    uint8_t *tmp = (uint8_t *)malloc(j->L + 16);
    for (;;) {
        uint64_t i0 = __atomic_fetch_add(j->next, 256, __ATOMIC_RELAXED);
        if (i0 >= j->n) break;
        uint64_t i1 = i0 + 256 < j->n ? i0 + 256 : j->n;
        for (uint64_t i = i0; i < i1; i++) {
            uint64_t seed = j->base_seed ^ (j->first + i);
            uint64_t rng = splitmix64(&seed);
            uint8_t *qdst = j->out + i * j->stride;
            if (j->window == 0) {
                /* query = L uniform bases; target = edited query */
                for (uint32_t p = 0; p < j->L; p += 32) {
                    uint64_t r = splitmix64(&rng);
                    for (uint32_t b = 0; b < 32 && p + b < j->L; b++, r >>= 2) qdst[p + b] = (uint8_t)BASES[r & 3];
                }
                uint8_t *tdst = qdst + ((j->L + 15) & ~15u);
                uint32_t tl = apply_edits(&rng, qdst, j->L, j->nedits, tdst, ed);
                j->q_off[i] = (uint64_t)(qdst - j->out); j->q_len[i] = j->L;
                j->t_off[i] = (uint64_t)(tdst - j->out); j->t_len[i] = tl;
            } else {
                /* target = window of uniform bases; query = edited substring */
                uint8_t *tdst = qdst + ((j->L + j->nedits + 15) & ~15u);
                for (uint32_t p = 0; p < j->window; p += 32) {
                    uint64_t r = splitmix64(&rng);
                    for (uint32_t b = 0; b < 32 && p + b < j->window; b++, r >>= 2) tdst[p + b] = (uint8_t)BASES[r & 3];
                }
                uint32_t start = (uint32_t)(splitmix64(&rng) % ((uint64_t)j->max_start + 1));
                if (start + j->L > j->window) start = j->window - j->L;
                memcpy(tmp, tdst + start, j->L);
                uint32_t ql = apply_edits(&rng, tmp, j->L, j->nedits, qdst, ed);
                j->q_off[i] = (uint64_t)(qdst - j->out); j->q_len[i] = ql;
                j->t_off[i] = (uint64_t)(tdst - j->out); j->t_len[i] = j->window;
            }
        }
    }
    free(ed); free(tmp);
    return NULL;
}

/* bytes needed per pair */
uint64_t wfagen_stride(uint32_t L, uint32_t nedits, uint32_t window)
{
    uint64_t a = ((uint64_t)L + 15) & ~15ull;
    if (window == 0) return a + (((uint64_t)L + nedits + 15) & ~15ull);
    return (((uint64_t)L + nedits + 15) & ~15ull) + (((uint64_t)window + 15) & ~15ull);
}

/* Generates pairs [first, first+n) of a config into out (n * stride bytes). */
void wfagen_pairs(uint64_t base_seed, uint64_t first, uint64_t n, uint32_t L, uint32_t nedits,
                  uint32_t window, uint32_t max_start, uint8_t *out,
                  uint64_t *q_off, uint32_t *q_len, uint64_t *t_off, uint32_t *t_len, int nthreads)
{
    uint64_t next = 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    gen_job job = {base_seed, first, n, L, nedits, window, max_start, out, wfagen_stride(L, nedits, window),
                   q_off, t_off, q_len, t_len, &next};
    pthread_t th[64];
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, gen_worker, &job);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
}
