/* philox.h -- Philox4x32-10 and the reset-draw helper shared by the oracle's reset functions.
 * TEST INFRASTRUCTURE (see oracle.h).  Independent restatement of the published algorithm
 * (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11). */
#ifndef RLP_ORACLE_PHILOX_H
#define RLP_ORACLE_PHILOX_H
#include <stdint.h>
#include <math.h>

typedef struct {
    uint32_t ctr[4];
    uint32_t key[2];
    uint32_t buf[4];
    int have;
} orc_rng;

static inline void orc_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void orc_rng_init(orc_rng *g, uint64_t seed, uint64_t env, uint32_t episode) {
    g->ctr[0] = (uint32_t)env; g->ctr[1] = (uint32_t)(env >> 32); g->ctr[2] = episode; g->ctr[3] = 0;
    g->key[0] = (uint32_t)seed; g->key[1] = (uint32_t)(seed >> 32);
    g->have = 0;
}

static inline double orc_u01(orc_rng *g) {
    if (g->have < 2) {
        orc_philox_block(g->ctr, g->key, g->buf);
        g->ctr[3] += 1;
        g->have = 4;
    }
    uint32_t a = g->buf[4 - g->have], b = g->buf[5 - g->have];
    g->have -= 2;
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

static inline double orc_uniform(orc_rng *g, double lo, double hi) { return fma(hi - lo, orc_u01(g), lo); }
#endif
