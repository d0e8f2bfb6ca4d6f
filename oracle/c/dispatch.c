/* dispatch.c -- batch drivers of the oracle (OpenMP over instances).  TEST INFRASTRUCTURE (see oracle.h). */
#include "oracle.h"
#include "philox.h"
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*step_one_fn)(const void *, const oracle_io *, int64_t, int64_t, uint32_t, uint64_t, int64_t);
typedef void (*reset_one_fn)(const void *, const oracle_io *, int64_t, int64_t, uint64_t, int64_t, int);

#define DECL(name)                                                                                          \
    void orc_##name##_step_one(const void *, const oracle_io *, int64_t, int64_t, uint32_t, uint64_t, int64_t); \
    void orc_##name##_reset_one(const void *, const oracle_io *, int64_t, int64_t, uint64_t, int64_t, int);
DECL(cartpole)
DECL(uav_att)
DECL(uav_pos)
DECL(fas)
DECL(soi)
DECL(ballbalancer)
DECL(twolink)
DECL(ugv)
DECL(ugvo)
DECL(uavrobust)
DECL(fas_discrete)

static step_one_fn step_of(int env_id) {
    switch (env_id) {
    case B200ENV_CARTPOLE: return orc_cartpole_step_one;
    case B200ENV_UAV_ATT: return orc_uav_att_step_one;
    case B200ENV_UAV_POS: return orc_uav_pos_step_one;
    case B200ENV_FAS: return orc_fas_step_one;
    case B200ENV_SOI: return orc_soi_step_one;
    case B200ENV_BALLBALANCER: return orc_ballbalancer_step_one;
    case B200ENV_TWOLINK: return orc_twolink_step_one;
    case B200ENV_UGV: return orc_ugv_step_one;
    case B200ENV_UGVO: return orc_ugvo_step_one;
    case B200ENV_UAVROBUST: return orc_uavrobust_step_one;
    case B200ENV_FAS_DISCRETE: return orc_fas_discrete_step_one;
    default: return 0;
    }
}
static reset_one_fn reset_of(int env_id) {
    switch (env_id) {
    case B200ENV_CARTPOLE: return orc_cartpole_reset_one;
    case B200ENV_UAV_ATT: return orc_uav_att_reset_one;
    case B200ENV_UAV_POS: return orc_uav_pos_reset_one;
    case B200ENV_FAS: return orc_fas_reset_one;
    case B200ENV_SOI: return orc_soi_reset_one;
    case B200ENV_BALLBALANCER: return orc_ballbalancer_reset_one;
    case B200ENV_TWOLINK: return orc_twolink_reset_one;
    case B200ENV_UGV: return orc_ugv_reset_one;
    case B200ENV_UGVO: return orc_ugvo_reset_one;
    case B200ENV_UAVROBUST: return orc_uavrobust_reset_one;
    case B200ENV_FAS_DISCRETE: return orc_fas_discrete_reset_one;
    default: return 0;
    }
}

int oracle_step(int env_id, int64_t n, const void *params, const oracle_io *io, uint32_t flags, uint64_t seed,
                int64_t off, int nthreads) {
    step_one_fn f = step_of(env_id);
    if (!f) return B200ENV_EENV;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t i = 0; i < n; ++i) f(params, io, n, i, flags, seed, off);
    return 0;
}

int oracle_reset(int env_id, int64_t n, const void *params, const oracle_io *io, const uint8_t *mask, uint64_t seed,
                 int64_t off) {
    reset_one_fn f = reset_of(env_id);
    if (!f) return B200ENV_EENV;
    for (int64_t i = 0; i < n; ++i)
        if (!mask || mask[i]) f(params, io, n, i, seed, off, 0);
    return 0;
}

int oracle_observe(int env_id, int64_t n, const void *params, const oracle_io *io) {
    reset_one_fn f = reset_of(env_id);
    if (!f) return B200ENV_EENV;
    for (int64_t i = 0; i < n; ++i) f(params, io, n, i, 0, 0, 1);
    return 0;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    orc_philox_block(ctr, key, out);
}
