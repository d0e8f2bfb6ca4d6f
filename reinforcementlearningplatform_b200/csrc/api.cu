// api.cu -- extern "C" surface of libb200env.so (see include/b200env.h).
// Dispatches on env id to the family launchers; no torch types, no state.
#include "common.cuh"

thread_local int g_b200_last_cuda_error = 0;

namespace {
struct Family {
    size_t params_bytes;
    int (*dims)(int, int *, int *, int *, int *);
    int (*step)(int, int64_t, const void *, const b200env_io *, uint32_t, uint64_t, int64_t, cudaStream_t);
    int (*reset)(int, int64_t, const void *, const b200env_io *, const uint8_t *, uint64_t, int64_t, cudaStream_t);
    int (*observe)(int, int64_t, const void *, const b200env_io *, cudaStream_t);
    // fused multi-step kernel, or NULL: b200env_rollout then launches `step` once per time step
    int (*rollout)(int, int64_t, const void *, const b200env_io *, const b200env_rollout_spec *, uint32_t, uint64_t,
                   int64_t, cudaStream_t);
};

#define FAM(name, P) Family{sizeof(P), name##_dims, name##_step, name##_reset, name##_observe, nullptr}
#define FAMR(name, P) Family{sizeof(P), name##_dims, name##_step, name##_reset, name##_observe, name##_rollout}
const Family *family(int env_id) {
    static Family table[B200ENV_COUNT] = {};
    static bool init = false;
    if (!init) {
        table[B200ENV_CARTPOLE] = FAMR(cartpole, b200_cartpole_params);
        table[B200ENV_UAV_ATT] = FAM(uav_att, b200_uav_params);
        table[B200ENV_UAV_POS] = FAM(uav_pos, b200_uav_params);
        table[B200ENV_FAS] = FAMR(fas, b200_fas_params);
        table[B200ENV_SOI] = FAMR(soi, b200_soi_params);
        table[B200ENV_BALLBALANCER] = FAMR(ballbalancer, b200_ballbalancer_params);
        table[B200ENV_TWOLINK] = FAMR(twolink, b200_twolink_params);
        table[B200ENV_UGV] = FAMR(ugv, b200_ugv_params);
        table[B200ENV_UGVO] = FAM(ugvo, b200_ugvo_params);
        table[B200ENV_UAVROBUST] = FAM(uavrobust, b200_uavrobust_params);
        table[B200ENV_FAS_DISCRETE] = FAMR(fas_discrete, b200_fas_discrete_params);
        init = true;
    }
    if (env_id < 0 || env_id >= B200ENV_COUNT) return nullptr;
    const Family *f = &table[env_id];
    return f->step ? f : nullptr;
}

int check_common(const Family *f, int dtype, int64_t n, const void *params, size_t bytes, const b200env_io *io) {
    if (!f) return B200ENV_EENV;
    if (dtype != B200ENV_F64 && dtype != B200ENV_F32) return B200ENV_EDTYPE;
    if (n <= 0 || n >= ((int64_t)1 << 31)) return B200ENV_ESIZE; // 32-bit instance index in the SoA accessors (common.cuh)
    if (!params || !io) return B200ENV_ENULL;
    if (bytes != f->params_bytes) return B200ENV_EPARAMS;
    return B200ENV_OK;
}
} // namespace

extern "C" {

const char *b200env_version(void) { return "b200env 0.1 (sm_100a)"; }

int b200env_last_cuda_error(void) { return g_b200_last_cuda_error; }

size_t b200env_params_bytes(int env_id) {
    const Family *f = family(env_id);
    return f ? f->params_bytes : 0;
}

int b200env_dims(int env_id, int variant, int *state_fields, int *obs_dim, int *action_dim, int *dis_dim) {
    const Family *f = family(env_id);
    if (!f) return B200ENV_EENV;
    return f->dims(variant, state_fields, obs_dim, action_dim, dis_dim);
}

int b200env_step(int env_id, int dtype, int64_t n_envs, const void *params, size_t params_bytes,
                 const b200env_io *io, uint32_t flags, uint64_t seed, int64_t env_index_offset, void *cuda_stream) {
    const Family *f = family(env_id);
    int rc = check_common(f, dtype, n_envs, params, params_bytes, io);
    if (rc) return rc;
    return f->step(dtype, n_envs, params, io, flags, seed, env_index_offset, (cudaStream_t)cuda_stream);
}

int b200env_state_layout(int env_id, int variant, int *block, int *slots) {
    const Family *f = family(env_id);
    if (!f) return B200ENV_EENV;
    int sf = 0;
    const int rc = f->dims(variant, &sf, nullptr, nullptr, nullptr);
    if (rc) return rc;
    const bool blocked = env_id == B200ENV_UAV_ATT || env_id == B200ENV_UAV_POS || env_id == B200ENV_UAVROBUST;
    if (block) *block = blocked ? B200_UAV_STATE_BLOCK : 0;
    if (slots) *slots = blocked ? B200_UAV_STATE_SLOTS : sf;
    return B200ENV_OK;
}

size_t b200env_state_elems(int env_id, int variant, int64_t n_envs) {
    int block = 0, slots = 0;
    if (n_envs <= 0 || b200env_state_layout(env_id, variant, &block, &slots)) return 0;
    if (!block) return (size_t)slots * (size_t)n_envs;
    return (size_t)((n_envs + block - 1) / block) * (size_t)block * (size_t)slots;
}

int b200env_rollout(int env_id, int dtype, int64_t n_envs, const void *params, size_t params_bytes,
                    const b200env_io *io, const b200env_rollout_spec *spec, uint32_t flags, uint64_t seed,
                    int64_t env_index_offset, void *cuda_stream) {
    const Family *f = family(env_id);
    int rc = check_common(f, dtype, n_envs, params, params_bytes, io);
    if (rc) return rc;
    if (!spec) return B200ENV_ENULL;
    if (spec->steps <= 0) return B200ENV_ESIZE;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (f->rollout) return f->rollout(dtype, n_envs, params, io, spec, flags, seed, env_index_offset, s);
    // no fused kernel for this family: one step launch per time step, rows addressed by the strides
    const size_t eio = (dtype == B200ENV_F32 || io->io_dtype == B200ENV_F32) ? 4 : 8;
    for (int64_t t = 0; t < spec->steps; ++t) {
        b200env_io it = *io;
        auto adv = [&](const void *ptr, int64_t stride, size_t es) -> void * {
            return ptr ? (void *)((const char *)ptr + (size_t)t * (size_t)stride * es) : nullptr;
        };
        it.action = adv(io->action, spec->action_stride, eio);
        it.dis = adv(io->dis, spec->dis_stride, eio);
        it.obs = adv(io->obs, spec->obs_stride, eio);
        it.next_obs = adv(io->next_obs, spec->next_obs_stride, eio);
        it.reward = adv(io->reward, spec->reward_stride, eio);
        it.done = (uint8_t *)adv(io->done, spec->done_stride, 1);
        it.flag = (int32_t *)adv(io->flag, spec->flag_stride, 4);
        rc = f->step(dtype, n_envs, params, &it, flags, seed, env_index_offset, s);
        if (rc) return rc;
    }
    return B200ENV_OK;
}

int b200env_reset(int env_id, int dtype, int64_t n_envs, const void *params, size_t params_bytes,
                  const b200env_io *io, const uint8_t *mask, uint64_t seed, int64_t env_index_offset,
                  void *cuda_stream) {
    const Family *f = family(env_id);
    int rc = check_common(f, dtype, n_envs, params, params_bytes, io);
    if (rc) return rc;
    return f->reset(dtype, n_envs, params, io, mask, seed, env_index_offset, (cudaStream_t)cuda_stream);
}

int b200env_observe(int env_id, int dtype, int64_t n_envs, const void *params, size_t params_bytes,
                    const b200env_io *io, void *cuda_stream) {
    const Family *f = family(env_id);
    int rc = check_common(f, dtype, n_envs, params, params_bytes, io);
    if (rc) return rc;
    return f->observe(dtype, n_envs, params, io, (cudaStream_t)cuda_stream);
}

} // extern "C"
