/* Batched MADDPG actor inference: the policy stage between two ticks.
 *
 * Replaces, for every controlled vehicle of every intersection at once, the reference's
 *     agent1_action = agent.action(state=[veh["state"][0]], sess)          main.py:44, 404, 563
 * i.e. the actor of model_agent_maddpg.py:23-49:
 *     LN(28) -> Dense 64 -> LN -> ReLU -> Dense 64 -> LN -> ReLU -> Dense 1 -> 3 * tanh
 * with tf.contrib.layers.layer_norm (tensorflow 1.12, un-vendored dependency): moments over the
 * feature axis, variance = mean((x - mean)^2), epsilon 1e-12,
 *     inv = rsqrt(var + eps) * gamma;  y = x * inv + (beta - mean * inv).
 * Uncontrolled vehicles get action 0 (main.py:401).  All arithmetic is fp32 like the reference's graph.
 *
 * One warp per intersection at a time.  Each lane owns two of the 64 hidden units and keeps their
 * weight columns in registers (28 + 28 + 64 + 64 values); the activations of a layer are exchanged
 * through a 64-float shared-memory line per warp and read back as broadcast float4s, so the inner
 * loops are FFMA with one LDS.128 per eight FFMAs.  Two agents are in flight per warp for latency.
 * Device only: there is no host version of this file (the numpy restatement for tests lives in
 * oracle/actor_oracle.py).
 */
#ifndef PVE_ACTOR_CUH
#define PVE_ACTOR_CUH

#include <stdint.h>

#include "pve_mcc.h"

/* flat weight layout (floats), see pve_actor_create */
enum { PVA_LN0_G = 0, PVA_LN0_B = 28, PVA_W1 = 56, PVA_B1 = PVA_W1 + 28 * 64, PVA_LN1_G = PVA_B1 + 64,
       PVA_LN1_B = PVA_LN1_G + 64, PVA_W2 = PVA_LN1_B + 64, PVA_B2 = PVA_W2 + 64 * 64, PVA_LN2_G = PVA_B2 + 64,
       PVA_LN2_B = PVA_LN2_G + 64, PVA_W3 = PVA_LN2_B + 64, PVA_B3 = PVA_W3 + 64, PVA_COUNT = PVA_B3 + 1 };
static_assert(PVA_COUNT == PVE_ACTOR_FLOATS, "actor parameter count");

#ifdef __CUDACC__
#define PVA_WARPS 4
#define PVA_ROWS 2            /* agents in flight per warp */
#define PVA_EPS 1e-12f

__device__ __forceinline__ float pva_warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

/* rows: [n_slots][28] stored rows (the scene's row0 buffer or any dense matrix); slot s is evaluated when
 * mask says so: meta != null -> control flag of meta[s] and (s mod slots_per_env) < n_veh[env];
 * meta == null -> every slot.  actions[s] = 3 tanh(...) (+ noise_scale * noise[s]) or 0. */
__global__ void __launch_bounds__(PVA_WARPS * 32, 2)
pve_actor_kernel(const float *__restrict__ W, const float *__restrict__ rows, const pve_veh_meta *__restrict__ meta,
                 const int32_t *__restrict__ n_veh, const float *__restrict__ noise, const float noise_scale,
                 float *__restrict__ actions, const int n_env, const int slots_per_env, const long long n_slots) {
    __shared__ __align__(16) float xs[PVA_WARPS][PVA_ROWS][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * PVA_WARPS + warp, nw = gridDim.x * PVA_WARPS;

    /* this lane's weight columns */
    float w1a[28], w1b[28], w2a[64], w2b[64];
#pragma unroll
    for (int i = 0; i < 28; ++i) { w1a[i] = W[PVA_W1 + i * 64 + lane]; w1b[i] = W[PVA_W1 + i * 64 + lane + 32]; }
#pragma unroll
    for (int i = 0; i < 64; ++i) { w2a[i] = W[PVA_W2 + i * 64 + lane]; w2b[i] = W[PVA_W2 + i * 64 + lane + 32]; }
    const float g0 = lane < 28 ? W[PVA_LN0_G + lane] : 0.f, be0 = lane < 28 ? W[PVA_LN0_B + lane] : 0.f;
    const float b1a = W[PVA_B1 + lane], b1b = W[PVA_B1 + lane + 32];
    const float g1a = W[PVA_LN1_G + lane], g1b = W[PVA_LN1_G + lane + 32];
    const float be1a = W[PVA_LN1_B + lane], be1b = W[PVA_LN1_B + lane + 32];
    const float b2a = W[PVA_B2 + lane], b2b = W[PVA_B2 + lane + 32];
    const float g2a = W[PVA_LN2_G + lane], g2b = W[PVA_LN2_G + lane + 32];
    const float be2a = W[PVA_LN2_B + lane], be2b = W[PVA_LN2_B + lane + 32];
    const float w3a = W[PVA_W3 + lane], w3b = W[PVA_W3 + lane + 32], b3 = W[PVA_B3];

    for (int env = gw; env < n_env; env += nw) {
        const size_t base = (size_t)env * (size_t)slots_per_env;
        const int nv = n_veh ? min(n_veh[env], slots_per_env) : slots_per_env;
        for (int k0 = 0; k0 < slots_per_env; k0 += 32) {
            const int k = k0 + lane;
            bool want = k < nv && (long long)(base + k) < n_slots;
            if (want && meta) want = ((meta[base + k].packed >> 24) & PVE_F_CONTROL) != 0;
            if (k < slots_per_env && !want && (long long)(base + k) < n_slots) actions[base + k] = 0.f;   /* main.py:401 */
            unsigned todo = __ballot_sync(0xffffffffu, want);
            while (todo) {
                int kk[PVA_ROWS];
                float x[PVA_ROWS];
#pragma unroll
                for (int r = 0; r < PVA_ROWS; ++r) {
                    kk[r] = todo ? k0 + __ffs(todo) - 1 : -1;
                    todo &= todo - 1;
                    x[r] = (kk[r] >= 0 && lane < 28) ? rows[(base + kk[r]) * PVE_OBS_W + lane] : 0.f;
                }
                float ha[PVA_ROWS], hb[PVA_ROWS];
                /* LN(28), NET:27 */
#pragma unroll
                for (int r = 0; r < PVA_ROWS; ++r) {
                    const float mean = pva_warp_sum(x[r]) * (1.f / 28.f);
                    const float d = lane < 28 ? x[r] - mean : 0.f;
                    const float var = pva_warp_sum(d * d) * (1.f / 28.f);
                    const float inv = rsqrtf(var + PVA_EPS) * g0;
                    xs[warp][r][lane] = x[r] * inv + (be0 - mean * inv);
                }
                __syncwarp();
                /* Dense 28 -> 64, NET:28 */
#pragma unroll
                for (int r = 0; r < PVA_ROWS; ++r) { ha[r] = 0.f; hb[r] = 0.f; }
#pragma unroll
                for (int i = 0; i < 28; i += 4)
#pragma unroll
                    for (int r = 0; r < PVA_ROWS; ++r) {
                        const float4 v = *reinterpret_cast<const float4 *>(&xs[warp][r][i]);
                        ha[r] = fmaf(v.x, w1a[i], ha[r]); hb[r] = fmaf(v.x, w1b[i], hb[r]);
                        ha[r] = fmaf(v.y, w1a[i + 1], ha[r]); hb[r] = fmaf(v.y, w1b[i + 1], hb[r]);
                        ha[r] = fmaf(v.z, w1a[i + 2], ha[r]); hb[r] = fmaf(v.z, w1b[i + 2], hb[r]);
                        ha[r] = fmaf(v.w, w1a[i + 3], ha[r]); hb[r] = fmaf(v.w, w1b[i + 3], hb[r]);
                    }
                __syncwarp();
                /* LN(64) + ReLU, NET:30-32 */
#pragma unroll
                for (int r = 0; r < PVA_ROWS; ++r) {
                    const float a = ha[r] + b1a, b = hb[r] + b1b;
                    const float mean = pva_warp_sum(a + b) * (1.f / 64.f);
                    const float da = a - mean, db = b - mean;
                    const float var = pva_warp_sum(da * da + db * db) * (1.f / 64.f);
                    const float rs = rsqrtf(var + PVA_EPS);
                    const float ia = rs * g1a, ib = rs * g1b;
                    xs[warp][r][lane] = fmaxf(a * ia + (be1a - mean * ia), 0.f);
                    xs[warp][r][lane + 32] = fmaxf(b * ib + (be1b - mean * ib), 0.f);
                }
                __syncwarp();
                /* Dense 64 -> 64, NET:34 */
#pragma unroll
                for (int r = 0; r < PVA_ROWS; ++r) { ha[r] = 0.f; hb[r] = 0.f; }
#pragma unroll
                for (int i = 0; i < 64; i += 4)
#pragma unroll
                    for (int r = 0; r < PVA_ROWS; ++r) {
                        const float4 v = *reinterpret_cast<const float4 *>(&xs[warp][r][i]);
                        ha[r] = fmaf(v.x, w2a[i], ha[r]); hb[r] = fmaf(v.x, w2b[i], hb[r]);
                        ha[r] = fmaf(v.y, w2a[i + 1], ha[r]); hb[r] = fmaf(v.y, w2b[i + 1], hb[r]);
                        ha[r] = fmaf(v.z, w2a[i + 2], ha[r]); hb[r] = fmaf(v.z, w2b[i + 2], hb[r]);
                        ha[r] = fmaf(v.w, w2a[i + 3], ha[r]); hb[r] = fmaf(v.w, w2b[i + 3], hb[r]);
                    }
                __syncwarp();
                /* LN(64) + ReLU, Dense 64 -> 1, 3 tanh: NET:36-47 */
#pragma unroll
                for (int r = 0; r < PVA_ROWS; ++r) {
                    const float a = ha[r] + b2a, b = hb[r] + b2b;
                    const float mean = pva_warp_sum(a + b) * (1.f / 64.f);
                    const float da = a - mean, db = b - mean;
                    const float var = pva_warp_sum(da * da + db * db) * (1.f / 64.f);
                    const float rs = rsqrtf(var + PVA_EPS);
                    const float ia = rs * g2a, ib = rs * g2b;
                    const float ra = fmaxf(a * ia + (be2a - mean * ia), 0.f);
                    const float rb = fmaxf(b * ib + (be2b - mean * ib), 0.f);
                    const float o = pva_warp_sum(fmaf(ra, w3a, rb * w3b)) + b3;
                    if (lane == 0 && kk[r] >= 0) {
                        float act = 3.f * tanhf(o);
                        if (noise) act += noise_scale * noise[base + kk[r]];     /* main.py:44 */
                        actions[base + kk[r]] = act;
                    }
                }
            }
        }
    }
}
#endif  /* __CUDACC__ */
#endif
