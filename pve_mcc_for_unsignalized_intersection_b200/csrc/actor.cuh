/* Batched MADDPG actor inference: the policy stage between two ticks.
 *
 * Replaces, for every controlled vehicle of every intersection at once, the reference's
 *     agent1_action = agent.action(state=[veh["state"][0]], sess)          main.py:44, 404, 563
 * i.e. the actor of model_agent_maddpg.py:23-49:
 *     LN(28) -> Dense 64 -> LN -> ReLU -> Dense 64 -> LN -> ReLU -> Dense 1 -> 3 * tanh
 * with tf.contrib.layers.layer_norm (tensorflow 1.12, un-vendored dependency): moments over the
 * feature axis, variance = mean((x - mean)^2), epsilon 1e-12,
 *     inv = rsqrt(var + eps) * gamma;  y = x * inv + (beta - mean * inv).
 * Uncontrolled vehicles get action 0 (main.py:401).  All arithmetic is fp32 like the reference's graph; the
 * network is ill-conditioned at the 1e-4 level (rounding only its inputs to fp32 moves some actions by
 * 9e-5), so reduced-precision tensor-core formats (tf32 / bf16) are not an option for parity.
 *
 * A CTA walks over its share of the intersections, appends the controlled vehicle slots to a ring in shared
 * memory (warp ballots) and, whenever 128 are queued, evaluates the network for them as a register-tiled
 * GEMM chain: thread t owns rows {16 r + t / 8} and eight hidden units (two groups of four), i.e. an 8 x 8
 * accumulator tile; activations (row-major, stride 68 floats) and weights (25 KB, copied once per CTA) are
 * read from shared memory as float4s, 16 LDS.128 per 256 FFMAs.  The first LayerNorm is computed by the
 * thread that loads the row, the other two inside the tile (the eight threads of a row are adjacent lanes:
 * three shuffle steps).  fp32 FFMA on purpose: see the conditioning note above; tensor cores are not used.
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
#define PVA_THREADS 128
#define PVA_TILE 128          /* vehicles per evaluation round */
#define PVA_AS 68             /* activation row stride in floats: rows 16 apart fall into different banks */
#define PVA_RING 1024         /* queued vehicle slots (>= PVA_TILE - 1 + one batch of intersections) */
#define PVA_BATCH 4           /* intersections per ticket (fewer for the large capacity classes) */
#define PVA_EPS 1e-12f
#define PVA_WPAD ((PVA_COUNT + 3) & ~3)
#define PVA_SMEM_BYTES ((PVA_WPAD + PVA_TILE * PVA_AS) * 4 + PVA_RING * 4)

/* The eight hidden units of thread column-group cg are {4 cg .. 4 cg + 3} and {32 + 4 cg .. 32 + 4 cg + 3}: the
 * eight groups of a warp then read 128 contiguous bytes of a weight row per LDS.128 (no bank conflicts).
 * acc[r][c] = bias[u(c)] + sum_k a[16 r + rg][k] * Wm[k][u(c)] */
template <int K, int R, int AS = PVA_AS>
__device__ __forceinline__ void pva_gemm_tile(const float *__restrict__ a, const float *__restrict__ Wm,
                                              const float *__restrict__ bias, float (&acc)[R][8], int rg, int cg) {
    {
        const float4 b0 = *reinterpret_cast<const float4 *>(bias + cg * 4), b1 = *reinterpret_cast<const float4 *>(bias + 32 + cg * 4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            acc[r][0] = b0.x; acc[r][1] = b0.y; acc[r][2] = b0.z; acc[r][3] = b0.w;
            acc[r][4] = b1.x; acc[r][5] = b1.y; acc[r][6] = b1.z; acc[r][7] = b1.w;
        }
    }
    const float *arow = a + rg * AS;
    const float *wcol = Wm + cg * 4;
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 4) {
        float4 av[R];
#pragma unroll
        for (int r = 0; r < R; ++r) av[r] = *reinterpret_cast<const float4 *>(arow + r * 16 * AS + k0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 b0 = *reinterpret_cast<const float4 *>(wcol + (k0 + kk) * 64);
            const float4 b1 = *reinterpret_cast<const float4 *>(wcol + (k0 + kk) * 64 + 32);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float v = kk == 0 ? av[r].x : kk == 1 ? av[r].y : kk == 2 ? av[r].z : av[r].w;
                acc[r][0] = fmaf(v, b0.x, acc[r][0]); acc[r][1] = fmaf(v, b0.y, acc[r][1]);
                acc[r][2] = fmaf(v, b0.z, acc[r][2]); acc[r][3] = fmaf(v, b0.w, acc[r][3]);
                acc[r][4] = fmaf(v, b1.x, acc[r][4]); acc[r][5] = fmaf(v, b1.y, acc[r][5]);
                acc[r][6] = fmaf(v, b1.z, acc[r][6]); acc[r][7] = fmaf(v, b1.w, acc[r][7]);
            }
        }
    }
}

/* LayerNorm over the 64 units of every row of the tile (8 per thread, 8 adjacent lanes per row) + ReLU */
template <int R>
__device__ __forceinline__ void pva_tile_ln_relu(float (&acc)[R][8], const float *__restrict__ gamma,
                                                 const float *__restrict__ beta, int cg) {
    float g[8], be[8];
    {
        const float4 g0 = *reinterpret_cast<const float4 *>(gamma + cg * 4), g1 = *reinterpret_cast<const float4 *>(gamma + 32 + cg * 4);
        const float4 e0 = *reinterpret_cast<const float4 *>(beta + cg * 4), e1 = *reinterpret_cast<const float4 *>(beta + 32 + cg * 4);
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        be[0] = e0.x; be[1] = e0.y; be[2] = e0.z; be[3] = e0.w; be[4] = e1.x; be[5] = e1.y; be[6] = e1.z; be[7] = e1.w;
    }
    float mean[R], rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = ((acc[r][0] + acc[r][1]) + (acc[r][2] + acc[r][3])) + ((acc[r][4] + acc[r][5]) + (acc[r][6] + acc[r][7]));
#pragma unroll
    for (int d = 1; d < 8; d <<= 1)
#pragma unroll
        for (int r = 0; r < R; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], d);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        mean[r] *= (1.f / 64.f);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float dlt = acc[r][c] - mean[r]; q = fmaf(dlt, dlt, q); }
        rs[r] = q;
    }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1)
#pragma unroll
        for (int r = 0; r < R; ++r) rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], d);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float inv0 = rsqrtf(rs[r] * (1.f / 64.f) + PVA_EPS);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float inv = inv0 * g[c];                                    /* x * inv + (beta - mean * inv) */
            acc[r][c] = fmaxf(fmaf(acc[r][c], inv, fmaf(-mean[r], inv, be[c])), 0.f);
        }
    }
}

/* one evaluation round: the n_valid (<= 16 R) queued vehicles at ring[head ...] -> their actions */
template <int R>
__device__ __forceinline__ void pva_round(const float *__restrict__ w, float *__restrict__ a, const int *__restrict__ ring,
                                          const int head, const int n_valid, const float *__restrict__ rows,
                                          const float *__restrict__ noise, const float noise_scale,
                                          float *__restrict__ actions) {
    const int tid = threadIdx.x, cg = tid & 7, rg = tid >> 3;
    /* first LayerNorm by the thread that loads the row (NET:27) */
    if (tid < 16 * R) {
        float x[28];
        const bool valid = tid < n_valid;
        const long long gs = valid ? (long long)ring[(head + tid) & (PVA_RING - 1)] : 0;
        const float4 *src = reinterpret_cast<const float4 *>(rows + gs * PVE_OBS_W);
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const float4 v = valid ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 28; i += 4) { s0 += x[i]; s1 += x[i + 1]; s2 += x[i + 2]; s3 += x[i + 3]; }
        const float mean = ((s0 + s1) + (s2 + s3)) * (1.f / 28.f);
        s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 28; i += 4) {
            const float d0 = x[i] - mean, d1 = x[i + 1] - mean, d2 = x[i + 2] - mean, d3 = x[i + 3] - mean;
            s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
        }
        const float rs = rsqrtf(((s0 + s1) + (s2 + s3)) * (1.f / 28.f) + PVA_EPS);
#pragma unroll
        for (int i = 0; i < 28; i += 4) {
            const float4 g = *reinterpret_cast<const float4 *>(w + PVA_LN0_G + i);
            const float4 b = *reinterpret_cast<const float4 *>(w + PVA_LN0_B + i);
            const float i0 = rs * g.x, i1 = rs * g.y, i2 = rs * g.z, i3 = rs * g.w;
            float4 y;
            y.x = fmaf(x[i], i0, fmaf(-mean, i0, b.x)); y.y = fmaf(x[i + 1], i1, fmaf(-mean, i1, b.y));
            y.z = fmaf(x[i + 2], i2, fmaf(-mean, i2, b.z)); y.w = fmaf(x[i + 3], i3, fmaf(-mean, i3, b.w));
            *reinterpret_cast<float4 *>(a + tid * PVA_AS + i) = y;
        }
    }
    __syncthreads();
    float acc[R][8];
    pva_gemm_tile<28, R>(a, w + PVA_W1, w + PVA_B1, acc, rg, cg);                  /* NET:28 */
    pva_tile_ln_relu<R>(acc, w + PVA_LN1_G, w + PVA_LN1_B, cg);                    /* NET:30-32 */
    __syncthreads();                                                               /* every thread has read its inputs */
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float *dst = a + (r * 16 + rg) * PVA_AS + cg * 4;
        *reinterpret_cast<float4 *>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4 *>(dst + 32) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
    __syncthreads();
    pva_gemm_tile<64, R>(a, w + PVA_W2, w + PVA_B2, acc, rg, cg);                  /* NET:34 */
    pva_tile_ln_relu<R>(acc, w + PVA_LN2_G, w + PVA_LN2_B, cg);                    /* NET:36-38 */
    {   /* Dense 64 -> 1, 3 tanh (NET:40-47) */
        const float4 u0 = *reinterpret_cast<const float4 *>(w + PVA_W3 + cg * 4);
        const float4 u1 = *reinterpret_cast<const float4 *>(w + PVA_W3 + 32 + cg * 4);
        float o[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
            o[r] = fmaf(acc[r][0], u0.x, fmaf(acc[r][1], u0.y, fmaf(acc[r][2], u0.z, acc[r][3] * u0.w)))
                   + fmaf(acc[r][4], u1.x, fmaf(acc[r][5], u1.y, fmaf(acc[r][6], u1.z, acc[r][7] * u1.w)));
#pragma unroll
        for (int d = 1; d < 8; d <<= 1)
#pragma unroll
            for (int r = 0; r < R; ++r) o[r] += __shfl_xor_sync(0xffffffffu, o[r], d);
        /* lane cg of the row's eight writes row r = cg */
        const int row = cg * 16 + rg;
        float mine = o[0];
#pragma unroll
        for (int r = 1; r < R; ++r) mine = cg == r ? o[r] : mine;
        if (cg < R && row < n_valid) {
            const long long gs = (long long)ring[(head + row) & (PVA_RING - 1)];
            float act = 3.f * tanhf(mine + w[PVA_B3]);
            if (noise) act += noise_scale * noise[gs];                             /* main.py:44 */
            actions[gs] = act;
        }
    }
    __syncthreads();                             /* the tile and the ring entries are free again */
}

/* rows: [n_slots][28] stored rows (the scene's row0 buffer or any dense matrix), n_slots = n_env * slots_per_env
 * (the last "intersection" of a dense matrix may be partial: n_slots bounds it).  Slot s is evaluated when
 * meta == null (every slot), or when its vehicle is live (s mod slots_per_env < n_veh[env]) and controlled
 * (flag of meta[s]).  actions[s] = 3 tanh(...) (+ noise_scale * noise[s]), or 0 for the other slots. */
__global__ void __launch_bounds__(PVA_THREADS, 3)
pve_actor_kernel(const float *__restrict__ W, const float *__restrict__ rows, const pve_veh_meta *__restrict__ meta,
                 const int32_t *__restrict__ n_veh, const float *__restrict__ noise, const float noise_scale,
                 float *__restrict__ actions, const int slots_per_env, const int n_env, const long long n_slots_max,
                 int *__restrict__ ticket, const int32_t *__restrict__ limit_dev, const int limit_mult,
                 const uint8_t *__restrict__ mask, const int slot_step) {
    /* device-side row count of a dense matrix (pve_actor_forward_n): rows >= limit_dev[0] * limit_mult are skipped */
    const long long n_slots = limit_dev ? min(n_slots_max, (long long)limit_dev[0] * limit_mult) : n_slots_max;
    extern __shared__ __align__(16) unsigned char pva_smem[];
    float *const w = reinterpret_cast<float *>(pva_smem);                        /* [PVA_WPAD] parameters */
    float *const a = w + PVA_WPAD;                                               /* [PVA_TILE][PVA_AS] activations */
    int *const ring = reinterpret_cast<int *>(a + PVA_TILE * PVA_AS);            /* [PVA_RING] queued slots */
    __shared__ int q_tail, next_env;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < PVA_COUNT; i += PVA_THREADS) w[i] = W[i];
    if (tid == 0) q_tail = 0;
    __syncthreads();

    /* Intersections are handed out PVA_BATCH at a time (ticket[0]); a CTA evaluates a tile whenever 128 vehicles
     * are queued and its last, partial tile (at the cost of its size class) when the pool is empty.  The last
     * CTA to leave re-arms the tickets. */
    int head = 0;                                /* ring positions are monotonic counters, used modulo PVA_RING */
    const int batch = (PVA_RING - PVA_TILE) / slots_per_env >= PVA_BATCH ? PVA_BATCH
                      : ((PVA_RING - PVA_TILE) / slots_per_env > 0 ? (PVA_RING - PVA_TILE) / slots_per_env : 1);
    for (;;) {
        if (tid == 0) next_env = atomicAdd(&ticket[0], batch);
        __syncthreads();
        const int envb = next_env < n_env ? next_env : n_env;
        const bool flush = envb == n_env;        /* the pool is empty: the partial tile */
        if (!flush) {
            /* which slots carry a controlled vehicle (main.py:401-403)? */
            const int enve = min(n_env, envb + batch);
            const int total = (enve - envb) * slots_per_env;
            const long long base = (long long)envb * slots_per_env;
            for (int s0 = 0; s0 < total; s0 += PVA_THREADS) {
                const int s = s0 + tid;
                const long long gs = base + s;
                bool want = s < total && gs < n_slots;
                if (want && mask) want = mask[gs] != 0;                      /* only the marked rows */
                if (want && meta) {
                    const int e = s / slots_per_env;
                    want = (s - e * slots_per_env) < n_veh[envb + e] && ((meta[gs].packed >> 24) & PVE_F_CONTROL) != 0;
                    if (!want) actions[gs] = 0.f;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, want);
                int at = 0;
                if (lane == 0 && bal) at = atomicAdd(&q_tail, __popc(bal));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (want) ring[(at + __popc(bal & ((1u << lane) - 1u))) & (PVA_RING - 1)] = (int)(gs * slot_step);   /* slot -> row of `rows` / element of `actions` (slot_step > 1: strided matrix) */
            }
            __syncthreads();
        }
        const int tail = q_tail;
        __syncthreads();                         /* everybody has read the tail before the next append moves it */
        while (tail - head >= (flush ? 1 : PVA_TILE)) {
            const int n_valid = min(PVA_TILE, tail - head);
            if (n_valid > 64) pva_round<8>(w, a, ring, head, n_valid, rows, noise, noise_scale, actions);
            else if (n_valid > 32) pva_round<4>(w, a, ring, head, n_valid, rows, noise, noise_scale, actions);
            else pva_round<2>(w, a, ring, head, n_valid, rows, noise, noise_scale, actions);
            head += n_valid;
        }
        if (flush) break;
    }
    if (tid == 0 && atomicAdd(&ticket[1], 1) == (int)gridDim.x - 1) { ticket[0] = 0; ticket[1] = 0; }
}
#endif  /* __CUDACC__ */
#endif
