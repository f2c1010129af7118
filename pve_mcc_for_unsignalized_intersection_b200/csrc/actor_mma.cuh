/* Tensor-core version of the batched actor (see actor.cuh for what is computed and why fp32 accuracy is
 * required): the two dense layers run on the tensor cores as bf16 x 3 split-precision products.
 *
 * Every fp32 operand is written as the sum of three bf16 numbers, x = xh + xm + xl (24 mantissa bits in
 * total, so the split is exact up to 2^-24 relative), and the product keeps the six terms of order <= 2:
 *     x w ~= xh wh + (xh wm + xm wh) + (xh wl + xm wm + xl wh)          (dropped terms <= 2^-32 relative)
 * accumulated in fp32 by  mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32.  The result is as close to the
 * float64 evaluation as the fp32 FFMA kernel is (tests/test_gpu_actor.py holds both to the same bound).
 *
 * Work split: a CTA of 8 warps evaluates 128 queued vehicles per round, 16 rows per warp, all 64 hidden units
 * (8 n-blocks).  The accumulator fragments of layer 1 (bias, LayerNorm and ReLU applied in registers; the
 * four lanes of a quad hold one row: two shuffle steps) ARE the A fragments of layer 2 -- no shared-memory
 * round trip between the layers.  Weights are split and laid out in B-fragment order once on the host
 * (pve_actor_create), 36 KB, copied to shared memory by every CTA; one LDS.64 per lane and fragment.
 * The outer loop (tickets, ring of controlled slots) is the one of actor.cuh.
 */
#ifndef PVE_ACTOR_MMA_CUH
#define PVE_ACTOR_MMA_CUH

#include "actor.cuh"

/* packed parameter block of the tensor-core kernel (32-bit words) */
enum { PVM_WF1 = 0,                                   /* [2 kk][8 j][3 split][32 lanes][2] bf16x2 */
       PVM_WF2 = PVM_WF1 + 2 * 8 * 3 * 64,            /* [4 kk][8 j][3 split][32 lanes][2] bf16x2 */
       PVM_VEC = PVM_WF2 + 4 * 8 * 3 * 64,            /* fp32 vectors, offsets below */
       PVM_LN0_G = PVM_VEC, PVM_LN0_B = PVM_LN0_G + 32, PVM_B1 = PVM_LN0_B + 32, PVM_LN1_G = PVM_B1 + 64,
       PVM_LN1_B = PVM_LN1_G + 64, PVM_B2 = PVM_LN1_B + 64, PVM_LN2_G = PVM_B2 + 64, PVM_LN2_B = PVM_LN2_G + 64,
       PVM_W3 = PVM_LN2_B + 64, PVM_B3 = PVM_W3 + 64, PVM_WORDS = PVM_B3 + 4 };

/* ---- host side: split + fragment order (called by pve_actor_create) -------------------------- */
static inline uint16_t pvm_bf16_rne(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float pvm_bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static inline void pvm_split3(float x, uint16_t out[3]) {
    out[0] = pvm_bf16_rne(x);
    const float r1 = x - pvm_bf16_to_f32(out[0]);
    out[1] = pvm_bf16_rne(r1);
    const float r2 = r1 - pvm_bf16_to_f32(out[1]);
    out[2] = pvm_bf16_rne(r2);
}
/* W: flat fp32 parameters in the order of include/pve_mcc.h; out: PVM_WORDS 32-bit words */
static inline void pvm_pack(const float *W, uint32_t *out) {
    memset(out, 0, sizeof(uint32_t) * PVM_WORDS);
    for (int layer = 0; layer < 2; ++layer) {
        const int K = layer ? 64 : 28, KK = layer ? 4 : 2;
        const float *Wm = W + (layer ? PVA_W2 : PVA_W1);               /* [K][64] row-major */
        uint32_t *dst = out + (layer ? PVM_WF2 : PVM_WF1);
        for (int kk = 0; kk < KK; ++kk)
            for (int j = 0; j < 8; ++j)
                for (int lane = 0; lane < 32; ++lane)
                    for (int reg = 0; reg < 2; ++reg) {
                        const int g = lane >> 2, t = lane & 3, n = 8 * j + g, k0 = 16 * kk + 2 * t + 8 * reg;
                        uint16_t e0[3] = {0, 0, 0}, e1[3] = {0, 0, 0};
                        if (k0 < K) pvm_split3(Wm[k0 * 64 + n], e0);
                        if (k0 + 1 < K) pvm_split3(Wm[(k0 + 1) * 64 + n], e1);
                        for (int s = 0; s < 3; ++s)
                            dst[(((kk * 8 + j) * 3 + s) * 32 + lane) * 2 + reg] = (uint32_t)e0[s] | ((uint32_t)e1[s] << 16);
                    }
    }
    float *v = (float *)out;
    memcpy(v + PVM_LN0_G, W + PVA_LN0_G, 28 * 4); memcpy(v + PVM_LN0_B, W + PVA_LN0_B, 28 * 4);
    memcpy(v + PVM_B1, W + PVA_B1, 64 * 4); memcpy(v + PVM_LN1_G, W + PVA_LN1_G, 64 * 4);
    memcpy(v + PVM_LN1_B, W + PVA_LN1_B, 64 * 4); memcpy(v + PVM_B2, W + PVA_B2, 64 * 4);
    memcpy(v + PVM_LN2_G, W + PVA_LN2_G, 64 * 4); memcpy(v + PVM_LN2_B, W + PVA_LN2_B, 64 * 4);
    memcpy(v + PVM_W3, W + PVA_W3, 64 * 4); v[PVM_B3] = W[PVA_B3];
}

#ifdef __CUDACC__
#include <cuda_bf16.h>

#define PVM_THREADS 256
#define PVM_TILE 128
#define PVM_AS 40             /* LayerNorm-0 tile stride (floats): conflict-free 64-bit fragment loads */
#define PVM_SMEM_BYTES (PVM_WORDS * 4 + PVM_TILE * PVM_AS * 4 + PVA_RING * 4)

__device__ __forceinline__ void pvm_mma(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

/* (x0, x1) -> three bf16x2 words (x0 in the low half: the lower k index) */
__device__ __forceinline__ void pvm_split_pair(float x0, float x1, uint32_t &h, uint32_t &m, uint32_t &l) {
    const __nv_bfloat162 bh = __floats2bfloat162_rn(x0, x1);
    const float r0 = x0 - __low2float(bh), r1 = x1 - __high2float(bh);
    const __nv_bfloat162 bm = __floats2bfloat162_rn(r0, r1);
    const float q0 = r0 - __low2float(bm), q1 = r1 - __high2float(bm);
    const __nv_bfloat162 bl = __floats2bfloat162_rn(q0, q1);
    h = *reinterpret_cast<const uint32_t *>(&bh);
    m = *reinterpret_cast<const uint32_t *>(&bm);
    l = *reinterpret_cast<const uint32_t *>(&bl);
}

/* acc[j] += A(16 x 16 k-slice, split) * B(k-slice kk, n-block j), six split products, small terms first */
__device__ __forceinline__ void pvm_kstep(float (&acc)[8][4], const uint32_t (&ah)[4], const uint32_t (&am)[4],
                                          const uint32_t (&al)[4], const uint32_t *__restrict__ wf, int kk, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint2 *bp = reinterpret_cast<const uint2 *>(wf) + ((kk * 8 + j) * 3) * 32 + lane;
        const uint2 bh = bp[0], bm = bp[32], bl = bp[64];
        pvm_mma(acc[j], al, bh); pvm_mma(acc[j], am, bm); pvm_mma(acc[j], ah, bl);
        pvm_mma(acc[j], am, bh); pvm_mma(acc[j], ah, bm);
        pvm_mma(acc[j], ah, bh);
    }
}

/* bias, LayerNorm over the 64 units of the two rows this lane holds (g and g + 8; a quad holds a row), ReLU */
__device__ __forceinline__ void pvm_ln_relu(float (&acc)[8][4], const float *__restrict__ bias, const float *__restrict__ gamma,
                                            const float *__restrict__ beta, int t) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 b = *reinterpret_cast<const float2 *>(bias + 8 * j + 2 * t);
        acc[j][0] += b.x; acc[j][1] += b.y; acc[j][2] += b.x; acc[j][3] += b.y;
        s0 += acc[j][0] + acc[j][1]; s1 += acc[j][2] + acc[j][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float m0 = s0 * (1.f / 64.f), m1 = s1 * (1.f / 64.f);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float d0 = acc[j][0] - m0, d1 = acc[j][1] - m0, d2 = acc[j][2] - m1, d3 = acc[j][3] - m1;
        q0 = fmaf(d0, d0, q0); q0 = fmaf(d1, d1, q0); q1 = fmaf(d2, d2, q1); q1 = fmaf(d3, d3, q1);
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 2); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float r0 = rsqrtf(q0 * (1.f / 64.f) + PVA_EPS), r1 = rsqrtf(q1 * (1.f / 64.f) + PVA_EPS);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 g = *reinterpret_cast<const float2 *>(gamma + 8 * j + 2 * t);
        const float2 e = *reinterpret_cast<const float2 *>(beta + 8 * j + 2 * t);
        const float i00 = r0 * g.x, i01 = r0 * g.y, i10 = r1 * g.x, i11 = r1 * g.y;
        acc[j][0] = fmaxf(fmaf(acc[j][0], i00, fmaf(-m0, i00, e.x)), 0.f);
        acc[j][1] = fmaxf(fmaf(acc[j][1], i01, fmaf(-m0, i01, e.y)), 0.f);
        acc[j][2] = fmaxf(fmaf(acc[j][2], i10, fmaf(-m1, i10, e.x)), 0.f);
        acc[j][3] = fmaxf(fmaf(acc[j][3], i11, fmaf(-m1, i11, e.y)), 0.f);
    }
}

/* one round of this warp: rows [16 warp, 16 warp + 16) of the tile, i.e. ring entries head + those */
__device__ __forceinline__ void pvm_warp_round(const uint32_t *__restrict__ pw, float *__restrict__ a, const int *__restrict__ ring,
                                               const int head, const int n_valid, const float *__restrict__ rows,
                                               const float *__restrict__ noise, const float noise_scale,
                                               float *__restrict__ actions) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const float *vec = reinterpret_cast<const float *>(pw);
    float *aw = a + warp * 16 * PVM_AS;
    /* first LayerNorm (NET:27): lanes l and l + 16 share row l (float4 pieces 0-3 / 4-6) */
    {
        const int r = lane & 15, half = lane >> 4, row = warp * 16 + r;
        const bool valid = row < n_valid;
        const long long gs = valid ? (long long)ring[(head + row) & (PVA_RING - 1)] : 0;
        const float4 *src = reinterpret_cast<const float4 *>(rows + gs * PVE_OBS_W) + half * 4;
        float4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            x[q] = (valid && half * 4 + q < 7) ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        float s = ((x[0].x + x[0].y) + (x[0].z + x[0].w)) + ((x[1].x + x[1].y) + (x[1].z + x[1].w))
                  + ((x[2].x + x[2].y) + (x[2].z + x[2].w)) + ((x[3].x + x[3].y) + (x[3].z + x[3].w));
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        const float mean = s * (1.f / 28.f);
        float q2 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (half * 4 + q < 7) {
                const float d0 = x[q].x - mean, d1 = x[q].y - mean, d2 = x[q].z - mean, d3 = x[q].w - mean;
                q2 = fmaf(d0, d0, q2); q2 = fmaf(d1, d1, q2); q2 = fmaf(d2, d2, q2); q2 = fmaf(d3, d3, q2);
            }
        q2 += __shfl_xor_sync(0xffffffffu, q2, 16);
        const float rs = rsqrtf(q2 * (1.f / 28.f) + PVA_EPS);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = (half * 4 + q) * 4;                    /* columns 28..31 come out as 0 (gamma = beta = 0) */
            const float4 gm = *reinterpret_cast<const float4 *>(vec + PVM_LN0_G + c);
            const float4 bt = *reinterpret_cast<const float4 *>(vec + PVM_LN0_B + c);
            const float i0 = rs * gm.x, i1 = rs * gm.y, i2 = rs * gm.z, i3 = rs * gm.w;
            float4 y;
            y.x = fmaf(x[q].x, i0, fmaf(-mean, i0, bt.x)); y.y = fmaf(x[q].y, i1, fmaf(-mean, i1, bt.y));
            y.z = fmaf(x[q].z, i2, fmaf(-mean, i2, bt.z)); y.w = fmaf(x[q].w, i3, fmaf(-mean, i3, bt.w));
            *reinterpret_cast<float4 *>(aw + r * PVM_AS + c) = y;
        }
    }
    __syncwarp();
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
    /* Dense 28 -> 64 (NET:28): A fragments from the tile */
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const float2 v00 = *reinterpret_cast<const float2 *>(aw + g * PVM_AS + 16 * kk + 2 * t);
        const float2 v10 = *reinterpret_cast<const float2 *>(aw + (g + 8) * PVM_AS + 16 * kk + 2 * t);
        const float2 v01 = *reinterpret_cast<const float2 *>(aw + g * PVM_AS + 16 * kk + 8 + 2 * t);
        const float2 v11 = *reinterpret_cast<const float2 *>(aw + (g + 8) * PVM_AS + 16 * kk + 8 + 2 * t);
        uint32_t ah[4], am[4], al[4];
        pvm_split_pair(v00.x, v00.y, ah[0], am[0], al[0]);
        pvm_split_pair(v10.x, v10.y, ah[1], am[1], al[1]);
        pvm_split_pair(v01.x, v01.y, ah[2], am[2], al[2]);
        pvm_split_pair(v11.x, v11.y, ah[3], am[3], al[3]);
        pvm_kstep(acc, ah, am, al, pw + PVM_WF1, kk, lane);
    }
    __syncwarp();                                                /* the tile rows may be overwritten next round */
    pvm_ln_relu(acc, vec + PVM_B1, vec + PVM_LN1_G, vec + PVM_LN1_B, t);             /* NET:28-32 */
    /* Dense 64 -> 64 (NET:34): the accumulator fragments of n-blocks 2 kk, 2 kk + 1 are the A fragment of slice kk */
    float acc2[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc2[j][0] = 0.f; acc2[j][1] = 0.f; acc2[j][2] = 0.f; acc2[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t ah[4], am[4], al[4];
        pvm_split_pair(acc[2 * kk][0], acc[2 * kk][1], ah[0], am[0], al[0]);
        pvm_split_pair(acc[2 * kk][2], acc[2 * kk][3], ah[1], am[1], al[1]);
        pvm_split_pair(acc[2 * kk + 1][0], acc[2 * kk + 1][1], ah[2], am[2], al[2]);
        pvm_split_pair(acc[2 * kk + 1][2], acc[2 * kk + 1][3], ah[3], am[3], al[3]);
        pvm_kstep(acc2, ah, am, al, pw + PVM_WF2, kk, lane);
    }
    pvm_ln_relu(acc2, vec + PVM_B2, vec + PVM_LN2_G, vec + PVM_LN2_B, t);            /* NET:34-38 */
    /* Dense 64 -> 1, 3 tanh (NET:40-47) */
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 u = *reinterpret_cast<const float2 *>(vec + PVM_W3 + 8 * j + 2 * t);
        o0 = fmaf(acc2[j][0], u.x, o0); o0 = fmaf(acc2[j][1], u.y, o0);
        o1 = fmaf(acc2[j][2], u.x, o1); o1 = fmaf(acc2[j][3], u.y, o1);
    }
    o0 += __shfl_xor_sync(0xffffffffu, o0, 1); o1 += __shfl_xor_sync(0xffffffffu, o1, 1);
    o0 += __shfl_xor_sync(0xffffffffu, o0, 2); o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
    if (t < 2) {                                                 /* lane t = 0 writes row g, t = 1 row g + 8 */
        const int row = warp * 16 + g + 8 * t;
        if (row < n_valid) {
            const long long gs = (long long)ring[(head + row) & (PVA_RING - 1)];
            float act = 3.f * tanhf((t ? o1 : o0) + vec[PVM_B3]);
            if (noise) act += noise_scale * noise[gs];                             /* main.py:44 */
            actions[gs] = act;
        }
    }
}

/* same contract as pve_actor_kernel (actor.cuh); PW = the packed block of pvm_pack */
__global__ void __launch_bounds__(PVM_THREADS, 2)
pve_actor_mma_kernel(const uint32_t *__restrict__ PW, const float *__restrict__ rows, const pve_veh_meta *__restrict__ meta,
                     const int32_t *__restrict__ n_veh, const float *__restrict__ noise, const float noise_scale,
                     float *__restrict__ actions, const int slots_per_env, const int n_env, const long long n_slots_max,
                     int *__restrict__ ticket, const int32_t *__restrict__ limit_dev, const int limit_mult,
                 const uint8_t *__restrict__ mask, const int slot_step) {
    /* device-side row count of a dense matrix (pve_actor_forward_n): rows >= limit_dev[0] * limit_mult are skipped */
    const long long n_slots = limit_dev ? min(n_slots_max, (long long)limit_dev[0] * limit_mult) : n_slots_max;
    extern __shared__ __align__(16) unsigned char pvm_smem[];
    uint32_t *const pw = reinterpret_cast<uint32_t *>(pvm_smem);
    float *const a = reinterpret_cast<float *>(pw + PVM_WORDS);
    int *const ring = reinterpret_cast<int *>(a + PVM_TILE * PVM_AS);
    __shared__ int q_tail, next_env;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < PVM_WORDS / 4; i += PVM_THREADS) reinterpret_cast<uint4 *>(pw)[i] = reinterpret_cast<const uint4 *>(PW)[i];
    if (tid == 0) q_tail = 0;
    __syncthreads();

    int head = 0;
    const int batch = (PVA_RING - PVM_TILE) / slots_per_env >= PVA_BATCH ? PVA_BATCH
                      : ((PVA_RING - PVM_TILE) / slots_per_env > 0 ? (PVA_RING - PVM_TILE) / slots_per_env : 1);
    for (;;) {
        if (tid == 0) next_env = atomicAdd(&ticket[0], batch);
        __syncthreads();
        const int envb = next_env < n_env ? next_env : n_env;
        const bool flush = envb == n_env;
        if (!flush) {
            const int enve = min(n_env, envb + batch);
            const int total = (enve - envb) * slots_per_env;
            const long long base = (long long)envb * slots_per_env;
            for (int s0 = 0; s0 < total; s0 += PVM_THREADS) {
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
        __syncthreads();
        while (tail - head >= (flush ? 1 : PVM_TILE)) {
            const int n_valid = min(PVM_TILE, tail - head);
            if ((tid >> 5) * 16 < n_valid)                   /* warps without rows skip the round */
                pvm_warp_round(pw, a, ring, head, n_valid, rows, noise, noise_scale, actions);
            head += n_valid;
            __syncthreads();                                 /* ring entries are free again */
        }
        if (flush) break;
    }
    if (tid == 0 && atomicAdd(&ticket[1], 1) == (int)gridDim.x - 1) { ticket[0] = 0; ticket[1] = 0; }
}
#endif  /* __CUDACC__ */
#endif
