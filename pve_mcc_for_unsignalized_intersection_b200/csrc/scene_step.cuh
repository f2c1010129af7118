/*
 * scene_step.cuh -- one tick of one intersection, executed by one CTA.
 *
 * Replaces, for lane_num = 12, the reference's step() x V (traffic_interaction_scene.py
 * "TIS" 1501-1539), scene_update() (TIS:222-376) with get_virtual_distance (TIS:733-803),
 * get_p (TIS:1250-1290), get_state (TIS:1292-1338), virtual_lane_search_closer (TIS:1340-1405),
 * check_lock (TIS:1469-1499), add_new_veh (TIS:378-433) and delete_vehicle() (TIS:435-444).
 *
 * The reference is sequential and order dependent; this file is an order-free reformulation
 * (SURVEY.md section 3.3, Q1-Q6):
 *   Q1  rear-end safety chain      -> two candidate next states per vehicle + a chain of
 *                                     1-bit boolean functions resolved per lane
 *   Q2  stale virtual-lane head    -> (head_lane, head_j) persisted in the header
 *   Q3  neighbour rows             -> this tick's row 0 if the neighbour precedes the ego in
 *                                     (lane, j) order, else last tick's row (ping-pong buffer)
 *   Q4  collision counters         -> hit bit per agent + "earlier hitters" / "all hitters" counts
 *   Q5  reward[-1] overrides       -> resolved per agent after all flags are known
 *   Q6  stable sorts               -> total orders (pos, slot) and (|delta|, rank)
 *
 * All state arithmetic is IEEE float64 with contraction disabled (nvcc -fmad=false), in the
 * reference's association order, so p, v, a stay bit-identical to the float64 reference.
 *
 * The body is written as barrier-separated phases (PVE_FOR_TID / PVE_END_TID).  nvcc compiles
 * it as the CUDA kernel.  tests/emul/ compiles THE SAME SOURCE with g++ as a sequential
 * emulation (each phase looped over tid) so that the CPU-only test tier can check the kernel
 * logic against the oracle.  The emulation is test infrastructure: the product library never
 * contains it.
 */
#pragma once
#include <math.h>
#include <stdint.h>

#include "pve_mcc.h"

#ifdef __CUDACC__
#define PVE_DEV __device__ __forceinline__
#define PVE_HD __host__ __device__ __forceinline__
#define PVE_FOR_TID(tid) { const int tid = (int)threadIdx.x;
#define PVE_END_TID } __syncthreads();
#define PVE_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#else
#define PVE_DEV static inline
#define PVE_HD static inline
#define PVE_FOR_TID(tid) for (int tid = 0; tid < NT; ++tid) {
#define PVE_END_TID }
static inline int pve_emul_atomic_add(int *p, int v) { int o = *p; *p = o + v; return o; }
#define PVE_ATOMIC_ADD(ptr, val) pve_emul_atomic_add((ptr), (val))
#endif

struct alignas(16) pve_v4 { uint32_t x, y, z, w; };

/* kernel parameters (by value; lives in the constant bank) */
struct PveParams {
    double dt, dt2, vm, vM, am, aM, v0, thr, lane_in, remove_p, lane_cw, abs_am, two_abs_am, aspan;
    double lane_len[3];
    double spawn_p[3];
    double vd_a1[2][4], vd_a2[2][4], vd_b[2][4];
    double rot_cos[4], rot_sin[4];
    int8_t l2l[PVE_NLANE][4];      /* lane2lane, TIS:153-166 */
    int8_t rev_dir[PVE_NLANE][4];  /* directions d whose lane2lane[d] contains this lane ... */
    int8_t rev_k[PVE_NLANE][4];    /* ... and its position k there */
    int32_t B, VC, AC, K;
    int64_t out_cap;
};

struct PveState {
    pve_env_header *hdr;
    double *p, *v, *a, *js;
    pve_veh_meta *meta;
    float *row0[2];           /* ping-pong: [phase] is read, [phase ^ 1] is written */
    int32_t *n_ctrl, *n_veh;  /* [B] copies of the header counts for the offset scan */
    double *stats;            /* [B][PVE_NSTAT] running per-intersection statistics */
    int32_t *agent_offset;    /* [B+1] rows of this tick (written by the scan kernel) */
};

enum { PVE_STAT_AGENT = 0, PVE_STAT_VEH, PVE_STAT_COLL, PVE_STAT_LOCK, PVE_STAT_JERK, PVE_STAT_RSUM,
       PVE_STAT_RSQ, PVE_STAT_REMOVED, PVE_STAT_STEPS, PVE_STAT_Q5U, PVE_NSTAT };

/* ---------------------------------------------------------------------------------------------
 * shared-memory layout, identical on host and device
 * ------------------------------------------------------------------------------------------- */
struct PveSmem {
    /* float64 [VC] */
    double *sp, *sv, *sa, *sjerk, *sjs;
    /* union region: step candidates [VC] x 6, later the virtual-lane arrays */
    double *cta0, *cta1, *cp0, *cv0, *cp1, *cv1;
    double *epos, *spos;          /* [EC] */
    uint16_t *eidx, *sidx;        /* [EC] */
    /* float64 [AC] */
    double *virdis;
    float *row0;                  /* [AC][28] */
    float *rew;                   /* [AC] */
    int32_t *suid;                /* [VC] */
    uint32_t *spk;                /* [VC] */
    int32_t *incb, *inct, *cpv;   /* [AC] */
    uint16_t *acnt, *surv;        /* [VC] exclusive counts: agents before k, survivors before k */
    uint16_t *vidx, *arank;       /* [AC] */
    int16_t *hdra;                /* [AC] agent index of vir_header, -1 if none */
    uint16_t *nn;                 /* [AC][6] vehicle slot of each neighbour, 0xFFFF if none */
    uint8_t *lane_of, *fbits, *ssel, *del, *slock, *ctl0; /* [VC] */
    int8_t *slocka;               /* [VC] */
    uint8_t *hit, *q5, *fin5, *status; /* [AC] */
    pve_env_header *hdr;
    int32_t *lane_off, *lane_aoff, *vl_base, *vl_cnt, *misc, *wsum;
    double *dsum;
};

enum { M_V = 0, M_A, M_NREM, M_PASSED, M_COLL, M_LOCK, M_NCTRL, M_Q5U, M_SURV, M_PSTEP, M_SPAWN0 /* 12 */,
       M_SPREF0 = M_SPAWN0 + 12 /* 13 */, M_COLLAG = M_SPREF0 + 13, M_OUTOK, M_COUNT };

PVE_HD size_t pve_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

PVE_HD size_t pve_smem_carve(PveSmem *m, unsigned char *base, int VC, int AC) {
    const int EC = 5 * AC;
    size_t o = 0;
#define CARVE(field, type, count) \
    do { o = pve_align_up(o, alignof(type) < 8 ? 8 : alignof(type)); \
         if (m) m->field = (type *)(base + o); o += sizeof(type) * (size_t)(count); } while (0)
    CARVE(hdr, pve_env_header, 1);
    o = pve_align_up(o, 16);
    CARVE(sp, double, VC); CARVE(sv, double, VC); CARVE(sa, double, VC);
    CARVE(sjerk, double, VC); CARVE(sjs, double, VC);
    /* union */
    size_t u0 = pve_align_up(o, 16);
    o = u0;
    CARVE(cta0, double, VC); CARVE(cta1, double, VC); CARVE(cp0, double, VC);
    CARVE(cv0, double, VC); CARVE(cp1, double, VC); CARVE(cv1, double, VC);
    size_t uA = o;
    o = u0;
    CARVE(epos, double, EC); CARVE(spos, double, EC); CARVE(eidx, uint16_t, EC); CARVE(sidx, uint16_t, EC);
    o = (o > uA) ? o : uA;
    CARVE(virdis, double, AC);
    o = pve_align_up(o, 16);
    CARVE(row0, float, (size_t)AC * PVE_OBS_W);
    CARVE(rew, float, AC);
    CARVE(suid, int32_t, VC); CARVE(spk, uint32_t, VC);
    CARVE(incb, int32_t, AC); CARVE(inct, int32_t, AC); CARVE(cpv, int32_t, AC);
    CARVE(acnt, uint16_t, VC + 1); CARVE(surv, uint16_t, VC + 1);
    CARVE(vidx, uint16_t, AC); CARVE(arank, uint16_t, AC); CARVE(hdra, int16_t, AC);
    CARVE(nn, uint16_t, (size_t)AC * PVE_NNBR);
    CARVE(lane_of, uint8_t, VC); CARVE(fbits, uint8_t, VC); CARVE(ssel, uint8_t, VC);
    CARVE(del, uint8_t, VC); CARVE(slock, uint8_t, VC); CARVE(ctl0, uint8_t, VC);
    CARVE(slocka, int8_t, VC);
    CARVE(hit, uint8_t, AC); CARVE(q5, uint8_t, AC); CARVE(fin5, uint8_t, AC); CARVE(status, uint8_t, AC);
    CARVE(lane_off, int32_t, 16); CARVE(lane_aoff, int32_t, 16); CARVE(vl_base, int32_t, 16);
    CARVE(vl_cnt, int32_t, 16); CARVE(misc, int32_t, M_COUNT); CARVE(wsum, int32_t, 40);
    CARVE(dsum, double, 40);
#undef CARVE
    return pve_align_up(o, 16);
}

/* ---------------------------------------------------------------------------------------------
 * block collectives.  Device: warp ballot / shuffle + one smem exchange.  Host emulation:
 * sequential loops over the same shared arrays.
 * ------------------------------------------------------------------------------------------- */
/* out[k] = number of set flags before k (k = 0..n), i.e. an exclusive scan; out[n] = total.
 * This is the stream-compaction index used for agent numbering and for vehicle removal. */
template <int NT>
PVE_DEV void pve_block_excl_scan(const uint8_t *flag, uint16_t *out, int n, int32_t *wsum) {
#ifdef __CUDACC__
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    int carry = 0;
    for (int base = 0; base < n; base += NT) {
        const int k = base + tid;
        const int f = (k < n) ? (flag[k] != 0) : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        const int wp = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { const int c = wsum[w]; woff += (w < warp) ? c : 0; tot += c; }
        if (k < n) out[k] = (uint16_t)(carry + woff + wp);
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) out[n] = (uint16_t)carry;
    __syncthreads();
#else
    (void)wsum;
    int c = 0;
    for (int k = 0; k < n; ++k) { out[k] = (uint16_t)c; c += (flag[k] != 0); }
    out[n] = (uint16_t)c;
#endif
}

/* deterministic sums of two shared arrays: r0 = sum x[k], r1 = sum x[k]^2, r2 = sum y[k] */
template <int NT>
PVE_DEV void pve_block_sums(const float *x, const double *y, int n, double *dsum) {
#ifdef __CUDACC__
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int k = tid; k < n; k += NT) { const double r = (double)x[k]; s0 += r; s1 += r * r; s2 += y[k]; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, d);
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) { dsum[3 + warp * 3] = s0; dsum[4 + warp * 3] = s1; dsum[5 + warp * 3] = s2; }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0, a1 = 0, a2 = 0;
        for (int w = 0; w < NW; ++w) { a0 += dsum[3 + w * 3]; a1 += dsum[4 + w * 3]; a2 += dsum[5 + w * 3]; }
        dsum[0] = a0; dsum[1] = a1; dsum[2] = a2;
    }
    __syncthreads();
#else
    double a0 = 0, a1 = 0, a2 = 0;
    for (int k = 0; k < n; ++k) { const double r = (double)x[k]; a0 += r; a1 += r * r; a2 += y[k]; }
    dsum[0] = a0; dsum[1] = a1; dsum[2] = a2;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * TIS:1250-1290 get_p for lane_num = 12 (yaw is never read by the caller)
 * ------------------------------------------------------------------------------------------- */
PVE_DEV void pve_world_xy(const PveParams &P, double p, int lane, double *x, double *y) {
    const double cw = P.lane_cw;
    const int m = lane % 3;
    double tx, ty;
    if (m == 1) {
        tx = p - 6 * cw; ty = 3 * cw;                                           /* TIS:1270 */
    } else {
        const double L = P.lane_len[m];
        if (p > L) {
            tx = p - L + 6 * cw; ty = (m == 0) ? cw : 5 * cw;                   /* TIS:1256, 1274 */
        } else if (p > 0) {
            const double r_a = (L - p) / L * 3.141593 / 2;                      /* TIS:1259, 1277 */
            double sn, cs;
#ifdef __CUDACC__
            sincos(r_a, &sn, &cs);
#else
            sn = sin(r_a); cs = cos(r_a);
#endif
            if (m == 0) { tx = 6 * cw - (7 * cw) * sn; ty = -6 * cw + (7 * cw) * cs; }   /* TIS:1262-1263 */
            else        { tx = 6 * cw + (-cw) * sn;    ty = 6 * cw + (-cw) * cs; }       /* TIS:1281-1282 */
        } else if (m == 0) {
            tx = -cw; ty = -6 * cw + p;                                         /* TIS:1267 */
        } else {
            tx = 5 * cw; ty = 6 * cw - p;                                       /* TIS:1286 */
        }
    }
    const double c = P.rot_cos[lane / 3], s = P.rot_sin[lane / 3];              /* TIS:1251 */
    *x = tx * c - ty * s;                                                       /* TIS:1287 */
    *y = ty * c + tx * s;                                                       /* TIS:1288 */
}

/* ---------------------------------------------------------------------------------------------
 * one tick of intersection b
 * ------------------------------------------------------------------------------------------- */
template <int NT>
PVE_DEV void pve_step_block(const PveParams &P, const PveState &S, const pve_outputs &O,
                            const int32_t *spawn_tick, const float *actions, const int phase,
                            const int b, unsigned char *smem_raw) {
    PveSmem m;
    pve_smem_carve(&m, smem_raw, P.VC, P.AC);
    const int VC = P.VC;
    const size_t vbase = (size_t)b * (size_t)VC;

    /* ---- L0: header -> shared -------------------------------------------------------------- */
    PVE_FOR_TID(tid)
        if (tid < PVE_HDR_BYTES / 16)
            ((pve_v4 *)m.hdr)[tid] = ((const pve_v4 *)(S.hdr + b))[tid];
        if (tid < M_COUNT) m.misc[tid] = 0;
        if (tid < 16) m.vl_cnt[tid] = 0;
    PVE_END_TID

    PVE_FOR_TID(tid)
        if (tid == 0) {
            int o = 0;
            for (int i = 0; i < PVE_NLANE; ++i) { m.lane_off[i] = o; o += m.hdr->lane_n[i]; }
            m.lane_off[PVE_NLANE] = o;
            m.misc[M_V] = o;
            m.hdr->tick += 1;                                                   /* TIS:223 */
            const int64_t lo = (int64_t)S.agent_offset[b], hi = (int64_t)S.agent_offset[b + 1];
            m.misc[M_OUTOK] = (hi <= P.out_cap && hi - lo <= P.AC) ? 1 : 0;
            if (!m.misc[M_OUTOK]) m.hdr->overflow += 1;      /* output rows do not fit: sticky flag */
        }
    PVE_END_TID
    const int V = m.misc[M_V];

    /* ---- A: load vehicles, both candidate next states (Q1) --------------------------------- */
    PVE_FOR_TID(tid)
        for (int k = tid; k < V; k += NT) {
            const double p = S.p[vbase + k], v = S.v[vbase + k], a = S.a[vbase + k];
            const pve_veh_meta mt = S.meta[vbase + k];
            const double act = (double)actions[vbase + k];
            int i = 0;
            while (i < PVE_NLANE - 1 && k >= m.lane_off[i + 1]) ++i;
            const int j = k - m.lane_off[i];
            const uint32_t fl = mt.packed >> 24;
            const bool ctrl = (fl & PVE_F_CONTROL) != 0;
            const int lock_a = (int)((fl >> 3) & 3u) - 1;
            double ta = fmin(P.aM, fmax(P.am, act));                             /* TIS:1502 */
            if ((fl & PVE_F_LOCK) && lock_a != 0 && p > 70.0) ta = a + (double)lock_a;   /* TIS:1503-1505 */
            const bool forced = (m.hdr->head_lane[i] == i && (int)m.hdr->head_j[i] == j)  /* TIS:1517 */
                                || (i % 3 == 2);                                /* TIS:1519 */
            const double ta0 = fmin(P.aM, fmax(P.am, forced ? P.aM : ta));       /* TIS:1521 */
            const double ta1 = forced ? P.aM : P.am;                             /* TIS:1516 */
            const double pv = p - v * P.dt;
            double v0n = fmin(P.vM, fmax(v + ta0 * P.dt, P.vm));                 /* TIS:1530 */
            double v1n = fmin(P.vM, fmax(v + ta1 * P.dt, P.vm));
            if (!ctrl) { v0n = P.v0; v1n = P.v0; }                               /* TIS:1535 */
            m.cta0[k] = ta0; m.cta1[k] = ta1;
            m.cp0[k] = pv - 0.5 * ta0 * P.dt2;                                   /* TIS:1528 */
            m.cp1[k] = pv - 0.5 * ta1 * P.dt2;
            m.cv0[k] = v0n; m.cv1[k] = v1n;
            m.sp[k] = p; m.sv[k] = v; m.sa[k] = a; m.sjs[k] = S.js[vbase + k];
            m.suid[k] = mt.uid; m.spk[k] = mt.packed;
            m.lane_of[k] = (uint8_t)i;
            m.ctl0[k] = ctrl ? 1 : 0;
            m.del[k] = 0; m.slock[k] = 0; m.slocka[k] = 0;
        }
    PVE_END_TID

    /* ---- B: F_k(s) = "rear-end override fires on k if its leader took candidate s" -------- */
    PVE_FOR_TID(tid)
        for (int k = tid; k < V; k += NT) {
            const int i = m.lane_of[k];
            int f = 0;
            if (k > m.lane_off[i] && m.ctl0[k] && m.ctl0[k - 1]) {               /* TIS:1509-1510 */
                const double v = m.sv[k], p = m.sp[k];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const double vf = s ? m.cv1[k - 1] : m.cv0[k - 1];
                    const double pf = s ? m.cp1[k - 1] : m.cp0[k - 1];
                    if (vf < v) {
                        const double d_safe = v * 0.4 + (v * v - vf * vf) / P.two_abs_am
                                              - (v - vf) * P.vm / P.abs_am;     /* TIS:1512-1514 */
                        if (p - pf < d_safe) f |= (1 << s);                      /* TIS:1515 */
                    }
                }
            }
            m.fbits[k] = (uint8_t)f;
        }
    PVE_END_TID

    /* ---- B2: resolve the chain front to back, one lane per thread ------------------------- */
    PVE_FOR_TID(tid)
        if (tid < PVE_NLANE) {
            int s = 0;
            for (int k = m.lane_off[tid]; k < m.lane_off[tid + 1]; ++k) {
                s = (m.fbits[k] >> s) & 1;
                m.ssel[k] = (uint8_t)s;
            }
        }
    PVE_END_TID

    /* ---- C: commit kinematics -------------------------------------------------------------- */
    PVE_FOR_TID(tid)
        for (int k = tid; k < V; k += NT) {
            const int s = m.ssel[k];
            const double a_new = s ? m.cta1[k] : m.cta0[k];
            m.sjerk[k] = a_new - m.sa[k];                                        /* TIS:1522 */
            m.sa[k] = a_new;                                                     /* TIS:1523 */
            m.sp[k] = s ? m.cp1[k] : m.cp0[k];
            m.sv[k] = s ? m.cv1[k] : m.cv0[k];
            uint32_t pk = m.spk[k];
            uint32_t step = pk & 0xFFFFu;
            step = step < 0xFFFFu ? step + 1 : step;                             /* TIS:1533 */
            const uint32_t fl = (pk >> 24) & (PVE_F_CONTROL | PVE_F_FINISH);     /* TIS:1506-1507 */
            m.spk[k] = (pk & 0x00FF0000u) | step | (fl << 24);
        }
    PVE_END_TID

    /* ---- agent numbering: controlled at step() time == gets outputs this tick ------------- */
    pve_block_excl_scan<NT>(m.ctl0, m.acnt, V, m.wsum);
    const int A = m.acnt[V];

    PVE_FOR_TID(tid)
        for (int k = tid; k < V; k += NT)
            if (m.ctl0[k]) m.vidx[m.acnt[k]] = (uint16_t)k;
        if (tid <= PVE_NLANE) m.lane_aoff[tid] = m.acnt[m.lane_off[tid]];
        for (int g = tid; g < A; g += NT) {
            m.incb[g] = 0; m.inct[g] = 0; m.q5[g] = 0; m.fin5[g] = 0; m.status[g] = 0; m.hit[g] = 0;
        }
    PVE_END_TID

    PVE_FOR_TID(tid)
        if (tid == 0) {
            /* capacity of each virtual lane: own agents + agents of the 4 conflicting lanes */
            int o = 0;
            for (int d = 0; d < PVE_NLANE; ++d) {
                m.vl_base[d] = o;
                if (m.hdr->lane_n[d] > 0) {                                      /* TIS:234 */
                    o += m.lane_aoff[d + 1] - m.lane_aoff[d];
                    if (d % 3 != 2)
                        for (int q = 0; q < 4; ++q) {
                            const int L = P.l2l[d][q];
                            o += m.lane_aoff[L + 1] - m.lane_aoff[L];
                        }
                }
            }
            m.vl_base[PVE_NLANE] = o;
        }
    PVE_END_TID

    /* ---- E: virtual-lane membership, one (agent, target lane) pair per work item ---------- */
    PVE_FOR_TID(tid)
        for (int it = tid; it < 5 * A; it += NT) {
            const int g = it / 5, s = it - g * 5;
            const int k = m.vidx[g];
            const int L = m.lane_of[k];
            int d = -1;
            double pos = 0;
            if (s == 0) {
                d = L; pos = m.sp[k];                                            /* TIS:242-249 */
            } else if (L % 3 != 2) {
                const int dd = P.rev_dir[L][s - 1], q = P.rev_k[L][s - 1];
                if (m.hdr->lane_n[dd] > 0) {                                     /* TIS:234, 259 */
                    const int mv = dd % 3;
                    const double delta = (m.sp[k] - P.vd_a1[mv][q]) + P.vd_a2[mv][q];    /* TIS:733-803 */
                    if (delta > 0) { d = dd; pos = P.vd_b[mv][q] + delta; }
                }
            }
            if (d >= 0) {
                const int e = m.vl_base[d] + PVE_ATOMIC_ADD(&m.vl_cnt[d], 1);
                m.epos[e] = pos; m.eidx[e] = (uint16_t)k;
            }
        }
    PVE_END_TID

    /* ---- F: stable sort by position (TIS:271) as a rank count on the key (pos, slot) ------ */
    PVE_FOR_TID(tid)
        for (int e = tid; e < m.vl_base[PVE_NLANE]; e += NT) {
            int d = 0;
            while (d < PVE_NLANE - 1 && e >= m.vl_base[d + 1]) ++d;
            const int base = m.vl_base[d], n = m.vl_cnt[d];
            if (e - base < n) {
                const double pos = m.epos[e];
                const int idx = m.eidx[e];
                int rank = 0;
                for (int t = 0; t < n; ++t) {
                    const double pt = m.epos[base + t];
                    const int it_ = m.eidx[base + t];
                    rank += (pt < pos || (pt == pos && it_ < idx)) ? 1 : 0;
                }
                m.spos[base + rank] = pos; m.sidx[base + rank] = (uint16_t)idx;
                if (m.lane_of[idx] == d) m.arank[m.acnt[idx]] = (uint16_t)rank;
            }
        }
    PVE_END_TID

    /* ---- G: per agent: neighbours, row 0, reward, collision test -------------------------- */
    PVE_FOR_TID(tid)
        for (int g = tid; g < A; g += NT) {
            const int k = m.vidx[g];
            const int d = m.lane_of[k];
            const int base = m.vl_base[d], n = m.vl_cnt[d], r = m.arank[g];
            const double pe = m.spos[base + r];
            /* vir_header / vir_dis, TIS:1349-1354 */
            if (r == 0) { m.hdra[g] = -1; m.virdis[g] = 100.0; }
            else { m.hdra[g] = (int16_t)m.acnt[m.sidx[base + r - 1]]; m.virdis[g] = pe - m.spos[base + r - 1]; }
            /* six nearest by |delta|, ties to the lower list index (stable sort, TIS:1389) */
            int lo = r - 1, hi = r + 1, run_cur = 0, run_end = -1;
            double run_d = 0, vd0 = 0;
            float *row = m.row0 + (size_t)g * PVE_OBS_W;
            row[0] = (float)pe; row[1] = (float)m.sv[k]; row[2] = (float)m.sa[k]; row[3] = (float)d;   /* TIS:1336 */
            int k0 = -1;
            for (int q = 0; q < PVE_NNBR; ++q) {
                if (run_cur > run_end && lo >= 0) {
                    run_end = lo; run_d = fabs(m.spos[base + lo] - pe);
                    int x = lo;
                    while (x - 1 >= 0 && fabs(m.spos[base + x - 1] - pe) == run_d) --x;
                    run_cur = x; lo = x - 1;
                }
                const bool has_lo = run_cur <= run_end, has_hi = hi < n;
                int pick = -1;
                if (has_lo && (!has_hi || run_d <= fabs(m.spos[base + hi] - pe))) pick = run_cur++;
                else if (has_hi) pick = hi++;
                float *o4 = row + 4 * (q + 1);
                if (pick >= 0) {
                    const int kn = m.sidx[base + pick];
                    const double vd = m.spos[base + pick];
                    m.nn[g * PVE_NNBR + q] = (uint16_t)kn;
                    o4[0] = (float)vd; o4[1] = (float)m.sv[kn]; o4[2] = (float)m.sa[kn];
                    o4[3] = (float)m.lane_of[kn];                                /* TIS:1330 */
                    if (q == 0) { k0 = kn; vd0 = vd; }
                } else {
                    m.nn[g * PVE_NNBR + q] = 0xFFFFu;
                    o4[0] = 0.f; o4[1] = 0.f; o4[2] = 0.f; o4[3] = 0.f;          /* TIS:1334 */
                }
            }
            /* reward, TIS:293-320 */
            const double p = m.sp[k], v = m.sv[k];
            double t_distance = 2, d_distance = 10;
            if (k0 >= 0) {
                d_distance = fabs(p - vd0);                                      /* TIS:300 */
                if (d_distance != 0) t_distance = (p - vd0) / (v - m.sv[k0] + 0.0001);   /* TIS:304 */
            }
            double r_ = 0;
            if (0 < t_distance && t_distance < 4) r_ += 1 / tanh(-t_distance / 4.0);     /* TIS:314 */
            const double jr = m.sjerk[k] / P.dt;
            r_ -= jr * jr / 3600.0 * 3.0;                                        /* TIS:316 */
            if (d_distance < 10) {
                const double x = d_distance / 10, x2 = x * x;
                r_ += log(x2 * x2 * x + 0.00001);                                /* TIS:318 */
            }
            r_ += (v - P.vm) / P.aspan * 2.0;                                    /* TIS:319 */
            m.rew[g] = (float)fmin(20.0, fmax(-20.0, r_));                       /* TIS:320 */
            m.sjs[k] += fabs(jr);                                                /* TIS:321 */
            /* collision test in world space, TIS:322-334 */
            if (k0 >= 0) {
                double ax, ay, bx, by;
                pve_world_xy(P, p, d, &ax, &ay);
                pve_world_xy(P, m.sp[k0], m.lane_of[k0], &bx, &by);
                const double dx = bx - ax, dy = by - ay;
                if (sqrt(dx * dx + dy * dy) < P.thr) {
                    m.hit[g] = 1;
                    const int g0 = m.acnt[k0];
                    PVE_ATOMIC_ADD(&m.inct[g0], 1);                              /* TIS:334 */
                    if (g < g0) PVE_ATOMIC_ADD(&m.incb[g0], 1);                  /* Q4 */
                }
            }
        }
    PVE_END_TID

    /* ---- H: removal / finish flags for every vehicle, TIS:335-359 ------------------------- */
    PVE_FOR_TID(tid)
        for (int k = tid; k < V; k += NT) {
            const int g = m.ctl0[k] ? (int)m.acnt[k] : -1;
            uint32_t pk = m.spk[k];
            const int prev = (int)((pk >> 16) & 0xFFu);
            const int rep = prev + (g >= 0 ? (int)m.hit[g] + m.incb[g] : 0);    /* seen at its turn */
            const int fin = prev + (g >= 0 ? (int)m.hit[g] + m.inct[g] : 0);    /* end of tick */
            uint32_t fl = pk >> 24;
            const double p = m.sp[k];
            if (g >= 0) { m.cpv[g] = rep; if (rep > 0) PVE_ATOMIC_ADD(&m.misc[M_COLL], rep); }   /* TIS:337-339 */
            if (g >= 0 && rep > 0) PVE_ATOMIC_ADD(&m.misc[M_COLLAG], 1);
            if (p < P.remove_p || rep > 0) {                                     /* TIS:341 */
                if (rep > 0) {
                    const int tgt = (g >= 0) ? g : (int)m.acnt[k] - 1;          /* reward[-1], TIS:346 */
                    if (tgt >= 0) m.q5[tgt] = 1; else PVE_ATOMIC_ADD(&m.misc[M_Q5U], 1);
                }
                m.del[k] = 1;                                                    /* TIS:348 */
                PVE_ATOMIC_ADD(&m.misc[M_NREM], 1);
                if (g >= 0) { m.status[g] = PVE_ST_DONE | PVE_ST_REMOVED; m.hdra[g] = -1; }   /* TIS:347-349 */
            } else if (p < 0 && (fl & PVE_F_CONTROL)) {                          /* TIS:350 */
                fl = (fl & ~(uint32_t)(PVE_F_CONTROL | PVE_F_LOCK)) | PVE_F_FINISH;      /* TIS:351-355 */
                m.status[g] = PVE_ST_DONE | PVE_ST_FINISHED;
                m.hdra[g] = -1;
                m.fin5[g] = 1;                                                   /* TIS:357 */
                PVE_ATOMIC_ADD(&m.misc[M_PASSED], 1);                            /* TIS:356 */
                PVE_ATOMIC_ADD(&m.misc[M_PSTEP], (int)(pk & 0xFFFFu));           /* TIS:359 */
            }
            const uint32_t c8 = fin > 255 ? 255u : (uint32_t)fin;
            m.spk[k] = (pk & 0xFFFFu) | (c8 << 16) | (fl << 24);
            if (!m.del[k] && (fl & PVE_F_CONTROL)) PVE_ATOMIC_ADD(&m.misc[M_NCTRL], 1);
        }
    PVE_END_TID

    /* ---- I: deadlock scan, TIS:365-370 + 1469-1499 ---------------------------------------- */
    PVE_FOR_TID(tid)
        for (int g = tid; g < A; g += NT) {
            const int k = m.vidx[g];
            if (!((m.spk[k] >> 24) & PVE_F_CONTROL) || m.del[k]) continue;
            int t = g, len = 0;
            for (int hop = 1; hop <= 10; ++hop) {                                /* TIS:1470-1478 */
                t = m.hdra[t];
                if (t < 0) break;
                if (t == g) { len = hop; break; }
            }
            if (len == 0) continue;
            m.slock[k] = 1;                                                      /* TIS:1482 */
            int mn = g;
            t = g;
            for (int hop = 0; hop < len; ++hop) { t = m.hdra[t]; mn = t < mn ? t : mn; }
            if (mn != g) continue;          /* the first member in (lane, j) order reports the ring */
            PVE_ATOMIC_ADD(&m.misc[M_LOCK], 1);
            double rd[10];
            int ro[10], rt[10];
            int nrec = 0;
            t = g;
            for (int hop = 0; hop < len; ++hop) {                                /* TIS:1481-1490 */
                const int nx = m.hdra[t];
                rd[nrec] = m.virdis[t]; ro[nrec] = t; rt[nrec] = nx; ++nrec;
                t = nx;
            }
            /* record_.sort(): lexicographic on (vir_dis, o_lane, o_j, t_lane, t_j); agent index
             * order equals (lane, j) order, TIS:1492 */
            for (int x = 1; x < nrec; ++x) {
                const double dx_ = rd[x]; const int ox = ro[x], tx_ = rt[x];
                int y = x - 1;
                while (y >= 0 && (dx_ < rd[y] || (dx_ == rd[y] && (ox < ro[y] || (ox == ro[y] && tx_ < rt[y]))))) {
                    rd[y + 1] = rd[y]; ro[y + 1] = ro[y]; rt[y + 1] = rt[y]; --y;
                }
                rd[y + 1] = dx_; ro[y + 1] = ox; rt[y + 1] = tx_;
            }
            double sum = 0;
            for (int x = 0; x < nrec; ++x) sum = sum + rd[x];
            if (rd[0] < P.thr || sum / (double)nrec < P.thr + 3) {               /* TIS:1495 */
                m.slocka[m.vidx[ro[0]]] = 1;                                     /* TIS:1496 */
                m.slocka[m.vidx[rt[0]]] = -1;                                    /* TIS:1497 */
            }
        }
    PVE_END_TID

    /* ---- removal by stream compaction (TIS:435-444): survivors keep their order ----------- */
    PVE_FOR_TID(tid)
        for (int k = tid; k < V; k += NT) m.fbits[k] = m.del[k] ? 0 : 1;
        for (int g = tid; g < A; g += NT) {
            m.virdis[g] = m.fin5[g] ? m.sjs[m.vidx[g]] : 0.0;                    /* TIS:358 */
            /* reward[-1] overrides in processing order: a later -10 beats the agent's own +5 (Q5) */
            if (m.q5[g]) m.rew[g] = -10.f;                                       /* TIS:346 */
            else if (m.fin5[g]) m.rew[g] = 5.f;                                  /* TIS:357 */
        }
    PVE_END_TID
    pve_block_excl_scan<NT>(m.fbits, m.surv, V, m.wsum);
    pve_block_sums<NT>(m.rew, m.virdis, A, m.dsum);

    /* ---- J: arrivals (TIS:378-433) and header update, one thread --------------------------- */
    PVE_FOR_TID(tid)
        if (tid == 0) {
            pve_env_header *h = m.hdr;
            const int tick = h->tick;
            int total = m.surv[V], nctrl = m.misc[M_NCTRL], nsp = 0;
            /* head of each rebuilt virtual lane, read by next tick's step() (Q2) */
            for (int d = 0; d < PVE_NLANE; ++d)
                if (h->lane_n[d] > 0) {
                    if (m.vl_cnt[d] > 0) {
                        const int kh = m.sidx[m.vl_base[d]];
                        h->head_lane[d] = (int8_t)m.lane_of[kh];
                        h->head_j[d] = (uint8_t)(kh - m.lane_off[m.lane_of[kh]]);
                    } else { h->head_lane[d] = -1; h->head_j[d] = 0; }
                }
            for (int i = 0; i < PVE_NLANE; ++i) {
                m.misc[M_SPREF0 + i] = nsp;
                const int surv_i = (int)m.surv[m.lane_off[i + 1]] - (int)m.surv[m.lane_off[i]];
                int sp = 0;
                if (tick >= h->next_spawn[i]) {                                  /* TIS:379 */
                    if (total + nsp < VC && nctrl + nsp < P.AC && surv_i < 255) sp = 1;
                    else h->overflow += 1;
                }
                m.misc[M_SPAWN0 + i] = sp;
                if (sp) {
                    const int rec = (int)h->veh_rec[i] + 1;                      /* TIS:430 */
                    h->veh_rec[i] = (uint16_t)rec;
                    h->next_spawn[i] = (rec < P.K) ? spawn_tick[((size_t)b * P.K + rec) * PVE_NLANE + i]
                                                   : PVE_NEVER;
                    nsp += 1;
                }
                h->lane_n[i] = (uint8_t)(surv_i + sp);
            }
            m.misc[M_SPREF0 + PVE_NLANE] = nsp;
            m.misc[M_SURV] = total;
            h->passed_veh += m.misc[M_PASSED];
            h->passed_step_total += m.misc[M_PSTEP];
            h->n_veh = total + nsp;
            h->n_ctrl = nctrl + nsp;
        }
    PVE_END_TID

    /* ---- K: write the state back, compacted ------------------------------------------------ */
    float *row0_next = S.row0[phase ^ 1] + vbase * PVE_OBS_W;
    PVE_FOR_TID(tid)
        for (int k = tid; k < V; k += NT)
            if (!m.del[k]) {
                const size_t o = vbase + (size_t)((int)m.surv[k] + m.misc[M_SPREF0 + m.lane_of[k]]);
                S.p[o] = m.sp[k]; S.v[o] = m.sv[k]; S.a[o] = m.sa[k]; S.js[o] = m.sjs[k];
                pve_veh_meta mt;
                mt.uid = m.suid[k];
                mt.packed = m.spk[k] | ((uint32_t)(m.slock[k] ? PVE_F_LOCK : 0) << 24)
                            | ((uint32_t)(m.slocka[k] + 1) << 27);
                S.meta[o] = mt;
            }
        if (tid < PVE_NLANE && m.misc[M_SPAWN0 + tid]) {                         /* TIS:395-427 */
            const int i = tid;
            const int np = (int)m.surv[m.lane_off[i + 1]] + m.misc[M_SPREF0 + i];
            const size_t o = vbase + (size_t)np;
            S.p[o] = P.spawn_p[i % 3]; S.v[o] = P.v0; S.a[o] = 0.0; S.js[o] = 0.0;
            pve_veh_meta mt;
            mt.uid = m.hdr->id_seq + m.misc[M_SPREF0 + i];
            mt.packed = ((uint32_t)PVE_F_CONTROL << 24) | (1u << 27);
            S.meta[o] = mt;
            pve_v4 z; z.x = 0; z.y = 0; z.z = 0; z.w = 0;
            for (int q = 0; q < PVE_OBS_W / 4; ++q) ((pve_v4 *)(row0_next + (size_t)np * PVE_OBS_W))[q] = z;
        }
        /* stored row 0 of every surviving agent -> next tick's neighbour rows / actor input */
        for (int it = tid; it < A * (PVE_OBS_W / 4); it += NT) {
            const int g = it / (PVE_OBS_W / 4), q = it - g * (PVE_OBS_W / 4);
            const int k = m.vidx[g];
            if (!m.del[k]) {
                const int np = (int)m.surv[k] + m.misc[M_SPREF0 + m.lane_of[k]];
                ((pve_v4 *)(row0_next + (size_t)np * PVE_OBS_W))[q] = ((const pve_v4 *)(m.row0 + (size_t)g * PVE_OBS_W))[q];
            }
        }
    PVE_END_TID

    PVE_FOR_TID(tid)
        if (tid == 0) m.hdr->id_seq += m.misc[M_SPREF0 + PVE_NLANE];           /* TIS:433 */
    PVE_END_TID

    /* ---- M: outputs ------------------------------------------------------------------------ */
    const int64_t obase = (int64_t)S.agent_offset[b];
    const bool out_ok = m.misc[M_OUTOK] != 0;
    const float *row0_prev = S.row0[phase] + vbase * PVE_OBS_W;
    PVE_FOR_TID(tid)
        if (tid < PVE_HDR_BYTES / 16) ((pve_v4 *)(S.hdr + b))[tid] = ((const pve_v4 *)m.hdr)[tid];
        if (tid == 0) {
            S.n_ctrl[b] = m.hdr->n_ctrl; S.n_veh[b] = m.hdr->n_veh;
            if (O.env_collisions) O.env_collisions[b] = m.misc[M_COLL];
            if (O.env_lock) O.env_lock[b] = m.misc[M_LOCK];
            if (O.env_removed) O.env_removed[b] = m.misc[M_NREM];
        }
        if (out_ok) {
            for (int g = tid; g < A; g += NT) {
                const int k = m.vidx[g];
                if (O.reward) O.reward[obase + g] = m.rew[g];
                if (O.ids) {
                    pve_v4 id; id.x = (uint32_t)b; id.y = m.lane_of[k];
                    id.z = (uint32_t)(k - m.lane_off[m.lane_of[k]]); id.w = (uint32_t)m.suid[k];
                    ((pve_v4 *)O.ids)[obase + g] = id;
                }
                if (O.cpv) O.cpv[obase + g] = m.cpv[g];
                if (O.status) O.status[obase + g] = m.status[g];
                if (O.jerk_sum) O.jerk_sum[obase + g] = (float)m.sjs[k];
            }
            /* 7 x 28 observation: row 0 = own row, row q+1 = neighbour q's stored row (Q3) */
            if (O.obs) {
                pve_v4 *dst = (pve_v4 *)O.obs + obase * (PVE_OBS_H * PVE_OBS_W / 4);
                const pve_v4 *prev4 = (const pve_v4 *)row0_prev;
                const pve_v4 *new4 = (const pve_v4 *)m.row0;
                for (int it = tid; it < A * 49; it += NT) {
                    const int g = it / 49, c = it - g * 49;
                    const int rr = c / 7, q = c - rr * 7;
                    pve_v4 val; val.x = 0; val.y = 0; val.z = 0; val.w = 0;
                    if (rr == 0) val = new4[g * 7 + q];
                    else {
                        const int kn = m.nn[g * PVE_NNBR + rr - 1];
                        if (kn != 0xFFFF) {
                            if (kn < (int)m.vidx[g]) val = new4[(int)m.acnt[kn] * 7 + q];   /* already updated */
                            else val = prev4[kn * 7 + q];                                    /* last tick's    */
                        }
                    }
                    dst[it] = val;
                }
            }
        }
    PVE_END_TID

    /* ---- per-intersection running statistics (end-of-rollout reduction, MAIN:407-415) ----- */
    PVE_FOR_TID(tid)
        if (tid == 0) {
            double *st = S.stats + (size_t)b * PVE_NSTAT;
            const double rs = m.dsum[0], rq = m.dsum[1];
            st[PVE_STAT_AGENT] += (double)A;
            st[PVE_STAT_VEH] += (double)V;
            st[PVE_STAT_COLL] += (double)m.misc[M_COLLAG];
            st[PVE_STAT_LOCK] += (double)m.misc[M_LOCK];
            st[PVE_STAT_JERK] += m.dsum[2];
            st[PVE_STAT_RSUM] += rs;
            st[PVE_STAT_RSQ] += rq;
            st[PVE_STAT_REMOVED] += (double)m.misc[M_NREM];
            st[PVE_STAT_STEPS] += 1.0;
            st[PVE_STAT_Q5U] += (double)m.misc[M_Q5U];
        }
    PVE_END_TID
}
