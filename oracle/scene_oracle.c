/*
 * scene_oracle.c -- CPU oracle for the PVE-MCC environment step (lane_num = 12).
 *
 * TEST INFRASTRUCTURE ONLY (see scene_oracle.h).  Sequential restatement of the reference
 * scene; every function cites the lines of /root/reference/traffic_interaction_scene.py
 * ("TIS") it follows.  It deliberately keeps the reference's processing order (lane asc,
 * j asc), its list-building and its stable sorts, so that it is an independent check of the
 * order-free formulation used by the CUDA kernels.
 *
 * Build: gcc -O2 -fPIC -shared -pthread -ffp-contract=off -fno-builtin-pow (see Makefile).
 * -ffp-contract=off keeps every a*b+c as two IEEE operations, like CPython does.
 */
#define _GNU_SOURCE          /* pthread_barrier_t */
#include "scene_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

typedef struct {
    double p, v, a, jerk, jerk_sum, vir_dis;
    double row0[ORC_OBS_W];
    int32_t collision, step, seq_in_lane, uid;
    int32_t hdr_lane, hdr_j;
    int8_t control, finish, done, lock, lock_a, del;
} veh_t;

typedef struct { double pos, v; int32_t lane, j, dir; } vl_ent;      /* [pos, lane, j, v, dir] */
typedef struct { double p; int32_t lane, j; } vq_ent;               /* self.virtual_lane item */

typedef struct {
    int32_t tick;
    double current_time;
    int32_t n[ORC_NLANE], off[ORC_NLANE];
    int32_t veh_rec[ORC_NLANE], head_lane[ORC_NLANE], head_j[ORC_NLANE];
    int32_t id_seq, passed_veh;
    int64_t passed_step_total;
    veh_t *veh;
} env_t;

typedef struct orc_scratch {          /* per-thread work arrays of env_tick */
    vq_ent *vq;
    vl_ent *vl;
    uint8_t *used;
    veh_t *tmp;
} scratch_t;
typedef struct orc_pool orc_pool;
static void pool_destroy(orc_scene *S);
static void scratch_free(scratch_t *W);

struct orc_scene {
    int32_t B, cap, K;
    orc_params prm;
    env_t *env;
    veh_t *pool;
    double *arrive;      /* [B][K][12] */
    int32_t *kvalid;     /* [B][12] */
    int32_t overflow;
    struct orc_pool *workers;          /* persistent worker threads of orc_step (null until first used) */
    int scratch_ok;
    struct orc_scratch scratch0;       /* the calling thread's scratch */
};

/* lane2lane for the 12-lane intersection, TIS:153-166 */
static const int8_t LANE2LANE[ORC_NLANE][4] = {
    {10, 3, 9, 7}, {10, 6, 3, 4}, {-1, -1, -1, -1},
    {1, 6, 0, 10}, {1, 9, 6, 7},  {-1, -1, -1, -1},
    {4, 9, 3, 1},  {4, 0, 9, 10}, {-1, -1, -1, -1},
    {7, 0, 6, 4},  {7, 3, 0, 1},  {-1, -1, -1, -1}};

static inline veh_t *VEH(env_t *e, int lane, int j) { return &e->veh[e->off[lane] + j]; }

static void set_offsets(env_t *e) {
    int o = 0;
    for (int i = 0; i < ORC_NLANE; i++) { e->off[i] = o; o += e->n[i]; }
}

static int total_veh(const env_t *e) {
    int t = 0;
    for (int i = 0; i < ORC_NLANE; i++) t += e->n[i];
    return t;
}

/* ------------------------------------------------------------------------------------------
 * TIS:1501-1539  step(i, j, eval_a)
 * ---------------------------------------------------------------------------------------- */
static void veh_step(const orc_params *P, env_t *e, int i, int j, double eval_a, vq_ent *vq, int *nvq) {
    veh_t *c = VEH(e, i, j);
    double target_a = fmin(P->aM, fmax(P->am, eval_a));                          /* 1502 */
    if (c->lock && c->lock_a != 0 && c->p > 70) target_a = c->a + c->lock_a;     /* 1503-1505 */
    c->lock = 0;                                                                 /* 1506 */
    c->lock_a = 0;                                                               /* 1507 */
    if (j > 0) {                                                                 /* 1509-1516 */
        const veh_t *f = VEH(e, i, j - 1);
        if (f->v < c->v && f->control && c->control) {
            double d_safe = c->v * 0.4 + (pow(c->v, 2) - pow(f->v, 2)) / (2 * fabs(P->am))
                            - (c->v - f->v) * P->vm / fabs(P->am);
            if (c->p - f->p < d_safe) target_a = P->am;
        }
    }
    if (e->head_lane[i] >= 0 && e->head_lane[i] == i && e->head_j[i] == j) target_a = P->aM;  /* 1517 */
    if (i % 3 == 2) target_a = P->aM;                                            /* 1519-1520 */
    target_a = fmin(P->aM, fmax(P->am, target_a));                               /* 1521 */
    c->jerk = target_a - c->a;                                                   /* 1522 */
    c->a = target_a;                                                             /* 1523 */
    c->p = c->p - c->v * P->dt - 0.5 * c->a * P->dt2;                            /* 1528-1529 */
    c->v = fmin(P->vM, fmax(c->v + c->a * P->dt, P->vm));                        /* 1530-1531 */
    c->step += 1;                                                                /* 1533 */
    if (!c->control) {
        c->v = P->v0;                                                            /* 1535 */
    } else {
        vq[*nvq].p = c->p; vq[*nvq].lane = i; vq[*nvq].j = j;                     /* 1539 */
        (*nvq)++;
    }
}

/* ------------------------------------------------------------------------------------------
 * TIS:733-803  get_virtual_distance(lane1, lane2, p1) for lane_num == 12.
 * Returns 1 and *vd when the vehicle joins lane2's virtual lane.
 * ---------------------------------------------------------------------------------------- */
static int virtual_distance(const orc_params *P, int lane1, int lane2, double p1, double *vd) {
    const double cw = P->lane_cw, thr = 0;
    const int8_t *l2l = LANE2LANE[lane2];
    double delta;
    if (lane2 % 3 == 1) {                                                        /* 733 */
        if (lane1 == l2l[0]) {
            delta = p1 - 3 * cw;                                                 /* 736 */
            if (delta > thr) { *vd = 9 * cw + delta; return 1; }                 /* 740 */
        } else if (lane1 == l2l[1]) {
            double beta_d = P->beta * 7 * cw;                                    /* 744 */
            delta = p1 - beta_d;
            if (delta > thr) { *vd = 6 * cw + P->cita + delta; return 1; }       /* 749 */
        } else if (lane1 == l2l[2]) {
            double alpha_d = P->alpha * 7 * cw;                                  /* 753 */
            delta = p1 - alpha_d;
            if (delta > thr) { *vd = 6 * cw - P->cita + delta; return 1; }       /* 758 */
        } else if (lane1 == l2l[3]) {
            delta = p1 - 9 * cw;                                                 /* 761 */
            if (delta > thr) { *vd = 3 * cw + delta; return 1; }                 /* 765 */
        } else if (p1 > 0) { *vd = p1; return 1; }                               /* 768 */
    } else if (lane2 % 3 == 0) {                                                 /* 771 */
        if (lane1 == l2l[0]) {
            delta = p1 - 6 * cw + P->cita;                                       /* 773 */
            if (delta > thr) { *vd = P->alpha * 7 * cw + delta; return 1; }      /* 777 */
        } else if (lane1 == l2l[1]) {
            delta = p1 - P->gama * 7 * cw;                                       /* 780 */
            if (delta > thr) { *vd = P->gama2 * 7 * cw + delta; return 1; }      /* 784 */
        } else if (lane1 == l2l[2]) {
            delta = p1 - P->gama2 * 7 * cw;                                      /* 787 */
            if (delta > thr) { *vd = P->gama * 7 * cw + delta; return 1; }       /* 791 */
        } else {
            delta = p1 - 6 * cw - P->cita;                                       /* 794 */
            if (delta > thr) { *vd = P->beta * 7 * cw + delta; return 1; }       /* 798 */
        }
    } else if (p1 > 0) { *vd = p1; return 1; }                                   /* 801 */
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * TIS:1250-1290  get_p(p, i, intention) for lane_num == 12; yaw is never read by the caller.
 * ---------------------------------------------------------------------------------------- */
static void world_xy(const orc_params *P, double p, int i, double *x, double *y) {
    const double cw = P->lane_cw;
    double tx, ty;
    if (i % 3 == 0) {
        const double L = P->lane_len[0];
        if (p > L) { tx = p - L + 6 * cw; ty = cw; }                             /* 1256 */
        else if (p > 0) {
            double r_a = (L - p) / L * 3.141593 / 2;                             /* 1259 */
            double p0x = 6 * cw, p0y = cw, prx = 6 * cw, pry = -6 * cw;
            tx = prx + (p0x - prx) * cos(r_a) - (p0y - pry) * sin(r_a);          /* 1262 */
            ty = pry + (p0y - pry) * cos(r_a) + (p0x - prx) * sin(r_a);          /* 1263 */
        } else { tx = -cw; ty = -6 * cw + p; }                                   /* 1267 */
    } else if (i % 3 == 1) {
        tx = p - 6 * cw; ty = 3 * cw;                                            /* 1270 */
    } else {
        const double L = P->lane_len[2];
        if (p > L) { tx = p - L + 6 * cw; ty = 5 * cw; }                         /* 1274 */
        else if (p > 0) {
            double r_a = (L - p) / L * 3.141593 / 2;                             /* 1277 */
            double p0x = 6 * cw, p0y = 5 * cw, prx = 6 * cw, pry = 6 * cw;
            tx = prx + (p0x - prx) * cos(r_a) + (p0y - pry) * sin(r_a);          /* 1281 */
            ty = pry + (p0y - pry) * cos(r_a) - (p0x - prx) * sin(r_a);          /* 1282 */
        } else { tx = 5 * cw; ty = 6 * cw - p; }                                 /* 1286 */
    }
    const double c = P->rot_cos[i / 3], s = P->rot_sin[i / 3];                   /* 1251 */
    *x = tx * c - ty * s;                                                        /* 1287 */
    *y = ty * c + tx * s;                                                        /* 1288 */
}

static int vl_find(const vl_ent *vl, int n, int lane, int j) {      /* list.index, first match */
    for (int t = 0; t < n; t++)
        if (vl[t].lane == lane && vl[t].j == j) return t;
    return -1;
}

/* stable insertion sort by position: `sorted(..., key=item[0])`, TIS:271 */
static void vl_sort(vl_ent *vl, int n) {
    for (int t = 1; t < n; t++) {
        vl_ent x = vl[t];
        int u = t - 1;
        while (u >= 0 && vl[u].pos > x.pos) { vl[u + 1] = vl[u]; u--; }
        vl[u + 1] = x;
    }
}

/* ------------------------------------------------------------------------------------------
 * TIS:1340-1405  virtual_lane_search_closer(i, j, vl, mode="closer", veh_num=6)
 * The first six non-ego entries of a STABLE sort by |pos - pos_ego| (TIS:1388-1397): picked
 * here by repeated selection of the smallest key, lowest list index first on ties.
 * ---------------------------------------------------------------------------------------- */
static void search_closer(env_t *e, int i, int j, const vl_ent *vl, int n, int closer[ORC_NN][2],
                          uint8_t *used) {
    int index = vl_find(vl, n, i, j);
    int cnt = 0;
    if (index >= 0) {
        veh_t *c = VEH(e, i, j);
        if (index == 0) { c->hdr_lane = -1; c->hdr_j = -1; c->vir_dis = 100; }               /* 1350 */
        else {
            c->hdr_lane = vl[index - 1].lane; c->hdr_j = vl[index - 1].j;                    /* 1353 */
            c->vir_dis = vl[index].pos - vl[index - 1].pos;                                  /* 1354 */
        }
        memset(used, 0, (size_t)n);
        while (cnt < ORC_NN) {
            int best = -1;
            double bestd = 0;
            for (int t = 0; t < n; t++) {
                if (used[t] || (vl[t].lane == i && vl[t].j == j)) continue;                  /* 1392 */
                double d = fabs(vl[t].pos - vl[index].pos) * 1;                              /* 1388 */
                if (best < 0 || d < bestd) { best = t; bestd = d; }
            }
            if (best < 0) break;
            used[best] = 1;
            closer[cnt][0] = vl[best].lane; closer[cnt][1] = vl[best].j;                     /* 1397 */
            cnt++;
        }
    }
    for (; cnt < ORC_NN; cnt++) { closer[cnt][0] = -1; closer[cnt][1] = -1; }                /* 1404 */
}

/* ------------------------------------------------------------------------------------------
 * TIS:1469-1499  check_lock(i, j)
 * ---------------------------------------------------------------------------------------- */
typedef struct { double d; int32_t o_lane, o_j, t_lane, t_j; } lock_rec;

static int lock_rec_less(const lock_rec *a, const lock_rec *b) {    /* list < list, lexicographic */
    if (a->d != b->d) return a->d < b->d;
    if (a->o_lane != b->o_lane) return a->o_lane < b->o_lane;
    if (a->o_j != b->o_j) return a->o_j < b->o_j;
    if (a->t_lane != b->t_lane) return a->t_lane < b->t_lane;
    return a->t_j < b->t_j;
}

static int check_lock(const orc_params *P, env_t *e, int i, int j) {
    int N = 10;
    int tl = i, tj = j;
    while (N) {
        N -= 1;
        const veh_t *t = VEH(e, tl, tj);
        int nl = t->hdr_lane, nj = t->hdr_j;                                     /* 1475 */
        tl = nl; tj = nj;
        if (tl == -1) break;
        if (tl == i && tj == j) {
            lock_rec rec[16];
            int nrec = 0;
            int flag = 1;
            while (flag) {                                                       /* 1481-1490 */
                veh_t *o = VEH(e, tl, tj);
                o->lock = 1;
                int ol = tl, oj = tj;
                tl = o->hdr_lane; tj = o->hdr_j;
                rec[nrec].d = o->vir_dis; rec[nrec].o_lane = ol; rec[nrec].o_j = oj;
                rec[nrec].t_lane = tl; rec[nrec].t_j = tj;
                nrec++;
                if (tl == i && tj == j) flag = 0;
            }
            for (int a = 1; a < nrec; a++) {                                     /* 1492 record_.sort() */
                lock_rec x = rec[a];
                int b = a - 1;
                while (b >= 0 && lock_rec_less(&x, &rec[b])) { rec[b + 1] = rec[b]; b--; }
                rec[b + 1] = x;
            }
            double sum = 0;
            for (int a = 0; a < nrec; a++) sum = sum + rec[a].d;                 /* 1495 sum(dis) */
            if (rec[0].d < P->collision_thr || sum / (double)nrec < P->collision_thr + 3) {
                VEH(e, rec[0].o_lane, rec[0].o_j)->lock_a = 1;                   /* 1496 */
                VEH(e, rec[0].t_lane, rec[0].t_j)->lock_a = -1;                  /* 1497 */
            }
            return 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * One tick of one environment: main.py:398-407 (step for every vehicle), TIS:222-376
 * (scene_update), TIS:435-444 (delete_vehicle).
 * ---------------------------------------------------------------------------------------- */

static void env_tick(orc_scene *S, int b, const float *act, orc_outputs *out, scratch_t *W) {
    const orc_params *P = &S->prm;
    env_t *e = &S->env[b];
    const int64_t base = out->agent_offset[b];
    int nvq = 0;

    for (int i = 0; i < ORC_NLANE; i++)                                          /* main.py:398-406 */
        for (int j = 0; j < e->n[i]; j++)
            veh_step(P, e, i, j, (double)act[e->off[i] + j], W->vq, &nvq);

    e->tick += 1;
    e->current_time += P->dt;                                                    /* 223 */
    int32_t collisions = 0, nrew = 0, ndel = 0, q5 = 0, nspawn = 0;
    veh_t spawned[ORC_NLANE];
    int8_t has_spawn[ORC_NLANE];
    memset(has_spawn, 0, sizeof has_spawn);

    for (int i = 0; i < ORC_NLANE; i++) {                                        /* 233 */
        if (e->n[i] > 0) {                                                       /* 234 */
            vl_ent *vl = W->vl;
            int n = 0;
            for (int q = 0; q < nvq; q++) {                                      /* 240-270 */
                const vq_ent *it = &W->vq[q];
                const veh_t *src = VEH(e, it->lane, it->j);
                if (it->lane == i) {
                    vl[n].pos = it->p; vl[n].lane = it->lane; vl[n].j = it->j;
                    vl[n].v = src->v; vl[n].dir = i; n++;                        /* 248 */
                } else {
                    int member = 0;
                    for (int k = 0; k < 4; k++) member |= (LANE2LANE[i][k] == it->lane);      /* 259 */
                    double vd;
                    if (member && virtual_distance(P, it->lane, i, it->p, &vd)) {
                        vl[n].pos = vd; vl[n].lane = it->lane; vl[n].j = it->j;
                        vl[n].v = src->v; vl[n].dir = it->lane; n++;             /* 268 */
                    }
                }
            }
            vl_sort(vl, n);                                                      /* 271 */
            if (n > 0) { e->head_lane[i] = vl[0].lane; e->head_j[i] = vl[0].j; }
            else { e->head_lane[i] = -1; e->head_j[i] = -1; }

            for (int j = 0; j < e->n[i]; j++) {                                  /* 274 */
                veh_t *c = VEH(e, i, j);
                double t_distance = 2, d_distance = 10;                          /* 280-281 */
                if (c->control) {                                                /* 282 */
                    const int64_t g = base + nrew;
                    int closer[ORC_NN][2];
                    /* ---- get_state, TIS:1292-1338 ---- */
                    int index = vl_find(vl, n, i, j);
                    search_closer(e, i, j, vl, n, closer, W->used);              /* 1324 */
                    double *obs = out->obs + g * (ORC_OBS_H * ORC_OBS_W);
                    memset(obs, 0, sizeof(double) * ORC_OBS_H * ORC_OBS_W);
                    obs[0] = vl[index].pos; obs[1] = vl[index].v; obs[2] = c->a; obs[3] = i;  /* 1336 */
                    for (int m = 0; m < ORC_NN; m++) {                           /* 1325-1335 */
                        if (closer[m][0] != -1) {
                            int ci = vl_find(vl, n, closer[m][0], closer[m][1]);
                            const veh_t *o = VEH(e, closer[m][0], closer[m][1]);
                            double *s = obs + 4 * (m + 1);
                            s[0] = vl[ci].pos; s[1] = vl[ci].v; s[2] = o->a; s[3] = closer[m][0];  /* 1330 */
                            memcpy(obs + (m + 1) * ORC_OBS_W, o->row0, sizeof o->row0);       /* 1332 */
                        }
                    }
                    memcpy(c->row0, obs, sizeof c->row0);                        /* 288 */
                    out->ids[2 * g] = i; out->ids[2 * g + 1] = j;                /* 291 */
                    out->uid[g] = c->uid;
                    for (int m = 0; m < ORC_NN; m++) {
                        out->nn[(g * ORC_NN + m) * 2] = closer[m][0];
                        out->nn[(g * ORC_NN + m) * 2 + 1] = closer[m][1];
                    }
                    /* ---- reward, TIS:293-320 ---- */
                    const int cl = closer[0][0], cj = closer[0][1];
                    if (cl >= 0) {
                        int ic = vl_find(vl, n, cl, cj);
                        d_distance = fabs(c->p - vl[ic].pos);                    /* 300 */
                        if (d_distance != 0)
                            t_distance = (c->p - vl[ic].pos) / (c->v - VEH(e, cl, cj)->v + 0.0001);  /* 304 */
                    }
                    double r_ = 0;
                    if (0 < t_distance && t_distance < 4) r_ += 1 / tanh(-t_distance / 4.0);   /* 314 */
                    r_ -= pow(c->jerk / P->dt, 2) / 3600.0 * 3.0;                /* 316 */
                    if (d_distance < 10) r_ += log(pow(d_distance / 10, 5) + 0.00001);         /* 318 */
                    r_ += (c->v - P->vm) / (double)(P->aM - P->am) * 2.0;        /* 319 */
                    out->reward[g] = fmin(20, fmax(-20, r_));                    /* 320 */
                    c->jerk_sum += fabs(c->jerk / P->dt);                        /* 321 */
                    /* ---- collision test in world space, TIS:322-334 ---- */
                    if (cl >= 0) {
                        veh_t *o = VEH(e, cl, cj);
                        double ax, ay, bx, by;
                        world_xy(P, c->p, i, &ax, &ay);
                        world_xy(P, o->p, cl, &bx, &by);
                        d_distance = sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));     /* 328 */
                        if (fabs(d_distance) < P->collision_thr) { c->collision += 1; o->collision += 1; }
                    }
                    /* without a neighbour d_distance stays 10 (TIS:281): no hit for thr <= 10 */
                    if (c->finish) c->control = 0;                               /* 335 */
                    collisions += c->collision;                                  /* 337 */
                    out->cpv[g] = c->collision;                                  /* 339 */
                    out->status[g] = 0;
                    out->jerk_sum[g] = 0;
                    nrew++;
                }
                if (c->p < P->remove_p || c->collision > 0) {                    /* 341 */
                    if (c->collision > 0) {
                        if (nrew > 0) out->reward[base + nrew - 1] = -10;        /* 346 */
                        else q5++;
                    }
                    c->done = 1; c->del = 1; ndel++;                             /* 347-348 */
                    c->hdr_lane = -1; c->hdr_j = -1;                             /* 349 */
                } else if (c->p < 0 && c->control) {                             /* 350 */
                    c->done = 1; c->finish = 1; c->control = 0;
                    c->hdr_lane = -1; c->hdr_j = -1; c->lock = 0;
                    e->passed_veh += 1;                                          /* 356 */
                    out->reward[base + nrew - 1] = 5;                            /* 357 */
                    out->jerk_sum[base + nrew - 1] = c->jerk_sum;                /* 358 */
                    out->status[base + nrew - 1] |= 4;
                    e->passed_step_total += c->step;                             /* 359 */
                }
            }
        }
        /* ---- add_new_veh(i), TIS:378-433 ---- */
        if (e->veh_rec[i] < S->kvalid[b * ORC_NLANE + i] &&
            e->current_time >= S->arrive[((size_t)b * S->K + e->veh_rec[i]) * ORC_NLANE + i]) {   /* 379 */
            if (total_veh(e) + nspawn + 1 > S->cap) {
                S->overflow = 1;
            } else {
                veh_t *nv = &spawned[i];
                memset(nv, 0, sizeof *nv);
                nv->p = P->lane_in + P->lane_len[i % 3];                         /* 395 */
                nv->v = P->v0; nv->vir_dis = 100; nv->hdr_lane = -1; nv->hdr_j = -1;
                nv->control = 1; nv->seq_in_lane = e->veh_rec[i]; nv->uid = e->id_seq;
                has_spawn[i] = 1; nspawn++;
                e->veh_rec[i] += 1;                                              /* 430 */
                e->id_seq += 1;                                                  /* 433 */
            }
        }
    }

    /* status bits for the agents, now that every flag is final; the agents of this tick are
     * exactly the vehicles queued by step(), in the same order */
    for (int q = 0; q < nvq; q++) {
        const veh_t *c = VEH(e, W->vq[q].lane, W->vq[q].j);
        out->status[base + q] |= (c->done ? 1 : 0) | (c->del ? 2 : 0);
    }

    /* ---- deadlock scan, TIS:365-370 (newly spawned vehicles have no header: no effect) ---- */
    int32_t lock = 0;
    for (int i = 0; i < ORC_NLANE; i++)
        for (int j = 0; j < e->n[i]; j++) {
            veh_t *c = VEH(e, i, j);
            if (c->control && !c->lock)
                if (check_lock(P, e, i, j)) lock += 1;
        }

    out->collisions[b] = collisions;
    out->lock[b] = lock;
    out->n_removed[b] = ndel;
    out->q5_undefined[b] = q5;

    /* ---- delete_vehicle, TIS:435-444, then lanes re-packed with this tick's arrivals ---- */
    int w = 0;
    int32_t nn_[ORC_NLANE];
    for (int i = 0; i < ORC_NLANE; i++) {
        int cnt = 0;
        for (int j = 0; j < e->n[i]; j++) {
            veh_t *c = VEH(e, i, j);
            if (!c->del) { W->tmp[w++] = *c; cnt++; }
        }
        if (has_spawn[i]) { W->tmp[w++] = spawned[i]; cnt++; }
        nn_[i] = cnt;
    }
    memcpy(e->veh, W->tmp, sizeof(veh_t) * (size_t)w);
    memcpy(e->n, nn_, sizeof nn_);
    set_offsets(e);
}

/* ------------------------------------------------------------------------------------------ */
orc_scene *orc_create(int32_t n_envs, int32_t veh_cap, const orc_params *prm) {
    orc_scene *S = (orc_scene *)calloc(1, sizeof *S);
    if (!S) return NULL;
    S->B = n_envs; S->cap = veh_cap; S->prm = *prm;
    S->env = (env_t *)calloc((size_t)n_envs, sizeof(env_t));
    S->pool = (veh_t *)calloc((size_t)n_envs * veh_cap, sizeof(veh_t));
    S->kvalid = (int32_t *)calloc((size_t)n_envs * ORC_NLANE, sizeof(int32_t));
    if (!S->env || !S->pool || !S->kvalid) { orc_destroy(S); return NULL; }
    for (int b = 0; b < n_envs; b++) {
        S->env[b].veh = S->pool + (size_t)b * veh_cap;
        for (int i = 0; i < ORC_NLANE; i++) { S->env[b].head_lane[i] = -1; S->env[b].head_j[i] = -1; }
    }
    return S;
}

void orc_destroy(orc_scene *S) {
    if (!S) return;
    pool_destroy(S);
    if (S->scratch_ok) scratch_free(&S->scratch0);
    free(S->env); free(S->pool); free(S->arrive); free(S->kvalid); free(S);
}

int32_t orc_overflow(const orc_scene *S) { return S->overflow; }

/* veh["control"] of every slot (0 past the live count): what main.py:399-405 reads to choose between the policy
 * and 0 */
void orc_control_mask(const orc_scene *S, uint8_t *out) {
    for (int b = 0; b < S->B; b++) {
        const env_t *e = &S->env[b];
        const int V = total_veh(e);
        for (int k = 0; k < S->cap; k++) out[(size_t)b * S->cap + k] = (k < V) ? (uint8_t)e->veh[k].control : 0;
    }
}

int64_t orc_count_agents(const orc_scene *S) {
    int64_t a = 0;
    for (int b = 0; b < S->B; b++) {
        const env_t *e = &S->env[b];
        int V = total_veh(e);
        for (int k = 0; k < V; k++) a += e->veh[k].control;
    }
    return a;
}

int32_t orc_reset(orc_scene *S, const double *arrive, const int32_t *kvalid, int32_t K, int32_t warmup) {
    free(S->arrive);
    size_t n = (size_t)S->B * K * ORC_NLANE;
    S->arrive = (double *)malloc(n * sizeof(double));
    if (!S->arrive) return -1;
    memcpy(S->arrive, arrive, n * sizeof(double));
    memcpy(S->kvalid, kvalid, (size_t)S->B * ORC_NLANE * sizeof(int32_t));
    S->K = K;
    S->overflow = 0;
    for (int b = 0; b < S->B; b++) {
        env_t *e = &S->env[b];
        veh_t *keep = e->veh;
        memset(e, 0, sizeof *e);
        e->veh = keep;
        for (int i = 0; i < ORC_NLANE; i++) { e->head_lane[i] = -1; e->head_j[i] = -1; }
    }
    if (warmup) {                                                                /* TIS:214-220 */
        orc_outputs out;
        memset(&out, 0, sizeof out);
        int64_t *off = (int64_t *)calloc((size_t)S->B + 1, sizeof(int64_t));
        int32_t *z = (int32_t *)calloc((size_t)S->B * 4, sizeof(int32_t));
        scratch_t W;
        W.vq = (vq_ent *)malloc(sizeof(vq_ent) * (size_t)S->cap);
        W.vl = (vl_ent *)malloc(sizeof(vl_ent) * (size_t)S->cap);
        W.used = (uint8_t *)malloc((size_t)S->cap);
        W.tmp = (veh_t *)malloc(sizeof(veh_t) * (size_t)S->cap);
        out.agent_offset = off;
        out.collisions = z; out.lock = z + S->B; out.n_removed = z + 2 * S->B; out.q5_undefined = z + 3 * S->B;
        for (int b = 0; b < S->B; b++) {
            env_t *e = &S->env[b];
            int any = 0;
            for (int i = 0; i < ORC_NLANE; i++) any |= (S->kvalid[b * ORC_NLANE + i] > 0);
            while (any && total_veh(e) == 0) env_tick(S, b, NULL, &out, &W);
        }
        free(W.vq); free(W.vl); free(W.used); free(W.tmp); free(off); free(z);
    }
    return 0;
}

int32_t orc_set_state(orc_scene *S, const orc_state_view *in) {
    for (int b = 0; b < S->B; b++) {
        env_t *e = &S->env[b];
        e->tick = in->tick[b];
        double t = 0;
        for (int k = 0; k < e->tick; k++) t += S->prm.dt;                         /* TIS:223 */
        e->current_time = t;
        int V = 0;
        for (int i = 0; i < ORC_NLANE; i++) {
            e->n[i] = in->lane_n[b * ORC_NLANE + i];
            e->veh_rec[i] = in->veh_rec[b * ORC_NLANE + i];
            e->head_lane[i] = in->head_lane[b * ORC_NLANE + i];
            e->head_j[i] = in->head_j[b * ORC_NLANE + i];
            V += e->n[i];
        }
        if (V > S->cap) return -1;
        set_offsets(e);
        e->id_seq = in->id_seq[b];
        e->passed_veh = in->passed_veh[b];
        e->passed_step_total = in->passed_step_total[b];
        for (int k = 0; k < V; k++) {
            size_t s = (size_t)b * S->cap + k;
            veh_t *c = &e->veh[k];
            memset(c, 0, sizeof *c);
            c->p = in->p[s]; c->v = in->v[s]; c->a = in->a[s]; c->jerk_sum = in->jerk_sum[s];
            c->collision = in->collision[s]; c->step = in->step[s];
            c->seq_in_lane = in->seq_in_lane[s]; c->uid = in->uid[s];
            c->control = in->flags[s] & 1; c->finish = (in->flags[s] >> 1) & 1; c->lock = (in->flags[s] >> 2) & 1;
            c->done = c->finish;
            c->lock_a = in->lock_a[s];
            c->hdr_lane = -1; c->hdr_j = -1; c->vir_dis = 100;
            memcpy(c->row0, in->row0 + s * ORC_OBS_W, sizeof c->row0);
        }
    }
    return 0;
}

int32_t orc_get_state(const orc_scene *S, orc_state_view *out) {
    for (int b = 0; b < S->B; b++) {
        const env_t *e = &S->env[b];
        out->tick[b] = e->tick;
        for (int i = 0; i < ORC_NLANE; i++) {
            out->lane_n[b * ORC_NLANE + i] = e->n[i];
            out->veh_rec[b * ORC_NLANE + i] = e->veh_rec[i];
            out->head_lane[b * ORC_NLANE + i] = e->head_lane[i];
            out->head_j[b * ORC_NLANE + i] = e->head_j[i];
        }
        out->id_seq[b] = e->id_seq;
        out->passed_veh[b] = e->passed_veh;
        out->passed_step_total[b] = e->passed_step_total;
        int V = total_veh(e);
        for (int k = 0; k < S->cap; k++) {
            size_t s = (size_t)b * S->cap + k;
            if (k < V) {
                const veh_t *c = &e->veh[k];
                out->p[s] = c->p; out->v[s] = c->v; out->a[s] = c->a; out->jerk_sum[s] = c->jerk_sum;
                out->collision[s] = c->collision; out->step[s] = c->step;
                out->seq_in_lane[s] = c->seq_in_lane; out->uid[s] = c->uid;
                out->flags[s] = (uint8_t)((c->control ? 1 : 0) | (c->finish ? 2 : 0) | (c->lock ? 4 : 0));
                out->lock_a[s] = c->lock_a;
                memcpy(out->row0 + s * ORC_OBS_W, c->row0, sizeof c->row0);
            } else {
                out->p[s] = out->v[s] = out->a[s] = out->jerk_sum[s] = 0;
                out->collision[s] = out->step[s] = out->seq_in_lane[s] = out->uid[s] = 0;
                out->flags[s] = 0; out->lock_a[s] = 0;
                memset(out->row0 + s * ORC_OBS_W, 0, sizeof(double) * ORC_OBS_W);
            }
        }
    }
    return 0;
}

/* environments are independent (main.py creates exactly one scene, main.py:230): a persistent pool of
 * worker threads takes blocks of 4 environments from a shared counter (created on the first multi-threaded
 * orc_step, reused by every later tick, joined by orc_destroy) */
struct orc_pool {
    int n;                              /* threads including the caller */
    pthread_t *th;
    pthread_barrier_t start, end;
    orc_scene *S;
    const float *actions;
    orc_outputs *out;
    atomic_int next;
    int quit;
};

static void pool_work(orc_pool *P, scratch_t *W) {
    orc_scene *S = P->S;
    const int blk = 4;
    for (;;) {
        const int b0 = atomic_fetch_add(&P->next, blk);
        if (b0 >= S->B) break;
        for (int b = b0; b < b0 + blk && b < S->B; b++)
            env_tick(S, b, P->actions + (size_t)b * S->cap, P->out, W);
    }
}

static int scratch_init(scratch_t *W, int cap) {
    W->vq = (vq_ent *)malloc(sizeof(vq_ent) * (size_t)cap);
    W->vl = (vl_ent *)malloc(sizeof(vl_ent) * (size_t)cap);
    W->used = (uint8_t *)malloc((size_t)cap);
    W->tmp = (veh_t *)malloc(sizeof(veh_t) * (size_t)cap);
    return (W->vq && W->vl && W->used && W->tmp) ? 0 : 1;
}
static void scratch_free(scratch_t *W) { free(W->vq); free(W->vl); free(W->used); free(W->tmp); }

static void *pool_main(void *arg) {
    orc_pool *P = (orc_pool *)arg;
    scratch_t W;
    scratch_init(&W, P->S->cap);
    for (;;) {
        pthread_barrier_wait(&P->start);
        if (P->quit) break;
        pool_work(P, &W);
        pthread_barrier_wait(&P->end);
    }
    scratch_free(&W);
    return NULL;
}

static void pool_destroy(orc_scene *S) {
    orc_pool *P = S->workers;
    if (!P) return;
    P->quit = 1;
    pthread_barrier_wait(&P->start);
    for (int t = 1; t < P->n; t++) pthread_join(P->th[t], NULL);
    pthread_barrier_destroy(&P->start); pthread_barrier_destroy(&P->end);
    free(P->th); free(P);
    S->workers = NULL;
}

static orc_pool *pool_get(orc_scene *S, int n) {
    if (S->workers && S->workers->n == n) return S->workers;
    pool_destroy(S);
    orc_pool *P = (orc_pool *)calloc(1, sizeof *P);
    if (!P) return NULL;
    P->n = n; P->S = S;
    P->th = (pthread_t *)calloc((size_t)n, sizeof(pthread_t));
    pthread_barrier_init(&P->start, NULL, (unsigned)n);
    pthread_barrier_init(&P->end, NULL, (unsigned)n);
    for (int t = 1; t < n; t++) pthread_create(&P->th[t], NULL, pool_main, P);
    S->workers = P;
    return P;
}

int32_t orc_step(orc_scene *S, const float *actions, orc_outputs *out, int32_t n_threads) {
    /* dense output rows: prefix sum of the controlled-vehicle counts */
    out->agent_offset[0] = 0;
    for (int b = 0; b < S->B; b++) {
        const env_t *e = &S->env[b];
        int V = total_veh(e), a = 0;
        for (int k = 0; k < V; k++) a += e->veh[k].control;
        out->agent_offset[b + 1] = out->agent_offset[b] + a;
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > (S->B + 3) / 4) n_threads = (S->B + 3) / 4;
    if (n_threads > 1024) n_threads = 1024;
    if (!S->scratch_ok) { if (scratch_init(&S->scratch0, S->cap)) return -1; S->scratch_ok = 1; }
    if (n_threads == 1) {
        for (int b = 0; b < S->B; b++) env_tick(S, b, actions + (size_t)b * S->cap, out, &S->scratch0);
    } else {
        orc_pool *P = pool_get(S, n_threads);
        if (!P) return -1;
        P->actions = actions; P->out = out;
        atomic_store(&P->next, 0);
        pthread_barrier_wait(&P->start);
        pool_work(P, &S->scratch0);
        pthread_barrier_wait(&P->end);
    }
    return S->overflow ? 1 : 0;
}
