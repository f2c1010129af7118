"""CPU oracle of the 4-lane intersection (``lane_num=4``).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product package never imports it.

A sequential restatement, in plain Python over flat per-vehicle records, of what the reference does for
``TrafficInteraction(arrive_time, 150, args, lane_num=4)`` -- every function cites the lines of
/root/reference/traffic_interaction_scene.py ("TIS") it follows.  Pure Python is enough here: a 4-lane intersection
holds ~30 vehicles and the parity tests run a few intersections for a few hundred ticks.

Parity status: PINNED against rollouts of the unmodified reference scene (tests/golden/rollout4_*.npz, minted by
tests/golden/make_golden_n3.py; checked by tests/test_oracle4_golden.py).

What differs from the 12-lane scene (oracle/scene_oracle.c):
  * a physical lane carries three routes; a vehicle's ``intention`` cycles with a scene-wide counter (TIS:387-388) and
    its route is ``direction[lane][intention]`` (TIS:73-78);
  * the virtual lane of a route also holds the vehicles of the other routes of the SAME lane that have not reached the
    junction yet (TIS:250-258);
  * agents are processed lane by lane, route by route, then by j (TIS:233-275), and for the left-turn routes
    ``get_state`` rewrites the virtual positions of one conflicting route relative to the ego and hands the rewritten
    list to the agents processed after it (TIS:286-287 with TIS:1301-1319): an order dependence of its own;
  * ``step`` tests the head of virtual lane ``i`` and the "right-turn lane" rule with the LANE index (TIS:1517-1520),
    so all of lane 2 accelerates at aM and the heads of routes 0-3 are the ones that matter.
"""
import math

NL, ND, OBS_W, NNB = 4, 12, 28, 6
DIRECTION = [[6, 7, 8], [0, 1, 2], [9, 10, 11], [3, 4, 5]]                    # TIS:73-78
LANE2LANE = [[10, 6, 9, 3, 7, 4, 8], [10, 6, 3, 4, 9, 5], [6, 10],            # TIS:58-71
             [1, 9, 0, 6, 10, 7, 11], [1, 9, 6, 7, 0, 8], [9, 1],
             [4, 0, 3, 9, 1, 10, 2], [4, 0, 9, 10, 3, 11], [0, 4],
             [7, 3, 6, 0, 4, 1, 5], [7, 3, 0, 1, 6, 2], [3, 7]]
ROUTE_LANE = {d: i for i in range(NL) for d in DIRECTION[i]}


class Veh:
    __slots__ = ("p", "v", "a", "jerk", "jerk_sum", "collision", "step", "uid", "control", "finish", "done", "lock",
                 "lock_a", "intention", "route", "row0", "hdr", "vir_dis", "delete")


class Scene4Oracle:
    # geometry tables as class attributes: oracle/scene8_oracle.py derives the 8-lane scene from this class
    NL, ND, NTYPE = NL, ND, 3
    DIRECTION, LANE2LANE = DIRECTION, LANE2LANE

    def __init__(self, vm=5, collision_thr=2, dis_ctl=150, deltaT=0.1, vM=13, am=-3, aM=3, v0=10, lane_cw=2.5):
        self.vm, self.vM, self.am, self.aM, self.v0, self.dt, self.cw = vm, vM, am, aM, v0, deltaT, lane_cw
        self.thr, self.dis_ctl = collision_thr, dis_ctl
        cw = lane_cw
        self.lane_in = dis_ctl - 2 * cw                                                      # TIS:53
        self.L = [3.1415 / 2 * 3 * cw, 4 * cw, 3.1415 / 2 * cw]                               # TIS:53-55
        self.remove_p = -dis_ctl + int((self.NL + 1) / 2) * cw                                    # TIS:341-342
        alpha = math.atan((4 - math.sqrt(2)) / (4 + math.sqrt(2)))                           # TIS:79
        alpha_ = math.atan((4 + math.sqrt(2)) / (4 - math.sqrt(2)))                          # TIS:80
        beta = math.atan(2 / math.sqrt(5))                                                   # TIS:81
        beta_ = math.atan(math.sqrt(5) / 2)                                                  # TIS:82
        gama = math.atan(1 / 2 * math.sqrt(2))                                               # TIS:83
        self.alpha, self.alpha_ = alpha, alpha_
        # get_virtual_distance (TIS:453-531) as (T, C): member iff p1 - T > 0, vd = abs(p1 - T) + C
        self.T = [[4 * cw - 3 * cw * math.cos(gama), (1.5 * 3.1415) * cw * (alpha_ / (0.5 * 3.1415)),      # TIS:456, 478
                   1.5 * 3.1415 * cw * beta / (0.5 * 3.1415), 1.5 * 3.1415 * cw * beta_ / (0.5 * 3.1415),  # TIS:468, 473
                   3 * cw * math.cos(gama), 0.0, 0.0],                                                   # TIS:483, 488, 493
                  [cw, 1.5 * 3.1415 * cw * gama / (0.5 * 3.1415),                                          # TIS:499, 504
                   1.5 * 3.1415 * cw * (0.5 * 3.1415 - gama) / (0.5 * 3.1415), 3 * cw, 0.0, 0.0],          # TIS:509, 514
                  [0.0, 0.0]]
        self.C = [[3 * cw * (0.5 * 3.1415 - gama), (1.5 * 3.1415) * cw * (alpha / (0.5 * 3.1415)),        # TIS:458, 480
                   1.5 * 3.1415 * cw * beta_ / (0.5 * 3.1415), 1.5 * 3.1415 * cw * beta / (0.5 * 3.1415),  # TIS:470, 475
                   1.5 * 3.1415 * cw * (gama / (0.5 * 3.1415)), 0.0, 0.0],                               # TIS:485
                  [3 * cw, 3 * cw * math.cos(gama), 4 * cw - 3 * cw * math.cos(gama), cw, 0.0, 0.0],       # TIS:501-516
                  [0.0, 0.0]]

    # ------------------------------------------------------------------------------------------------------------
    def reset(self, arrive, warmup=True):
        """TIS:195-220.  ``arrive``: [K][4] seconds."""
        self.arr = [[float(x) for x in row] for row in arrive]
        # rows usable per lane: the positive, non-decreasing prefix (zero padding ends a table; the reference would go on
        # spawning one vehicle per tick and then raise IndexError)
        self.kvalid = []
        for i in range(self.NL):
            k = 0
            while k < len(self.arr) and self.arr[k][i] > 0 and (k == 0 or self.arr[k][i] >= self.arr[k - 1][i]):
                k += 1
            self.kvalid.append(k)
        self.time, self.tick = 0, 0                      # TIS:196 (an int that becomes a float on the first += 0.1)
        self.lanes = [[] for _ in range(self.NL)]
        self.veh_rec = [0] * self.NL
        self.id_seq = self.passed = self.passed_steps = self.intention_re = 0
        self.vlist = [[] for _ in range(self.ND)]             # virtual_lane_4: entries [pos, lane, j, v, tag]
        self.agents = []                                 # self.virtual_lane: [p, lane, j, intention]
        if warmup:
            while not any(self.lanes) and any(self.veh_rec[i] < self.kvalid[i] for i in range(self.NL)):      # TIS:214-220
                self._scene_update()

    def control_mask(self):
        return [v.control for i in range(self.NL) for v in self.lanes[i]]

    # ------------------------------------------------------------------------------------------------------------
    def _move(self, i, j, act):
        """TIS:1501-1539 ``step``."""
        me = self.lanes[i][j]
        ta = min(self.aM, max(self.am, act))                                                  # TIS:1502
        if me.lock and me.lock_a != 0 and me.p > 70:                                          # TIS:1503-1505
            ta = me.a + me.lock_a
        me.lock, me.lock_a = False, 0                                                         # TIS:1506-1507
        if j > 0:
            fr = self.lanes[i][j - 1]
            if fr.v < me.v and fr.control and me.control:                                     # TIS:1509-1510
                d_safe = me.v * 0.4 + (pow(me.v, 2) - pow(fr.v, 2)) / (2 * abs(self.am)) \
                    - (me.v - fr.v) * self.vm / abs(self.am)                                  # TIS:1512-1514
                if me.p - fr.p < d_safe:
                    ta = self.am
        if self.vlist[i] and self.vlist[i][0][1] == i and self.vlist[i][0][2] == j:           # TIS:1517 (list of ROUTE i)
            ta = self.aM
        if i in (2, 5, 8, 11):                                                                # TIS:1519: the lane index
            ta = self.aM
        ta = min(self.aM, max(self.am, ta))                                                   # TIS:1521
        me.jerk = ta - me.a
        me.a = ta
        me.p = me.p - me.v * self.dt - 0.5 * me.a * pow(self.dt, 2)                           # TIS:1528-1529
        me.v = min(self.vM, max(me.v + me.a * self.dt, self.vm))                              # TIS:1530-1531
        me.step += 1
        if not me.control:
            me.v = self.v0                                                                    # TIS:1535
        else:
            self.agents.append([me.p, i, j, me.intention])                                    # TIS:1539

    # ------------------------------------------------------------------------------------------------------------
    def _vd(self, other_route, ego_route, p1):
        """TIS:453-531: the other vehicle's place on the ego route's virtual lane, or None."""
        r = ego_route % self.NTYPE
        k = self.LANE2LANE[ego_route].index(other_route)
        delta = p1 - self.T[r][k]
        if delta > 0:
            return abs(delta) + self.C[r][k] if self.T[r][k] != 0.0 or self.C[r][k] != 0.0 else p1
        return None

    def _xy(self, p, i, m):
        """TIS:896-1062 ``get_p`` for lane_num = 4 (the yaw is never read)."""
        cw = self.cw
        if m == 1:                                                                            # straight
            return [(-1 * p + 2 * cw, -1 * cw), (p - 2 * cw, 1 * cw), (cw, -1 * p + 2 * cw), (-1 * cw, p - 2 * cw)][i]
        Lm = self.L[m]
        if p > Lm:                                                                            # before the junction
            u = p - Lm + 2 * cw
            return [(-1 * u, -1 * cw), (1 * u, 1 * cw), (1 * cw, -1 * u), (-1 * cw, 1 * u)][i]
        if m == 0:
            if p > 0:                                                                         # left-turn arc, radius 3 cw
                b = p / (3 * cw)
                s, c = math.sin(b) * 3 * cw, math.cos(b) * 3 * cw
                return [(1 * (c - 2 * cw), 1 * (2 * cw - s)), (-1 * (c - 2 * cw), -1 * (2 * cw - s)),
                        (1 * (s - 2 * cw), -1 * (2 * cw - c)), (-1 * (s - 2 * cw), 1 * (2 * cw - c))][i]
            q = -1 * p + 2 * cw
            return [(cw, q), (-1 * cw, -1 * q), (-1 * q, cw), (1 * q, -1 * cw)][i]
        if p > 0:                                                                             # right-turn arc, radius cw
            b = p / cw
            s, c = math.sin(b) * cw, math.cos(b) * cw
            return [(-1 * (2 * cw - c), -1 * (2 * cw - s)), (1 * (2 * cw - c), 1 * (2 * cw - s)),
                    (1 * (2 * cw - s), -1 * (2 * cw - c)), (-1 * (2 * cw - s), 1 * (2 * cw - c))][i]
        q = -1 * p + 2 * cw
        return [(-1 * cw, -1 * q), (1 * cw, q), (q, -1 * cw), (-1 * q, 1 * cw)][i]

    # ------------------------------------------------------------------------------------------------------------
    def _observe(self, i, j, route):
        """TIS:1292-1338 ``get_state`` + TIS:1340-1405 ``virtual_lane_search_closer`` ("closer", 6).
        Rewrites self.vlist[route] (TIS:286-287) and returns (observation rows, six neighbours)."""
        ori = self.vlist[route]
        new = [e[:] for e in ori]
        idx = next(k for k, e in enumerate(ori) if e[1] == i and e[2] == j)
        if self.NL == 4 and route % 3 == 0:                                                   # TIS:1301-1319
            tag = self.LANE2LANE[route][1]
            k3 = 3 * self.cw
            for s, e in enumerate(ori):
                if e[4] == tag:
                    ori_p = e[0] + (self.alpha_ - self.alpha) * 3 * self.cw
                    ego = ori[idx][0]
                    if ego < ori_p:
                        far = ori_p - self.alpha_ * 3 * self.cw + self.alpha * 3 * self.cw
                        new[s][0] = far
                        if far < ego:
                            new[s][0] = ego + 1
                    else:
                        far = ori_p + self.alpha_ * 3 * self.cw - self.alpha * 3 * self.cw
                        new[s][0] = far
                        if far > ego:
                            new[s][0] = ego - 1
            del k3
        me = self.lanes[i][j]
        pe = new[idx][0]
        if idx == 0:                                                                          # TIS:1349-1354
            me.hdr, me.vir_dis = (-1, -1), 100
        else:
            me.hdr, me.vir_dis = (new[idx - 1][1], new[idx - 1][2]), new[idx][0] - new[idx - 1][0]
        order = sorted(range(len(new)), key=lambda k: abs(new[k][0] - pe))                    # TIS:1388-1390 (stable)
        near = [k for k in order if k != idx][:NNB]
        rows = [[0.0] * OBS_W for _ in range(NNB + 1)]
        first = [pe, new[idx][3], me.a, me.route]
        nbr = []
        for n in range(NNB):
            if n < len(near):
                e = new[near[n]]
                car = self.lanes[e[1]][e[2]]
                first += [e[0], e[3], car.a, car.route]                                       # TIS:1330-1331
                rows[n + 1] = list(car.row0)                                                  # TIS:1332 (Q3)
                nbr.append((e[1], e[2]))
            else:
                first += [0, 0, 0, 0]
                nbr.append((-1, -1))
        rows[0] = [float(x) for x in first]
        self.vlist[route] = new                                                               # TIS:287
        return rows, nbr

    # ------------------------------------------------------------------------------------------------------------
    def _scene_update(self):
        """TIS:222-376."""
        self.time += self.dt                                                                  # TIS:223
        self.tick += 1
        out = {"ids": [], "uid": [], "obs": [], "reward": [], "cpv": [], "nn": [], "jerks": [], "collisions": 0, "lock": 0}
        dele = []
        for i in range(self.NL):
            if self.lanes[i]:                                                                 # TIS:234
                for m, route in enumerate(self.DIRECTION[i]):
                    if route == -1:                                                           # TIS:236-237 (lane_num = 8)
                        continue
                    lst = []
                    for p, l, jj, it in self.agents:                                          # TIS:240-270
                        car = self.lanes[l][jj]
                        if l == i:
                            if self.DIRECTION[l][it] == route:
                                lst.append([p, l, jj, car.v, route])
                            elif car.p - self.L[car.intention] > 0:                           # TIS:252-258
                                lst.append([car.p - self.L[car.intention] + self.L[m], l, jj, car.v, route])
                        elif self.DIRECTION[l][it] in self.LANE2LANE[route]:
                            vd = self._vd(self.DIRECTION[l][it], route, p)
                            if vd is not None:
                                lst.append([vd, l, jj, car.v, self.DIRECTION[l][it]])
                    lst.sort(key=lambda e: e[0])                                              # TIS:271 (stable)
                    self.vlist[route] = lst
                    for j, me in enumerate(self.lanes[i]):
                        if me.intention != m:
                            continue
                        if me.control:
                            rows, nbr = self._observe(i, j, route)                            # TIS:286-287
                            me.row0 = rows[0]
                            out["ids"].append((i, j)); out["uid"].append(me.uid); out["obs"].append(rows); out["nn"].append(nbr)
                            t_dist, d_dist = 2, 10
                            c0 = nbr[0]
                            if c0[0] >= 0:                                                    # TIS:294-308
                                e = next(e for e in self.vlist[route] if e[1] == c0[0] and e[2] == c0[1])
                                d_dist = abs(me.p - e[0])
                                if d_dist != 0:
                                    t_dist = (me.p - e[0]) / (me.v - self.lanes[c0[0]][c0[1]].v + 0.0001)
                            r = 0
                            if 0 < t_dist < 4:
                                r += 1 / math.tanh(-t_dist / 4.0)                             # TIS:314
                            r -= pow(me.jerk / self.dt, 2) / 3600.0 * 3.0                     # TIS:316
                            if d_dist < 10:
                                r += math.log(pow(d_dist / 10, 5) + 0.00001)                  # TIS:318
                            r += (me.v - self.vm) / float(self.aM - self.am) * 2.0            # TIS:319
                            out["reward"].append(min(20, max(-20, r)))
                            me.jerk_sum += abs(me.jerk / self.dt)                             # TIS:321
                            if c0[0] >= 0:                                                    # TIS:322-331
                                other = self.lanes[c0[0]][c0[1]]
                                a, b = self._xy(me.p, i, me.intention), self._xy(other.p, c0[0], other.intention)
                                d_dist = math.sqrt((b[0] - a[0]) ** 2 + (b[1] - a[1]) ** 2)
                            if abs(d_dist) < self.thr:                                        # TIS:332-334
                                me.collision += 1
                                self.lanes[c0[0]][c0[1]].collision += 1
                            if me.finish:
                                me.control = False
                            out["collisions"] += me.collision                                 # TIS:337
                            out["cpv"].append(me.collision)
                        if me.p < self.remove_p or me.collision > 0:                          # TIS:341-349
                            if me.collision > 0:
                                out["reward"][-1] = -10
                            me.done = True
                            dele.append((i, j))
                            me.hdr = (-1, -1)
                        elif me.p < 0 and me.control:                                         # TIS:350-359
                            me.done = me.finish = True
                            me.control = False
                            me.hdr = (-1, -1)
                            me.lock = False
                            self.passed += 1
                            out["reward"][-1] = 5
                            out["jerks"].append(me.jerk_sum)
                            self.passed_steps += me.step
            self._spawn(i)                                                                    # TIS:361
        self.agents = []                                                                      # TIS:364
        for i in range(self.NL):                                                                   # TIS:365-370
            for j, me in enumerate(self.lanes[i]):
                if me.control and not me.lock and self._check_lock(i, j):
                    out["lock"] += 1
        out["done"] = [self.lanes[i][j].done for i, j in out["ids"]]
        out["removed"] = [(i, j) in dele for i, j in out["ids"]]
        out["n_removed"] = len(dele)
        self._delete = dele
        return out

    def _spawn(self, i):
        """TIS:378-433."""
        if self.veh_rec[i] < self.kvalid[i] and self.time >= self.arr[self.veh_rec[i]][i]:
            v = Veh()
            v.intention = self._draw_intention(i)
            v.route = self.DIRECTION[i][v.intention]
            v.p = sum([self.lane_in, self.L[v.intention]])                                    # TIS:395
            v.v, v.a, v.jerk, v.jerk_sum = self.v0, 0, 0, 0
            v.collision = v.step = 0
            v.control, v.finish, v.done, v.lock, v.lock_a = True, False, False, False, 0
            v.hdr, v.vir_dis, v.row0, v.uid = (-1, -1), 100, [0.0] * OBS_W, self.id_seq
            self.lanes[i].append(v)
            self.veh_rec[i] += 1
            self.id_seq += 1

    def _draw_intention(self, i):
        it = self.intention_re % 3                                                            # TIS:387
        self.intention_re += 1                                                                # TIS:388
        return it

    def _check_lock(self, i, j):
        """TIS:1469-1499."""
        t = (i, j)
        for _ in range(10):
            t = self.lanes[t[0]][t[1]].hdr
            if t[0] == -1:
                return False
            if t == (i, j):
                rec = []
                while True:
                    car = self.lanes[t[0]][t[1]]
                    car.lock = True
                    o, t = t, car.hdr
                    rec.append([car.vir_dis, o[0], o[1], t[0], t[1]])
                    if t == (i, j):
                        break
                rec.sort()
                dis = [r[0] for r in rec]
                if rec[0][0] < self.thr or sum(dis) / float(len(dis)) < self.thr + 3:         # TIS:1495
                    self.lanes[rec[0][1]][rec[0][2]].lock_a = 1
                    self.lanes[rec[0][3]][rec[0][4]].lock_a = -1
                return True
        return False

    # ------------------------------------------------------------------------------------------------------------
    def step(self, actions):
        """One tick as main.py:398-407 + 441 drives it: ``actions`` per vehicle in (lane, j) order."""
        k = 0
        for i in range(self.NL):
            for j in range(len(self.lanes[i])):
                self._move(i, j, float(actions[k]))
                k += 1
        assert k == len(actions)
        out = self._scene_update()
        for i, j in sorted(self._delete, key=lambda x: -x[1]):                                # TIS:435-444
            self.lanes[i].pop(j)
        return out

    def snapshot(self):
        vs = [v for i in range(self.NL) for v in self.lanes[i]]
        s = {k: [getattr(v, k) for v in vs] for k in ("p", "v", "a", "jerk_sum", "collision", "step", "uid", "control",
                                                      "finish", "lock", "lock_a", "intention")}
        s["row0"] = [list(v.row0) for v in vs]
        s.update(tick=self.tick, lane_n=[len(x) for x in self.lanes], veh_rec=list(self.veh_rec), id_seq=self.id_seq,
                 passed_veh=self.passed, passed_step_total=self.passed_steps, intention_re=self.intention_re,
                 head_lane=[(l[0][1] if l else -1) for l in self.vlist], head_j=[(l[0][2] if l else -1) for l in self.vlist])
        return s
