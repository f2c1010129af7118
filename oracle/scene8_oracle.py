"""CPU oracle of the 8-lane intersection (``lane_num=8``, two lanes per approach).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product package never imports it.

Derived from the 4-lane oracle (oracle/scene4_oracle.py), whose control flow is the reference's own for every
``lane_num != 12`` (TIS:233-361): what changes is the geometry -- every function below cites the lines of
/root/reference/traffic_interaction_scene.py ("TIS") it follows.

  * 16 routes: lane ``i`` carries ``direction[i] = [left, straight, -]`` (even lanes) or ``[-, straight, right]`` (odd
    lanes), TIS:136-145; ``-1`` entries are skipped (TIS:236-237);
  * a new vehicle's intention is ``intention[i][random.randint(0, 1)]`` after ``random.seed()`` (TIS:382, 389-390): the
    reference is not reproducible here, so the draw of the k-th arrival of lane i is an INPUT (``draws[k][i]`` in
    {0, 1}); the golden rollouts are minted with ``random.randint`` patched to return exactly that table;
  * get_virtual_distance TIS:537-660, get_p TIS:1061-1249; get_state has no rewrite for 8 lanes (TIS:1301);
  * ``step`` still tests ``i in [2, 5, 8, 11]`` and ``virtual_lane_4[i]`` with the LANE index (TIS:1517-1520): lanes 2
    and 5 always accelerate, and the heads that matter are those of routes 0-7.

Parity status: PINNED against rollouts of the unmodified reference scene with ``lane_num=8``
(tests/golden/rollout8_*.npz, minted by tests/golden/make_golden_n3.py; checked by tests/test_oracle4_golden.py).
"""
import math

from oracle.scene4_oracle import Scene4Oracle

DIRECTION8 = [[0, 1, -1], [-1, 2, 3], [4, 5, -1], [-1, 6, 7], [8, 9, -1], [-1, 10, 11], [12, 13, -1], [-1, 14, 15]]   # TIS:136-145
LANE2LANE8 = [[14, 4, 13, 12, 9, 10, 5], [14, 13, 8, 4, 5, 6, 12], [14, 13, 8, 4, 5, 6, 7], [14],                      # TIS:107-123
              [2, 8, 1, 0, 13, 14, 9], [2, 1, 12, 8, 9, 10, 0], [2, 1, 12, 8, 9, 10, 11], [2],
              [6, 12, 5, 4, 1, 2, 13], [6, 5, 0, 12, 13, 14, 4], [6, 5, 0, 12, 13, 14, 15], [6],
              [10, 0, 9, 8, 5, 6, 1], [10, 9, 4, 0, 1, 2, 8], [10, 9, 4, 0, 1, 2, 3], [10]]
INTENTION8 = [[0, 1], [1, 2], [0, 1], [1, 2], [0, 1], [1, 2], [0, 1], [1, 2]]                                          # TIS:125-134


def geometry8(cw, dis_ctl=150):
    """lane_in, L[3], and get_virtual_distance as (T, C1, C2) per route type (route % 4) and position k in lane2lane:
    member iff p1 - T > 0, vd = (abs(p1 - T) + C1) - C2 -- the reference's own operation order (TIS:537-660)."""
    lane_in = dis_ctl - 4 * cw                                                                # TIS:101
    L = [3.1415 / 2 * 5 * cw, 8 * cw, 3.1415 / 2 * cw]                                        # TIS:101-103
    s24 = math.sqrt(24)
    T = [[8 * cw - s24 * cw, math.atan(3 / 4) * 5 * cw, 4 * cw, math.atan(4 / 3) * 5 * cw, 4 * cw, s24 * cw, 0.0],     # TIS:542-576
         [3 * cw, 3 * cw, math.atan(3 / 4) * 5 * cw, math.atan(4 / 3) * 5 * cw, 5 * cw, 5 * cw, 0.0],                  # TIS:581-615
         [cw, cw, math.atan(1 / s24) * 5 * cw, math.atan(s24) * 5 * cw, 7 * cw, 7 * cw, 0.0],                          # TIS:621-655
         [0.0]]                                                                                                       # TIS:658
    C1 = [[math.atan(s24) * 5 * cw, math.atan(4 / 3) * 5 * cw, math.atan(4 / 3) * 5 * cw, math.atan(3 / 4) * 5 * cw,
           math.atan(3 / 4) * 5 * cw, math.atan(1 / s24) * 5 * cw, 0.0],
          [7 * cw, 5 * cw, 4 * cw, 4 * cw, 3 * cw, cw, 0.0],
          [7 * cw, 5 * cw, s24 * cw, 8 * cw, 3 * cw, cw, 0.0],                                 # k = 3: abs(d) + 8 cw - sqrt(24) cw, TIS:638
          [0.0]]
    C2 = [[0.0] * 7, [0.0] * 7, [0.0, 0.0, 0.0, s24 * cw, 0.0, 0.0, 0.0], [0.0]]
    return lane_in, L, T, C1, C2


def world_xy8(cw, L, p, i, m):
    """TIS:1061-1249 ``get_p`` for lane_num = 8 (the yaw is never read)."""
    a4 = 4 * cw
    if i % 2 == 0:                                        # lanes 0, 2, 4, 6: left turn (m = 0) or straight (m = 1)
        if m == 1:                                                                            # TIS:1082-1085, 1129-1132, ...
            return [(p - a4, 1 * cw), (-1 * cw, p - a4), (-1 * p + a4, -1 * cw), (1 * cw, -1 * p + a4)][i // 2]
        if p > L[0]:                                                                          # TIS:1066-1069, ...
            u = p - L[0] + a4
            return [(1 * u, 1 * cw), (-1 * cw, 1 * u), (-1 * u, -1 * cw), (1 * cw, -1 * u)][i // 2]
        if p > 0:
            b = p / (5 * cw)                                                                  # TIS:1071
            s, c = math.sin(b) * 5 * cw, math.cos(b) * 5 * cw
            # lane 0: delta_y = sin, delta_x = cos (TIS:1072-1075); lane 2: delta_x = sin, delta_y = cos (TIS:1119-1122)
            return [(-1 * (c - a4), -1 * (a4 - s)), (-1 * (s - a4), 1 * (a4 - c)),
                    (1 * (c - a4), 1 * (a4 - s)), (1 * (s - a4), -1 * (a4 - c))][i // 2]
        q = -1 * p + a4                                                                       # TIS:1078-1079, ...
        return [(-1 * cw, -1 * q), (1 * q, -1 * cw), (cw, q), (-1 * q, 1 * cw)][i // 2]
    # lanes 1, 3, 5, 7: straight (m = 1) or right turn (m = 2)
    if m == 1:                                                                                # TIS:1088-1091, ...
        return [(p - a4, 3 * cw), (-3 * cw, p - a4), (-1 * p + a4, -3 * cw), (3 * cw, -1 * p + a4)][i // 2]
    if p > L[2]:                                                                              # TIS:1094-1097, ...
        u = p - L[2] + a4
        return [(1 * u, 3 * cw), (-3 * cw, 1 * u), (-1 * u, -3 * cw), (3 * cw, -1 * u)][i // 2]
    if p > 0:
        b = p / cw                                                                            # TIS:1099
        s, c = math.sin(b) * cw, math.cos(b) * cw
        # lane 1: delta_y = sin, delta_x = cos (TIS:1100-1103); lane 3: delta_x = sin, delta_y = cos (TIS:1147-1150)
        return [(1 * (a4 - c), 1 * (a4 - s)), (-1 * (a4 - s), 1 * (a4 - c)),
                (-1 * (a4 - c), -1 * (a4 - s)), (1 * (a4 - s), -1 * (a4 - c))][i // 2]
    q = -1 * p + a4                                                                           # TIS:1106-1107, ...
    return [(3 * cw, q), (-1 * q, 3 * cw), (-3 * cw, -1 * q), (q, -3 * cw)][i // 2]


class Scene8Oracle(Scene4Oracle):
    NL, ND, NTYPE = 8, 16, 4
    DIRECTION, LANE2LANE = DIRECTION8, LANE2LANE8

    def __init__(self, vm=5, collision_thr=2, dis_ctl=150, deltaT=0.1, vM=13, am=-3, aM=3, v0=10, lane_cw=2.5):
        super().__init__(vm, collision_thr, dis_ctl, deltaT, vM, am, aM, v0, lane_cw)
        self.lane_in, self.L, self.T, self.C1, self.C2 = geometry8(lane_cw, dis_ctl)
        self.remove_p = -dis_ctl + int((self.NL + 1) / 2) * lane_cw                           # TIS:341-342
        self.draws = None

    def reset(self, arrive, draws=None, warmup=True):
        """``draws[k][i]`` in {0, 1}: what ``random.randint(0, 1)`` returns for the k-th arrival of lane i (TIS:390)."""
        self.draws = [[int(x) for x in row] for row in draws]
        assert len(self.draws) >= len(arrive)
        super().reset(arrive, warmup=warmup)

    def _draw_intention(self, i):
        it = INTENTION8[i][self.draws[self.veh_rec[i]][i]]                                    # TIS:390
        self.intention_re += 1                                                                # TIS:392
        return it

    def _vd(self, other_route, ego_route, p1):
        """TIS:537-660."""
        r = ego_route % 4
        k = self.LANE2LANE[ego_route].index(other_route)
        delta = p1 - self.T[r][k]
        if delta > 0:
            if self.T[r][k] == 0.0 and self.C1[r][k] == 0.0:
                return p1
            return abs(delta) + self.C1[r][k] - self.C2[r][k]
        return None

    def _xy(self, p, i, m):
        return world_xy8(self.cw, self.L, p, i, m)
