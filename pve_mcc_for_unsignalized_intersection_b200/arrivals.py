"""Arrival-time tables: synthetic generator and conversion to integer spawn ticks.

The reference reads ``arvTimeNewVeh`` (float64 ``[K, 12]``, per-lane ascending arrival times in
seconds, zero-padded tail) from MATLAB files (main.py:228-229, 388-389) and spawns a vehicle on
lane ``i`` when ``current_time >= arrive_time[veh_rec[i]][i]`` (traffic_interaction_scene.py:379),
where ``current_time`` is a float64 that accumulates ``+= 0.1`` once per ``scene_update``
(traffic_interaction_scene.py:223).

On the device the comparison is done on integers: ``spawn_tick[k, i]`` is the first tick ``n``
whose accumulated clock satisfies the reference's comparison, so spawn indexing is bit-exact by
construction.  Example: an arrival at 1.0 s spawns at tick 11, because ten
accumulated additions of 0.1 give 0.9999999999999999.
"""
import numpy as np

NLANE = 12
NEVER = np.int32(2**31 - 1)
MAX_TICK = 4_000_000          # arrivals later than this many ticks (111 h) are treated as NEVER


def reference_clock(n_ticks, delta_t=0.1):
    """``clock[n]`` = the reference's ``current_time`` after ``n`` scene updates (TIS:223)."""
    clock = np.zeros(n_ticks + 1, dtype=np.float64)
    # ufunc.accumulate is a strictly sequential left fold, i.e. the same additions in the same
    # order as ``t += delta_t`` (checked against a Python loop in tests/test_host_logic.py)
    np.add.accumulate(np.full(n_ticks, delta_t, dtype=np.float64), out=clock[1:])
    return clock


def to_spawn_ticks(arrive_time, delta_t=0.1, max_tick=None):
    """Convert arrival seconds ``[..., K, 12]`` to int32 spawn ticks of the same shape.

    The zero-padded tail of a lane -- its first entry that is not positive, or smaller than its predecessor -- and
    everything after it maps to ``NEVER``.  The reference would instead spawn one vehicle per tick from the padding and
    then raise ``IndexError`` at row ``K``; no shipped run reaches that point, and "table exhausted = no
    more arrivals" is the defined behaviour here.  Two EQUAL consecutive arrival times are valid arrivals: the reference
    spawns them on consecutive ticks (one vehicle per lane and tick, TIS:379), and so does the device.

    The last dimension is the number of lanes of the table: 12, or 8 / 4 for the 8- and 4-lane intersections (``lane_num``).
    """
    arr = np.asarray(arrive_time, dtype=np.float64)
    assert arr.shape[-1] in (NLANE, 8, 4), arr.shape
    K = arr.shape[-2]
    if max_tick is None:
        finite_max = float(arr.max()) if arr.size else 0.0
        max_tick = min(int(finite_max / delta_t) + 8, MAX_TICK)
    clock = reference_clock(max_tick, delta_t)
    # first n with clock[n] >= arr  (clock is strictly increasing)
    ticks = np.searchsorted(clock, arr, side="left").astype(np.int64)
    ticks = np.minimum(ticks, int(NEVER))
    ticks[ticks > max_tick] = int(NEVER)
    # padding: once a lane stops ascending, it never spawns again
    bad = arr <= 0
    if K > 1:
        bad[..., 1:, :] |= arr[..., 1:, :] < arr[..., :-1, :]
    bad = np.logical_or.accumulate(bad, axis=-2)
    ticks[bad] = int(NEVER)
    return ticks.astype(np.int32)


def synthetic_arrivals(n_envs, rate_veh_per_hour, horizon_s, seed=0, min_headway=1.0, rows=None):
    """Poisson-like arrival tables with the statistics of the shipped fixtures.

    Per lane, independently: ``headway = max(min_headway, Exp(mean = 3600 / rate))``, first
    arrival drawn the same way, cumulative sum.  Returns float64 ``[n_envs, K, 12]`` whose rows
    cover at least ``horizon_s`` seconds on every lane.
    """
    mean = 3600.0 / float(rate_veh_per_hour)
    eff = min_headway + mean * np.exp(-min_headway / mean)      # E[max(h0, X)]
    if rows is None:
        rows = int(horizon_s / eff * 1.25) + 24
    rng = np.random.Generator(np.random.Philox(key=[seed, 0x5EED]))
    while True:
        head = rng.exponential(mean, size=(n_envs, rows, NLANE))
        np.maximum(head, min_headway, out=head)
        arr = np.cumsum(head, axis=1)
        if float(arr[:, -1, :].min()) > horizon_s:
            return arr
        rows = int(rows * 1.3) + 8


def stress_arrivals(n_envs, horizon_s, headway=1.0):
    """Worst-case occupancy: every lane receives a vehicle every ``headway`` seconds."""
    rows = int(horizon_s / headway) + 4
    col = headway * np.arange(1, rows + 1, dtype=np.float64)
    return np.broadcast_to(col[None, :, None], (n_envs, rows, NLANE)).copy()
