"""Synthetic workloads of BASELINE.json configs 4 and 5:
dense random-obstacle maps and query batches.  Deterministic (numpy Generator
seeds), no reference data needed."""
import numpy as np

WALLS = np.array([[-22, -22, 22, -20], [-22, 20, 22, 22], [-22, -20, -20, 20], [20, -20, 22, 20]],
                 dtype=np.float32)   # crl_kino/utils/__init__.py:18-21


def _aabb_intersect(a, b):
    """crl_kino/utils/__init__.py aabb_intersect_aabb rule, boxes as [l, b, r, t]."""
    return not (a[2] < b[0] or a[0] > b[2] or a[3] < b[1] or a[1] > b[3])


def random_rectangles(rng, n_obs=8, size_range=(3.0, 20.0), min_gap=3.0, bounds=(-20.0, 20.0)):
    """The generation rule of the reference's obs_generate (utils/__init__.py:26-56)."""
    out = []
    lo, hi = bounds
    guard = 0
    while len(out) < n_obs and guard < 100000:
        guard += 1
        l, b = rng.uniform(lo, hi - 5, 2)
        w, h = rng.uniform(size_range[0], size_range[1], 2)
        r, t = l + w, b + h
        if not 40 <= w * h <= 150:
            continue
        if r > hi or t > hi:
            continue
        box = (l, b, r, t)
        if all(not _aabb_intersect(box, (o[0] - min_gap, o[1] - min_gap, o[2] + min_gap, o[3] + min_gap))
               for o in out):
            out.append(box)
    return np.array(out, dtype=np.float64)


def dense_map(n_cells=10000, cell=0.25, seed=0):
    """cfg4 map: obs_generate-style rectangles tiled into `cell`-sized RectObs cells
    until n_cells of them exist, plus the 4 boundary walls -> float32 [n_cells + 4, 4]."""
    rng = np.random.default_rng(seed)
    cells = []
    guard = 0
    while len(cells) < n_cells and guard < 64:
        guard += 1
        for (l, b, r, t) in random_rectangles(rng, n_obs=8):
            nx = max(1, int(np.floor((r - l) / cell)))
            ny = max(1, int(np.floor((t - b) / cell)))
            xs = l + cell * np.arange(nx)
            ys = b + cell * np.arange(ny)
            gx, gy = np.meshgrid(xs, ys, indexing="ij")
            c = np.stack([gx.ravel(), gy.ravel(), gx.ravel() + cell, gy.ravel() + cell], 1)
            cells.append(c)
            if sum(len(x) for x in cells) >= n_cells:
                break
        if sum(len(x) for x in cells) >= n_cells:
            break
    cells = np.concatenate(cells)[:n_cells]
    return np.ascontiguousarray(np.concatenate([cells.astype(np.float32), WALLS]))


def random_states(rng, n, bounds=(-20.0, 20.0), moving=False):
    s = np.zeros((n, 5))
    s[:, 0] = rng.uniform(bounds[0], bounds[1], n)
    s[:, 1] = rng.uniform(bounds[0], bounds[1], n)
    s[:, 2] = rng.uniform(-np.pi, np.pi, n)
    if moving:
        s[:, 3] = rng.uniform(-0.1, 1.0, n)
        s[:, 4] = rng.uniform(-np.pi, np.pi, n)
    return s


def free_states(clearance_fn, rng, n, min_clearance=0.5, moving=False, batch=4096):
    """Rejection-sample n states whose clearance exceeds min_clearance (the
    reference's own reset rule, differential_gym.py:229,238).  clearance_fn maps
    [B, 5] float64 -> [B] clearance (numpy)."""
    out = []
    have = 0
    while have < n:
        c = random_states(rng, batch, moving=moving)
        ok = clearance_fn(c) > min_clearance
        out.append(c[ok])
        have += int(ok.sum())
    return np.concatenate(out)[:n]


def query_pairs(clearance_fn, n, seed=1, min_dist=5.0):
    """cfg4 queries: (start, goal) pairs at rest, clearance > 0.5, ||dxy|| > min_dist."""
    rng = np.random.default_rng(seed)
    starts, goals = [], []
    have = 0
    while have < n:
        s = free_states(clearance_fn, rng, n, 0.5)
        g = free_states(clearance_fn, rng, n, 0.5)
        s[:, 2:] = 0.0
        g[:, 2:] = 0.0
        ok = np.linalg.norm(s[:, :2] - g[:, :2], axis=1) > min_dist
        starts.append(s[ok]); goals.append(g[ok])
        have += int(ok.sum())
    return np.concatenate(starts)[:n], np.concatenate(goals)[:n]


def sample_stream(state_bounds, goal, n, seed, goal_sample_rate=5):
    """Pre-drawn equivalent of RRT.sample (rrt.py:146-154): n candidate states,
    each the goal with probability (goal_sample_rate + 1) / 101, else uniform."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(state_bounds[:, 0], state_bounds[:, 1], (n, 5))
    pick_goal = rng.integers(0, 101, n) <= goal_sample_rate
    u[pick_goal] = np.asarray(goal, dtype=np.float64)
    return u


DENSE_DWA = dict(dt=0.2, predict_time=1.0, v_res=0.4 / 63, w_res=0.4 * np.pi / 63)   # 64 x 64 nominal grid
